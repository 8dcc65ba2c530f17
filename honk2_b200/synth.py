"""Seeded synthetic waveforms and weight 'hardening' shared by tests and bench
(the reference ships no data; benchmarks use synthetic 1 s / 16 kHz
clips shaped like dataset/gsc_dataset.py:166-174 output)."""
import numpy as np
import torch


def broadband(B, N=16000, seed=0, dtype=np.float32):
    """0.1 * u_b * N(0,1), u_b ~ U(0,1) per utterance."""
    rng = np.random.default_rng(seed)
    amp = rng.uniform(0.05, 1.0, size=(B, 1))
    return (0.1 * amp * rng.standard_normal((B, N))).astype(dtype)


def speechlike(B, N=16000, seed=0, sr=16000, dtype=np.float32):
    """Sum of 3-6 AM chirps (100-3800 Hz) + N(0, 0.01^2) floor, random onset/offset,
    zero-padded tail (mimics gsc_dataset.py:170 padding)."""
    rng = np.random.default_rng(seed)
    t = np.arange(N) / sr
    out = np.zeros((B, N), dtype=np.float64)
    for b in range(B):
        y = 0.01 * rng.standard_normal(N)
        for _ in range(int(rng.integers(3, 7))):
            f0, f1 = rng.uniform(100, 3800, size=2)
            am = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(2, 12) * t + rng.uniform(0, 6.28)))
            phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / t[-1])
            y += rng.uniform(0.02, 0.3) * am * np.sin(phase + rng.uniform(0, 6.28))
        on = int(rng.integers(0, N // 4))
        off = int(rng.integers(N // 2, N + 1))
        y[:on] = 0.0
        y[off:] = 0.0
        out[b] = y
    return out.astype(dtype)


def noisy_dataset_like(B, N=16000, seed=0, dtype=np.float32):
    """speechlike + 0.1 * U(0, 0.1) positive-mean noise (hey_snips_dataset.py:69,86 style)."""
    rng = np.random.default_rng(seed + 7919)
    return (speechlike(B, N, seed, dtype=np.float64) + 0.1 * rng.uniform(0, 0.1, size=(B, N))).astype(dtype)


def edge_cases(N=16000, dtype=np.float32):
    """all-zeros (-> features exactly 0), single impulse, full-scale square wave."""
    z = np.zeros(N)
    imp = np.zeros(N)
    imp[N // 3] = 1.0
    sq = np.where((np.arange(N) // 40) % 2 == 0, 1.0, -1.0)
    return np.stack([z, imp, sq]).astype(dtype)


def harden_(state_dict, seed=1234, scale=1.5):
    """In-place 'hardened' parity weights: default init, then
    running_mean ~ N(0, 0.5^2), running_var ~ U(0.5, 2), output/lin_1 bias = 0 and conv
    weights x scale, so argmax spreads over classes and BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        if k.endswith("running_mean"):
            v.copy_(0.5 * torch.randn(v.shape, generator=g))
        elif k.endswith("running_var"):
            v.copy_(0.5 + 1.5 * torch.rand(v.shape, generator=g))
        elif k in ("layers.output.bias", "layers.lin_1.bias"):
            v.zero_()
        elif "conv_" in k and k.endswith("weight"):
            v.mul_(scale)
    return state_dict


def calibrate_output_(state_dict, pooled, seed=4321, spread=4.0):
    """Make argmax parity meaningful (with random weights every
    utterance maps to one class because the pooled features barely move between utterances).
    Given pooled features [n, C] of a calibration batch (from the oracle), replace the output
    layer by W = spread * N(0,1) / std_c and b = -W @ mean, so the logits are driven by the
    per-utterance DEVIATION of the pooled features and spread over all classes."""
    g = torch.Generator().manual_seed(seed)
    w = state_dict["layers.output.weight"]
    mean = pooled.mean(0)
    std = pooled.std(0).clamp_min(1e-6)
    new_w = spread * torch.randn(w.shape, generator=g) / std / (w.shape[1] ** 0.5)
    w.copy_(new_w)
    state_dict["layers.output.bias"].copy_(-(new_w @ mean))
    return state_dict
