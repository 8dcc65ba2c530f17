"""Generates tests/golden/*.npz (run in the build container, where /root/reference exists):

  mfcc_golden.npz   seeded synthetic waveforms -> features from oracle/mfcc_ref.py (the numpy
                    restatement of AudioProcessor.compute_mfccs), plus the torchaudio cross-check
                    deviation measured at generation time.  librosa is not installable here, so
                    these pin the RESTATEMENT, not librosa itself ("parity unpinned" at that
                    boundary, see oracle/mfcc_ref.py).
  model_golden.npz  for every zoo config: logits of the UNMODIFIED reference modules
                    (/root/reference/model/resnet.py, cnn.py imported through
                    oracle/reference_loader.py) on a fixed feature batch, for default-init
                    weights under the config's seed and for the hardened weights
                    (honk2_b200.synth.harden_), plus a float64 checksum of the weights so tests
                    can prove they rebuilt the same parameters.

    python -m oracle.make_golden
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from honk2_b200 import synth  # noqa: E402
from honk2_b200.zoo import MODEL_ZOO, model_config  # noqa: E402
from oracle import mfcc_ref, model_ref, reference_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_waves():
    """name -> float waveforms [n, N]; regenerated (not stored) by the tests."""
    return {
        "broadband": synth.broadband(4, seed=11),
        "speechlike": synth.speechlike(3, seed=12),
        "noisy": synth.noisy_dataset_like(2, seed=13),
        "edge": synth.edge_cases(),
        "odd_len": synth.broadband(2, N=12345, seed=14),
        "f64": synth.broadband(1, seed=15, dtype=np.float64),
    }


def weight_checksum(sd):
    return float(sum(v.double().abs().sum().item() for k, v in sorted(sd.items()) if v.is_floating_point()))


def golden_features(T=101, B=4, seed=21):
    """Feature-like inputs (2*ln(mel) range) for the model goldens."""
    w = synth.speechlike(B, N=160 * (T - 1), seed=seed) + synth.broadband(B, N=160 * (T - 1), seed=seed + 1)
    return mfcc_ref.compute_mfccs_batch(w)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    # ---------------- MFCC
    out = {}
    try:
        import torchaudio
        mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=480, hop_length=160, f_min=20.0,
                                                   f_max=4000.0, n_mels=40, power=2.0, center=True,
                                                   pad_mode="reflect", norm="slaney", mel_scale="slaney")
    except Exception:
        mel = None
    for name, waves in golden_waves().items():
        feats = mfcc_ref.compute_mfccs_batch(waves)
        out[f"{name}_feat"] = feats
        out[f"{name}_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(waves).tobytes()).digest(), np.uint8)
        if mel is not None and name in ("broadband", "speechlike", "noisy"):
            m = mel(torch.from_numpy(waves.astype(np.float32)))
            ta = (2.0 * torch.log(m)).transpose(1, 2).numpy()
            live = feats != 0  # exact-zero mel energies stay 0 in the reference (audio_processor.py:27)
            dev = (np.abs(ta - feats) / np.maximum(np.abs(feats), 1.0))[live]
            print(f"mfcc {name}: restatement vs torchaudio max scaled dev {dev.max():.3e}")
            out[f"{name}_torchaudio_dev"] = np.float64(dev.max())
    np.savez_compressed(os.path.join(GOLDEN, "mfcc_golden.npz"), **out)

    # ---------------- models (reference modules)
    feats = golden_features()
    x = torch.from_numpy(feats)
    x9 = torch.from_numpy(golden_features(T=301, B=2, seed=31))
    out = {"feats": feats, "feats_long": x9.numpy()}
    for name in MODEL_ZOO:
        kind, cfg = model_config(name)
        seed = MODEL_ZOO[name]["seed"]
        for variant in ("default", "hardened"):
            ref = reference_loader.build_model(kind, cfg, seed)
            sd = ref.state_dict()
            if variant == "hardened":
                synth.harden_(sd)
            with torch.no_grad():
                y = ref(x).numpy()
            out[f"{name}/{variant}/logits"] = y
            out[f"{name}/{variant}/wsum"] = np.float64(weight_checksum(sd))
            # restated forward must agree with the real module
            y2 = model_ref.forward(kind, sd, cfg, x).numpy()
            assert np.array_equal(y, y2) or np.allclose(y, y2, rtol=0, atol=1e-6), (name, np.abs(y - y2).max())
            if kind == "ResNet" and variant == "hardened":
                with torch.no_grad():
                    out[f"{name}/{variant}/logits_long"] = ref(x9).numpy()
            print(f"{name:18s} {variant:9s} argmax {np.argmax(y, 1).tolist()}  |logit|max {np.abs(y).max():.3f}")
    np.savez_compressed(os.path.join(GOLDEN, "model_golden.npz"), **out)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
