"""Per-launch timing of the hot path for bench.py's `roofline` object: the front-end is
bracketed with torch CUDA events, the model's layer launches by the native LaunchProfiler
(kws_model_set_profile / kws_model_profile_read)."""
import ctypes as C

import torch

from . import _native


def conv_flops(model, T, F):
    """Algorithmic FLOPs per utterance of the dominant (C->C) convolution launches:
    2 * Cout * Cin * kh * kw * Ho * Wo per layer, padded taps counted, channel padding not."""
    from .model import CNN, ResNet
    if isinstance(model, ResNet):
        ph, pw = model.pool if model.pool else (1, 1)
        H, W = T // ph, F // pw
        C_ = model.n_maps
        return model.n_layers * 2.0 * C_ * C_ * 9 * H * W, model.n_layers, "conv3x3 (C->C, dilated) + ReLU + skip + BN"
    if isinstance(model, CNN) and "conv_1" in model.layers:
        c1 = model.layers["conv_1"]
        from .torch_utils import calculate_conv_output_size, calculate_pool_output_size
        c0 = model.layers["conv_0"]
        s = calculate_conv_output_size([T, F], c0.kernel_size, stride=c0.stride)
        s = calculate_pool_output_size(s, model.layers["pool_0"].kernel_size
                                       if isinstance(model.layers["pool_0"].kernel_size, (list, tuple))
                                       else [model.layers["pool_0"].kernel_size] * 2)
        o = calculate_conv_output_size(s, c1.kernel_size, stride=c1.stride)
        f1 = 2.0 * c1.out_channels * c1.in_channels * c1.kernel_size[0] * c1.kernel_size[1] * o[0] * o[1]
        if getattr(model, "precision", "fp32") == "bf16":
            # cnn_tc_fused_kernel: conv_0 (before pooling) and conv_1 are ONE launch
            s0 = calculate_conv_output_size([T, F], c0.kernel_size, stride=c0.stride)
            f0 = 2.0 * c0.out_channels * c0.kernel_size[0] * c0.kernel_size[1] * s0[0] * s0[1]
            return f0 + f1, 1, "conv_0 + ReLU + max-pool + conv_1 + bias + ReLU in ONE launch per sub-batch"
        return f1, 1, "conv_1 + bias + ReLU"
    return 0.0, 0, "none"


def profile_layers(model, audio_processor, wave_sets, steps):
    """Run `steps` forward_wave passes with per-launch events; returns milliseconds and launch
    counts for the dominant convolution kernel, the front-end and everything else."""
    dev = wave_sets[0].device
    lib, st = model._state(dev)
    B, N = wave_sets[0].shape
    T = audio_processor.n_frames(N)
    per_utt, _, kernel = conv_flops(model, T, audio_processor.n_mels)
    _native.check(lib.kws_model_set_profile(st["handle"], 1), "kws_model_set_profile")
    fe_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    feats = torch.empty((B, T, audio_processor.n_mels), dtype=torch.float32, device=dev)
    try:
        for i in range(steps):
            w = wave_sets[i % len(wave_sets)]
            e0.record()
            audio_processor.compute_mfccs_batch(w, out=feats)   # no allocation inside the bracket
            e1.record()
            model(feats)
            torch.cuda.synchronize(dev)
            fe_ms += e0.elapsed_time(e1)
        conv_ms, other_ms = C.c_double(), C.c_double()
        conv_n, other_n = C.c_int64(), C.c_int64()
        _native.check(lib.kws_model_profile_read(st["handle"], C.byref(conv_ms), C.byref(conv_n), C.byref(other_ms),
                                                 C.byref(other_n)), "kws_model_profile_read")
    finally:
        lib.kws_model_set_profile(st["handle"], 0)
    launches = int(conv_n.value)
    path = lib.kws_model_kernel_path(st["handle"], T, audio_processor.n_mels, model._precision_id())
    path = path.decode() if path else "unknown"
    if launches == steps and kernel.startswith("conv3x3") and path in ("resnet_tc_sweep_kernel", "resnet_tc_fused_kernel"):
        kernel = ("%s: conv_0 + %d x (conv3x3 + ReLU + skip + BN) + mean + Linear in ONE launch "
                  "(FLOPs counted: the C->C convolutions)" % (path, model.n_layers))
    else:
        kernel = "%s: %s" % (path, kernel)
    return {"conv_ms": conv_ms.value, "conv_launches": launches, "other_ms": other_ms.value,
            "other_launches": int(other_n.value), "frontend_ms": fe_ms,
            "total_ms": conv_ms.value + other_ms.value + fe_ms,
            "conv_flops_per_launch": per_utt * B * steps / max(launches, 1), "conv_kernel": kernel, "kernel_path": path}
