"""On-device parity report of a fast precision mode against the fp32 CUDA-core mode (which is itself pinned to the
reference modules at <= 5e-6, tests/test_gpu_parity.py): maximum logit error in the north-star form
``|a - b| / max(|b|_inf per row, 1e-3)`` and the argmax agreement rate, on a model whose output layer is calibrated so
that the classes are actually spread (with random weights every utterance lands in one class and
"100 % agreement" would say nothing).  Used by bench.py's ``parity`` object and by the GPU tests; everything runs
through the product path (no oracle import)."""
import torch

from . import synth


def pooled_features(name, feats, harden=True):
    """BatchNorm'd global-mean features [n, C] of ResNet `name` (resnet.py:55-58) for CUDA features `feats`, read
    through the fp32 path with an identity output layer: same seed and same construction order => the convolution
    weights equal those of ``build_model(name)`` (the output Linear is the last module to draw from the RNG)."""
    from . import build_model
    from .zoo import model_config
    _, cfg = model_config(name)
    C = cfg["n_feature_maps"]
    probe = build_model(name, n_labels=C, precision="fp32")
    sd = probe.state_dict()
    if harden:
        synth.harden_(sd)
    sd["layers.output.weight"].copy_(torch.eye(C))
    sd["layers.output.bias"].zero_()
    probe = probe.to(feats.device)
    with torch.no_grad():
        out = torch.cat([probe(feats[i:i + 1024]) for i in range(0, feats.shape[0], 1024)])
    return out.cpu()


def calibrated_model(name, cal_feats, precision="fp32", harden=True, spread=4.0):
    """Zoo model `name` (hardened weights) whose output layer is replaced by ``synth.calibrate_output_`` on the pooled
    features of `cal_feats`, so that argmax spreads over the classes."""
    from . import build_model
    m = build_model(name, precision=precision)
    sd = m.state_dict()
    if harden:
        synth.harden_(sd)
    synth.calibrate_output_(sd, pooled_features(name, cal_feats, harden=harden), spread=spread)
    return m.to(cal_feats.device)


def compare_logits(got, ref):
    """Both [n, L] tensors on one device -> dict of the north-star parity figures."""
    got, ref = got.double(), ref.double()
    den = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-3)
    err = ((got - ref).abs() / den).amax(dim=1)
    agree = got.argmax(1) == ref.argmax(1)
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]) / den[:, 0]
    out = {"n": int(ref.shape[0]), "max_logit_err": float(err.max()), "mean_logit_err": float(err.mean()),
           "argmax_agree": float(agree.double().mean()), "argmax_disagree_rows": int((~agree).sum()),
           "classes_hit": int(ref.argmax(1).unique().numel()),
           "min_top2_margin": float(margin.min())}
    if (~agree).any():
        out["max_margin_of_disagreeing_rows"] = float(margin[~agree].max())
    return out


def parity_report(model, audio_processor, waves, precision, ref_precision="fp32", sub_batch=2048):
    """Run `model` on CUDA waveforms [n, N] in `precision` and in `ref_precision`; -> compare_logits dict."""
    keep = model.precision
    outs = {}
    try:
        with torch.no_grad():
            for prec in (precision, ref_precision):
                model.precision = prec
                outs[prec] = torch.cat([model.forward_wave(waves[i:i + sub_batch], audio_processor)
                                        for i in range(0, waves.shape[0], sub_batch)])
    finally:
        model.precision = keep
    rep = compare_logits(outs[precision], outs[ref_precision])
    rep.update({"mode": precision, "against": ref_precision + " CUDA-core path (pinned to the reference modules)",
                "tolerance": "fp32-grade modes: max_logit_err <= 1e-3 and argmax_agree == 1.0; bf16: reported"})
    return rep
