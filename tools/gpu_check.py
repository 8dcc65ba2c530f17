"""Stage-by-stage parity report on a B200 (debug aid; prints one line per check and never
stops at the first failure).  Usage: python tools/gpu_check.py [--precision fp32|bf16]"""
import argparse
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import honk2_b200  # noqa: E402
from conftest import logit_err, mfcc_err, scaled_err  # noqa: E402
from honk2_b200 import AudioProcessor, find_cls, synth  # noqa: E402
from honk2_b200.zoo import MODEL_ZOO, model_config  # noqa: E402
from oracle import mfcc_ref, model_ref  # noqa: E402
from oracle.make_golden import golden_waves  # noqa: E402


def check(label, fn):
    try:
        print(f"{label:58s} {fn()}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{label:58s} EXCEPTION {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--skip-zoo", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    fe = AudioProcessor()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "mfcc_golden.npz"))

    for name, w in golden_waves().items():
        def f(w=w, name=name):
            got = fe.compute_mfccs_batch(torch.from_numpy(w.astype(np.float32)).to(dev)).cpu().numpy()
            ref = gold[f"{name}_feat"]
            e = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
            worst = np.unravel_index(np.argmax(e), e.shape)
            fe_err, n_floor, floor_ok = mfcc_err(got, ref)
            return f"resolved err {fe_err:.3e} floor bins {n_floor} ok={floor_ok} | raw scaled err max {e.max():.3e} mean {e.mean():.3e} worst@{worst} got {got[worst]:.5f} ref {ref[worst]:.5f}"
        check(f"mfcc {name}", f)

    feats = torch.from_numpy(mfcc_ref.compute_mfccs_batch(synth.noisy_dataset_like(5, seed=9)))

    def model_case(kind, cfg, label, x=feats, seed=0):
        def f():
            torch.manual_seed(seed)
            c = dict(cfg, precision=args.precision)
            m = find_cls(f"model.{kind}")(c).eval()
            sd = m.state_dict()
            synth.harden_(sd)
            ref = model_ref.forward(kind, sd, cfg, x).numpy()
            m = m.to(dev)
            with torch.no_grad():
                y = m(x.to(dev)).cpu().numpy()
            return (f"logit err {logit_err(y, ref):.3e}  argmax agree {np.mean(y.argmax(1) == ref.argmax(1)):.2f} "
                    f"|ref|max {np.abs(ref).max():.2f} finite {np.isfinite(y).all()}")
        check(label, f)

    for C_ in (45, 19):
        for n_layers in (0, 1, 2, 3, 4):
            model_case("ResNet", {"n_feature_maps": C_, "n_layers": n_layers, "use_dilation": False, "n_labels": 12},
                       f"resnet C={C_} layers={n_layers} d=1 nopool")
        for n_layers in (4, 7, 13, 16):
            model_case("ResNet", {"n_feature_maps": C_, "n_layers": n_layers, "use_dilation": True, "n_labels": 12},
                       f"resnet C={C_} layers={n_layers} dilated nopool")
        for pool in ([4, 3], [2, 2], [3, 5]):
            model_case("ResNet", {"n_feature_maps": C_, "n_layers": 2, "use_dilation": False, "n_labels": 12,
                                  "pool": pool}, f"resnet C={C_} layers=2 pool={pool}")
    if args.precision == "fp32":
        model_case("ResNet", {"n_feature_maps": 30, "n_layers": 3, "use_dilation": True, "n_labels": 5},
                   "resnet C=30 layers=3 dilated labels=5")
    if not args.skip_zoo:
        for name in MODEL_ZOO:
            kind, cfg = model_config(name)
            if args.precision == "bf16" and kind == "CNN":
                continue
            model_case(kind, cfg, f"zoo {name}", seed=MODEL_ZOO[name]["seed"])


if __name__ == "__main__":
    main()
