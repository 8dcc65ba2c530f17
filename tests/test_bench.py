"""bench.py's driver contract: the reference arm runs anywhere (oracle/ only, never the product package), the GPU arm
refuses to run without a GPU, and (GPU) the default arm prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line_and_stays_off_the_product_package():
    # a probe appended through PYTHONSTARTUP-free means: sitecustomize is not ours to touch, so ask the child itself
    code = ("import runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', "
            "'--ref-batch', '4']; runpy.run_path(%r, run_name='__main__'); "
            "print('LOADED_PRODUCT', any(m == 'honk2_b200' or m.startswith('honk2_b200.') for m in sys.modules), file=sys.stderr)"
            % BENCH)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("res15 utterances/sec") and d["unit"] == "utterances/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["model_config"] == "res15" and d["config"]["batch_per_gpu"] == 8192
    assert "LOADED_PRODUCT False" in r.stderr, "the reference arm must not import honk2_b200"


def test_reference_arm_on_other_ranks_is_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1"])
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    r = _run(["--steps", "2", "--warmup", "3", "--batch", "592", "--no-cpu-baseline", "--no-other-configs"], timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (BASE_KEYS | {"clocks", "roofline", "e2e_pcm16", "parity", "parity_mode"}) <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "bf16" and d["gpu_launches"] >= 2 * 3
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 592 * 16000 * 4
    assert d["e2e_pcm16"]["h2d_bytes_per_step"] == 592 * 16000 * 2 and d["e2e_pcm16"]["logits_bit_identical_to_f32_path"] is True
    ro = d["roofline"]
    assert ro["bound"] == "tensor" and ro["unit"] == "TFLOP/s" and 0 < ro["frac"] < 1 and abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-9
    assert d["parity_mode"]["precision"] == "bf16x3" and d["parity_mode"]["parity"]["argmax_agree"] == 1.0
    assert d["parity_mode"]["parity"]["max_logit_err"] <= 1e-3
