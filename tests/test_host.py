"""CPU-side tests of the host layer: the plugin registry, the size calculators, the model zoo
against the reference's JSON configs, the module surface (state_dict keys, parameter counts,
error behaviour without a GPU) and the C-ABI library's exported symbols."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import honk2_b200
from honk2_b200 import _native, build, find_cls, register_cls
from honk2_b200.class_registry import Registry
from honk2_b200.torch_utils import calculate_conv_output_size, calculate_pool_output_size
from honk2_b200.zoo import MODEL_ZOO, model_config
from oracle import reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = reference_loader.REFERENCE_ROOT

# parameter counts measured on the reference modules
PARAMS = {"res8": 110307, "res15": 237882, "res26": 438357, "res8_narrow": 19905, "res15_narrow": 42648,
          "res26_narrow": 78387, "hey_snips_res26": 437897, "cnn-trad-fpool3": 1376044, "cnn-trad-pool2": 493708,
          "cnn-one-fpool3": 1366754, "cnn-one-fstride4": 1320038, "cnn-one-fstride8": 861308,
          "cnn-tstride2": 950360, "cnn-tstride4": 550718, "cnn-tstride8": 374984, "cnn-tpool2": 1092600,
          "cnn-tpool3": 823384}


# ---- registry (utils/class_registry.py:4-14, utils/trie.py:4-32) -------------------------------

def test_registry_contract():
    assert find_cls("model.ResNet") is honk2_b200.ResNet
    assert find_cls("model.CNN") is honk2_b200.CNN
    assert find_cls("model.DoesNotExist") is None            # trie.py:28-29: no exception
    assert find_cls("model") is None and find_cls("model.ResNet.extra") is None
    assert find_cls("nope", default_value=7) == 7

    @register_cls("test_ns.Thing")
    class A:  # noqa: D401
        pass

    @register_cls("test_ns.Thing")   # re-registering OVERWRITES (trie.py:21)
    class B:
        pass
    assert find_cls("test_ns.Thing") is B


def test_install_into_foreign_registry():
    """The plug-in hook: our classes re-registered into another registry (honk2's own).  Only model.* by default:
    run_utils.py is imported by the training script too."""
    other = Registry()

    def other_register(identifier):
        def deco(cls):
            other.add(identifier, cls)
            return cls
        return deco
    done = honk2_b200.install_into(other_register)
    assert sorted(done) == ["model.CNN", "model.ResNet"]
    assert other.get("model.ResNet") is honk2_b200.ResNet and other.get("model.CNN") is honk2_b200.CNN
    assert other.get("metric.Acc") is None and other.get("loss_fn.ce_loss") is None
    done = honk2_b200.install_into(other_register, prefixes=("metric.", "loss_fn.", "data_loader."))
    assert sorted(done) == ["data_loader.AudioDataLoader", "loss_fn.ce_loss", "metric.Acc", "metric.PerClassAcc"]
    assert other.get("metric.Acc") is honk2_b200.metric.Acc


@pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")
def test_install_into_reference_registry_overrides_model_classes():
    ref_find = reference_loader.load()
    import utils as ref_utils
    from metric.metric_utils import MetricType, collect_metrics    # (imports, hence registers, the reference's metrics)
    touched = ("model.ResNet", "model.CNN", "metric.Acc", "metric.PerClassAcc", "loss_fn.ce_loss",
               "data_loader.AudioDataLoader")
    original = {k: ref_find(k) for k in touched}
    try:
        assert sorted(honk2_b200.install_into(ref_utils.register_cls)) == ["model.CNN", "model.ResNet"]
        kind, cfg = model_config("res8")
        m = ref_find(f"model.{kind}")(cfg)           # run/test.py:61-64
        assert isinstance(m, honk2_b200.ResNet)
        assert ref_find("metric.Acc") is original["metric.Acc"]        # metrics and loss stay the reference's
        # opting in to the device-resident metrics keeps the reference's contracts
        honk2_b200.install_into(ref_utils.register_cls, prefixes=("metric.", "loss_fn."))
        acc, pca = ref_find("metric.Acc")(), ref_find("metric.PerClassAcc")()
        assert isinstance(acc, honk2_b200.metric.Acc)
        assert acc.get_type() == MetricType.MACRO                        # run/train.py:104
        assert pca.get_type() == MetricType.MICRO
        acc.get_metric = lambda: 0.5
        pca.get_metric = lambda: {0: 1.0, 1: 0.25}
        out = collect_metrics({"Acc": acc, "PerClassAcc": pca}, ["yes", "no"])    # metric_utils.py:39-52
        assert out == {"metric_Acc": 0.5, "metric_PerClassAcc": {"yes": 1.0, "no": 0.25}}
        # the loss keeps an autograd graph when the logits have one (run/train.py:142-143)
        logits = torch.randn(4, 3, requires_grad=True)
        loss = ref_find("loss_fn.ce_loss")(logits, torch.tensor([0, 1, 2, 1]))
        loss.backward()
        assert logits.grad is not None and torch.isfinite(logits.grad).all()
        assert torch.allclose(loss, torch.nn.functional.cross_entropy(logits, torch.tensor([0, 1, 2, 1])))
    finally:
        for k, v in original.items():
            if v is not None:
                ref_utils.register_cls(k)(v)


def test_models_do_not_load_the_native_library_on_the_cpu_side(tmp_path):
    """Constructing a model, reading its state_dict and deleting it must not dlopen the library (bench.py's CPU arm
    and checkpoint tools touch only that surface)."""
    code = ("import honk2_b200, gc\n"
            "from honk2_b200 import _native\n"
            "m = honk2_b200.build_model('res8'); sd = m.state_dict(); n = m.num_params(); del m; gc.collect()\n"
            "ap = honk2_b200.AudioProcessor(); del ap; gc.collect()\n"
            "assert _native.loaded() is None\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libhonk2_b200' not in maps, 'native library was mapped'\n"
            "print('ok', n)\n")
    r = subprocess.run([os.sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok 110307"), r.stdout + r.stderr


# ---- size calculators (utils/torch_utils.py:29-65) ----------------------------------------------

@pytest.mark.parametrize("k,s", [((20, 8), (1, 1)), ((32, 8), (1, 4)), ((16, 8), (8, 1)), ((3, 3), (2, 2))])
def test_conv_output_size_matches_torch(k, s):
    out = torch.nn.Conv2d(1, 2, k, stride=s)(torch.zeros(1, 1, 101, 40))
    assert calculate_conv_output_size([101, 40], list(k), stride=list(s)) == list(out.shape[2:])


@pytest.mark.parametrize("k", [(1, 3), (2, 2), (3, 3), (2, 3), (1, 1)])
def test_pool_output_size_matches_torch(k):
    out = torch.nn.MaxPool2d(k)(torch.zeros(1, 1, 82, 33))
    assert calculate_pool_output_size([82, 33], list(k)) == list(out.shape[2:])


@pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")
def test_calculators_match_reference():
    reference_loader.load()
    import utils as ref_utils
    for size in ([101, 40], [82, 11], [25, 13]):
        for k, s in (([10, 4], [1, 1]), ([5, 4], [2, 1]), ([3, 3], 1)):
            assert calculate_conv_output_size(size, k, stride=s) == ref_utils.calculate_conv_output_size(size, k, stride=s)
            assert calculate_pool_output_size(size, k) == ref_utils.calculate_pool_output_size(size, k)


# ---- zoo vs the reference's shipped configs -----------------------------------------------------

REF_CONFIGS = {"res8": "resnet/res8.json", "res8_narrow": "resnet/res8_narrow.json", "res15": "resnet/res15.json",
               "res15_narrow": "resnet/res15_narrow.json", "res26": "resnet/res26.json",
               "res26_narrow": "resnet/res26_narrow.json", "hey_snips_res26": "hey_snips/res26.json",
               "dev": "gsc_dev_config.json"}
REF_CONFIGS.update({n: f"cnn/{n}.json" for n in MODEL_ZOO if n.startswith("cnn-")})


@pytest.mark.skipif(not reference_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("name", sorted(REF_CONFIGS))
def test_zoo_equals_reference_json(name):
    with open(os.path.join(REF, "config", REF_CONFIGS[name])) as f:
        ref = json.load(f)
    assert MODEL_ZOO[name]["name"] == ref["model"]["name"]
    assert MODEL_ZOO[name]["config"] == ref["model"]["config"]
    assert MODEL_ZOO[name]["seed"] == ref["seed"]
    ds = ref[ref["dataset"]["name"]] if isinstance(ref.get("dataset"), dict) and "name" in ref["dataset"] else None
    if ds is not None and "target_class" in ds:   # run/test.py:54-58
        n = len(ds["target_class"]) + int(ds.get("unknown_class", False)) + int(ds.get("silence_class", False))
        assert MODEL_ZOO[name]["n_labels"] == n


# ---- module surface ------------------------------------------------------------------------------

@pytest.mark.parametrize("name", sorted(PARAMS))
def test_param_counts(name):
    m = honk2_b200.build_model(name)
    assert m.num_params() == PARAMS[name]
    assert m.num_trainable_params() == PARAMS[name]


def test_state_dict_keys_and_shapes():
    m = honk2_b200.build_model("res15")
    sd = m.state_dict()
    assert len(sd) == 55
    assert sd["layers.conv_0.weight"].shape == (45, 1, 3, 3)
    assert sd["layers.conv_13.weight"].shape == (45, 45, 3, 3)
    assert sd["layers.bn_13.running_var"].shape == (45,) and "layers.bn_13.weight" not in sd
    assert sd["layers.output.weight"].shape == (12, 45) and sd["layers.output.bias"].shape == (12,)
    assert m.layers["conv_13"].dilation == (16, 16) and m.layers["conv_13"].padding == (16, 16)
    c = honk2_b200.build_model("cnn-trad-fpool3").state_dict()
    assert c["layers.lin_0.weight"].shape == (32, 37376) and c["layers.conv_1.weight"].shape == (64, 64, 10, 4)
    assert "layers.dnn_1.weight" not in c and c["layers.lin_1.weight"].shape == (12, 128)
    print(m)   # run/test.py:73


def test_missing_config_key_raises_keyerror():
    with pytest.raises(KeyError):
        honk2_b200.ResNet({"n_layers": 6, "use_dilation": False, "n_labels": 12})   # resnet.py:14
    with pytest.raises(KeyError):
        honk2_b200.CNN({"time": 101, "frequency": 40, "n_labels": 12})              # cnn.py:20


def test_hey_snips_avg_pool_key_is_ignored():
    m = honk2_b200.build_model("hey_snips_res26")
    assert "pool" not in m.layers and m.pool is None     # resnet.py:29 looks for "pool" only
    assert m.layers["conv_24"].dilation == (128, 128)


def test_no_cpu_fallback():
    m = honk2_b200.build_model("res8")
    with pytest.raises(honk2_b200.NativeError):
        m(torch.zeros(2, 101, 40))
    with pytest.raises(honk2_b200.NativeError):
        honk2_b200.AudioProcessor().compute_mfccs_batch(torch.zeros(2, 16000))


def test_audio_processor_signature():
    ap = honk2_b200.AudioProcessor()     # constructed with no arguments, audio_data_loader.py:14
    assert (ap.sr, ap.n_mels, ap.f_max, ap.f_min, ap.n_fft, ap.hop_length) == (16000, 40, 4000, 20, 480, 160)
    assert honk2_b200.AudioProcessor(f_max=None).f_max == 8000
    assert ap.n_frames(16000) == 101 and ap.n_frames(144000) == 901
    with pytest.raises(ValueError):
        ap.compute_mfccs(np.zeros(16000, dtype=np.int16))
    with pytest.raises(NotImplementedError):
        ap.compute_pcen(np.zeros(10, dtype=np.float32))


def test_stream_window_targets_match_reference_bookkeeping():
    """StreamingDataset.__getitem__ (dataset_utils.py:47-95), restated in oracle/stream_ref.py: window slices and
    majority-vote targets (lowest label wins ties) for random labelled segment streams."""
    from honk2_b200.streaming import n_stream_windows, stream_window_targets
    from oracle import stream_ref
    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(80):
        n_labels = int(rng.integers(2, 6))
        n_seg = int(rng.integers(3, 10))
        lens = rng.integers(1, 50, n_seg)
        labs = rng.integers(0, n_labels, n_seg)
        window = int(rng.integers(4, 40))
        shift = int(rng.integers(1, min(10, window) + 1))
        n = n_stream_windows(int(lens.sum()), window, shift)
        assert n == max(0, int((int(lens.sum()) - window) / shift))       # dataset_utils.py:31
        if n <= 0:
            continue
        segs = [rng.standard_normal(l) for l in lens]
        stream = np.concatenate(segs)
        ref = list(stream_ref.iterate_windows(segs, list(labs), n_labels, window, shift, n))
        got = stream_window_targets(lens, labs, n_labels, window, shift)
        assert [t for _, t in ref] == list(got)
        for k, (w, _) in enumerate(ref):
            assert np.array_equal(w, stream[k * shift:k * shift + window])
        checked += 1
    assert checked >= 40
    # ties: equal halves -> the lower label index (strict ">" while enumerating, dataset_utils.py:76-79)
    assert list(stream_window_targets([4, 4, 8], [3, 1, 2], 4, 8, 4)) == [1, 1]
    with pytest.raises(ValueError):
        stream_window_targets([4, 4], [0, 5], 4, 4, 2)
    with pytest.raises(ValueError):
        stream_window_targets([4, 4], [0, 1], 2, 4, 1, n_windows=100)


def test_stream_frontend_has_no_cpu_path():
    ap = honk2_b200.AudioProcessor()
    assert ap.n_stream_windows(16000 * 10, 16000, 160) == 900
    with pytest.raises(honk2_b200.NativeError):
        ap.compute_mfccs_stream(torch.zeros(40000))


# ---- the C-ABI library ---------------------------------------------------------------------------

def header_symbols():
    text = open(os.path.join(ROOT, "include", "honk2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kws_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(native_lib, n), f"{n} is declared in include/honk2_b200.h but not exported"
    assert set(names) == set(_native.SIGNATURES), "ctypes table and header disagree"
    assert native_lib.kws_abi_version() == _native.ABI_VERSION == 2


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_argument_errors_without_gpu(native_lib):
    """Pure argument validation paths return KWS_ERR_INVALID and set the error text."""
    assert native_lib.kws_model_forward(None, None, 1, 101, 40, None, 0, None, 0, None) == 1
    assert b"model is null" in native_lib.kws_last_error()
    assert native_lib.kws_frontend_n_frames(None, 16000) == 0
    assert native_lib.kws_model_workspace_bytes(None, 1, 101, 40, 0) == 0
    assert native_lib.kws_mfcc_stream_scratch_bytes(None, 10, 16000, 160) == 0
    assert native_lib.kws_mfcc_stream_forward(None, None, 1, 16000, 160, None, None, 0, None) == 1
    assert b"frontend is null" in native_lib.kws_last_error()
    out = ctypes.c_void_p()
    assert native_lib.kws_frontend_create(16000, 40, 20.0, 4000.0, 512, 160, ctypes.byref(out)) == 1
    assert b"n_fft=480" in native_lib.kws_last_error()


def test_reference_checkpoint_loading(tmp_path):
    """utils/workspace.py:28-70: a checkpoint dict with pickled extras and DataParallel's `module.` prefix."""
    import torch
    import honk2_b200
    from honk2_b200 import load_checkpoint, strip_data_parallel_prefix
    src = honk2_b200.build_model("res8", seed=7)
    sd = {k: v.clone() + (0.25 if v.is_floating_point() else 0) for k, v in src.state_dict().items()}
    ckpt = {"model_state_dict": {"module." + k: v for k, v in sd.items()}, "loss_fn": len, "metrics": {"acc": object},
            "optimizer_state_dict": {"state": {}}, "lr_scheduler_state_dict": {}, "epoch": 17}
    path = tmp_path / "best_model.pt"
    torch.save(ckpt, path)
    dst = honk2_b200.build_model("res8", seed=1)
    rest = load_checkpoint(dst, str(path))
    assert rest["epoch"] == 17 and "model_state_dict" not in rest and "optimizer_state_dict" not in rest
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # bare state_dict, no prefix; a partial prefix is left alone (strict loading then fails loudly)
    load_checkpoint(dst, dict(src.state_dict()))
    mixed = {"module.a": 1, "b": 2}
    assert strip_data_parallel_prefix(mixed) is mixed
    import pytest
    with pytest.raises(RuntimeError):
        load_checkpoint(dst, {"model_state_dict": {"module.layers.conv_0.weight": sd["layers.conv_0.weight"]}})
