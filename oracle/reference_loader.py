"""CPU ORACLE support (test infrastructure, NOT product code) -- imports the UNMODIFIED honk2
model classes from /root/reference, in the build container only.

`import model` in the reference pulls in utils/__init__.py:1 -> utils/audio_processor.py:1-3,
which imports ``librosa`` and ``pcen``; neither is installed here and neither is touched by
model/resnet.py or model/cnn.py, so two empty stub modules are pre-seeded.
/root/reference does not exist on the GPU box: callers must check `available()` and fall
back to oracle/model_ref.py + the committed tests/golden fixtures.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HONK2_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "resnet.py"))


def load():
    """Return the reference's `find_cls` (utils/class_registry.py:13) with model.* registered."""
    if not available():
        raise RuntimeError(f"reference checkout not present at {REFERENCE_ROOT}")
    for name in ("librosa", "pcen"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.__honk2_stub__ = True
            sys.modules[name] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import model  # noqa: F401  (registers model.ResNet / model.CNN, model/__init__.py:1-2)
    import utils as ref_utils
    return ref_utils.find_cls


def build_model(kind, config, seed):
    """find_cls("model.<kind>")(config) under torch.manual_seed(seed), eval mode
    (run/test.py:60-64, run/run_utils.py:15-18, run/test.py:21)."""
    import torch
    find_cls = load()
    torch.manual_seed(seed)
    m = find_cls(f"model.{kind}")(config)
    m.eval()
    return m
