import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def mfcc_golden():
    return np.load(os.path.join(GOLDEN, "mfcc_golden.npz"))


@pytest.fixture(scope="session")
def model_golden():
    return np.load(os.path.join(GOLDEN, "model_golden.npz"))


@pytest.fixture(scope="session")
def golden_waves():
    from oracle.make_golden import golden_waves as gw
    return gw()


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale and nvcc is present) and load the C-ABI library."""
    from honk2_b200 import _native, build
    if build.is_stale():
        build.build()
    return _native.load()


def scaled_err(a, b):
    """max |a-b| / max(|b|, 1)  -- the MFCC tolerance form of SURVEY.md section 8d."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0


def logit_err(a, b):
    """max over rows of |a-b|_inf / max(|b|_inf per row, 1e-3) -- fp32 logits tolerance form."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-3)
    return float((np.abs(a - b) / den).max())
