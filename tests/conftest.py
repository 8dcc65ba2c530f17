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


FLOOR_DB = 2.0 * np.log(1e9)   # feature distance (2*ln) of a 1e-9 power ratio


def mfcc_err(got, ref):
    """Floor-aware MFCC comparison -> (scaled error over resolved bins, number of floor bins,
    floor bins consistent?).

    A bin whose reference mel power lies more than 1e-9 below the strongest bin of the SAME
    frame is a numerical-floor bin: for exactly periodic synthetic inputs (square wave, pure
    tone with an integer number of periods per window) the reference value there is float64
    rounding noise of its own FFT (2*ln(1e-31) = -143 for the square wave), which no
    independent implementation can reproduce; SURVEY.md section 7.2/8d asks for a separate
    bound on these.  They are excluded from the 1e-4 check and instead required to be floor
    bins in the CUDA output too (at least 1e-8 below the frame maximum).  Frames that are
    exactly zero (feature 0 everywhere, audio_processor.py:27) have no floor bins."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    fmax = ref.max(axis=-1, keepdims=True)
    zero_frame = np.all(ref == 0, axis=-1, keepdims=True)
    floor = (ref < fmax - FLOOR_DB) & ~zero_frame & (ref != 0)
    e = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    err = float(e[~floor].max()) if (~floor).any() else 0.0
    ok = bool(np.all(got[floor] < (np.broadcast_to(fmax, ref.shape)[floor] - 2.0 * np.log(1e8)))) if floor.any() else True
    return err, int(floor.sum()), ok
