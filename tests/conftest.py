"""Shared fixtures.  `-m "not gpu"` runs everywhere (oracle vs golden vectors, host logic,
ABI surface, gloo plumbing); `-m gpu` are the parity tests proper and need a B200."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

# BASELINE.json tolerances (relative to the magnitude of the float64 reference values).
# "tf32" is the default mode: forward contractions error-compensated (3xTF32), backward
# single-pass with rounded operands; "tf32x1" is every contraction single-pass on raw operands.
RTOL = {"fp32": 1e-5, "tf32": 2e-3, "tf32x1": 2e-3}
# forward values of a contraction: the compensated mode must reach fp32-grade accuracy (that is
# what keeps leaky_relu / maxpool routing identical to the float64 reference)
FWD_RTOL = {"fp32": 1e-5, "tf32": 2e-5, "tf32x1": 2e-3}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(autouse=True)
def set_np_seed():
    np.random.seed(0)  # the reference's autouse fixture (tests/test_smallpebble.py:33-36)
    yield
    from smallpebble_b200 import core

    core._reset_rounding()  # library-global TF32 store rounding: never leaks between tests


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def scaled_err(actual, expected) -> float:
    """max |actual - expected| / max |expected|  (the 'stated relative tolerance' metric)."""
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"shape {a.shape} != {e.shape}"
    if e.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(e))), 1e-30)
    return float(np.max(np.abs(a - e))) / scale


def rel_l2(actual, expected) -> float:
    """||actual - expected||_2 / ||expected||_2."""
    a = np.asarray(actual, dtype=np.float64).ravel()
    e = np.asarray(expected, dtype=np.float64).ravel()
    assert a.shape == e.shape
    return float(np.linalg.norm(a - e) / max(np.linalg.norm(e), 1e-30))


def assert_close(actual, expected, rtol, what=""):
    err = scaled_err(actual, expected)
    assert err <= rtol, f"{what}: scaled max error {err:.3e} > {rtol:.1e}"


@pytest.fixture(params=["fp32", "tf32", "tf32x1"])
def precision(request):
    import smallpebble_b200 as sp

    old = sp.get_precision()
    sp.set_precision(request.param)
    yield request.param
    sp.set_precision(old)
