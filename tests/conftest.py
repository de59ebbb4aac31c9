import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
# The parity tests hold eval / evalp / pdf to the reference's BITS: they run the exact tier (also inherited by the C++ test
# programs and plugin binaries the tests start).  The default 1e-5 tier has its own tests (test_gpu_parity.py::test_fast_tier_*,
# test_facade_cpp.py), which switch modes explicitly.
os.environ.setdefault("DJB200_PRECISION", "bits")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import api
    return api.PortOracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import api
    if not api.ref_available():
        if not api.build_ref():
            pytest.skip("reference not available (no /root/reference and no prebuilt oracle/_ref)")
    return api.RefOracle()


@pytest.fixture(scope="session")
def djb():
    """The product: built in-tree; fails loudly if the CUDA library is missing."""
    import dj_brdf_b200
    from dj_brdf_b200 import build
    if not build.LIB.exists():
        build.build_library()
    dj_brdf_b200.load()
    return dj_brdf_b200


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        ia = a.view(np.uint32 if a.dtype == np.float32 else np.uint64)
        ib = b.view(ia.dtype)
        return (ia == ib) | (np.isnan(a) & np.isnan(b))
    return a == b


def rel_err(a, b):
    """|a-b| / max(|b|, tiny); exact zeros, NaNs and infs must match in place."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | same_inf
    denom = np.maximum(np.abs(b), 1e-30)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / denom
    e[ok] = 0.0
    e[np.isnan(e)] = np.inf
    return e
