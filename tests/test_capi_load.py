"""CPU: the C-ABI library loads, exports every symbol include/djb200.h declares, and its host-side pieces (params
factories, argument checking, error strings) behave like the reference.  No compute calls without a GPU -- and on a
machine without one every compute entry point must fail loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from tests import cases
from tests.conftest import bits_equal

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "djb200.h").read_text()
    return sorted(set(re.findall(r"DJB200_API\s+[\w\s\*]+?\b(djb200_\w+)\s*\(", text)))


def test_header_symbols_are_exported(djb):
    from dj_brdf_b200 import capi
    lib = C.CDLL(str(capi.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 30, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/djb200.h but not exported by libdjb200.so"
    assert set(capi.EXPORTED_SYMBOLS) <= set(syms)


def test_params_factories_match_oracle(djb, port):
    for pname, P in cases.param_sets(port).items():
        if pname == "offcentre":
            got = djb.params.pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)
        elif pname == "standard":
            got = djb.params.standard()
        else:
            a1, a2, ph = P[3], P[4], P[5]
            got = djb.params.elliptic(float(a1), float(a2), float(ph))
        assert bits_equal(got, P).all(), pname
    rng = np.random.default_rng(0)
    for _ in range(200):
        a1, a2, ph = rng.uniform(0.01, 1.5), rng.uniform(0.01, 1.5), rng.uniform(-4, 4)
        assert bits_equal(djb.params.elliptic(a1, a2, ph), port.params_elliptic(a1, a2, ph)).all()
        ax, ay, rho = rng.uniform(0.01, 1.5), rng.uniform(0.01, 1.5), rng.uniform(-0.95, 0.95)
        tx, ty = rng.uniform(-1, 1), rng.uniform(-1, 1)
        assert bits_equal(djb.params.pdfparams(ax, ay, rho, tx, ty), port.params_pdfparams(ax, ay, rho, tx, ty)).all()


def test_invalid_arguments(djb):
    with pytest.raises(djb.DjbError, match="Invalid ellipse radii"):  # DJB_ASSERT, dj_brdf.h:1453
        djb.params.elliptic(0.0, 0.3)
    with pytest.raises(djb.DjbError, match="Invalid correlation"):  # dj_brdf.h:1467
        djb.params.pdfparams(0.3, 0.3, 1.0)
    with pytest.raises(djb.DjbError):
        djb.merl("/nonexistent/file.binary")  # djb::exc "Failed to open", dj_brdf.h:970
    with pytest.raises(djb.DjbError):
        djb.tabular(djb.ggx(), 2)  # DJB_ASSERT(res > 2), dj_brdf.h:2218


def test_no_cpu_fallback(djb):
    if djb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    wi, wo, _ = cases.pairs(16)
    with pytest.raises(djb.DjbError) as e:
        djb.ggx().eval(wi, wo, djb.params.isotropic(0.1))
    assert e.value.status == 2  # DJB200_ERR_NO_DEVICE
    with pytest.raises(djb.DjbError):
        djb.merl.index(wi, wo)
    with pytest.raises(djb.DjbError):
        djb.nmap2leanmap(cases.synthetic_nmap(8, 8))


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or any CPU implementation)."""
    for p in list((ROOT / "dj_brdf_b200").rglob("*.py")) + list((ROOT / "dj_brdf_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            text = p.read_text()
            assert "oracle" not in text.replace("# oracle", "").lower() or p.name == "build.py", p
