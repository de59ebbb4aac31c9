"""CPU: the exactness claims behind shortcuts that only the device code takes (csrc/djb_lean.cuh), checked EXHAUSTIVELY over
the float ranges they cover with the oracle port (which is bit-identical to the reference), plus the sync of the generated
preset table with the reference header."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import api

ROOT = Path(__file__).resolve().parents[1]


def floats_between(lo, hi):
    a, b = np.array([lo, hi], np.float32).view(np.uint32)
    return np.arange(a, b + 1, dtype=np.uint32).view(np.float32)


def test_erf_saturates_exactly_from_4(port):
    """erf_lean returns +-1 for x^2 >= 16 without evaluating the exponential: djb::erf must be exactly 1.0f there (and for
    the first time at x = 3.9195216, so the threshold has margin)."""
    x = floats_between(3.9, 10.5)
    y = port.erf(x)
    assert (y[x >= np.float32(4.0)] == np.float32(1.0)).all()
    assert (port.erf(-x[x >= np.float32(4.0)]) == np.float32(-1.0)).all()
    first = x[np.argmax(y == np.float32(1.0))]
    assert abs(float(first) - 3.9195216) < 1e-6 and (y[x >= first] == np.float32(1.0)).all()
    big = np.array([10.5, 50.0, 1e4, 1e19, 3e38, np.inf], np.float32)
    assert (port.erf(big) == np.float32(1.0)).all()


def test_beckmann_sigma_std_equals_cosine_past_the_shortcut(port):
    """beck_sigma_std_lean returns c itself when -(c / s)^2 < -16.5 and c > 0 (cot > 4.06): every float c in that range must
    give sigma_std_radial(c) == c in the reference's arithmetic."""
    c = floats_between(0.96, 0.99999994)
    s = np.sqrt(1.0 - c.astype(np.float64) ** 2)
    nu = (c / s.astype(np.float32)).astype(np.float32)
    sel = -(nu * nu) < np.float32(-16.5)
    assert sel.sum() > 400_000
    got = port.radial_query("sigma_std", c[sel], ndf=api.NDF_BECKMANN)
    assert np.array_equal(got.view(np.uint32), c[sel].view(np.uint32))


def test_beckmann_p22_is_zero_past_103_5(port):
    """lean_ndf / the compacting kernel treat r^2 > 103.5 as D == 0: exp(-r^2) / pi must round to +0 there."""
    r2 = np.concatenate([floats_between(103.5, 104.5)[1:], np.array([110, 200, 1e4, 3e38, np.inf], np.float32)])
    got = port.radial_query("p22", r2, ndf=api.NDF_BECKMANN)
    assert (got == 0).all() and not np.signbit(got).any()
    # the threshold is conservative: the smallest subnormal results end near r^2 = 102.83
    below = port.radial_query("p22", floats_between(102.0, 102.8), ndf=api.NDF_BECKMANN)
    assert (below > 0).all()


def test_generated_preset_table_is_in_sync_with_the_reference(tmp_path):
    ref = api.REF_ROOT / "dj_brdf.h"
    if not ref.exists():
        pytest.skip("/root/reference not present")
    committed = (ROOT / "dj_brdf_b200" / "csrc" / "djb_presets.inc").read_text()
    script = (ROOT / "tools" / "extract_presets.py").read_text().replace(
        'OUT = Path(__file__).resolve().parent.parent / "dj_brdf_b200" / "csrc" / "djb_presets.inc"',
        f'OUT = Path(r"{tmp_path / "presets.inc"}")')
    (tmp_path / "extract.py").write_text(script)
    subprocess.run([sys.executable, str(tmp_path / "extract.py"), str(ref)], check=True, capture_output=True)
    assert (tmp_path / "presets.inc").read_text() == committed


def test_restated_glibc_float_functions_are_bit_identical_to_libm(tmp_path):
    """csrc/djb_glibcf.h restates glibc's logf / expf / powf (what the reference's erfinv / qf2_radial call) operation for
    operation; the same header compiled for the host must reproduce the platform's libm bit for bit.  Strided here (about 6e6
    arguments per function); stride 1 -- every float of the sampling path's domains, 2e9 per function -- was run when the
    tables were written: 0 mismatches (profiles/r02_glibcf_exhaustive.txt)."""
    exe = tmp_path / "glibcf_check"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-pthread", str(ROOT / "tests" / "cpp" / "glibcf_check.cpp"), "-o", str(exe)],
                   check=True, capture_output=True)
    r = subprocess.run([str(exe), "331"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count(" 0 mismatches") == 3, r.stdout


def test_lean_double_exp_log_match_libm(tmp_path):
    """dj_brdf_b200/csrc/djb_dmath.cuh (the double exp / log / pow / sqrt / acos / atan / atan2 / sincos of the analytic and table BRDF
    kernels) compiled for the host against libm: ulp bounds over 2e7 arguments per function, identical special cases, and the nine
    one-float coordinate maps against the reference's expressions over every 16th float of their domains (`dmath_check full` checks
    every float: profiles/r02_n_dmath_exhaustive.txt)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "dmath_check"
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-pthread", f"-I{root / 'dj_brdf_b200/csrc'}", str(root / "tests/cpp/dmath_check.cpp"),
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_dmath_tables_match_their_generator():
    """dj_brdf_b200/csrc/djb_dmath_tables.inc is exactly what tools/gen_dmath_tables.py prints (mpmath at 60 digits, each value rounded
    once to double and written as a hex float): the tables of the double exp / log / atan and the asin coefficients are reproducible."""
    import subprocess
    import sys
    from pathlib import Path
    pytest = __import__("pytest")
    pytest.importorskip("mpmath")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "tools/gen_dmath_tables.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == (root / "dj_brdf_b200/csrc/djb_dmath_tables.inc").read_text()


# ---- float forms of "double operation rounded to float" in the table / analytic kernels (csrc/djb_device.cuh, round 2) ------------
def test_utia_cell_index_as_float_product_plus_remainder():
    """floor_div (djb_device.cuh): (int)floor((double)x / d) for the floats x of utia::eval's ranges -- theta in [0, 90) over
    d = 15, phi in [0, 360) over d = 7.5 (dj_brdf.h:1090-1100) -- equals the float product with 1 / d, floored, corrected by the
    float remainder x - q d (one fused operation)."""
    for d, hi in ((15.0, 90.0), (7.5, 360.0)):
        # every 37th float of the range, and every float within 4096 ulps of a multiple of d (where a quotient can land on the
        # wrong side); the full range was run once with stride 1: no difference
        x = floats_between(0.0, hi)[::37]
        near = []
        for m in np.arange(d, hi + d / 2, d, dtype=np.float64):
            c = np.array([m], np.float32).view(np.uint32)[0]
            near.append(np.arange(c - 4096, c + 4097, dtype=np.uint32).view(np.float32))
        x = np.concatenate([x] + near)
        x = x[x < np.float32(hi)]
        want = np.floor(x.astype(np.float64) / d).astype(np.int32)
        inv = np.float32(1.0) / np.float32(d)
        q = np.floor(x * inv).astype(np.int32)                       # floorf(x * inv_d), float product
        r = (x.astype(np.float64) - d * q.astype(np.float64)).astype(np.float32)  # fmaf(-d, q, x): exact product, one rounding
        q = np.where(r < 0, q - 1, np.where(r >= np.float32(d), q + 1, q))
        assert np.array_equal(q, want), (d, int((q != want).sum()))
        # the correction is needed: the plain float product alone is wrong somewhere in the range
        assert (np.floor(x * inv).astype(np.int32) != want).any()


def test_utia_float_forms_of_double_expressions():
    """utia_eval1: (float)(15.0 * k) == 15.0f * (float)k for the cell indices; (double)acc > 0.0375 <=> acc >= 0.0375f for every
    float; (float)((double)x +- 360.0) == x +- 360.0f."""
    k = np.arange(0, 50)
    for d in (15.0, 7.5):
        assert np.array_equal((d * k).astype(np.float32), np.float32(d) * k.astype(np.float32))
    a = floats_between(0.03, 0.045)
    assert np.array_equal(a.astype(np.float64) > 0.0375, a >= np.float32(0.0375))
    x = np.concatenate([floats_between(1e-3, 400.0)[::7], -floats_between(1e-3, 400.0)[::7]])
    for s in (360.0, -360.0):
        assert np.array_equal((x.astype(np.float64) + s).astype(np.float32), x + np.float32(s))


def test_sqrt_half_and_float_square_roots():
    """sqrt_half (params construction): (float)sqrt(0.5 * (double)x) == sqrtf(0.5f * x) for floats of magnitude >= 1e-30;
    (float)sqrt((double)x) == sqrtf(x) (the double rounding of a square root is innocuous) -- strided over all positive floats."""
    x = floats_between(1e-30, 3e38)[::257]
    assert np.array_equal(np.sqrt(0.5 * x.astype(np.float64)).astype(np.float32), np.sqrt(np.float32(0.5) * x))
    assert np.array_equal(np.sqrt(x.astype(np.float64)).astype(np.float32), np.sqrt(x))


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_shadowing_gate_of_centred_lobes_is_open_whenever_both_directions_are_up(port, ndf):
    """fast_pdf_try (djb_lean.cuh): for centred lobes (params::elliptic, alpha <= 1e3) and pairs with i.z > 0, o.z > 0,
    i.z o.z > 1e-30, |i|, |o| <= 2 the pdf kernels take `gaf(h, i, o) > 0` for granted instead of evaluating sigma(i).  Here the
    reference's own gaf (the port, bit-identical to it) over directions from normal to extreme grazing incidence (z down to 1e-14),
    unit and non-unit lengths, roughness from 1e-4 to 1e3, isotropic to 1 : 1e4 anisotropy: positive, finite, every time."""
    rng = np.random.default_rng(2024)
    n = 200_000

    def dirs():
        z = np.where(rng.random(n) < 0.5, rng.random(n), 10.0 ** rng.uniform(-14.0, 0.0, n))
        phi = rng.uniform(0, 2 * np.pi, n)
        r = np.sqrt(np.maximum(1.0 - z * z, 0.0))
        d = np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)
        return (d * rng.uniform(0.5, 1.99, (n, 1))).astype(np.float32)  # |d| <= 2, not necessarily 1

    def params_centred(P):  # djb_lean.cuh; P = n.xyz a1 a2 phi_a ax ay rho srho tx ty
        return (P[10] == 0 and P[11] == 0 and P[0] == 0 and P[1] == 0 and P[2] == 1 and 0 < P[6] <= 1e3 and 0 < P[7] <= 1e3
                and abs(P[8]) < 1)

    worst, excluded = np.inf, 0
    for a1, a2, ph in [(0.1, 0.1, 0.0), (1e-4, 1e-4, 0.0), (1e-4, 1.0, 0.7), (1e3, 1e3, 0.0), (1e3, 0.1, 2.0), (0.02, 0.8, 1.1),
                       (5.0, 5e-4, 3.0), (900.0, 0.3, 0.4), (0.5, 0.05, 2.5)]:
        P = port.params_elliptic(a1, a2, ph)
        wi, wo = dirs(), dirs()
        up = (wi[:, 2] > 0) & (wo[:, 2] > 0) & (wi[:, 2] * wo[:, 2] > np.float32(1e-30))
        assert up.mean() > 0.99
        h = wi + wo
        h /= np.linalg.norm(h, axis=1, keepdims=True)
        G = port.component("gaf", ndf, P, h.astype(np.float32), wi, wo, shadow=True)
        if not params_centred(P):
            # e.g. (1e3, 0.1, 2.0): params::elliptic rounds rho to 1.0000001, sqrt(1 - rho^2) is NaN and the reference's G is 0 for every
            # pair -- such a material switches the shortcut off for its launch, and the kernels evaluate G as the reference does
            excluded += 1
            continue
        assert np.isfinite(G[up]).all() and (G[up] > 0).all(), (a1, a2, ph, int((~(G[up] > 0)).sum()))
        worst = min(worst, float(G[up].min()))
    assert excluded == 3  # rho rounds to 1 for (1e-4, 1, 0.7) and (5, 5e-4, 3), above 1 for (1e3, 0.1, 2): six materials were checked
    assert worst > 1e-36  # far from the underflow threshold
