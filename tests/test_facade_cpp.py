"""The C++ host side (include/djb200_facade.hpp, include/compat/dj_brdf.h): programs written against the reference's
`djb::` interface build against the facade and -- on the GPU box -- produce the oracle's numbers through the C-ABI."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import api
from tests import cases
from tests.conftest import bits_equal, rel_err

ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "dj_brdf_b200"


def compile_cpp(src, out, incs, std="-std=c++11"):
    cmd = ["g++", "-O2", std, *[f"-I{i}" for i in incs], str(src), f"-L{LIBDIR}", "-ldjb200",
           f"-Wl,-rpath,{LIBDIR}", "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def bins(djb, tmp_path_factory):
    d = tmp_path_factory.mktemp("cpp")
    return dict(
        check=compile_cpp(ROOT / "tests/cpp/facade_check.cpp", d / "facade_check", [ROOT / "include/compat"]),
        example=compile_cpp(ROOT / "examples/merl_params_batch.cpp", d / "merl_params_batch", [ROOT / "include"]),
        dir=d)


def test_facade_programs_compile(bins):
    assert bins["check"].exists() and bins["example"].exists()


def test_reference_example_compiles_unchanged(djb, tmp_path):
    """examples/merl_params.cpp of the reference, byte for byte, against include/compat/dj_brdf.h."""
    src = api.REF_ROOT / "examples" / "merl_params.cpp"
    if not src.exists():
        pytest.skip("/root/reference not present")
    compile_cpp(src, tmp_path / "ref_merl_params", [ROOT / "include/compat"], std="-std=gnu++11")


def test_api_surface_has_the_reference_signatures():
    """tests/cpp/api_surface.cpp pins every public member the facade provides to the reference's exact signature; it must
    compile against both headers"""
    src = ROOT / "tests/cpp/api_surface.cpp"
    r = subprocess.run(["g++", "-std=gnu++11", "-fsyntax-only", f"-I{ROOT / 'include/compat'}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if (api.REF_ROOT / "dj_brdf.h").exists():
        r = subprocess.run(["g++", "-std=gnu++11", "-fsyntax-only", "-DUSE_REFERENCE", f"-I{api.REF_ROOT}", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_params_host_api_matches_reference(djb, tmp_path):
    """params factories / setters / getters are host arithmetic on both sides: identical hex-float text, no GPU needed"""
    if not (api.REF_ROOT / "dj_brdf.h").exists():
        pytest.skip("/root/reference not present")
    src = ROOT / "tests/cpp/params_host_check.cpp"
    ours = compile_cpp(src, tmp_path / "params_b200", [ROOT / "include/compat"], std="-std=gnu++11")
    r = subprocess.run(["g++", "-O3", "-ffp-contract=off", "-DNVERBOSE", "-DUSE_REFERENCE", f"-I{api.REF_ROOT}", str(src), "-o",
                        str(tmp_path / "params_ref")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = subprocess.run([str(ours)], capture_output=True, text=True)
    b = subprocess.run([str(tmp_path / "params_ref")], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    assert a.stdout == b.stdout and a.stdout.count("\n") == 242 + 61 + 29 + 17 * 17


def test_facade_fails_loudly_without_gpu(djb, bins):
    if djb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    wi, wo, u = cases.pairs(32)
    inp = bins["dir"] / "in.bin"
    with open(inp, "wb") as f:
        f.write(np.int32(32).tobytes() + wi.tobytes() + wo.tobytes() + u.tobytes())
    r = subprocess.run([str(bins["check"]), str(inp), str(bins["dir"] / "out.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bits", "1e-5"])
def test_facade_matches_oracle(bins, port, precision):
    """The reference's C++ interface on the GPU, in both precision modes of eval / pdf (DJB200_PRECISION; "1e-5" is the default)."""
    n = 20_000
    wi, wo, u = cases.pairs(n, stream=300)
    inp, outp = bins["dir"] / "in.bin", bins["dir"] / "out.bin"
    with open(inp, "wb") as f:
        f.write(np.int32(n).tobytes() + wi.tobytes() + wo.tobytes() + u.tobytes())
    r = subprocess.run([str(bins["check"]), str(inp), str(outp)], capture_output=True, text=True,
                       env=dict(os.environ, DJB200_PRECISION=precision))
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(outp, dtype=np.float32)
    pos = 0

    def take(*shape):
        nonlocal pos
        k = int(np.prod(shape))
        a = raw[pos:pos + k].reshape(shape)
        pos += k
        return a

    fr = api.Fresnel.schlick([0.9, 0.5, 0.2])
    P = port.params_elliptic(0.1, 0.4, 0.7)
    two = [port.params_elliptic(0.3, 0.3, 0.0), port.params_pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)]
    for ndf in (api.NDF_GGX, api.NDF_BECKMANN):
        ev, pdf, smp, sc, ev2 = take(n, 3), take(n), take(n, 3), take(16, 3), take(2, n, 3)
        want = port.eval(ndf, P, wi, wo, fr)
        assert rel_err(ev, want).max() <= 1e-5 and np.array_equal(ev == 0, want == 0)
        assert rel_err(pdf, port.pdf(ndf, P, wi, wo, fr)).max() <= 1e-5
        wsm = port.sample(ndf, P, u, wo)
        if precision == "bits":
            same = bits_equal(smp, wsm).all(axis=1).mean()
            assert same >= (0.9999 if ndf == api.NDF_GGX else 0.999)
        else:  # the 1e-5 tier of sample: tests/test_gpu_parity.py::test_fast_tier_sample_*
            err = np.abs(smp.astype(np.float64) - wsm).max(axis=1)
            assert (err <= 1e-5).mean() >= (1.0 if ndf == api.NDF_GGX else 0.999) and err.max() <= 2e-2
        assert bits_equal(sc, ev[:16]).all(), "scalar virtual calls must equal the batch"
        for m in range(2):
            assert rel_err(ev2[m], port.eval(ndf, two[m], wi, wo, fr)).max() <= 1e-5
    back = take(5)
    E = port.params_to_lrep(two[1])[0]
    d = np.array([0.01, 0.02, 0.03, 0.04, 0.001], np.float32)
    e1, e2 = np.float32(E[0] + d[0]), np.float32(E[1] + d[1])  # operator+= advances E1, E2 first (dj_brdf.h:2011-2020)
    E2 = np.array([e1, e2, E[2] + (d[2] + np.float32(2.0) * e1 * d[0]), E[3] + (d[3] + np.float32(2.0) * e2 * d[1]),
                   E[4] + ((d[4] + e1 * d[1]) + e2 * d[0])], np.float32)
    wantp = port.lrep_to_params(E2[None])[0]
    assert np.allclose(back, [wantp[6], wantp[7], wantp[8], wantp[10], wantp[11]], rtol=1e-5, atol=1e-7)
    # djb::tabular fitted from an analytic GGX and evaluated / sampled through the brdf base class
    tev, tsm, te0, tal = take(n, 3), take(n, 3), take(3), take(2)
    fit = port.fit_tabular(api.Source.microfacet(api.NDF_GGX), 90)
    want = port.tabular_query("eval", fit, wi, wo, None, nthreads=8)
    assert rel_err(tev, want).max() <= 1e-5 and np.array_equal(tev == 0, want == 0)
    assert bits_equal(tsm, port.tabular_query("sample", fit, u, wo, None, nthreads=8)).all(axis=1).mean() >= 0.9995
    assert bits_equal(te0, tev[0]).all()
    assert np.allclose(tal, fit["alpha"], rtol=1e-6)
    # the scalar members, one facade call per item, against the batched Python mirror of the same entry points (which
    # tests/test_gpu_round2.py pins to the reference): same kernels, so the same bits
    import dj_brdf_b200 as djb
    nm = 64
    rows = take(nm, 8 + 21 + 6)
    q, v, t = rows[:, :8], rows[:, 8:29].reshape(nm, 7, 3), rows[:, 29:]
    c = wo[:nm, 2].copy()
    sn = np.sqrt(np.float32(1.0) - c * c).astype(np.float32)
    u1, u2 = u[:nm, 0].copy(), u[:nm, 1].copy()
    g, b = djb.ggx(), djb.beckmann()
    sg, ab = djb.sgd("gold-metallic-paint"), djb.abc("gold-metallic-paint")
    a, o = np.ascontiguousarray(wi[:nm]), np.ascontiguousarray(wo[:nm])
    want_q = [g.qf1(u1), g.qf2_radial(u1, c, sn), g.qf3_radial(u2, q[:, 1].copy()), b.qf1(u1), b.qf2_radial(u1, c, sn),
              b.qf3_radial(u2, q[:, 4].copy()), ab.gaf(a, a, o)]
    for k, w in enumerate(want_q):
        assert bits_equal(q[:, k], w).all(), ("scalar member", k)
    want_v = [sg.ndf(a), sg.gaf(a, a, o), sg.g1(o), sg.fresnel_term(c), ab.ndf(a), ab.fresnel_term(c)]
    for k, w in enumerate(want_v):
        assert bits_equal(v[:, k], w).all(), ("vec3 member", k)
    ta = djb.tabular_anisotropic(djb.beckmann(), 8, 10)
    phi, theta = (np.float32(6.2) * u2).astype(np.float32), (np.float32(1.5) * u1).astype(np.float32)
    want_t = [ta.pdf1(phi), ta.cdf1(phi), ta.qf1(u1), ta.pdf2(theta, phi), ta.cdf2(theta, phi), ta.qf2(u1, phi)]
    for k, w in enumerate(want_t):
        assert bits_equal(t[:, k], w).all(), ("table member", k)
    assert take(1)[0] == 2.0, "microfacet::qf2 and radial::qf2_radial must throw djb::exc as in the reference"
    assert pos == raw.size


@pytest.mark.gpu
def test_merl_params_example(bins, port):
    d = bins["dir"]
    tabs = {f"mat{s}": cases.smooth_merl_table(s) for s in (41, 42, 43)}
    files = []
    for name, t in tabs.items():
        p = d / f"{name}.binary"
        api.write_merl_file(p, t)
        files.append(str(p))
    out = d / "params.txt"
    r = subprocess.run([str(bins["example"]), "-o", str(out), *files], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = out.read_text().strip().splitlines()
    assert lines[0] == "# MERL Beckmann GGX" and len(lines) == 4
    for line, (name, t) in zip(lines[1:], tabs.items()):
        want = port.fit_tabular(api.Source.merl(t), 90)["alpha"]
        nm, b, g = line.split()
        assert nm == name and b == f"{want[0]:.3f}" and g == f"{want[1]:.3f}", (line, want)
