"""CPU: the oracle port (oracle/djb_oracle*.c) against the committed golden vectors, which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  Everything must match to the bit -- this is what pins the
oracle on machines where /root/reference is absent (the GPU box)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import api
from tests import cases
from tests.conftest import bits_equal

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD / "eval_golden.npz")


@pytest.fixture(scope="module")
def f():
    return np.load(GOLD / "fit_golden.npz")


def ndf_id(name):
    return api.NDF_GGX if name == "ggx" else api.NDF_BECKMANN


def test_params_factories(port, g):
    for pname, P in cases.param_sets(port).items():
        assert bits_equal(P, g[f"params/{pname}"]).all(), pname


@pytest.mark.parametrize("nname", ["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "iso0.5", "aniso", "aniso2", "offcentre", "standard"])
def test_eval_pdf_sample(port, g, nname, pname):
    ndf, P = ndf_id(nname), g[f"params/{pname}"]
    wi, wo, u = g["wi"], g["wo"], g["u"]
    assert bits_equal(port.eval(ndf, P, wi, wo), g[f"{nname}/{pname}/eval"]).all()
    assert bits_equal(port.evalp(ndf, P, wi, wo), g[f"{nname}/{pname}/evalp"]).all()
    assert bits_equal(port.pdf(ndf, P, wi, wo), g[f"{nname}/{pname}/pdf"]).all()
    assert bits_equal(port.sample(ndf, P, u, wo), g[f"{nname}/{pname}/sample"]).all()


@pytest.mark.parametrize("nname", ["ggx", "beckmann"])
def test_fresnel_shadow_is_null(port, g, nname):
    ndf, P = ndf_id(nname), g["params/aniso"]
    wi, wo, u = g["wi"], g["wo"], g["u"]
    for fname, fr in cases.fresnels().items():
        for shadow in (True, False):
            tag = f"{nname}/aniso/{fname}/shadow{int(shadow)}"
            assert bits_equal(port.eval(ndf, P, wi, wo, fr, shadow), g[f"{tag}/eval"]).all(), tag
            assert bits_equal(port.pdf(ndf, P, wi, wo, fr, shadow), g[f"{tag}/pdf"]).all(), tag
    w, i, p = port.evalp_is(ndf, P, u, wo, cases.fresnels()["schlick"])
    assert bits_equal(w, g[f"{nname}/aniso/schlick/evalp_is_w"]).all()
    assert bits_equal(i, g[f"{nname}/aniso/schlick/evalp_is_i"]).all()
    assert bits_equal(p, g[f"{nname}/aniso/schlick/evalp_is_pdf"]).all()
    assert bits_equal(port.eval(ndf, None, wi, wo), g[f"{nname}/null_params/eval"]).all()


def test_frames_and_tables(port, g):
    wi, wo = g["wi"], g["wo"]
    n = len(g["io_to_hd/h"])
    h, d = port.io_to_hd(wi[:n], wo[:n])
    assert bits_equal(h, g["io_to_hd/h"]).all() and bits_equal(d, g["io_to_hd/d"]).all()
    i2, o2 = port.hd_to_io(h, d)
    assert bits_equal(i2, g["hd_to_io/i"]).all() and bits_equal(o2, g["hd_to_io/o"]).all()
    assert (port.merl_index(wi, wo) == g["merl/index"]).all()
    t = cases.random_merl_table(int(g["merl/table_seed"][0]))
    if cases.sha(t) != str(g["merl/table_sha256"][0]):
        pytest.skip("numpy PCG64 stream differs on this machine; golden table cannot be regenerated")
    assert bits_equal(port.merl_eval(t, wi, wo), g["merl/eval"]).all()
    ut = cases.random_utia_table(int(g["utia/table_seed"][0]))
    assert cases.sha(ut) == str(g["utia/table_sha256"][0])
    assert bits_equal(port.utia_eval(ut, wi, wo), g["utia/eval"]).all()


def test_lean(port, g):
    for bias in (0.0, 25.0):
        nm = g[f"lean/bias{int(bias)}/nmap"]
        l1, l2 = port.nmap2leanmap(nm, 1e-5, bias)
        assert bits_equal(l1, g[f"lean/bias{int(bias)}/l1"]).all() and bits_equal(l2, g[f"lean/bias{int(bias)}/l2"]).all()
    P = port.lrep_to_params(g["lrep/E"])
    assert bits_equal(P, g["lrep/params"]).all()
    assert bits_equal(port.params_to_lrep(P), g["lrep/E_back"]).all()


def _cmp_fit(got, f, prefix):
    for k, v in got.items():
        assert bits_equal(v, f[f"{prefix}/{k}"]).all(), f"{prefix}/{k}"


@pytest.mark.parametrize("nname", ["ggx", "beckmann"])
def test_fit_analytic(port, f, nname):
    """tests/plot_qf.cpp and plot_cdf.cpp of the reference build tabular(beckmann|ggx, 180)."""
    ndf = ndf_id(nname)
    _cmp_fit(port.fit_tabular(api.Source.microfacet(ndf), 90), f, f"iso/{nname}/res90")
    _cmp_fit(port.fit_tabular(api.Source.microfacet(ndf), 180), f, f"iso/{nname}/res180")
    _cmp_fit(port.fit_tabular(api.Source.microfacet(ndf), 64, shadow=False), f, f"iso/{nname}/res64_noshadow")
    _cmp_fit(port.fit_tabular_anisotropic(api.Source.microfacet(ndf), 16, 20, nthreads=4), f, f"aniso/{nname}/16x20")


def test_fit_tables(port, f):
    for seed in (21, 22, 23):
        t = cases.smooth_merl_table(seed)
        assert cases.sha(t) == str(f[f"iso/merl{seed}/table_sha256"][0])
        _cmp_fit(port.fit_tabular(api.Source.merl(t), 90), f, f"iso/merl{seed}/res90")
    _cmp_fit(port.fit_tabular_anisotropic(api.Source.merl(cases.smooth_merl_table(21)), 12, 16, nthreads=4), f,
             "aniso/merl21/12x16")
    ut = cases.random_utia_table(12)
    _cmp_fit(port.fit_tabular_anisotropic(api.Source.utia(ut), 14, 18, nthreads=4), f, "aniso/utia12/14x18")
    _cmp_fit(port.fit_tabular(api.Source.utia(ut), 48), f, "iso/utia12/res48")


def test_survey_spot_values(port):
    """Known answers printed by the reference during the survey."""
    def d(theta, phi):
        s = np.float32(np.sin(np.float64(np.float32(theta))))
        return np.array([[np.float32(np.float64(s) * np.cos(np.float64(np.float32(phi)))),
                          np.float32(np.float64(s) * np.sin(np.float64(np.float32(phi)))),
                          np.float32(np.cos(np.float64(np.float32(theta))))]], np.float32)
    i, o = d(0.3, 0.1), d(0.5, 2.0)
    P = port.params_elliptic(0.1, 0.1, 0.0)
    assert abs(port.eval(api.NDF_GGX, P, i, o)[0, 0] - 0.181454852) < 2e-7
    assert abs(port.pdf(api.NDF_GGX, P, i, o)[0] - 0.173391879) < 2e-7
    assert abs(port.eval(api.NDF_BECKMANN, P, i, o)[0, 0] - 0.0131159481) < 2e-8
    assert abs(port.pdf(api.NDF_BECKMANN, P, i, o)[0] - 0.0125301415) < 2e-8
    u = np.array([[0.3, 0.7]], np.float32)
    assert np.abs(port.sample(api.NDF_GGX, P, u, o)[0] - [0.282283455, -0.468485653, 0.837159991]).max() < 1e-6
    assert np.abs(port.sample(api.NDF_BECKMANN, P, u, o)[0] - [0.230195954, -0.343669266, 0.910440087]).max() < 1e-6


@pytest.fixture(scope="module")
def x():
    return np.load(GOLD / "extra_golden.npz")


def test_sgd_abc_presets(port, x):
    """djb::sgd / djb::abc for all 100 materials: the product's coefficient tables (host-only preset lookup) through the
    port's eval must reproduce the reference's output bit for bit."""
    import dj_brdf_b200 as djb
    wi, wo = x["analytic/wi"], x["analytic/wo"]
    assert list(x["sgd/names"]) == djb.sgd.names() and list(x["abc/names"]) == djb.abc.names()
    for k, name in enumerate(djb.sgd.names()):
        assert bits_equal(port.sgd_eval(djb.sgd(name).coefficients(), wi, wo), x["sgd/eval"][k]).all(), name
    for k, name in enumerate(djb.abc.names()):
        assert bits_equal(port.abc_eval(djb.abc(name).coefficients(), wi, wo), x["abc/eval"][k]).all(), name


@pytest.mark.parametrize("kind,name", [("sgd", "gold-metallic-paint"), ("abc", "blue-metallic-paint")])
def test_fit_from_analytic_source(port, x, kind, name):
    import dj_brdf_b200 as djb
    m = getattr(djb, kind)(name)
    p = port.fit_tabular(getattr(api.Source, kind)(name, m.coefficients()), 90)
    for k, v in p.items():
        assert bits_equal(v, x[f"fit/{kind}/{name}/{k}"]).all(), k


def test_tabular_anisotropic_sampling(port, x):
    """the port's sampling tables and sample / evalp_is against the reference's (golden) ones"""
    er, ar, tag = 14, 18, "aniso_sampling/utia12/14x18"
    pt = port.aniso_sampling_tables(x[f"{tag}/p22"], er, ar)
    for k in ("pdf1", "cdf1", "qf1", "pdf2", "cdf2", "qf2"):
        assert bits_equal(pt[k], x[f"{tag}/{k}"]).all(), k
    assert list(x[f"{tag}/sizes"]) == [ar, ar, pt["n_qf1"], er * ar, er * ar, pt["n_qf2"]]
    fit = dict(p22=x[f"{tag}/p22"], sigma=x[f"{tag}/sigma"], fresnel=x[f"{tag}/fresnel"])
    u, wo, P = x["aniso_sampling/u"], x["aniso_sampling/wo"], x["aniso_sampling/params"]
    assert bits_equal(port.tabular_aniso_sample_query("sample", fit, pt, er, ar, u, wo, P), x[f"{tag}/sample"]).all()
    w, i, pdf = port.tabular_aniso_sample_query("evalp_is", fit, pt, er, ar, u, wo, P)
    assert bits_equal(w, x[f"{tag}/evalp_is_w"]).all() and bits_equal(i, x[f"{tag}/evalp_is_i"]).all()
    assert bits_equal(pdf, x[f"{tag}/evalp_is_pdf"]).all()


def test_lean_shading_params(port, x):
    E, alpha = x["lean_shading/E"], x["lean_shading/alpha"]
    for tag, kw in (("lean", dict()), ("mip", dict(lean_filtering=False)), ("scaled", dict(dmap_scale=1.5))):
        P = port.lean_shading_params(E, alpha, **kw)
        assert bits_equal(P, x[f"lean_shading/{tag}/params"]).all(), tag
        wi, wo = x["lean_shading/wi"], x["lean_shading/wo"]
        got = np.concatenate([port.evalp(api.NDF_BECKMANN, P[k], wi[k:k + 1], wo[k:k + 1]) for k in range(len(P))])
        assert bits_equal(got, x[f"lean_shading/{tag}/evalp"]).all(), tag


def test_dmap2nmap(port, x):
    for tag in "abc":
        assert np.array_equal(port.dmap2nmap(x[f"dmap/{tag}/dmap"], float(x[f"dmap/{tag}/scale"])), x[f"dmap/{tag}/nmap"]), tag


@pytest.mark.parametrize("nname", ["ggx", "beckmann"])
def test_microfacet_components(port, x, nname):
    wi, wo, h, xy, cs = (x[f"components/{k}"] for k in ("wi", "wo", "h", "xy", "cos"))
    P, f = x["components/params"], api.Fresnel.unpolarized([1.5, 1.8, 2.4])
    args = dict(ndf=(h,), gaf=(h, wi, wo), g1=(h, wo), sigma=(wo,), p22=(xy,), vp22=(xy, wo), vndf=(h, wo), fresnel=(cs,))
    for what, a in args.items():
        assert bits_equal(port.component(what, ndf_id(nname), P, *a, fresnel=f), x[f"components/{nname}/{what}"]).all(), what


def test_port_anisotropic_fit_at_90x90_matches_the_reference(port):
    """The port at the size the reference's plugin uses (90 x 90, 8010 unknowns) against the reference's output
    (tests/golden/fixture_golden.npz, generated by tests/golden/make_fixture_golden.py): bit for bit."""
    g = np.load(GOLD / "fixture_golden.npz")
    p = port.fit_tabular_anisotropic(api.Source.utia(cases.random_utia_table(12)), 90, 90, nthreads=8)
    for k in ("p22", "sigma", "fresnel", "beckmann", "ggx"):
        assert bits_equal(p[k], g[f"utia12_90x90/{k}"]).all(), k
