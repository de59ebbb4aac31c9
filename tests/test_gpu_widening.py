"""GPU parity of the SURVEY.md section 8f rows: djb::sgd / djb::abc (N3) and what the Mitsuba plugins do with them
(tabular(sgd / abc, 90), mitsuba/dj_sgd.cpp:29-30, dj_abc.cpp:30-32).  CUDA through the C-ABI against the oracle port
and the golden vectors produced by the unmodified reference."""
from pathlib import Path

import numpy as np
import pytest

from oracle import api
from tests import cases
from tests.conftest import bits_equal, rel_err
from tests.test_gpu_fit import check_fit

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
REL_TOL = 1e-5


def close(got, want, what, min_bits):
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(got == 0, want == 0), f"{what}: zero pattern differs"
    e = rel_err(got, want)
    assert float(e.max()) <= REL_TOL, f"{what}: max rel err {float(e.max()):.3e}"
    rate = bits_equal(got, want).mean()
    assert rate >= min_bits, f"{what}: bit-identical rate {rate:.6f}"


def test_sgd_abc_eval_all_presets_vs_golden(djb):
    """every material of both tables, on the golden pairs (device double exp / pow / acos differ from glibc's by an ulp
    now and then: >= 99 % bit-identical, all within 1e-5)"""
    x = np.load(GOLD / "extra_golden.npz")
    wi, wo = x["analytic/wi"], x["analytic/wo"]
    for k, name in enumerate(djb.sgd.names()):
        close(djb.sgd(name).eval(wi, wo), x["sgd/eval"][k], f"sgd {name}", 0.97)
    for k, name in enumerate(djb.abc.names()):
        close(djb.abc(name).eval(wi, wo), x["abc/eval"][k], f"abc {name}", 0.97)


@pytest.mark.parametrize("kind,name", [("sgd", "alum-bronze"), ("sgd", "white-fabric"), ("abc", "aluminium"),
                                       ("abc", "beige-fabric")])
def test_sgd_abc_eval_vs_port_at_scale(djb, port, kind, name):
    import torch
    wi, wo, _ = cases.pairs(cases.N_PARITY, stream=48)
    m = getattr(djb, kind)(name)
    want = getattr(port, kind + "_eval")(m.coefficients(), wi, wo, nthreads=8)
    close(m.eval(wi, wo), want, f"{kind} {name} host", 0.995)
    got = m.eval(torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda()).cpu().numpy()
    close(got, want, f"{kind} {name} device", 0.995)
    # brdf::evalp = eval * i.z (dj_brdf.h:803-806)
    assert bits_equal(m.evalp(wi, wo), m.eval(wi, wo) * wi[:, 2:3]).all()


@pytest.mark.parametrize("kind,name", [("sgd", "gold-metallic-paint"), ("abc", "blue-metallic-paint")])
def test_fit_from_analytic_source(djb, port, kind, name):
    x = np.load(GOLD / "extra_golden.npz")
    m = getattr(djb, kind)(name)
    t = djb.tabular(m, 90)
    want = port.fit_tabular(getattr(api.Source, kind)(name, m.coefficients()), 90)
    check_fit(t, want, f"{kind}/{name} vs port")
    check_fit(t, {k: x[f"fit/{kind}/{name}/{k}"] for k in want}, f"{kind}/{name} vs golden")


# ---- LEAN-filtered shading, fused (N1: mitsuba/dj_beckmannconductor.cpp:283-319) ------------------------------------------
def per_pair(fn, ndf, P, a, b, **kw):
    return np.concatenate([fn(ndf, P[k], a[k:k + 1], b[k:k + 1], **kw) for k in range(len(P))])


@pytest.mark.parametrize("kw", [dict(), dict(lean_filtering=False), dict(dmap_scale=1.5), dict(bias=0.0)],
                         ids=["lean", "naive_mip", "dmapscale", "unbiased"])
def test_lean_shading_params(djb, port, kw):
    import torch
    E, alpha = cases.lean_texels(100_000)
    want = port.lean_shading_params(E, alpha, **kw)
    got = djb.beckmann.lean_shading_params(E, alpha, **kw)
    # device double sin / cos / sqrt / atan vs glibc: an ulp now and then in phi_a / ax; everything within 1e-6
    assert bits_equal(got, want).all(axis=1).mean() >= 0.999
    assert np.abs(got - want).max() <= 2e-6 * max(1.0, float(np.abs(want).max()))
    dev = djb.beckmann.lean_shading_params(torch.from_numpy(E).cuda(), torch.from_numpy(alpha).cuda(), **kw).cpu().numpy()
    assert bits_equal(dev, got).all()
    a0 = np.array([0.1, 0.3, 0.4], np.float32)
    got0, want0 = djb.beckmann.lean_shading_params(E, a0, **kw), port.lean_shading_params(E, a0, **kw)
    assert bits_equal(got0, want0).all(axis=1).mean() >= 0.999


@pytest.mark.parametrize("fname", ["ideal", "schlick", "unpolarized", "spline"])
def test_lean_shading_queries_match_unfused_path(djb, port, fname):
    """fused kernel == params construction + PER_PAIR query (bit for bit: same device code on both routes), and both
    == the oracle on the golden records"""
    import torch
    from tests.test_gpu_parity import mk_fresnel
    f = api.Fresnel.ideal() if fname == "ideal" else cases.fresnels()[fname]
    b = djb.beckmann(mk_fresnel(djb, f))
    n = 200_000
    E, alpha = cases.lean_texels(n, seed=11)
    wi, wo, u = cases.pairs(n, stream=710)
    tE, ta, twi, two, tu = (torch.from_numpy(v).cuda() for v in (E, alpha, wi, wo, u))
    P = djb.beckmann.lean_shading_params(tE, ta)
    assert bits_equal(b.evalp_lean(twi, two, tE, ta).cpu().numpy(), b.evalp(twi, two, P).cpu().numpy()).all()
    assert bits_equal(b.pdf_lean(twi, two, tE, ta).cpu().numpy(), b.pdf(twi, two, P).cpu().numpy()).all()
    fw, fi, fp = b.evalp_is_lean(tu, two, tE, ta)
    uw, ui, up = b.evalp_is(tu, two, P)
    assert bits_equal(fw.cpu().numpy(), uw.cpu().numpy()).all() and bits_equal(fi.cpu().numpy(), ui.cpu().numpy()).all()
    assert bits_equal(fp.cpu().numpy(), up.cpu().numpy()).all()
    # host arrays through the staging pipeline give the same numbers
    assert bits_equal(b.evalp_lean(wi, wo, E, alpha), b.evalp_lean(twi, two, tE, ta).cpu().numpy()).all()
    # against the oracle on a subset (oracle params, so a params ulp does not blur the query comparison)
    m = 3000
    Po = port.lean_shading_params(E[:m], alpha[:m])
    same = bits_equal(P.cpu().numpy()[:m], Po).all(axis=1)
    got = b.evalp_lean(wi[:m], wo[:m], E[:m], alpha[:m])
    want = per_pair(port.evalp, api.NDF_BECKMANN, Po, wi[:m], wo[:m], fresnel=f)
    close(got[same], want[same], f"evalp_lean {fname}", 0.999)
    gp = b.pdf_lean(wi[:m], wo[:m], E[:m], alpha[:m])
    wp = per_pair(port.pdf, api.NDF_BECKMANN, Po, wi[:m], wo[:m])
    close(gp[same], wp[same], "pdf_lean", 0.999)


def test_lean_shading_vs_golden(djb):
    x = np.load(GOLD / "extra_golden.npz")
    E, alpha, wi, wo = (x[f"lean_shading/{k}"] for k in ("E", "alpha", "wi", "wo"))
    b = djb.beckmann()
    for tag, kw in (("lean", dict()), ("mip", dict(lean_filtering=False)), ("scaled", dict(dmap_scale=1.5))):
        P = djb.beckmann.lean_shading_params(E, alpha, **kw)
        same = bits_equal(P, x[f"lean_shading/{tag}/params"]).all(axis=1)
        assert same.mean() >= 0.995, tag
        close(b.evalp_lean(wi, wo, E, alpha, **kw)[same], x[f"lean_shading/{tag}/evalp"][same], f"{tag} evalp", 0.999)
        close(b.pdf_lean(wi, wo, E, alpha, **kw)[same], x[f"lean_shading/{tag}/pdf"][same], f"{tag} pdf", 0.999)


# ---- dmap2nmap (N4: utils/dmap2nmap.cpp:13-44) ----------------------------------------------------------------------------
def test_dmap2nmap_bit_exact(djb, port):
    import torch
    x = np.load(GOLD / "extra_golden.npz")
    for tag in "abc":
        assert np.array_equal(djb.dmap2nmap(x[f"dmap/{tag}/dmap"], float(x[f"dmap/{tag}/scale"])), x[f"dmap/{tag}/nmap"]), tag
    rng = np.random.default_rng(8)
    for h, w, sc in ((1, 1, 0.1), (2, 1, 0.3), (7, 1023, 0.5), (1024, 2048, 0.01), (513, 511, 0.1)):
        d = rng.integers(0, 256, (h, w), dtype=np.uint8)
        want = port.dmap2nmap(d, sc)
        assert np.array_equal(djb.dmap2nmap(d, sc), want), (h, w, "host")
        assert np.array_equal(djb.dmap2nmap(torch.from_numpy(d).cuda(), sc).cpu().numpy(), want), (h, w, "device")
    # the full chain of the reference's asset tools on the device: dmap -> nmap -> LEAN maps
    d = rng.integers(96, 160, (256, 256), dtype=np.uint8)
    nm = djb.dmap2nmap(torch.from_numpy(d).cuda(), 0.01)
    l1, l2 = djb.nmap2leanmap(nm, 1e-5, 25.0)
    w1, w2 = port.nmap2leanmap(port.dmap2nmap(d, 0.01), 1e-5, 25.0)
    assert bits_equal(l1.cpu().numpy(), w1).all() and bits_equal(l2.cpu().numpy(), w2).all()


# ---- djb::radial's public scalar queries (dj_brdf.h:307-310) --------------------------------------------------------------
def test_radial_scalar_queries(djb, port):
    rng = np.random.default_rng(4)
    args = dict(p22=rng.uniform(0, 30, 50_000), sigma_std=rng.uniform(-1, 1, 50_000), cdf=rng.uniform(0, 20, 50_000),
                qf=rng.uniform(1e-4, 1 - 1e-4, 50_000))
    args["sigma_std"][:3] = [1.0, 0.0, -1.0]
    args = {k: v.astype(np.float32) for k, v in args.items()}
    for ndf, cls in ((api.NDF_GGX, djb.ggx), (api.NDF_BECKMANN, djb.beckmann)):
        for what, x in args.items():
            got, want = getattr(cls(), what + "_radial")(x), port.radial_query(what, x, ndf=ndf)
            assert rel_err(got, want).max() <= REL_TOL and bits_equal(got, want).mean() >= 0.999, (ndf, what)
    t = djb.tabular(djb.beckmann(), 90)
    fit = dict(p22=t.m_p22, sigma=t.m_sigma, qf=t.m_qf, cdf=t.m_cdf)
    for what, x in args.items():
        got, want = getattr(t, what + "_radial")(x), port.radial_query(what, x, fit=fit)
        assert rel_err(got, want).max() <= REL_TOL and bits_equal(got, want).mean() >= 0.999, ("tabular", what)


# ---- djb::microfacet's public component queries (dj_brdf.h:258-272) -------------------------------------------------------
@pytest.mark.parametrize("nname", ["ggx", "beckmann"])
def test_microfacet_components(djb, port, nname):
    import torch
    from tests.test_gpu_parity import mk_fresnel
    x = np.load(GOLD / "extra_golden.npz")
    ndf = api.NDF_GGX if nname == "ggx" else api.NDF_BECKMANN
    f = api.Fresnel.unpolarized([1.5, 1.8, 2.4])
    b = (djb.ggx if nname == "ggx" else djb.beckmann)(mk_fresnel(djb, f))
    wi, wo, h, xy, cs = (x[f"components/{k}"] for k in ("wi", "wo", "h", "xy", "cos"))
    P = x["components/params"]
    calls = dict(ndf=lambda: b.ndf(h, P), gaf=lambda: b.gaf(h, wi, wo, P), g1=lambda: b.g1(h, wo, P), sigma=lambda: b.sigma(wo, P),
                 p22=lambda: b.p22(xy, P), vp22=lambda: b.vp22(xy, wo, P), vndf=lambda: b.vndf(h, wo, P),
                 fresnel=lambda: b.fresnel_term(cs))
    for what, call in calls.items():
        close(call(), x[f"components/{nname}/{what}"], f"{nname} {what} vs golden", 0.995)
    # at scale against the port, device arrays
    wi, wo, u = cases.pairs(cases.N_PARITY, stream=34)
    h = ((wi + wo) / np.linalg.norm(wi + wo, axis=1, keepdims=True)).astype(np.float32)
    th, two, twi = (torch.from_numpy(v).cuda() for v in (h, wo, wi))
    Pa = cases.param_sets(port)["aniso"]
    close(b.vndf(th, two, Pa).cpu().numpy(), port.component("vndf", ndf, Pa, h, wo, fresnel=f), f"{nname} vndf", 0.999)
    close(b.gaf(th, twi, two, Pa).cpu().numpy(), port.component("gaf", ndf, Pa, h, wi, wo, fresnel=f), f"{nname} gaf", 0.999)
    close(b.sigma(two).cpu().numpy(), port.component("sigma", ndf, None, wo), f"{nname} sigma NULL params", 0.999)
