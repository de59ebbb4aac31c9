"""GPU parity, round 2: the cases VERDICT r01 found unpinned.

* the CUDA path against the UNMODIFIED reference itself (oracle/_ref/libdjbref.so travels to the GPU box), not only
  against the port -- eval / pdf / sample / evalp_is / MERL lookups / the isotropic fit;
* chi-square test of `sample` against `pdf` (GGX, Beckmann, tabular);
* fresnel::sgd as a microfacet Fresnel term;
* one MERL handle + one microfacet descriptor driven from four host threads at once (SURVEY 8b threading).
"""
import threading
from pathlib import Path

import numpy as np
import pytest

from oracle import api
from tests import cases
from tests.conftest import bits_equal, rel_err
from tests.test_gpu_parity import check_close, mk_brdf

pytestmark = pytest.mark.gpu


# ---- directly against the reference --------------------------------------------------------------------------------
@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "aniso", "offcentre"])
def test_queries_against_the_unmodified_reference(djb, ref, ndf, pname):
    wi, wo, u = cases.pairs(100_000, stream=300)
    P = cases.param_sets(ref)[pname]
    f = api.Fresnel.schlick([0.95, 0.64, 0.54])
    b = mk_brdf(djb, ndf, f)
    check_close(b.eval(wi, wo, P), ref.eval(ndf, P, wi, wo, f), f"eval {pname}", min_bit_rate=0.9999)
    check_close(b.pdf(wi, wo, P), ref.pdf(ndf, P, wi, wo, f), f"pdf {pname}", min_bit_rate=0.9999)
    same = bits_equal(b.sample(u, wo, P), ref.sample(ndf, P, u, wo, f)).all(axis=1)
    assert same.mean() >= (0.99999 if ndf == api.NDF_GGX else 0.9999), same.mean()
    gw, gi, gp = b.evalp_is(u, wo, P)
    ww, wi_, wp = ref.evalp_is(ndf, P, u, wo, f)
    ok = bits_equal(gi, wi_).all(axis=1)
    assert ok.mean() >= 0.9995, ok.mean()
    check_close(gw[ok], ww[ok], "evalp_is weight")
    check_close(gp[ok], wp[ok], "evalp_is pdf")


def test_merl_lookup_against_the_unmodified_reference(djb, ref):
    wi, wo, _ = cases.pairs(200_000, stream=310)
    table = cases.synthetic_merl_table(0.15)
    assert np.array_equal(djb.merl.index(wi, wo), ref.merl_index(wi, wo)), "MERL cell index differs from the reference"
    got, want = djb.merl(table).eval(wi, wo), ref.merl_eval(table, wi, wo)
    assert bits_equal(got, want).all(), "MERL eval differs from the reference"
    assert (want == 0).all(axis=1).any() and (want > 0).all(axis=1).any()  # both the 'negative -> 0' rule and real cells


def test_isotropic_fit_against_the_unmodified_reference(djb, ref):
    from tests.test_gpu_fit import check_fit
    tab = cases.synthetic_merl_table(0.2, kind="ggx")
    want = ref.fit_tabular(api.Source.merl(tab), 90)
    check_fit(djb.tabular(djb.merl(tab), 90), want, "synthetic GGX table vs reference")


# ---- fresnel::sgd inside a microfacet BRDF (dj_brdf.h:1316-1328, djb_device.cuh FK_SGD) -----------------------------
@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_sgd_fresnel_term_in_a_microfacet_brdf(djb, port, ndf):
    wi, wo, u = cases.pairs(100_000, stream=320)
    f0, f1 = [0.12, 0.31, 0.66], [0.02, 0.2, 0.9]
    f = api.Fresnel(api.F_SGD, f0 + f1)
    cls = djb.ggx if ndf == api.NDF_GGX else djb.beckmann
    b = cls(djb.fresnel.sgd(f0, f1), True)
    for pname in ("iso0.5", "aniso"):
        P = cases.param_sets(port)[pname]
        check_close(b.eval(wi, wo, P), port.eval(ndf, P, wi, wo, f), f"sgd-fresnel eval {pname}", min_bit_rate=0.999)
        check_close(b.evalp(wi, wo, P), port.evalp(ndf, P, wi, wo, f), f"sgd-fresnel evalp {pname}", min_bit_rate=0.999)
        gw, gi, gp = b.evalp_is(u, wo, P)
        ww, wi_, wp = port.evalp_is(ndf, P, u, wo, f)
        ok = bits_equal(gi, wi_).all(axis=1)
        assert ok.mean() >= 0.9995
        check_close(gw[ok], ww[ok], f"sgd-fresnel evalp_is weight {pname}")
    cs = np.zeros((4097, 3), np.float32)
    cs[:, 0] = np.linspace(0, 1, 4097, dtype=np.float32)
    assert bits_equal(b.fresnel_term(cs), port.component("fresnel", ndf, None, cs, fresnel=f)).mean() >= 0.999


# ---- chi-square: the sampler draws from the density `pdf` reports -----------------------------------------------------
def _chi2_sample_vs_pdf(b, P, wo, n=1 << 20, nz=24, nphi=48, sub=12, seed=0):
    """Histogram of n sampled directions over an (i.z, phi) grid of the upper hemisphere against n * integral of pdf over
    each cell (midpoint rule with sub x sub points per cell: equal-area cells, so the integral is mean(pdf) * cell area).
    Directions sampled below the horizon have pdf 0 in the reference (G1(i) = 0) and are not binned.  Returns the
    statistic, the degrees of freedom and the p-value.  n = 2^20: the reference's sampler is itself approximate (GGX's
    rational quantile, dj_brdf.h:2138-2146; the 1e-5 tolerance of Beckmann's Newton search, :1939) -- run on the CPU
    oracle, 4e6 samples already expose that (p = 1e-2 .. 1e-11 for the reference itself) while 1e6 do not (p = 0.17 .. 0.79)."""
    from scipy import stats
    rng = np.random.default_rng(seed)
    u = rng.random((n, 2), dtype=np.float32)
    o = np.broadcast_to(np.asarray(wo, np.float32), (n, 3)).copy()
    i = np.asarray(b.sample(u, o, P))
    up = i[:, 2] > 0
    zi = np.minimum((i[up, 2] * nz).astype(np.int64), nz - 1)
    ph = np.arctan2(i[up, 1], i[up, 0]) % (2 * np.pi)
    pi_ = np.minimum((ph / (2 * np.pi) * nphi).astype(np.int64), nphi - 1)
    obs = np.bincount(zi * nphi + pi_, minlength=nz * nphi).astype(np.float64)
    # expected counts
    zs = (np.arange(nz * sub) + 0.5) / (nz * sub)
    ps = (np.arange(nphi * sub) + 0.5) / (nphi * sub) * 2 * np.pi
    Z, PH = np.meshgrid(zs, ps, indexing="ij")
    r = np.sqrt(1 - Z * Z)
    q = np.stack([r * np.cos(PH), r * np.sin(PH), Z], -1).reshape(-1, 3).astype(np.float32)
    oo = np.broadcast_to(np.asarray(wo, np.float32), q.shape).copy()
    pdf = np.asarray(b.pdf(q, oo, P)).astype(np.float64).reshape(nz, sub, nphi, sub)
    cell = (1.0 / nz) * (2 * np.pi / nphi)  # d(z) d(phi) is the solid angle measure
    exp = n * pdf.mean(axis=(1, 3)).reshape(-1) * cell
    # The reference's pdf is gated by G > 0 (dj_brdf.h:1719), i.e. it is 0 for directions behind the mean normal of an
    # off-centre lobe although the sampler does produce them: cells the gate touches are left out.
    gated = ~(pdf > 0).all(axis=(1, 3)).reshape(-1)
    keep = (exp >= 25) & ~gated
    pool = ~keep & ~gated
    # cells with small expectation are pooled into one
    o_k, e_k = np.append(obs[keep], obs[pool].sum()), np.append(exp[keep], exp[pool].sum())
    if e_k[-1] < 25:
        o_k, e_k = o_k[:-1], e_k[:-1]
    stat = float(((o_k - e_k) ** 2 / e_k).sum())
    dof = len(e_k)  # n is not conditioned on: every cell is free (mass below the horizon is not binned)
    return stat, dof, float(stats.chi2.sf(stat, dof)), float(obs[~gated].sum() / n), float(exp[~gated].sum() / n)


@pytest.mark.parametrize("tier", ["bits", "1e-5"])
@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("case", ["iso", "aniso", "offcentre"])
def test_sample_follows_pdf_chi_square(djb, ndf, case, tier):
    cls = djb.ggx if ndf == api.NDF_GGX else djb.beckmann
    b = cls()
    P = {"iso": djb.params.isotropic(0.35), "aniso": djb.params.elliptic(0.25, 0.6, 0.7),
         "offcentre": djb.params.pdfparams(0.4, 0.3, 0.3, 0.1, -0.15)}[case]
    wo = np.array([np.sin(0.7) * np.cos(0.4), np.sin(0.7) * np.sin(0.4), np.cos(0.7)], np.float32)
    try:
        djb.set_precision(tier)  # both tiers of sample and pdf
        stat, dof, p, mass_obs, mass_exp = _chi2_sample_vs_pdf(b, P, wo)
    finally:
        djb.set_precision("bits")
    # the binned mass and the integral of the pdf over the upper hemisphere agree, and the histogram is a plausible draw
    assert abs(mass_obs - mass_exp) < 2e-3, (mass_obs, mass_exp)
    assert p > 1e-4, f"chi2 = {stat:.1f} on {dof} cells, p = {p:.2e}"


def test_sampler_chi_square_detects_a_wrong_density(djb):
    """The test has teeth: the histogram of GGX samples is rejected against the Beckmann pdf of the same roughness."""
    P = djb.params.isotropic(0.35)
    wo = np.array([np.sin(0.7), 0.0, np.cos(0.7)], np.float32)

    class mixed:
        def sample(self, u, o, p): return djb.ggx().sample(u, o, p)
        def pdf(self, i, o, p): return djb.beckmann().pdf(i, o, p)
    stat, dof, p, _, _ = _chi2_sample_vs_pdf(mixed(), P, wo, n=1 << 20)
    assert p < 1e-12


def test_tabular_sample_follows_pdf_chi_square(djb):
    tab = djb.tabular(djb.ggx(), 90)  # fitted from an analytic GGX (alpha = 1): sampling goes through the fitted qf table
    P = djb.params.isotropic(0.4)
    wo = np.array([np.sin(0.5), 0.0, np.cos(0.5)], np.float32)
    stat, dof, p, mass_obs, mass_exp = _chi2_sample_vs_pdf(tab, P, wo, nz=16, nphi=32)
    # djb::tabular samples the NDF (not the visible normals) through a 90-entry piecewise-linear quantile table whose
    # inverse is not exactly the tabulated density (the reference's own plot_qf test shows it: SURVEY section 4), so the
    # histogram is held to a relative bound per cell rather than to the chi-square distribution
    assert abs(mass_obs - mass_exp) < 2e-2, (mass_obs, mass_exp)
    assert stat / dof < 60.0, (stat, dof)


# ---- threading: concurrent calls on the same handles ----------------------------------------------------
def test_four_host_threads_share_one_merl_handle_and_one_descriptor(djb, port):
    n = 50_000
    table = cases.random_merl_table(5)
    m = djb.merl(table)
    f = api.Fresnel.schlick([0.9, 0.5, 0.2])
    g = mk_brdf(djb, api.NDF_GGX, f)
    bk = mk_brdf(djb, api.NDF_BECKMANN, f)
    P = cases.param_sets(port)["aniso"]
    jobs = []
    for t in range(4):
        wi, wo, u = cases.pairs(n, stream=400 + 8 * t)
        jobs.append((wi, wo, u, port.merl_eval(table, wi, wo), port.eval(api.NDF_GGX, P, wi, wo, f),
                     port.pdf(api.NDF_BECKMANN, P, wi, wo, f), port.sample(api.NDF_GGX, P, u, wo)))
    errors = []
    barrier = threading.Barrier(4)

    def work(t):
        try:
            wi, wo, u, w_merl, w_eval, w_pdf, w_smp = jobs[t]
            barrier.wait()
            for rep in range(25):  # host buffers: every call stages, launches and copies back on the calling thread
                assert bits_equal(m.eval(wi, wo), w_merl).all(), "merl"
                check_close(g.eval(wi, wo, P), w_eval, "eval")
                check_close(bk.pdf(wi, wo, P), w_pdf, "pdf")
                assert bits_equal(g.sample(u, wo, P), w_smp).all(axis=1).mean() > 0.9999, "sample"
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            errors.append((t, repr(e)))

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors


# ---- the fits at the sizes the reference uses, against outputs of the reference itself (tests/golden/fixture_golden.npz) ------------
FIXTURE_GOLD = Path(__file__).resolve().parent / "golden" / "fixture_golden.npz"
FIXTURES = Path(__file__).resolve().parent / "_fixtures"  # unpacked by tests/golden/make_fixture_golden.py; not committed (licensed data)


def test_anisotropic_fit_90x90_against_the_reference(djb):
    """djb::tabular_anisotropic(utia, 90, 90) -- the size mitsuba/dj_brdf.cpp:238 uses and bench.py times -- on the seeded synthetic
    UTIA table: 8010 unknowns, 4 power iterations, the 8100-entry tables and both 5-parameter fits against the reference's own output."""
    from tests.test_gpu_fit import check_aniso
    g = np.load(FIXTURE_GOLD)
    want = {k: g[f"utia12_90x90/{k}"] for k in ("p22", "sigma", "fresnel", "beckmann", "ggx")}
    t = djb.tabular_anisotropic(djb.utia(cases.random_utia_table(12)), 90, 90)
    check_aniso(t, want, "utia12 90x90 vs reference")


def test_shipped_fixtures_on_the_device(djb):
    """The two measured materials the reference ships (mitsuba/dj_matpreview.zip), loaded from their files by the device loaders
    and fitted at the reference's sizes: examples/merl_params.cpp on blue-metallic-paint.binary, mitsuba/dj_brdf.cpp:238 on
    m064_fabric099.bin.  Golden = the reference's outputs; the alphas are the ones SURVEY.md section 8c records."""
    from tests.test_gpu_fit import check_aniso, check_fit
    merl_path, utia_path = FIXTURES / "blue-metallic-paint.binary", FIXTURES / "m064_fabric099.bin"
    if not (merl_path.exists() and utia_path.exists()):
        pytest.skip("tests/_fixtures not unpacked (python tests/golden/make_fixture_golden.py where /root/reference exists)")
    g = np.load(FIXTURE_GOLD)
    t = djb.tabular(djb.merl(str(merl_path)), 90)
    check_fit(t, {k: g[f"merl_fixture/{k}"] for k in ("p22", "sigma", "cdf", "qf", "fresnel", "alpha")}, "MERL fixture vs reference")
    assert abs(t.alpha_beckmann - 0.417621881) <= 1e-4 and abs(t.alpha_ggx - 0.172961175) <= 1e-4
    a = djb.tabular_anisotropic(djb.utia(str(utia_path)), 90, 90)
    check_aniso(a, {k: g[f"utia_fixture/{k}"] for k in ("p22", "sigma", "fresnel", "beckmann", "ggx")}, "UTIA fixture vs reference")
    assert np.abs(np.asarray(a.beckmann) - np.array([1.75892246, 1.35479379, -0.00675898697, 0.0259098969, 0.0165748745])).max() <= 1e-4
    assert np.abs(np.asarray(a.ggx)[:2] - np.array([0.727996469, 0.568536818])).max() <= 1e-4


# ---- LEAN maps as the renderer consumes them (half RGBA + mip pyramid) -------------------------------------------
@pytest.mark.parametrize("shape", [(64, 96), (37, 53), (1, 7), (128, 128)])
def test_leanmap_half_rgba_mip_pyramid_bit_exact(djb, port, shape):
    import torch
    h, w = shape
    nm = cases.synthetic_nmap(h, w, seed=9)
    for bias in (0.0, 25.0):
        l1, l2 = port.nmap2leanmap(nm, 1e-5, bias)
        for lm in (l1, l2):
            want = port.leanmap_half_mips(lm)
            got_h = djb.leanmap_half_mips(lm)
            got_d = djb.leanmap_half_mips(torch.from_numpy(lm).cuda())
            assert len(got_h) == len(want) == len(got_d)
            assert want[-1].shape == (1, 1, 4)
            for L, (a, b, c) in enumerate(zip(got_h, got_d, want)):
                assert a.shape == c.shape, (L, a.shape, c.shape)
                assert np.array_equal(a.view(np.uint16), c.view(np.uint16)), ("host", shape, bias, L)
                assert np.array_equal(b.cpu().numpy().view(np.uint16), c.view(np.uint16)), ("device", shape, bias, L)
    # a bounded chain, and the whole asset chain on the device: dmap -> nmap -> LEAN maps -> half mips
    assert len(djb.leanmap_half_mips(l1, levels=3)) == min(3, len(want))
    d = torch.from_numpy(np.random.default_rng(1).integers(96, 160, (64, 64), dtype=np.uint8)).cuda()
    g1, g2 = djb.nmap2leanmap(djb.dmap2nmap(d, 0.01), 1e-5, 25.0)
    w1, _ = port.nmap2leanmap(port.dmap2nmap(d.cpu().numpy(), 0.01), 1e-5, 25.0)
    for a, c in zip(djb.leanmap_half_mips(g1), port.leanmap_half_mips(w1)):
        assert np.array_equal(a.cpu().numpy().view(np.uint16), c.view(np.uint16))


# ---- the remaining public scalar members (VERDICT r01 missing #5) against reference-generated golden vectors ---------------
MEMBER_GOLDEN = Path(__file__).parent / "golden" / "member_golden.npz"


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_quantile_members_against_the_reference(djb, ndf):
    """beckmann / ggx ::qf1, qf2_radial, qf3_radial (dj_brdf.h:366-369, 384-389) -- the pieces `sample` is made of."""
    g = np.load(MEMBER_GOLDEN)
    name = "ggx" if ndf == api.NDF_GGX else "beckmann"
    b = (djb.ggx if ndf == api.NDF_GGX else djb.beckmann)()
    floor = 1.0 if ndf == api.NDF_GGX else 0.999  # Beckmann's search: glibc powf / expf / logf re-run on the device
    for what, got in (("qf1", b.qf1(g["q/u"])), ("qf2_radial", b.qf2_radial(g["q/u"], g["q/cos"], g["q/sin"])),
                      ("qf3_radial", b.qf3_radial(g["q/u3"], g[f"q/{name}/qf2_radial"]))):
        want = g[f"q/{name}/{what}"]
        same = bits_equal(got, want)
        assert same.mean() >= floor, (name, what, same.mean())
        assert rel_err(got, want).max() <= 1e-3, (name, what)
    with pytest.raises(djb.DjbError):  # microfacet::qf2 / qf3 throw in the reference (dj_brdf.h:1783-1791)
        b.qf2(g["q/u"], g["a/i"])


@pytest.mark.parametrize("name", ["gold-metallic-paint", "alum-bronze", "blue-fabric"])
def test_sgd_abc_members_against_the_reference(djb, name):
    """sgd ::ndf / gaf / g1 / fresnel and abc ::ndf / gaf / fresnel (dj_brdf.h:506-509, 531-533)."""
    g = np.load(MEMBER_GOLDEN)
    h, i, o, cs = g["a/h"], g["a/i"], g["a/o"], g["a/cos"]
    s, a = djb.sgd(name), djb.abc(name)
    for kind, m, whats in (("sgd", s, ("ndf", "gaf", "g1", "fresnel")), ("abc", a, ("ndf", "gaf", "fresnel"))):
        for what in whats:
            got = {"ndf": lambda: m.ndf(h), "gaf": lambda: m.gaf(h, i, o), "g1": lambda: m.g1(i),
                   "fresnel": lambda: m.fresnel_term(cs)}[what]()
            want = g[f"a/{kind}/{name}/{what}"]
            assert got.shape == want.shape, (kind, what, got.shape, want.shape)
            same = bits_equal(got, want)
            assert same.mean() >= 0.99, (kind, what, same.mean())   # device double pow / exp / acos vs glibc: last-bit cases
            assert rel_err(got, want).max() <= 1e-5, (kind, what)


def test_tabular_anisotropic_table_members_against_the_reference(djb):
    """tabular_anisotropic::pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 (dj_brdf.h:450-455, 2766-2824) on a handle built from the
    reference's own p22 / sigma tables; arguments run past one period and past the horizon."""
    g = np.load(MEMBER_GOLDEN)
    er, ar = 14, 18
    r = dict(m_p22=g["t/utia12/14x18/p22"], m_sigma=g["t/utia12/14x18/sigma"],
             m_fresnel_points=np.ascontiguousarray(g["t/utia12/14x18/fresnel"]), residuals=np.zeros(4, np.float32),
             beckmann=g["t/utia12/14x18/beckmann"], ggx=g["t/utia12/14x18/ggx"])
    t = djb.tabular_anisotropic(None, er, ar, _result=r)
    phi, theta, u = g["t/phi"], g["t/theta"], g["t/u"]
    for what, got in (("pdf1", t.pdf1(phi)), ("cdf1", t.cdf1(phi)), ("qf1", t.qf1(u)), ("pdf2", t.pdf2(theta, phi)),
                      ("cdf2", t.cdf2(theta, phi)), ("qf2", t.qf2(u, phi))):
        want = g[f"t/utia12/14x18/{what}"]
        same = bits_equal(got, want)
        assert same.mean() >= 0.999, (what, same.mean(), np.abs(got - want).max())
        assert np.abs(got - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max())), what


def test_members_directly_against_the_unmodified_reference(djb, ref):
    """The same members on fresh inputs, reference run on the spot (oracle/_ref travels to the GPU box)."""
    n = 20_000
    u = (api.uniforms(n, 530) * np.float32(0.99998) + np.float32(0.00001)).astype(np.float32)
    c = (np.float32(1e-3) + np.float32(0.999) * api.uniforms(n, 531)).astype(np.float32)
    s = np.sqrt(np.maximum(0.0, 1.0 - c.astype(np.float64) ** 2)).astype(np.float32)
    for ndf, b, floor in ((api.NDF_GGX, djb.ggx(), 1.0), (api.NDF_BECKMANN, djb.beckmann(), 0.999)):
        assert bits_equal(b.qf1(u), ref.member_query("qf1", u, ndf=ndf)).mean() >= floor
        want = ref.member_query("qf2_radial", u, c, s, ndf=ndf)
        assert bits_equal(b.qf2_radial(u, c, s), want).mean() >= floor
        assert bits_equal(b.qf3_radial(u[::-1].copy(), want), ref.member_query("qf3_radial", u[::-1].copy(), want, ndf=ndf)).mean() >= floor
    h, i, o = api.directions(n, 540), api.directions(n, 542), api.directions(n, 544)
    name = "violet-acrylic"
    m = djb.sgd(name)
    for what, got, args in (("ndf", m.ndf(h), (h,)), ("gaf", m.gaf(h, i, o), (h, i, o)), ("g1", m.g1(o), (o,))):
        want = ref.member_query(what, *args, sgd=name)
        assert rel_err(got, want).max() <= 1e-5 and bits_equal(got, want).mean() >= 0.99, what
    m = djb.abc(name)
    for what, got, args in (("ndf", m.ndf(h), (h,)), ("gaf", m.gaf(h, i, o), (h, i, o))):
        want = ref.member_query(what, *args, abc=name)
        assert rel_err(got, want).max() <= 1e-5 and bits_equal(got, want).mean() >= 0.99, what


# ---- the double functions of csrc/djb_dmath.cuh as the device compiles them ----------------------------------------------------------
def _ulps(got, want):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.abs(got - want) / np.spacing(np.abs(want))


def test_device_dmath_against_libm(djb):
    """exp_t / log_t / sqrt_d / acos_d / atan_t / atan2_t / sincos_d / div_core / pow_pos_t evaluated on the device
    (djb200_debug_dmath: MUFU seeds, fused operations, shared-memory tables) against numpy's libm over 2e6 arguments each -- the
    device build has the accuracy tests/cpp/dmath_check.cpp establishes for the host build of the same source."""
    import ctypes as C
    import torch
    from dj_brdf_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(77)
    n = 2_000_000
    u, v = rng.random(n), rng.random(n)

    def run(fn, x, y=None):
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        yd = torch.from_numpy(np.ascontiguousarray(y)).cuda() if y is not None else None
        out = torch.empty(len(x), dtype=torch.float64, device="cuda")
        capi.check(lib.djb200_debug_dmath(C.c_int(fn), C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr() if yd is not None else None),
                                          C.c_int64(len(x)), C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    half = np.arange(n) % 2 == 0
    x = np.where(half, (2 * u - 1) * 700.0, (2 * u - 1) * 2.0)
    assert _ulps(run(0, x), np.exp(x)).max() <= 2.0
    x = np.where(half, 10.0 ** ((2 * u - 1) * 300.0), 0.5 + u)
    lx = np.log(x)
    assert (np.abs(run(1, x) - lx) / (2.0 ** -52 * np.maximum(1.0, np.abs(lx)))).max() <= 2.0  # absolute criterion (djb_dmath.cuh)
    x = np.where(half, 10.0 ** ((2 * u - 1) * 29.0), u)
    assert _ulps(run(2, x), np.sqrt(x)).max() <= 1.0
    x = np.where(half, 2 * u - 1, 1.0 - u ** 4)
    assert _ulps(run(3, x), np.arccos(x)).max() <= 2.0
    x = np.where(half, 10.0 ** ((2 * u - 1) * 25.0) * np.sign(v - 0.5), (2 * u - 1) * 4.0)
    assert _ulps(run(4, x), np.arctan(x)).max() <= 2.0
    y, x = (2 * u - 1), (2 * v - 1) * np.where(half, 1.0, 1e-3)
    assert _ulps(run(5, y, x), np.arctan2(y, x)).max() <= 2.0
    x = np.where(half, (2 * u - 1) * 1e5, (2 * u - 1) * 7.0)
    assert _ulps(run(6, x), np.sin(x)).max() <= 2.0
    assert _ulps(run(7, x), np.cos(x)).max() <= 2.0
    a, b = (2 * u - 1) * 10.0, np.where(half, 10.0 ** ((2 * v - 1) * 29.0), 0.1 + v)
    assert _ulps(run(8, a, b), a / b).max() <= 1.0
    pb, pe = 1e-3 + 4.0 * u, (2 * v - 1) * np.where(half, 500.0, 3.0)
    with np.errstate(over="ignore", under="ignore"):
        want = np.power(pb, pe)
    ok = (want > 1e-300) & (want < 1e300)
    got = run(9, pb, pe)
    err = np.abs(got[ok] - want[ok]) / (want[ok] * 2.0 ** -52 * (1.0 + np.abs(pe * np.log(pb))[ok]))
    k = int(np.nanargmax(err)) if np.isfinite(err).any() else 0
    assert np.isfinite(err).all() and err.max() <= 8.0, (err[k], pb[ok][k], pe[ok][k], got[ok][k], want[ok][k], int((~np.isfinite(err)).sum()))
    # special arguments take the library's path
    x = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-320, 800.0, -800.0, 1.5, 1e31])
    with np.errstate(all="ignore"):
        for fn, f in ((0, np.exp), (1, np.log), (2, np.sqrt), (3, np.arccos), (4, np.arctan), (6, np.sin), (7, np.cos)):
            got, want = run(fn, x), f(x)
            same = (got == want) | (np.isnan(got) & np.isnan(want)) | (_ulps(got, want) <= 2.0)
            if fn == 1:  # log_t's criterion is absolute: log_t(1.0) is 1.7e-18, not 0
                same |= np.abs(got - want) <= 2.0 ** -52 * np.maximum(1.0, np.abs(want))
            assert same.all(), (fn, got, want)
