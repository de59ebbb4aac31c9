"""GPU parity: the CUDA path (through the C-ABI, via the Python host mirror) against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star): eval / pdf within 1e-5 relative with
identical zero / NaN pattern; MERL indices, LEAN maps and lrep params bit exact."""
import numpy as np
import pytest

from oracle import api
from tests import cases
from tests.conftest import bits_equal, rel_err

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: eval/pdf <= 1e-5 relative FP32


def mk_fresnel(djb, f):
    k = f.kind
    if k == api.F_IDEAL:
        return djb.fresnel.ideal()
    if k == api.F_SCHLICK:
        return djb.fresnel.schlick(f.data[:3])
    if k == api.F_UNPOLARIZED:
        return djb.fresnel.unpolarized(f.data[:3])
    if k == api.F_SPLINE:
        return djb.fresnel.spline(f.data.reshape(-1, 3))
    raise ValueError(k)


def mk_brdf(djb, ndf, f, shadow=True):
    cls = djb.ggx if ndf == api.NDF_GGX else djb.beckmann
    return cls(mk_fresnel(djb, f), shadow)


def check_close(got, want, what, tol=REL_TOL, min_bit_rate=0.0):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    e = rel_err(got, want)
    same = bits_equal(got, want)
    # identical zero pattern
    assert np.array_equal(got == 0, want == 0), f"{what}: zero pattern differs"
    worst = float(e.max()) if e.size else 0.0
    assert worst <= tol, f"{what}: max rel err {worst:.3e} > {tol} (bit-identical {same.mean():.6f})"
    assert same.mean() >= min_bit_rate, f"{what}: bit-identical rate {same.mean():.6f} < {min_bit_rate}"
    return same.mean(), worst


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "iso0.5", "aniso", "aniso2", "offcentre", "standard"])
def test_eval_pdf_parity(djb, port, ndf, pname):
    wi, wo, _ = cases.pairs(cases.N_PARITY)
    P = cases.param_sets(port)[pname]
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    check_close(b.eval(wi, wo, P), port.eval(ndf, P, wi, wo), f"eval {pname}", min_bit_rate=0.9999)
    check_close(b.evalp(wi, wo, P), port.evalp(ndf, P, wi, wo), f"evalp {pname}", min_bit_rate=0.9999)
    check_close(b.pdf(wi, wo, P), port.pdf(ndf, P, wi, wo), f"pdf {pname}", min_bit_rate=0.9999)


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("fname", ["schlick", "unpolarized", "spline"])
@pytest.mark.parametrize("shadow", [True, False])
def test_eval_fresnel_and_shadow(djb, port, ndf, fname, shadow):
    wi, wo, _ = cases.pairs(50_000, stream=16)
    f = cases.fresnels()[fname]
    P = cases.param_sets(port)["aniso"]
    b = mk_brdf(djb, ndf, f, shadow)
    check_close(b.eval(wi, wo, P), port.eval(ndf, P, wi, wo, f, shadow), f"eval {fname} shadow={shadow}",
                min_bit_rate=0.9999)
    check_close(b.pdf(wi, wo, P), port.pdf(ndf, P, wi, wo, f, shadow), f"pdf {fname} shadow={shadow}",
                min_bit_rate=0.9999)


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_null_params_is_standard(djb, port, ndf):
    wi, wo, _ = cases.pairs(10_000, stream=32)
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    check_close(b.eval(wi, wo, None), port.eval(ndf, None, wi, wo), "eval NULL params")


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_edge_inputs(djb, port, ndf):
    wi, wo, u = cases.edge_pairs()
    b = mk_brdf(djb, ndf, api.Fresnel.schlick([0.9, 0.5, 0.2]))
    f = api.Fresnel.schlick([0.9, 0.5, 0.2])
    for pname, P in cases.param_sets(port).items():
        got, want = b.eval(wi, wo, P), port.eval(ndf, P, wi, wo, f)
        assert bits_equal(got, want).all(), (pname, got, want)
        assert bits_equal(b.pdf(wi, wo, P), port.pdf(ndf, P, wi, wo, f)).all(), pname


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_broadcast_16_materials(djb, port, ndf):
    """config 2 layout: every pair under each of 16 anisotropic materials, material-major output."""
    wi, wo, u = cases.pairs(20_000, stream=48)
    mats = cases.c2_materials(port)
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    got = b.eval(wi, wo, mats)
    assert got.shape == (16, len(wi), 3)
    gotp = b.pdf(wi, wo, mats)
    gots = b.sample(u, wo, mats)
    for m in range(16):
        check_close(got[m], port.eval(ndf, mats[m], wi, wo), f"eval material {m}")
        check_close(gotp[m], port.pdf(ndf, mats[m], wi, wo), f"pdf material {m}")
        same = bits_equal(gots[m], port.sample(ndf, mats[m], u, wo))
        assert same.mean() > (0.999 if ndf == api.NDF_GGX else 0.9995), (m, same.mean())


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_per_pair_params(djb, port, ndf):
    n = 4096
    wi, wo, _ = cases.pairs(n, stream=64)
    rng = np.random.default_rng(3)
    blocks = np.stack([port.params_elliptic(float(a), float(b), float(c)) for a, b, c in
                       zip(rng.uniform(0.05, 0.8, n), rng.uniform(0.05, 0.8, n), rng.uniform(0, 3.1, n))])
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    got = b.eval(wi, wo, blocks, per_pair=True)
    want = np.stack([port.eval(ndf, blocks[k], wi[k:k + 1], wo[k:k + 1])[0] for k in range(n)])
    check_close(got, want, "per-pair eval")


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "aniso", "offcentre"])
def test_sample_parity(djb, port, ndf, pname):
    """sample is not under the 1e-5 bar: the mirrored-rounding bit-match rate is
    reported and bounded, and the mismatching tail must stay small in angle."""
    _, wo, u = cases.pairs(cases.N_PARITY, stream=80)
    P = cases.param_sets(port)[pname]
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    got, want = b.sample(u, wo, P), port.sample(ndf, P, u, wo)
    same = bits_equal(got, want).all(axis=1)
    # GGX has no single-precision libm call: it must match to the bit almost everywhere.
    # Beckmann goes through logf / expf / powf: the device runs glibc's own algorithms (csrc/djb_glibcf.h), so it matches too
    # (round 1, which rounded double evaluations instead, matched 99.92-99.97 %).
    assert same.mean() >= (0.99999 if ndf == api.NDF_GGX else 0.9999), same.mean()
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max(axis=1)
    assert np.quantile(err, 0.999) < 1e-3, np.quantile(err, 0.999)


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_evalp_is_parity(djb, port, ndf):
    _, wo, u = cases.pairs(50_000, stream=96)
    P = cases.param_sets(port)["aniso"]
    f = api.Fresnel.schlick([0.9, 0.5, 0.2])
    b = mk_brdf(djb, ndf, f)
    gw, gi, gp = b.evalp_is(u, wo, P)
    ww, wi_, wp = port.evalp_is(ndf, P, u, wo, f)
    ok = bits_equal(gi, wi_).all(axis=1)
    assert ok.mean() >= (0.9999 if ndf == api.NDF_GGX else 0.9995), ok.mean()
    # where the sampled direction agrees to the bit, weight and pdf must meet the eval/pdf bar
    check_close(gw[ok], ww[ok], "evalp_is weight")
    check_close(gp[ok], wp[ok], "evalp_is pdf")


def test_sample_pdf_consistency(djb):
    """Size-independent property: E[1/pdf] over sampled directions == measure of the sampled domain;
    cheaper equivalent used here: weights F*G/G1 are in [0, 1] and pdf > 0 wherever the weight is > 0."""
    import torch
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand(n, 2, device="cuda", generator=g)
    z = 1 - 0.95 * torch.rand(n, device="cuda", generator=g)
    ph = 6.2831853 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(1 - z * z)
    wo = torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
    for cls in (djb.ggx, djb.beckmann):
        w, i, pdf = cls().evalp_is(u, wo, djb.params.elliptic(0.2, 0.5, 0.3))
        w, pdf = w.cpu().numpy(), pdf.cpu().numpy()
        assert np.isfinite(w).all() and (w >= 0).all() and (w <= 1.0 + 1e-5).all()
        assert (pdf[w[:, 0] > 0] > 0).all()


def test_device_and_host_paths_agree(djb, port):
    import torch
    wi, wo, u = cases.pairs(100_000, stream=112)
    mats = cases.c2_materials(port)[:3]
    b = djb.beckmann()
    host = b.eval(wi, wo, mats)
    dev = b.eval(torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda(), mats).cpu().numpy()
    assert bits_equal(host, dev).all()


def test_io_hd_roundtrip_and_parity(djb, port):
    wi, wo, _ = cases.pairs(100_000, stream=128)
    h, d = djb.brdf.io_to_hd(wi, wo)
    hr, dr = port.io_to_hd(wi, wo)
    assert bits_equal(h, hr).mean() > 0.99999 and bits_equal(d, dr).mean() > 0.9999
    i2, o2 = djb.brdf.hd_to_io(hr, dr)
    ir, or_ = port.hd_to_io(hr, dr)
    assert bits_equal(i2, ir).mean() > 0.9999 and bits_equal(o2, or_).mean() > 0.9999
    assert np.abs(i2 - wi).max() < 1e-5


def test_merl_index_bit_exact(djb, port):
    n = 2_000_000
    wi, wo, _ = cases.pairs(n, stream=144)
    got = djb.merl.index(wi, wo)
    want = port.merl_index(wi, wo, nthreads=8)
    mism = int((got != want).sum())
    # north_star: bit exact index arithmetic.  The only tolerated source of difference is a device
    # double acos/atan2 that rounds to another float than glibc's (expected ~0 in 2e6).
    assert mism <= 2, f"{mism} MERL index mismatches in {n}"


def test_merl_eval(djb, port):
    table = cases.synthetic_merl_table()
    m = djb.merl(table)
    wi, wo, _ = cases.pairs(500_000, stream=160)
    got = m.eval(wi, wo)
    want = port.merl_eval(table, wi, wo, nthreads=8)
    idx_ok = djb.merl.index(wi, wo) == port.merl_index(wi, wo, nthreads=8)
    assert bits_equal(got[idx_ok], want[idx_ok]).all()
    assert idx_ok.mean() > 0.999999
    assert (want == 0).all(axis=1).any(), "the synthetic table must exercise the below-horizon branch"
    ewi, ewo, _ = cases.edge_pairs()
    assert bits_equal(m.eval(ewi, ewo), port.merl_eval(table, ewi, ewo)).all()


def test_merl_random_table_bit_exact(djb, port):
    rng = np.random.default_rng(0)
    table = rng.uniform(-0.05, 3.0, 3 * api.MERL_CELLS)  # full double mantissas: exercises the scaling
    m = djb.merl(table)
    wi, wo, _ = cases.pairs(100_000, stream=176)
    got, want = m.eval(wi, wo), port.merl_eval(table, wi, wo, nthreads=8)
    assert bits_equal(got, want).mean() > 0.99999


def test_utia_eval(djb, port):
    rng = np.random.default_rng(1)
    raw = rng.uniform(-0.5, 60.0, 3 * 6 * 48 * 6 * 48)
    t = djb.utia(raw)
    wi, wo, _ = cases.pairs(100_000, stream=192)
    got, want = t.eval(wi, wo), port.utia_eval(raw, wi, wo, nthreads=8)
    check_close(got, want, "utia eval", tol=1e-5, min_bit_rate=0.999)


@pytest.mark.parametrize("bias", [0.0, 25.0])
@pytest.mark.parametrize("shape", [(64, 48), (33, 17), (512, 1024)])
def test_lean_maps_bit_exact(djb, port, bias, shape):
    nm = cases.synthetic_nmap(*shape)
    g1, g2 = djb.nmap2leanmap(nm, 1e-5, bias)
    w1, w2 = port.nmap2leanmap(nm, 1e-5, bias)
    assert bits_equal(g1, w1).all() and bits_equal(g2, w2).all()


def test_lean_device_path_and_params(djb, port):
    import torch
    nm = cases.synthetic_nmap(256, 256)
    g1, g2 = djb.nmap2leanmap(torch.from_numpy(nm).cuda(), 0.05, 0.0)
    w1, w2 = port.nmap2leanmap(nm, 0.05, 0.0)
    assert bits_equal(g1.cpu().numpy(), w1).all() and bits_equal(g2.cpu().numpy(), w2).all()
    P = djb.leanmap_to_params(w1, w2, 0.0)
    E = np.stack([w1[0].ravel(), w1[1].ravel(), w2[0].ravel(), w2[1].ravel(), w2[2].ravel()], 1)
    assert bits_equal(P, port.lrep_to_params(E)).all()
    assert bits_equal(djb.beckmann.lrep_to_params(E), port.lrep_to_params(E)).all()
    assert bits_equal(djb.beckmann.params_to_lrep(P), port.params_to_lrep(P)).all()


def test_empty_and_ragged(djb):
    z = np.zeros((0, 3), np.float32)
    assert djb.ggx().eval(z, z, djb.params.isotropic(0.1)).shape == (0, 3)
    assert djb.merl.index(z, z).shape == (0,)
    for n in (1, 2, 3, 5, 31, 33, 257):
        wi, wo, u = cases.pairs(n, stream=208)
        assert djb.ggx().eval(wi, wo, djb.params.isotropic(0.3)).shape == (n, 3)
        assert djb.beckmann().sample(u, wo, djb.params.isotropic(0.3)).shape == (n, 3)


def test_merl_filter_certifies_only_exact_cells_at_full_size(djb):
    """BASELINE config 3 size (1e8 lookups): the FP32 filter of the MERL lookup may only certify a cell that equals
    the exact (double-arithmetic) one; everything else must be handed to the exact path.  Also on the two input
    families where the filter is weakest (half vector / difference vector near their poles)."""
    import torch
    n = 100_000_000
    g = torch.Generator(device="cuda").manual_seed(11)

    def dirs(k):
        z = 1.0 - 0.999 * torch.rand(k, device="cuda", generator=g)
        ph = 6.283185307179586 * torch.rand(k, device="cuda", generator=g)
        r = torch.sqrt(torch.clamp(1 - z * z, min=0))
        return torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()

    wi, wo = dirs(n), dirs(n)
    s = djb.merl_filter_stats(wi, wo)
    assert s["certified_wrong"] == 0 and s["h_mismatch"] == 0, s
    assert s["rejected"] < 0.01 * n, s
    assert s["max_d_error"] < 1e-6 and s["max_acos_error"] < 1e-6, s  # budgets in kernels_merl.cu are 2e-6 each
    m = 20_000_000
    wo2 = dirs(m)
    spec = wo2.clone()
    spec[:, :2] *= -1
    for base, sg in ((spec, 1e-2), (wo2, 1e-2), (spec, 5e-2), (wo2, 5e-2)):
        w = base + sg * torch.randn(m, 3, device="cuda", generator=g)
        w = (w / w.norm(dim=1, keepdim=True)).contiguous()
        s = djb.merl_filter_stats(w, wo2)
        assert s["certified_wrong"] == 0 and s["h_mismatch"] == 0, s
    # and the product path itself: eval == table[index] for every lookup (index kernel and eval kernel agree)
    rng = np.random.default_rng(0)
    tab = rng.uniform(-0.05, 3.0, 3 * 90 * 90 * 180)
    mm = djb.merl(tab)
    out = mm.eval(wi[:m], wo[:m])
    idx = djb.merl.index(wi[:m], wo[:m]).long()
    cells = np.stack([(tab[:1458000] * (1.00 / 1500.0)).astype(np.float32),
                      (tab[1458000:2916000] * (1.15 / 1500.0)).astype(np.float32),
                      (tab[2916000:] * (1.66 / 1500.0)).astype(np.float32)], 1)
    cells[(cells < 0).any(axis=1)] = 0
    assert torch.equal(torch.from_numpy(cells).cuda()[idx], out)


@pytest.mark.parametrize("ndf", ["ggx", "beckmann"])
def test_lean_kernels_match_mirrored_kernels_at_scale(djb, ndf):
    """The lean FP32 eval / evalp / pdf kernels (csrc/djb_lean.cuh) against the mirrored-rounding kernels that follow
    the reference's double sub-expressions literally: 2e7 pairs x 16 materials = 3.2e8 results per query."""
    import ctypes as C
    import torch
    from dj_brdf_b200 import capi
    lib = capi.load()
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    def dirs():
        z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
        ph = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
        r = torch.sqrt(torch.clamp(1 - z * z, min=0))
        return torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
    wi, wo = dirs(), dirs()
    rng = np.random.default_rng(1)
    mats = np.stack([djb.params.elliptic(float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))),
                                         float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))), float(rng.uniform(0, np.pi)))
                     for _ in range(15)] + [djb.params.pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)])
    b = (djb.ggx if ndf == "ggx" else djb.beckmann)(djb.fresnel.schlick([0.9, 0.5, 0.2]))
    try:
        for q in ("eval", "pdf"):
            lib.djb200_debug_force_generic(C.c_int(0))
            lean = getattr(b, q)(wi, wo, mats)
            lib.djb200_debug_force_generic(C.c_int(1))
            ref = getattr(b, q)(wi, wo, mats)
            same = (lean.view(torch.int32) == ref.view(torch.int32)) | (torch.isnan(lean) & torch.isnan(ref))
            rate = same.float().mean().item()
            assert torch.equal(lean == 0, ref == 0), f"{ndf} {q}: zero pattern differs"
            bad = ~same
            rel = ((lean[bad] - ref[bad]).abs() / ref[bad].abs().clamp_min(1e-30)).max().item() if bad.any() else 0.0
            assert rate >= 0.99999 and rel <= 1e-5, f"{ndf} {q}: bit-identical {rate:.7f}, worst rel {rel:.2e}"
            del lean, ref, same, bad
        u = torch.rand(n, 2, device="cuda", generator=g)
        lib.djb200_debug_force_generic(C.c_int(0))
        lean = b.sample(u, wo, mats)
        lib.djb200_debug_force_generic(C.c_int(1))
        ref = b.sample(u, wo, mats)
        same = ((lean.view(torch.int32) == ref.view(torch.int32)) | (torch.isnan(lean) & torch.isnan(ref))).all(dim=-1)
        rate = same.float().mean().item()
        assert rate >= 0.9999, f"{ndf} sample: lean vs mirrored bit-identical {rate:.6f}"
        # evalp_is and the PER_PAIR layout (roughness from textures) on a smaller set
        k = 2_000_000
        pp = torch.from_numpy(np.ascontiguousarray(mats[np.arange(k) % 16])).cuda()
        outs = {}
        for mode in (0, 1):
            lib.djb200_debug_force_generic(C.c_int(mode))
            w, iv, pd = b.evalp_is(u[:k], wo[:k], mats[3])
            outs[mode] = (w, iv, pd, b.eval(wi[:k], wo[:k], pp), b.pdf(wi[:k], wo[:k], pp), b.sample(u[:k], wo[:k], pp))
        names = ("evalp_is weight", "evalp_is i", "evalp_is pdf", "per-pair eval", "per-pair pdf", "per-pair sample")
        for name, x, y in zip(names, outs[0], outs[1]):
            same = (x.view(torch.int32) == y.view(torch.int32)) | (torch.isnan(x) & torch.isnan(y))
            assert same.float().mean().item() >= 0.9999, f"{ndf} {name}: lean vs mirrored bit-identical {same.float().mean().item():.6f}"
    finally:
        lib.djb200_debug_force_generic(C.c_int(0))


def test_host_pipeline_many_chunks(djb, port, monkeypatch):
    """DJB200_MEM_HOST arrays are staged through the device in chunks (3-slot copy / compute pipeline, one pitched
    D2H copy per chunk and output).  With a 1 MB slot the 100k-pair calls below take ~20 chunks each: results must be
    the same as with device-resident arrays, for every output shape (rgb, scalar, three outputs, PER_PAIR params)."""
    import torch
    monkeypatch.setenv("DJB200_CHUNK_MB", "1")
    n = 100_003  # not a multiple of the chunk size
    wi, wo, u = cases.pairs(n, stream=320)
    mats = cases.c2_materials(port)[:5]
    dwi, dwo, du = (torch.from_numpy(x).cuda() for x in (wi, wo, u))
    for cls in (djb.ggx, djb.beckmann):
        b = cls(djb.fresnel.schlick([0.9, 0.5, 0.2]))
        assert bits_equal(b.eval(wi, wo, mats), b.eval(dwi, dwo, mats).cpu().numpy()).all()
        assert bits_equal(b.pdf(wi, wo, mats), b.pdf(dwi, dwo, mats).cpu().numpy()).all()
        assert bits_equal(b.sample(u, wo, mats), b.sample(du, dwo, mats).cpu().numpy()).all()
        hw, hi, hp = b.evalp_is(u, wo, mats[0])
        gw, gi, gp = b.evalp_is(du, dwo, mats[0])
        assert bits_equal(hw, gw.cpu().numpy()).all() and bits_equal(hi, gi.cpu().numpy()).all() and bits_equal(hp, gp.cpu().numpy()).all()
    rng = np.random.default_rng(4)
    blocks = np.stack([port.params_elliptic(float(a), float(c), float(d)) for a, c, d in
                       zip(rng.uniform(0.05, 0.8, n), rng.uniform(0.05, 0.8, n), rng.uniform(0, 3.1, n))])
    host = djb.ggx().eval(wi, wo, blocks, per_pair=True)
    dev = djb.ggx().eval(dwi, dwo, torch.from_numpy(blocks).cuda()).cpu().numpy()
    assert bits_equal(host, dev).all()
    table = cases.random_merl_table(3)
    m = djb.merl(table)
    assert bits_equal(m.eval(wi, wo), m.eval(dwi, dwo).cpu().numpy()).all()
    assert (djb.merl.index(wi, wo) == djb.merl.index(dwi, dwo).cpu().numpy()).all()


@pytest.mark.parametrize("fname", ["ideal", "schlick"])
def test_beckmann_compaction_is_bit_identical(djb, port, fname):
    """mf_beck_compact_kernel (shadowing work compacted across the warp) against the plain lean kernel: the same functions
    run on the same operands on other lanes, so every result must be the same float -- ragged sizes (tail warps, a single
    pair), edge directions, 2 .. 40 materials, and against the oracle."""
    import ctypes as C
    import torch
    from dj_brdf_b200 import capi
    lib = capi.load()
    f = api.Fresnel.ideal() if fname == "ideal" else cases.fresnels()["schlick"]
    b = mk_brdf(djb, api.NDF_BECKMANN, f)
    ewi, ewo, _ = cases.edge_pairs()
    try:
        for n, nm in ((1, 2), (31, 3), (33, 16), (1_000_003, 16), (200_000, 40), (20_001, 300)):  # 300: two params chunks
            wi, wo, _ = cases.pairs(n, stream=900 + nm)
            if n > 1000:
                wi, wo = np.concatenate([wi, ewi]), np.concatenate([wo, ewo])
            mats = cases.c2_materials(port, nm, seed=nm)
            mats[-1] = port.params_pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)
            twi, two = torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda()
            for q in ("eval", "evalp", "pdf"):
                lib.djb200_debug_beckmann_compaction(C.c_int(1))
                on = getattr(b, q)(twi, two, mats).cpu().numpy()
                lib.djb200_debug_beckmann_compaction(C.c_int(0))
                off = getattr(b, q)(twi, two, mats).cpu().numpy()
                assert bits_equal(on, off).all(), (n, nm, q)
        # against the oracle with compaction on (default)
        lib.djb200_debug_beckmann_compaction(C.c_int(1))
        wi, wo, _ = cases.pairs(50_000, stream=77)
        mats = cases.c2_materials(port, 16)
        got_e, got_p = b.eval(wi, wo, mats), b.pdf(wi, wo, mats)
        for m in range(16):
            check_close(got_e[m], port.eval(api.NDF_BECKMANN, mats[m], wi, wo, f), f"compact eval m{m}", min_bit_rate=0.9999)
            check_close(got_p[m], port.pdf(api.NDF_BECKMANN, mats[m], wi, wo, f), f"compact pdf m{m}", min_bit_rate=0.9999)
    finally:
        lib.djb200_debug_beckmann_compaction(C.c_int(1))


# ---- the 1e-5 tier (the library's default; the tests above run the exact tier, see conftest.py) ---------------------------------
@pytest.fixture
def tier_1e5(djb):
    djb.set_precision("1e-5")
    yield djb
    djb.set_precision("bits")


def check_1e5(got, want, what):
    """north_star: eval / pdf <= 1e-5 relative FP32, identical zero / NaN pattern.  Returns (worst relative error, bit-identical rate)."""
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(got == 0, want == 0), f"{what}: zero pattern differs"
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN pattern differs"
    assert np.array_equal(np.signbit(got[got == 0]), np.signbit(want[want == 0])), f"{what}: sign of zero differs"
    e = rel_err(got, want)
    worst = float(e.max()) if e.size else 0.0
    assert worst <= REL_TOL, f"{what}: max rel err {worst:.3e} > {REL_TOL}"
    return worst, float(bits_equal(got, want).mean())


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "iso0.5", "aniso", "aniso2", "offcentre", "standard"])
def test_fast_tier_against_oracle(tier_1e5, port, ndf, pname):
    """The default tier against the CPU oracle: 200 k pairs + the edge pairs, ideal and Schlick Fresnel, shadowing on / off,
    single material (plain lean kernel) -- <= 1e-5 relative, zero / NaN / signed-zero pattern identical."""
    djb = tier_1e5
    assert djb.get_precision() == "1e-5"
    wi, wo, _ = cases.pairs(cases.N_PARITY)
    ewi, ewo, _ = cases.edge_pairs()
    wi, wo = np.concatenate([wi, ewi]), np.concatenate([wo, ewo])
    P = cases.param_sets(port)[pname]
    for f in (api.Fresnel.ideal(), api.Fresnel.schlick([0.9, 0.5, 0.2])):
        for shadow in (True, False):
            b = mk_brdf(djb, ndf, f, shadow)
            tag = f"{pname} fresnel {f.kind} shadow {shadow}"
            check_1e5(b.eval(wi, wo, P), port.eval(ndf, P, wi, wo, f, shadow), "eval " + tag)
            check_1e5(b.evalp(wi, wo, P), port.evalp(ndf, P, wi, wo, f, shadow), "evalp " + tag)
            check_1e5(b.pdf(wi, wo, P), port.pdf(ndf, P, wi, wo, f, shadow), "pdf " + tag)


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
def test_fast_tier_broadcast_16_materials_against_oracle(tier_1e5, port, ndf):
    """configs[1] layout (16 materials, the warp-compacting Beckmann kernel) in the default tier against the oracle."""
    djb = tier_1e5
    wi, wo, _ = cases.pairs(100_000, stream=16)
    mats = cases.c2_materials(port)
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    got, gotp = b.eval(wi, wo, mats), b.pdf(wi, wo, mats)
    worst = 0.0
    for m in range(16):
        worst = max(worst, check_1e5(got[m], port.eval(ndf, mats[m], wi, wo), f"eval material {m}")[0])
        worst = max(worst, check_1e5(gotp[m], port.pdf(ndf, mats[m], wi, wo), f"pdf material {m}")[0])
    print(f"1e-5 tier, 16 materials, ndf {ndf}: worst relative error {worst:.2e}")


@pytest.mark.parametrize("ndf", ["ggx", "beckmann"])
def test_fast_tier_vs_exact_tier_at_scale(djb, ndf):
    """The 1e-5 tier against the exact tier on the device: 2e7 pairs x 16 materials = 3.2e8 results per query (eval, evalp, pdf;
    Schlick Fresnel; BROADCAST and PER_PAIR layouts): worst relative difference <= 1e-5, zero / NaN pattern identical."""
    import torch
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(7)
    def dirs():
        z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
        ph = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
        r = torch.sqrt(torch.clamp(1 - z * z, min=0))
        return torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
    wi, wo = dirs(), dirs()
    rng = np.random.default_rng(1)
    mats = np.stack([djb.params.elliptic(float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))),
                                         float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))), float(rng.uniform(0, np.pi)))
                     for _ in range(15)] + [djb.params.pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)])
    b = (djb.ggx if ndf == "ggx" else djb.beckmann)(djb.fresnel.schlick([0.9, 0.5, 0.2]))
    report = {}
    try:
        for q in ("eval", "evalp", "pdf"):
            djb.set_precision("1e-5")
            fast = getattr(b, q)(wi, wo, mats)
            djb.set_precision("bits")
            ref = getattr(b, q)(wi, wo, mats)
            assert torch.equal(fast == 0, ref == 0), f"{ndf} {q}: zero pattern differs"
            assert torch.equal(torch.isnan(fast), torch.isnan(ref)), f"{ndf} {q}: NaN pattern differs"
            rel = ((fast - ref).abs() / ref.abs().clamp_min(1e-30))
            rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
            worst = rel.max().item()
            report[q] = worst
            assert worst <= 1e-5, f"{ndf} {q}: worst relative difference {worst:.3e}"
            del fast, ref, rel
        # pdf over centred lobes only: the launch decides the shadowing gate without sigma(i) (djb_lean.cuh, fast_pdf_try)
        djb.set_precision("1e-5")
        fast = b.pdf(wi, wo, mats[:15])
        djb.set_precision("bits")
        ref = b.pdf(wi, wo, mats[:15])
        assert torch.equal(fast == 0, ref == 0), f"{ndf} pdf, centred lobes: zero pattern differs"
        assert torch.equal(torch.isnan(fast), torch.isnan(ref))
        rel = ((fast - ref).abs() / ref.abs().clamp_min(1e-30))
        rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
        report["pdf, centred lobes"] = rel.max().item()
        assert report["pdf, centred lobes"] <= 1e-5, (ndf, report)
        del fast, ref, rel
        k = 2_000_000
        pp = torch.from_numpy(np.ascontiguousarray(mats[np.arange(k) % 16])).cuda()
        for q in ("eval", "pdf"):
            djb.set_precision("1e-5")
            fast = getattr(b, q)(wi[:k], wo[:k], pp, per_pair=True)
            djb.set_precision("bits")
            ref = getattr(b, q)(wi[:k], wo[:k], pp, per_pair=True)
            assert torch.equal(fast == 0, ref == 0)
            rel = ((fast - ref).abs() / ref.abs().clamp_min(1e-30))
            rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
            report["per-pair " + q] = rel.max().item()
            assert report["per-pair " + q] <= 1e-5, (ndf, q, report)
    finally:
        djb.set_precision("bits")
    print(f"1e-5 tier vs exact tier, {ndf}: worst relative difference per query {report}")


# ---- the 1e-5 tier of sample ---------------------------------------------------------------------------------------------
# The sampled direction is compared component-wise (absolute: unit vectors).  GGX: every sample within 1e-5.  Beckmann: the
# quantile search stops at the reference's own test |CDF(b) - u| < 1e-5; where the two tiers' values straddle that threshold the
# fast tier makes one trip more or less and the sample moves by the reference's convergence tolerance -- measured 2.2e-4 of the
# samples beyond 1e-5, 4.4e-6 beyond 1e-4 (3.2e8 samples, profiles/r02_g_fast_sample.md).  The (0, 0, 1) fallback pattern
# (o_std.z <= 0) is identical: the gate is formed by the exact tier's operations.
def sample_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    up = np.array([0.0, 0.0, 1.0])
    assert np.array_equal((got == up).all(axis=-1), (want == up).all(axis=-1)), "the (0, 0, 1) pattern differs"
    assert not np.isnan(got).any()
    return np.abs(got - want).max(axis=-1)


@pytest.mark.parametrize("ndf", [api.NDF_GGX, api.NDF_BECKMANN], ids=["ggx", "beckmann"])
@pytest.mark.parametrize("pname", ["iso0.1", "iso0.5", "aniso", "aniso2", "offcentre", "standard"])
def test_fast_tier_sample_against_oracle(tier_1e5, port, ndf, pname):
    djb = tier_1e5
    _, wo, u = cases.pairs(cases.N_PARITY)
    _, ewo, eu = cases.edge_pairs()
    wo, u = np.concatenate([wo, ewo]), np.concatenate([u, eu])
    P = cases.param_sets(port)[pname]
    b = mk_brdf(djb, ndf, api.Fresnel.ideal())
    err = sample_err(b.sample(u, wo, P), port.sample(ndf, P, u, wo))
    if ndf == api.NDF_GGX:
        assert err.max() <= 1e-5, (pname, err.max())
    else:
        assert (err <= 1e-5).mean() >= 0.999 and (err <= 1e-4).mean() >= 0.9999 and err.max() <= 2e-2, \
            (pname, (err <= 1e-5).mean(), (err <= 1e-4).mean(), err.max())
    assert np.median(err) <= 2.5e-7


@pytest.mark.parametrize("ndf", ["ggx", "beckmann"])
def test_fast_tier_sample_vs_exact_tier_at_scale(djb, port, ndf):
    """2e7 pairs x 16 materials (one of them off-centre) = 3.2e8 samples per distribution, the 1e-5 tier against the exact tier."""
    import torch
    n = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(11)
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
    ph = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    wo = torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
    u = torch.rand(n, 2, device="cuda", generator=g)
    mats = cases.c2_materials(port)
    mats[15] = djb.params.pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)
    b = (djb.ggx if ndf == "ggx" else djb.beckmann)()
    try:
        djb.set_precision("1e-5")
        fast = b.sample(u, wo, mats)
        djb.set_precision("bits")
        exact = b.sample(u, wo, mats)
    finally:
        djb.set_precision("bits")
    assert not torch.isnan(fast).any()
    up = torch.tensor([0.0, 0.0, 1.0], device="cuda")
    assert torch.equal((fast == up).all(-1), (exact == up).all(-1)), "the (0, 0, 1) pattern differs"
    d = (fast - exact).abs().amax(dim=-1).reshape(-1)
    f5, f4, worst = (d <= 1e-5).double().mean().item(), (d <= 1e-4).double().mean().item(), d.max().item()
    print(f"1e-5 tier sample vs exact tier, {ndf}: within 1e-5 {f5:.7f}, within 1e-4 {f4:.7f}, max {worst:.3e}, mean {d.double().mean().item():.2e}")
    if ndf == "ggx":
        assert worst <= 1e-5
    else:
        assert f5 >= 0.9995 and f4 >= 0.99998 and worst <= 5e-2
    assert d.double().mean().item() <= 3e-7


@pytest.mark.parametrize("ndf", ["ggx", "beckmann"])
def test_fast_tier_many_materials_and_host_arrays(djb, port, ndf):
    """The default tier through the other plumbing: 300 params blocks (more than one shared-memory stage of 256, not in the kernel
    arguments), host (numpy) arrays through the staged pipeline, an odd pair count -- against the exact tier on the same inputs."""
    n = 10_007
    wi, wo, u = cases.pairs(n, stream=700)
    rng = np.random.default_rng(9)
    mats = np.stack([djb.params.elliptic(float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))),
                                         float(np.exp(rng.uniform(np.log(0.02), np.log(0.8)))), float(rng.uniform(0, np.pi)))
                     for _ in range(299)] + [djb.params.pdfparams(0.3, 0.2, 0.4, 0.3, -0.4)])
    b = (djb.ggx if ndf == "ggx" else djb.beckmann)(djb.fresnel.schlick([0.9, 0.5, 0.2]))
    try:
        djb.set_precision("bits")
        want = {q: getattr(b, q)(wi, wo, mats) for q in ("eval", "pdf")}
        want_s = b.sample(u, wo, mats)
        djb.set_precision("1e-5")
        for q in ("eval", "pdf"):
            got = getattr(b, q)(wi, wo, mats)
            assert got.shape == want[q].shape and got.shape[0] == 300
            check_1e5(got, want[q], f"{ndf} {q}, 300 materials, host arrays")
        err = sample_err(b.sample(u, wo, mats), want_s)
        assert (err <= 1e-5).mean() >= (1.0 if ndf == "ggx" else 0.999), (err <= 1e-5).mean()
    finally:
        djb.set_precision("bits")
