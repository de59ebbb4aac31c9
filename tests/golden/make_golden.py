#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref/libdjbref.so, compiled in
place from /root/reference by oracle/Makefile).  Runs only where /root/reference exists.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Inputs are stored next to the outputs (numpy's float64 sin/cos may differ in the last bit between CPUs, so the
seeded generator is NOT re-run at test time for these files).  Tables are regenerated from PCG64 streams, whose
doubles are platform independent; their SHA-256 is stored so that a drifting generator is detected, not mis-tested.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import api  # noqa: E402
from tests import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
N = 1024


sha, random_merl_table, random_utia_table, smooth_merl_table = (cases.sha, cases.random_merl_table,
                                                                 cases.random_utia_table, cases.smooth_merl_table)


def extra(ref):
    """Golden vectors of the section-8f rows added after the first two files (kept in a file of their own so that
    adding a row never rewrites the earlier vectors): djb::sgd / djb::abc."""
    e = {}
    wi, wo, _ = cases.pairs(96, stream=77)
    wi[:4, 2] = [0.0, -0.1, 1.0, 1e-4]
    e["analytic/wi"], e["analytic/wo"] = wi, wo
    import dj_brdf_b200 as djb  # host-only calls: the preset names (the coefficient tables are what is being pinned)
    e["sgd/names"] = np.array(djb.sgd.names())
    e["abc/names"] = np.array(djb.abc.names())
    e["sgd/eval"] = np.stack([ref.sgd_eval(n, wi, wo) for n in djb.sgd.names()])
    e["abc/eval"] = np.stack([ref.abc_eval(n, wi, wo) for n in djb.abc.names()])
    for kind, name in (("sgd", "gold-metallic-paint"), ("abc", "blue-metallic-paint")):
        m = getattr(djb, kind)(name)
        r = ref.fit_tabular(getattr(api.Source, kind)(name, m.coefficients()), 90)
        for k, v in r.items():
            e[f"fit/{kind}/{name}/{k}"] = v
    # tabular_anisotropic sampling: private tables (opened harness) + sample / evalp_is through the reference's object
    ro = api.RefOracle(opened=True)
    src, er, ar = api.Source.utia(cases.random_utia_table(12)), 14, 18
    t = ro.aniso_sampling_tables(src, er, ar)
    for k in ("pdf1", "cdf1", "qf1", "pdf2", "cdf2", "qf2", "p22", "sigma"):
        e[f"aniso_sampling/utia12/14x18/{k}"] = t[k]
    e["aniso_sampling/utia12/14x18/sizes"] = np.array(t["sizes"])
    _, swo, su = cases.pairs(512, stream=400)
    e["aniso_sampling/wo"], e["aniso_sampling/u"] = swo, su
    P = ref.params_pdfparams(0.7, 0.5, 0.3, 0.1, -0.1)
    e["aniso_sampling/params"] = P
    e["aniso_sampling/utia12/14x18/sample"] = ref.tabular_aniso_query("sample", src, er, ar, su, swo, P)
    w, i, pdf = ref.tabular_aniso_query("evalp_is", src, er, ar, su, swo, P)
    e["aniso_sampling/utia12/14x18/evalp_is_w"], e["aniso_sampling/utia12/14x18/evalp_is_i"] = w, i
    e["aniso_sampling/utia12/14x18/evalp_is_pdf"] = pdf
    e["aniso_sampling/utia12/14x18/fresnel"] = ref.fit_tabular_anisotropic(src, er, ar)["fresnel"]
    # LEAN-filtered shading: per-shading-point params (mitsuba/dj_beckmannconductor.cpp:283-314) + the queries on them
    E, alpha = cases.lean_texels(768, seed=9)
    lwi, lwo, lu = cases.pairs(768, stream=700)
    e["lean_shading/E"], e["lean_shading/alpha"] = E, alpha
    e["lean_shading/wi"], e["lean_shading/wo"], e["lean_shading/u"] = lwi, lwo, lu
    for tag, kw in (("lean", dict()), ("mip", dict(lean_filtering=False)), ("scaled", dict(dmap_scale=1.5))):
        P = ref.lean_shading_params(E, alpha, **kw)
        e[f"lean_shading/{tag}/params"] = P
        e[f"lean_shading/{tag}/evalp"] = np.concatenate(
            [ref.evalp(api.NDF_BECKMANN, P[k], lwi[k:k + 1], lwo[k:k + 1]) for k in range(len(P))])
        e[f"lean_shading/{tag}/pdf"] = np.concatenate(
            [ref.pdf(api.NDF_BECKMANN, P[k], lwi[k:k + 1], lwo[k:k + 1]) for k in range(len(P))])
    # dmap2nmap (utils/dmap2nmap.cpp compiled in place)
    rng = np.random.default_rng(3)
    for tag, (h, w, sc) in (("a", (37, 53, 0.1)), ("b", (64, 96, 0.01)), ("c", (3, 5, 1.0))):
        d = rng.integers(0, 256, (h, w), dtype=np.uint8)
        e[f"dmap/{tag}/dmap"], e[f"dmap/{tag}/scale"], e[f"dmap/{tag}/nmap"] = d, np.float32(sc), ref.dmap2nmap(d, sc)
    # the public component queries of djb::microfacet (dj_brdf.h:258-272)
    cwi, cwo, cu = cases.pairs(512, stream=33)
    ch = ((cwi + cwo) / np.linalg.norm(cwi + cwo, axis=1, keepdims=True)).astype(np.float32)
    cxy = np.concatenate([(cu * 4 - 2).astype(np.float32), np.zeros((len(cu), 1), np.float32)], 1)
    ccos = np.concatenate([cu[:, :1], np.zeros((len(cu), 2), np.float32)], 1).astype(np.float32)
    e["components/wi"], e["components/wo"], e["components/h"], e["components/xy"], e["components/cos"] = cwi, cwo, ch, cxy, ccos
    cP = cases.param_sets(ref)["offcentre"]
    e["components/params"] = cP
    cf = api.Fresnel.unpolarized([1.5, 1.8, 2.4])
    cargs = dict(ndf=(ch,), gaf=(ch, cwi, cwo), g1=(ch, cwo), sigma=(cwo,), p22=(cxy,), vp22=(cxy, cwo), vndf=(ch, cwo), fresnel=(ccos,))
    for ndf, nname in ((api.NDF_GGX, "ggx"), (api.NDF_BECKMANN, "beckmann")):
        for what, a in cargs.items():
            e[f"components/{nname}/{what}"] = ref.component(what, ndf, cP, *a, fresnel=cf)
    np.savez_compressed(OUT / "extra_golden.npz", **e)
    print("extra_golden.npz", (OUT / "extra_golden.npz").stat().st_size, "bytes")


def main():
    ref = api.RefOracle()
    if "--extra-only" in sys.argv:
        extra(ref)
        return
    extra(ref)
    wi, wo, u = cases.pairs(N)
    ewi, ewo, eu = cases.edge_pairs()
    wi = np.concatenate([wi, ewi]); wo = np.concatenate([wo, ewo]); u = np.concatenate([u, eu])
    g = {"wi": wi, "wo": wo, "u": u}
    psets = cases.param_sets(ref)
    fres = cases.fresnels()
    for pname, P in psets.items():
        g[f"params/{pname}"] = P
    for ndf, nname in ((api.NDF_GGX, "ggx"), (api.NDF_BECKMANN, "beckmann")):
        for pname, P in psets.items():
            g[f"{nname}/{pname}/eval"] = ref.eval(ndf, P, wi, wo)
            g[f"{nname}/{pname}/evalp"] = ref.evalp(ndf, P, wi, wo)
            g[f"{nname}/{pname}/pdf"] = ref.pdf(ndf, P, wi, wo)
            g[f"{nname}/{pname}/sample"] = ref.sample(ndf, P, u, wo)
        P = psets["aniso"]
        for fname, f in fres.items():
            for shadow in (True, False):
                tag = f"{nname}/aniso/{fname}/shadow{int(shadow)}"
                g[f"{tag}/eval"] = ref.eval(ndf, P, wi, wo, f, shadow)
                g[f"{tag}/pdf"] = ref.pdf(ndf, P, wi, wo, f, shadow)
        w, i, p = ref.evalp_is(ndf, P, u, wo, fres["schlick"])
        g[f"{nname}/aniso/schlick/evalp_is_w"] = w
        g[f"{nname}/aniso/schlick/evalp_is_i"] = i
        g[f"{nname}/aniso/schlick/evalp_is_pdf"] = p
        g[f"{nname}/null_params/eval"] = ref.eval(ndf, None, wi, wo)
    h, d = ref.io_to_hd(wi[:N], wo[:N])
    g["io_to_hd/h"], g["io_to_hd/d"] = h, d
    i2, o2 = ref.hd_to_io(h, d)
    g["hd_to_io/i"], g["hd_to_io/o"] = i2, o2
    g["merl/index"] = ref.merl_index(wi, wo)
    t = random_merl_table(11)
    g["merl/table_seed"] = np.array([11]); g["merl/table_sha256"] = np.array([sha(t)])
    g["merl/eval"] = ref.merl_eval(t, wi, wo)
    ut = random_utia_table(12)
    g["utia/table_seed"] = np.array([12]); g["utia/table_sha256"] = np.array([sha(ut)])
    g["utia/eval"] = ref.utia_eval(ut, wi, wo)
    # LEAN
    for bias in (0.0, 25.0):
        nm = cases.synthetic_nmap(37, 53, seed=5)
        l1, l2 = ref.nmap2leanmap(nm, 1e-5, bias)
        g[f"lean/bias{int(bias)}/nmap"] = nm
        g[f"lean/bias{int(bias)}/l1"], g[f"lean/bias{int(bias)}/l2"] = l1, l2
    E = np.stack([l1[0].ravel() - 25.0, l1[1].ravel() - 25.0, l2[0].ravel(), l2[1].ravel(), l2[2].ravel() - 625.0], 1)
    E = np.ascontiguousarray(E[:512], np.float32)
    g["lrep/E"] = E
    g["lrep/params"] = ref.lrep_to_params(E)
    g["lrep/E_back"] = ref.params_to_lrep(g["lrep/params"])
    np.savez_compressed(OUT / "eval_golden.npz", **g)

    # fits: the reference's tests/plot_qf.cpp / plot_cdf.cpp objects (analytic NDFs through the fit) plus tables
    f = {}
    for ndf, nname in ((api.NDF_GGX, "ggx"), (api.NDF_BECKMANN, "beckmann")):
        for res in (90, 180):
            r = ref.fit_tabular(api.Source.microfacet(ndf), res)
            for k, v in r.items():
                f[f"iso/{nname}/res{res}/{k}"] = v
        r = ref.fit_tabular(api.Source.microfacet(ndf), 64, shadow=False)
        for k, v in r.items():
            f[f"iso/{nname}/res64_noshadow/{k}"] = v
        r = ref.fit_tabular_anisotropic(api.Source.microfacet(ndf), 16, 20)
        for k, v in r.items():
            f[f"aniso/{nname}/16x20/{k}"] = v
    for seed in (21, 22, 23):
        t = smooth_merl_table(seed)
        f[f"iso/merl{seed}/table_sha256"] = np.array([sha(t)])
        r = ref.fit_tabular(api.Source.merl(t), 90)
        for k, v in r.items():
            f[f"iso/merl{seed}/res90/{k}"] = v
    t = smooth_merl_table(21)
    r = ref.fit_tabular_anisotropic(api.Source.merl(t), 12, 16)
    for k, v in r.items():
        f[f"aniso/merl21/12x16/{k}"] = v
    ut = random_utia_table(12)
    r = ref.fit_tabular_anisotropic(api.Source.utia(ut), 14, 18)
    for k, v in r.items():
        f[f"aniso/utia12/14x18/{k}"] = v
    r = ref.fit_tabular(api.Source.utia(ut), 48)
    for k, v in r.items():
        f[f"iso/utia12/res48/{k}"] = v
    np.savez_compressed(OUT / "fit_golden.npz", **f)
    for p in ("eval_golden.npz", "fit_golden.npz"):
        print(p, (OUT / p).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
