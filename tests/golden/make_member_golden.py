#!/usr/bin/env python
"""Golden vectors of the public scalar members added last (dj_brdf.h:366-369, 384-389, 450-455, 506-509, 531-533), generated
from the UNMODIFIED reference (oracle/_ref/libdjbref.so, libdjbref_open.so).  Runs only where /root/reference exists.

    python tests/golden/make_member_golden.py        # rewrites tests/golden/member_golden.npz

Inputs are stored next to the outputs.  The tabular_anisotropic lookups are made on the reference's own object; its p22 /
sigma / Fresnel tables are stored too, so the device handle under test is built from the very same tables.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import api  # noqa: E402
from tests import cases  # noqa: E402

OUT = Path(__file__).resolve().parent
N = 4096


def main():
    assert api.build_ref(), "needs /root/reference"
    ref, ro = api.RefOracle(), api.RefOracle(opened=True)
    g = {}
    # quantile pieces: the variates as `sample` clamps them (dj_brdf.h:1680-1681), view angles over the hemisphere
    u = (api.uniforms(N, 500) * np.float32(0.99998) + np.float32(0.00001)).astype(np.float32)
    u[:6] = [0.00001, 0.99999, 0.5, 0.25, 0.75, 0.49999997]
    c = (np.float32(1e-3) + np.float32(0.999) * api.uniforms(N, 501)).astype(np.float32)
    c[:3] = [1.0, 0.99999, 1e-4]
    s = np.sqrt(np.maximum(0.0, 1.0 - c.astype(np.float64) ** 2)).astype(np.float32)
    g["q/u"], g["q/cos"], g["q/sin"] = u, c, s
    for ndf, name in ((api.NDF_GGX, "ggx"), (api.NDF_BECKMANN, "beckmann")):
        g[f"q/{name}/qf1"] = ref.member_query("qf1", u, ndf=ndf)
        q2 = ref.member_query("qf2_radial", u, c, s, ndf=ndf)
        g[f"q/{name}/qf2_radial"] = q2
        u3 = (api.uniforms(N, 502) * np.float32(0.99998) + np.float32(0.00001)).astype(np.float32)
        g["q/u3"] = u3
        g[f"q/{name}/qf3_radial"] = ref.member_query("qf3_radial", u3, q2, ndf=ndf)
    # sgd / abc per-channel terms
    h, i, o = api.directions(512, 510), api.directions(512, 512), api.directions(512, 514)
    h[:3, 2] = [1.0, 0.5, 1e-3]
    cs = np.linspace(0, 1, 512, dtype=np.float32)
    g["a/h"], g["a/i"], g["a/o"], g["a/cos"] = h, i, o, cs
    for name in ("gold-metallic-paint", "alum-bronze", "blue-fabric"):
        for what, args in (("ndf", (h,)), ("gaf", (h, i, o)), ("g1", (i,)), ("fresnel", (cs,))):
            g[f"a/sgd/{name}/{what}"] = ref.member_query(what, *args, sgd=name)
        for what, args in (("ndf", (h,)), ("gaf", (h, i, o)), ("fresnel", (cs,))):
            g[f"a/abc/{name}/{what}"] = ref.member_query(what, *args, abc=name)
    # tabular_anisotropic table queries on the reference's object
    src, er, ar = api.Source.utia(cases.random_utia_table(12)), 14, 18
    fit = ref.fit_tabular_anisotropic(src, er, ar)
    for k in ("p22", "sigma", "fresnel", "beckmann", "ggx"):
        g[f"t/utia12/14x18/{k}"] = fit[k]
    phi = (api.uniforms(N, 520) * np.float32(8.0) - np.float32(1.0)).astype(np.float32)  # beyond one period on both sides
    theta = (api.uniforms(N, 521) * np.float32(1.7)).astype(np.float32)                    # crosses pi/2
    uu = api.uniforms(N, 522)
    uu[:2] = [0.0, 1.0]
    g["t/phi"], g["t/theta"], g["t/u"] = phi, theta, uu
    for what, a, b in (("pdf1", phi, None), ("cdf1", phi, None), ("qf1", uu, None), ("pdf2", theta, phi), ("cdf2", theta, phi),
                       ("qf2", uu, phi)):
        g[f"t/utia12/14x18/{what}"] = ro.tabular_aniso_lookup(src, er, ar, what, a, b)
    np.savez_compressed(OUT / "member_golden.npz", **g)
    print("member_golden.npz", (OUT / "member_golden.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
