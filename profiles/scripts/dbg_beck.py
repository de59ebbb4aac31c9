import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi
from oracle import api
from tests import cases
lib = capi.load()
port = api.PortOracle()
wi, wo, _ = cases.pairs(200000)
for pname, P in cases.param_sets(port).items():
    b = djb.beckmann()
    for q in ("eval", "pdf"):
        lib.djb200_debug_force_generic(C.c_int(0)); lean = getattr(b, q)(wi, wo, P)
        lib.djb200_debug_force_generic(C.c_int(1)); gen = getattr(b, q)(wi, wo, P)
        want = getattr(port, q)(api.NDF_BECKMANN, P, wi, wo)
        lean = lean.reshape(len(wi), -1)[:, 0]; gen = gen.reshape(len(wi), -1)[:, 0]; want = want.reshape(len(wi), -1)[:, 0]
        dl = lean.view(np.uint32) != want.view(np.uint32); dg = gen.view(np.uint32) != want.view(np.uint32)
        rel = np.abs(lean.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
        zp = ((lean == 0) != (want == 0)).sum()
        print(f"{pname:10s} {q:5s} lean!=oracle {dl.mean():.2e} generic!=oracle {dg.mean():.2e} max rel {rel.max():.2e} zero-pattern diffs {zp}")
        if dl.any():
            k = np.argmax(rel)
            print("   worst:", wi[k], wo[k], "lean", lean[k], "oracle", want[k], "generic", gen[k])
lib.djb200_debug_force_generic(C.c_int(0))

wi, wo, u = cases.pairs(200000, stream=80)
for ndfname, ndf, cls in (("ggx", api.NDF_GGX, djb.ggx), ("beckmann", api.NDF_BECKMANN, djb.beckmann)):
    for pname in ("iso0.1", "aniso", "offcentre"):
        P = cases.param_sets(port)[pname]
        b = cls()
        lib.djb200_debug_force_generic(C.c_int(0)); lean = b.sample(u, wo, P)
        lib.djb200_debug_force_generic(C.c_int(1)); gen = b.sample(u, wo, P)
        want = port.sample(ndf, P, u, wo)
        sl = (lean.view(np.uint32) == want.view(np.uint32)).all(axis=1).mean()
        sg = (gen.view(np.uint32) == want.view(np.uint32)).all(axis=1).mean()
        slg = (gen.view(np.uint32) == lean.view(np.uint32)).all(axis=1).mean()
        print(f"sample {ndfname:8s} {pname:10s} lean==oracle {sl:.6f} generic==oracle {sg:.6f} lean==generic {slg:.6f}")
lib.djb200_debug_force_generic(C.c_int(0))
