import sys, ctypes as C, numpy as np
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi
from oracle import api
from tests import cases
lib = capi.load(); port = api.PortOracle()
wi, wo, _ = cases.pairs(200000)
P = cases.param_sets(port)["iso0.1"]
b = djb.beckmann()
want = port.eval(api.NDF_BECKMANN, P, wi, wo)[:, 0]
for flags, name in ((0, "all lean"), (1, "sigma generic"), (2, "p22 generic"), (3, "both generic")):
    lib.djb200_debug_force_generic(C.c_int(10 + flags))
    got = b.eval(wi, wo, P)[:, 0]
    d = got.view(np.uint32) != want.view(np.uint32)
    print(f"{name:14s}: mismatch {d.mean():.2e}", "magnitudes of mismatching:", np.sort(np.abs(want[d]))[[0, len(want[d]) // 2, -1]] if d.any() else "")
