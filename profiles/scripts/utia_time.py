"""utia eval timing (2e7 pairs, the bench's table): prints ms per launch"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi
lib = capi.load()
g = torch.Generator(device="cuda").manual_seed(1234)
n = 20_000_000
def dirs():
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
    phi = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
wi, wo = dirs(), dirs()
out = torch.empty(n, 3, device="cuda")
stream = torch.cuda.current_stream(); sptr = C.c_void_p(stream.cuda_stream)
pv = lambda t: C.c_void_p(t.data_ptr())
ut = djb.utia(np.random.default_rng(3).uniform(0.0, 40.0, 3 * 6 * 48 * 6 * 48))
f = lambda: capi.check(lib.djb200_utia_eval(ut._h, pv(wi), pv(wo), C.c_int64(n), pv(out), C.c_int(capi.MEM_DEVICE), sptr))
f(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(5): f()
b.record(stream); torch.cuda.synchronize()
print(sys.argv[1] if len(sys.argv) > 1 else "", "utia eval %.3f ms" % (a.elapsed_time(b) / 5), float(out.sum()))
