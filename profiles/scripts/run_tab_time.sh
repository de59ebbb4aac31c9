cat > /tmp/tt.py <<'PY'
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
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): f()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
pv = lambda t: C.c_void_p(t.data_ptr())
for src in ("ggx", "beckmann"):
    t = djb.tabular(getattr(djb, src)(), 90)
    h, _ = t._first_arg()
    ms = timed(lambda: capi.check(lib.djb200_tabular_eval(h, None, C.c_int64(0), C.c_int(0), pv(wi), pv(wo), C.c_int64(n), pv(out), C.c_int(1), sptr)))
    print(src, "tabular eval via C-ABI: %.3f ms" % ms)
    ms = timed(lambda: t.eval(wi, wo))
    print(src, "tabular eval via python: %.3f ms" % ms)
    ms = timed(lambda: t.pdf(wi, wo)); print(src, "pdf %.3f ms" % ms)
    u = torch.rand(n, 2, device="cuda", generator=g)
    ms = timed(lambda: t.sample(u, wo)); print(src, "sample %.3f ms" % ms)
PY
python /tmp/tt.py
