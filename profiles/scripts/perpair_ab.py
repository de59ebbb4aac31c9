"""A/B of the two tiers on the kernels that do not compact their work: PER_PAIR params and LEAN-texel params (Beckmann, GGX)."""
import numpy as np, torch, ctypes as C
import dj_brdf_b200 as djb
from oracle import api
from tests import cases
port = api.PortOracle()
n = 20_000_000
g = torch.Generator(device="cuda").manual_seed(3)
def dirs():
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
    ph = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    return torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
wi, wo = dirs(), dirs()
mats = cases.c2_materials(port)
pp = torch.from_numpy(np.ascontiguousarray(mats[np.arange(n) % 16])).cuda()
def timeit(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); f(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3
E = torch.randn(n, 5, device="cuda", generator=g) * 0.05
E[:, 0:2] += 25.0; E[:, 2:4] = E[:, 2:4].abs() + 625.0 + 0.02; E[:, 4] += 625.0
for name, cls in (("beckmann", djb.beckmann), ("ggx", djb.ggx)):
    b = cls(djb.fresnel.schlick([0.9, 0.5, 0.2]))
    for mode in ("bits", "1e-5"):
        djb.set_precision(mode)
        t_pp = timeit(lambda: b.eval(wi, wo, pp, per_pair=True))
        t_pdf = timeit(lambda: b.pdf(wi, wo, pp, per_pair=True))
        t_one = timeit(lambda: b.eval(wi, wo, mats[3]))
        line = f"{name} {mode}: per-pair eval {t_pp:.3f} ms, per-pair pdf {t_pdf:.3f} ms, one material eval {t_one:.3f} ms"
        if name == "beckmann":
            t_lean = timeit(lambda: b.evalp_lean(wi, wo, E, [0.1, 0.2, 0.3]))
            line += f", lean evalp {t_lean:.3f} ms"
        print(line, flush=True)
djb.set_precision("bits")
