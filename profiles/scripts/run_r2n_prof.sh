# round 2 (n): ncu --set full of the three FP64-bound widened kernels (sgd eval, utia eval, tabular eval), source counters kept
mkdir -p gpurun_out
cat > /tmp/wk.py <<'PY'
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
sg = djb.sgd("gold-metallic-paint")
ut = djb.utia(np.random.default_rng(3).uniform(0.0, 40.0, 3 * 6 * 48 * 6 * 48))
tab = djb.tabular(djb.ggx(), 90)
th, _k = tab._first_arg()
for _ in range(2):
    capi.check(lib.djb200_sgd_eval(C.byref(sg._data), pv(wi), pv(wo), C.c_int64(n), pv(out), C.c_int(capi.MEM_DEVICE), sptr))
    capi.check(lib.djb200_utia_eval(ut._h, pv(wi), pv(wo), C.c_int64(n), pv(out), C.c_int(capi.MEM_DEVICE), sptr))
    capi.check(lib.djb200_tabular_eval(th, None, C.c_int64(0), C.c_int(capi.PARAMS_BROADCAST), pv(wi), pv(wo), C.c_int64(n), pv(out), C.c_int(capi.MEM_DEVICE), sptr))
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"analytic_eval_kernel|utia_eval_kernel|tabular_query_kernel" -s 3 -c 3 -f \
    -o gpurun_out/prof_r02_n3_widened env PYTHONPATH=$PWD python /tmp/wk.py > gpurun_out/ncu_r02_n3.log 2>&1
tail -2 gpurun_out/ncu_r02_n3.log
ls -la gpurun_out/prof_r02_n3_widened.ncu-rep
