# round 2 (n): racecheck / memcheck / synccheck of the kernels that now stage the djb_dmath.cuh table in shared memory or use
# 256-bit loads: sgd / abc eval (plain and general path), utia eval (32-byte entries), tabular / tabular_anisotropic queries,
# the LEAN-source fused shading kernels, a fit from a UTIA / SGD source, djb200_debug_dmath
mkdir -p gpurun_out
cat > /tmp/san_n.py <<'PY'
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi
from oracle import api
from tests import cases
lib = capi.load()
for n in (1, 33, 3001):
    wi, wo, u = cases.pairs(n, stream=70 + n % 7)
    if n > 1000:
        ewi, ewo, eu = cases.edge_pairs()
        wi, wo, u = np.concatenate([wi, ewi]), np.concatenate([wo, ewo]), np.concatenate([u, eu])
    twi, two, tu = torch.from_numpy(wi).cuda(), torch.from_numpy(wo).cuda(), torch.from_numpy(u).cuda()
    for name in ("gold-metallic-paint", "alumina-oxide", "aluminium", "white-fabric"):
        djb.sgd(name).eval(twi, two); djb.sgd(name).eval(wi, wo)
        djb.abc(name).eval(twi, two)
    ut = djb.utia(cases.random_utia_table(12))
    ut.eval(twi, two); ut.eval(wi, wo)
    t = djb.tabular(djb.ggx(), 90)
    t.eval(twi, two); t.pdf(twi, two); t.sample(tu, two); t.evalp_is(tu, two)
    ta = djb.tabular_anisotropic(djb.ggx(), 20, 24)
    ta.eval(twi, two); ta.pdf(twi, two); ta.sample(tu, two)
djb.tabular(djb.sgd("gold-metallic-paint"), 32)
djb.tabular(djb.utia(cases.random_utia_table(5)), 32)
x = torch.linspace(-2.0, 2.0, 4097, dtype=torch.float64, device="cuda")
y = torch.linspace(0.5, 3.0, 4097, dtype=torch.float64, device="cuda")
o = torch.empty_like(x)
for fn in range(10):
    capi.check(lib.djb200_debug_dmath(C.c_int(fn), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_int64(x.numel()), C.c_void_p(o.data_ptr()), None))
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in racecheck memcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool python /tmp/san_n.py 2>&1 | grep -E "SUMMARY|sanitizer workload|Error|hazard|Race|Traceback|rror:" | head -12
done > gpurun_out/sanitizer_n.log 2>&1
cat gpurun_out/sanitizer_n.log
