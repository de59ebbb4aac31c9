cat > /tmp/e2e.py <<'PY'
import torch, time, numpy as np, ctypes as C, os
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi
lib=capi.load()
n=12_500_000; M=16
P=np.stack([djb.params.elliptic(0.1+0.01*k,0.4,0.7) for k in range(M)])
h_wi=torch.rand(n,3).pin_memory(); h_wo=torch.rand(n,3).pin_memory(); h_out=torch.empty(M*n*3).pin_memory()
h_wi[:,2]+=0.1; h_wo[:,2]+=0.1
d=djb.ggx()._desc()
def call():
    capi.check(lib.djb200_microfacet_eval(C.byref(d),C.c_void_p(P.ctypes.data),C.c_int64(M),C.c_int(0),C.c_void_p(h_wi.data_ptr()),C.c_void_p(h_wo.data_ptr()),C.c_int64(n),C.c_void_p(h_out.data_ptr()),C.c_int(0),None))
call()
t=time.perf_counter(); call(); call(); dt=(time.perf_counter()-t)/2
print(f"eval host call {dt*1e3:.1f} ms -> D2H {M*n*12/dt/1e9:.1f} GB/s, {M*n/dt/1e9:.2f} G evals/s", flush=True)
PY
for mb in 256; do DJB200_TRACE=1 DJB200_CHUNK_MB=$mb PYTHONPATH=$PWD python /tmp/e2e.py 2>&1 | tail -3; done
DJB200_TRACE=1 python bench.py --steps 1 --warmup 1 --pairs 25000000 --no-extras --no-cpu-baseline 2> gpurun_out/e2e_trace.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"
head -30 gpurun_out/e2e_trace.log; tail -5 gpurun_out/e2e_trace.log
