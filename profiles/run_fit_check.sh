set -x
python -m pytest tests/test_gpu_fit.py -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch
import dj_brdf_b200 as djb
from tests import cases
tabs=[cases.smooth_merl_table(100+s) for s in range(8)]
src=[djb.merl(t) for t in tabs]
srcs=[src[k%8] for k in range(128)]
for it in (4,50):
    djb.tabular.fit_batch(srcs[:4],90,True,it)
    torch.cuda.synchronize(); t=time.perf_counter()
    r=djb.tabular.fit_batch(srcs,90,True,it)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(f"iters={it}: 128 fits in {dt*1e3:.2f} ms -> {128/dt:.0f} fits/s; alpha_ggx[0]={r[0].alpha_ggx}")
PY
