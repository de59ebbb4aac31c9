set -x
python -m pytest tests/test_gpu_fit.py -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch
import dj_brdf_b200 as djb
from tests import cases
ut=cases.random_utia_table(12)
src=djb.utia(ut)
djb.tabular_anisotropic(src,20,20)
for it in (4,):
    torch.cuda.synchronize(); t=time.perf_counter()
    r=djb.tabular_anisotropic(src,90,90,True,it)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(f"aniso 90x90 iters={it}: {dt*1e3:.1f} ms; beckmann={r.beckmann} ggx={r.ggx} resid={r.residuals}")
PY
