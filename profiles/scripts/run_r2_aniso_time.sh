cat > /tmp/anisot.py <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from tests import cases
ut = djb.utia(cases.random_utia_table(12))
ts=[]
for _ in range(12):
    torch.cuda.synchronize(); t=time.perf_counter()
    a = djb.tabular_anisotropic(ut, 90, 90)
    torch.cuda.synchronize(); ts.append(round((time.perf_counter()-t)*1e3,2))
print('aniso 90x90 ms', ts)
import os
os.environ["DJB200_TRACE"]="1"
PY
python /tmp/anisot.py
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,utilization.gpu --format=csv
