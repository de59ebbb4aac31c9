# anisotropic fit: parity tests, then wall time and the per-kernel launch list of a 90 x 90 UTIA fit
python -m pytest tests/test_gpu_fit.py tests/test_gpu_widening.py tests/test_plugins.py -m gpu -q --tb=short 2>&1 | tail -25
cat > /tmp/anisot.py <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from tests import cases
ut = djb.utia(cases.random_utia_table(12))
for _ in range(4):
    torch.cuda.synchronize(); t=time.perf_counter()
    a = djb.tabular_anisotropic(ut, 90, 90)
    torch.cuda.synchronize(); print('aniso 90x90', round((time.perf_counter()-t)*1e3,3),'ms', a.beckmann)
PY
python /tmp/anisot.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_d_aniso_launches.csv python /tmp/anisot.py > /dev/null 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r02_d_aniso_launches.csv').read().split('\n')
k=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[k:]))
idx=[i for i,r in enumerate(rows) if r['Kernel Name'].startswith('aniso_pre_kernel')]
agg={}; tot=0
for r in rows[idx[-1]:]:
    n=r['Kernel Name'].split('(')[0]; t=float(r['Metric Value'])/1e3
    agg.setdefault(n,[0,0]); agg[n][0]+=t; agg[n][1]+=1; tot+=t
for n,(t,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(f"{n:40s} {c:3d} launches {t:9.1f} us")
print('total',tot)
PY
