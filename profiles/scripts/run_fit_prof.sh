cat > /tmp/fitp.py <<'PY'
import time, numpy as np, torch
import dj_brdf_b200 as djb
from tests import cases
tabs=[djb.merl(cases.smooth_merl_table(100+s)) for s in range(8)]
srcs=[tabs[k%8] for k in range(128)]
djb.tabular.fit_batch(srcs[:2],90,True,4)
for it in (4,50):
    torch.cuda.synchronize(); t=time.perf_counter()
    r=djb.tabular.fit_batch(srcs,90,True,it)
    torch.cuda.synchronize(); print(it, (time.perf_counter()-t)*1e3,'ms')
PY
PYTHONPATH=$PWD ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fit_tabular --csv python /tmp/fitp.py 2>&1 | grep -E "fit_tabular|ms$" | cut -c1-300
PYTHONPATH=$PWD ncu --set full --clock-control none --import-source on -k regex:fit_tabular -s 1 -c 1 -f -o gpurun_out/prof_fit_iso python /tmp/fitp.py > /dev/null 2>&1
