import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from tests import cases
tabs=[djb.merl(cases.smooth_merl_table(100+s)) for s in range(8)]
srcs=[tabs[k%8] for k in range(128)]
djb.tabular.fit_batch(srcs[:2],90,True,4)
for it in (4,50):
    torch.cuda.synchronize(); t=time.perf_counter()
    r=djb.tabular.fit_batch(srcs,90,True,it)
    torch.cuda.synchronize(); print(it, (time.perf_counter()-t)*1e3,'ms')
