import numpy as np, torch, time
import dj_brdf_b200 as djb
from dj_brdf_b200 import fit_sharded as fs, workloads
src = djb.utia(np.random.default_rng(12).uniform(-0.5, 60.0, 3 * 6 * 48 * 6 * 48))
def t5(tag):
    ts=[]
    for _ in range(5):
        tm={}
        t0=time.perf_counter(); fs.tabular_anisotropic_sharded(src, 90, 90, True, 4, timing=tm); ts.append((time.perf_counter()-t0)*1e3)
    print(tag, [round(t,2) for t in ts], tm, flush=True)
t5('fresh')
specs = workloads.fit_table_specs(16)
tabs=[djb.merl(workloads.fit_table(s)) for s in specs]
djb.tabular.fit_packed(djb.tabular.source_array(tabs), 90, True, 50)
t5('after 16 iso fits (split mode)')
many = djb.tabular.source_array([tabs[k % 16] for k in range(1184)])
djb.tabular.fit_packed(many, 90, True, 50)
t5('after 1184 iso fits')
x = torch.empty(20_000_000_000 // 4, device='cuda')
t5('after a 20 GB torch tensor')
