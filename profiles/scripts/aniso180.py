import numpy as np, torch, time
import dj_brdf_b200 as djb
from dj_brdf_b200 import fit_sharded as fs
src = djb.utia(np.random.default_rng(12).uniform(-0.5, 60.0, 3 * 6 * 48 * 6 * 48))
for res in (90, 180):
    for _ in range(3):
        tm={}
        t0=time.perf_counter(); f=fs.tabular_anisotropic_sharded(src, res, res, True, 4, timing=tm); dt=time.perf_counter()-t0
    print(res, 'wall ms', dt*1e3, tm, f.beckmann)
whole = djb.tabular_anisotropic(src, 180, 180)
print('plain entry', whole.beckmann, np.array_equal(whole.m_p22, f.m_p22), np.array_equal(whole.m_sigma, f.m_sigma))
