# round 2, first profiling pass: ncu --set full of the Beckmann sampling kernel (after the glibc-exact rewrite), of GGX sample, and of the
# three fit kernels VERDICT r01 asked for (fit_tabular_kernel, aniso_matvec_kernel, aniso_sigma_kernel).
mkdir -p gpurun_out
cat > /tmp/bs.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
import bench
g = torch.Generator(device="cuda").manual_seed(1234)
n = 20_000_000
z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
phi = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
wo = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
u = torch.rand(n, 2, device="cuda", generator=g).contiguous()
a1, a2, ph = bench.materials(16)
P = np.stack([djb.params.elliptic(float(a), float(b), float(c)) for a, b, c in zip(a1, a2, ph)])
for b in (djb.beckmann(), djb.ggx()):
    for _ in range(2):
        b.sample(u, wo, P)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"mf_lean_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r02_a_bsample \
    python /tmp/bs.py > gpurun_out/ncu_r02_a_bsample.log 2>&1
tail -1 gpurun_out/ncu_r02_a_bsample.log
ncu --set full --clock-control none --import-source on -k regex:"mf_lean_kernel" -s 3 -c 1 -f -o gpurun_out/prof_r02_a_gsample \
    python /tmp/bs.py > gpurun_out/ncu_r02_a_gsample.log 2>&1
tail -1 gpurun_out/ncu_r02_a_gsample.log
cat > /tmp/fitp.py <<'PY'
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
ut = djb.utia(cases.random_utia_table(12))
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    a = djb.tabular_anisotropic(ut, 90, 90)
    torch.cuda.synchronize(); print('aniso 90x90', (time.perf_counter()-t)*1e3,'ms')
PY
ncu --set full --clock-control none --import-source on -k regex:"fit_tabular_kernel" -s 2 -c 1 -f -o gpurun_out/prof_r02_a_fit_iso python /tmp/fitp.py > gpurun_out/ncu_r02_a_fit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aniso_matvec_kernel|aniso_sigma_kernel" -s 5 -c 5 -f -o gpurun_out/prof_r02_a_fit_aniso python /tmp/fitp.py >> gpurun_out/ncu_r02_a_fit.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_a_fit_launches.csv python /tmp/fitp.py > /dev/null 2>&1
tail -3 gpurun_out/ncu_r02_a_fit.log
ls -la gpurun_out/*.ncu-rep
