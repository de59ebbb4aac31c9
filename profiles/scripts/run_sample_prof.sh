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
b = djb.beckmann()
for _ in range(2):
    b.sample(u, wo, P)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"mf_beck_sample_pool|mf_lean_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r01_e_bsample \
    python /tmp/bs.py > gpurun_out/ncu_bsample.log 2>&1
tail -1 gpurun_out/ncu_bsample.log
