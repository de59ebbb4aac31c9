mkdir -p gpurun_out
cat > /tmp/tb.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
g = torch.Generator(device="cuda").manual_seed(1234)
n = 20_000_000
def dirs():
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
    phi = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
wi, wo = dirs(), dirs()
t = djb.tabular(djb.ggx(), 90)
for _ in range(2):
    t.eval(wi, wo)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:tabular_query_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_h_tab \
    env PYTHONPATH=$PWD python /tmp/tb.py > gpurun_out/ncu_tab.log 2>&1
tail -1 gpurun_out/ncu_tab.log
