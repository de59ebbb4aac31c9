#!/bin/bash
# round-1, session e: the section-8f rows (sgd / abc, tabular_anisotropic sampling, fused LEAN shading, plugins, dmap2nmap)
# on one B200 -- full GPU test suite, smoke, both bench arms, the ncu launch list of the bench command, and one
# `ncu --set full` capture of the new kernels (fused LEAN shading evalp, sgd / abc eval, dmap2nmap, sampling-table build).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01_e_bench_n1_reference.json 2> gpurun_out/bench_e.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r01_e_bench_n1.json 2>> gpurun_out/bench_e.err
cat gpurun_out/r01_e_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r01_e_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_e.log 2>&1
cat > /tmp/new_kernels.py <<'PY'
import numpy as np, torch
import dj_brdf_b200 as djb
g = torch.Generator(device="cuda").manual_seed(1)
n = 20_000_000
def dirs():
    z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
    phi = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()
wi, wo = dirs(), dirs()
sl = torch.randn(n, 2, device="cuda", generator=g) * 0.25
v = torch.rand(n, 3, device="cuda", generator=g)
vx, vy = 1e-5 + 0.08 * v[:, 0], 1e-5 + 0.08 * v[:, 1]
E = torch.stack([sl[:, 0] + 25, sl[:, 1] + 25, sl[:, 0] ** 2 + vx, sl[:, 1] ** 2 + vy,
                 sl[:, 0] * sl[:, 1] + (1.4 * v[:, 2] - 0.7) * torch.sqrt(vx * vy) + 625], 1).contiguous()
al = torch.rand(n, 3, device="cuda", generator=g); al[:, :2] = 0.03 + 0.47 * al[:, :2]; al[:, 2] *= 3.14159
b = djb.beckmann()
for _ in range(2):
    b.evalp_lean(wi, wo, E, al)
    djb.sgd("gold-metallic-paint").eval(wi, wo)
    djb.abc("blue-metallic-paint").eval(wi, wo)
    d = torch.randint(0, 256, (8192, 8192), dtype=torch.uint8, device="cuda", generator=g)
    djb.dmap2nmap(d, 0.01)
t = djb.tabular_anisotropic(djb.ggx(), 90, 90)
t.sample(torch.rand(1000, 2, device="cuda"), wo[:1000])
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on \
    -k regex:'mf_lean_kernel|analytic_eval_kernel|dmap2nmap_kernel|aniso_sampling_tables_kernel' -c 12 -f \
    -o gpurun_out/prof_r01_e_new env PYTHONPATH=$PWD python /tmp/new_kernels.py > gpurun_out/ncu_new_e.log 2>&1
tail -2 gpurun_out/ncu_new_e.log
tail -2 gpurun_out/bench_e.err
