# round 2 (j): final state -- whole GPU suite, smoke, both bench arms, launch list, ncu --set full of the fit kernels (single + split launch)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/r02_j_tests.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_j_bench_n1_reference.json 2> gpurun_out/r02_j_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_j_bench_n1.json 2> gpurun_out/r02_j_bench.err
tail -3 gpurun_out/r02_j_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_j_bench_n1.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9,'e2e',d['e2e']['value']/1e9, 'roofline',d['roofline']['kernel'],d['roofline']['frac'])
for k,v in d['kernels'].items(): print(k, round(v['ms'],2), round(v['algo_gbs']/6437.9,3))
for k in ('fit','aniso_fit','lean_shading','sgd','c1'):
    print(k, {a:b for a,b in d[k].items() if isinstance(b,(int,float))}, d[k].get('device_full'), d[k].get('grid_180x180'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_j_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r02_j_bench.log 2>&1
cat > /tmp/fitp.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import dj_brdf_b200 as djb
from tests import cases
tabs=[djb.merl(cases.smooth_merl_table(100+s)) for s in range(8)]
srcs=[tabs[k%8] for k in range(128)]
djb.tabular.fit_packed(srcs,90,True,50)      # single launch: one CTA per material
djb.tabular.fit_packed(srcs[:16],90,True,50) # split mode: 6 launches, 8 CTAs per material
ut = djb.utia(cases.random_utia_table(12))
a = djb.tabular_anisotropic(ut, 90, 90)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"fit_tabular_kernel" -c 7 -f -o gpurun_out/prof_r02_j_fit_iso python /tmp/fitp.py > gpurun_out/ncu_r02_j_fit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aniso_matvec_kernel|aniso_sigma_kernel" -c 5 -f -o gpurun_out/prof_r02_j_fit_aniso python /tmp/fitp.py >> gpurun_out/ncu_r02_j_fit.log 2>&1
tail -2 gpurun_out/ncu_r02_j_fit.log
