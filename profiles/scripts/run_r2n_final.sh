# round 2 (n): final state after the djb_dmath work -- whole GPU suite, smoke, both bench arms, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/r02_n_tests.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_n_bench_n1_reference.json 2> gpurun_out/r02_n_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_n_bench_n1.json 2> gpurun_out/r02_n_bench.err
tail -3 gpurun_out/r02_n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_n_bench_n1.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9,'e2e',d['e2e']['value']/1e9, 'roofline',d['roofline']['kernel'],d['roofline']['frac'])
for k,v in d['kernels'].items(): print(k, round(v['ms'],2), round(v['algo_gbs']/6437.9,3))
for k in ('fit','aniso_fit','lean_shading','sgd','c1'):
    print(k, {a:b for a,b in d[k].items() if isinstance(b,(int,float))}, d[k].get('device_full'), d[k].get('grid_180x180'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_n_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r02_n_bench.log 2>&1
