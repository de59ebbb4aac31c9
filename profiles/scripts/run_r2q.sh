# round 2 (q): last GPU seconds of the round -- smoke() and the bench (all legs but the host-buffer and CPU ones) on the final code
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_q_bench_n1.json 2> gpurun_out/r02_q_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_q_bench_n1.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9, {k: round(v['ms'],2) for k,v in d['kernels'].items()})
for k in ('sgd','abc','utia','tabular_eval','tabular_anisotropic_sample','lean_shading','merl','lean'):
    print(k, round(d[k]['ms'],3))
"
