# round 2 (c): full single-GPU pass: smoke, GPU tests, bench (both arms), launch list of the bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c_bench_n1_reference.json 2> gpurun_out/r02_c_ref.err
python bench.py > gpurun_out/r02_c_bench_n1.json 2> gpurun_out/r02_c_bench.err; tail -5 gpurun_out/r02_c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c_bench_n1.json'))
print('value G/s', d['value']/1e9, 'ms/step', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value']/1e9, 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],4))
for k,v in d['kernels'].items(): print(' ', k, round(v['ms'],2))
for k in ('c1','merl','lean','fit','aniso_fit'):
    print(k, json.dumps(d.get(k))[:900])
print('cpu', d['cpu_baseline'])
PY
