# round 2 (g2): tests of the 1e-5 tier (sample + the looser ill-conditioning gate of G), then the step + lean-shading timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_widening.py tests/test_facade_cpp.py tests/test_plugins.py -m gpu -q --tb=short -k "fast_tier or chi_square or lean or facade or plugin" -s 2>&1 | grep -v "^$" | tail -25
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/r02_g_bench_quick.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_g_bench_quick.json').read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9)
for k,v in d['kernels'].items(): print(k, round(v['ms'],2), round(v['algo_gbs']/6437.9,3))
print('lean_shading', d['lean_shading']['ms'], d['lean_shading']['two_pass_ms'])
"
DJB200_PRECISION=bits python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('bits: ms/step',d['ms_per_step']); print('lean_shading', d['lean_shading']['ms'], d['lean_shading']['two_pass_ms'])
"
