# quick check of the microfacet kernels: sampling / parity tests, then the six-kernel step without the extras
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sample or evalp_is or broadcast or reference" 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9)
for k,v in d['kernels'].items(): print(k, round(v['ms'],2))
"
