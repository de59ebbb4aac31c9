# round 2 (e): the 1e-5 tier -- whole GPU suite (exact tier + the fast-tier tests), then the six-kernel step in the default tier
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('ms/step',d['ms_per_step'],'G/s',d['value']/1e9)
for k,v in d['kernels'].items(): print(k, round(v['ms'],2), round(v['algo_gbs']/6437.9,3))
"
DJB200_PRECISION=bits python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('exact tier: ms/step',d['ms_per_step'],'G/s',d['value']/1e9)
for k,v in d['kernels'].items(): print(k, round(v['ms'],2))
"
