# round 2 (i): the mask-screening compaction kernel -- parity (A/B against the plain kernel, both tiers, facade), then timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_facade_cpp.py tests/test_gpu_round2.py -m gpu -q --tb=short -x 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('1e-5: ms/step',d['ms_per_step'],'G/s',d['value']/1e9)
for k,v in d['kernels'].items(): print(k, round(v['ms'],2), round(v['algo_gbs']/6437.9,3))
"
DJB200_PRECISION=bits python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('bits: ms/step',d['ms_per_step']); print({k: round(v['ms'],2) for k,v in d['kernels'].items()})
"
