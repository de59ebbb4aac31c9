mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.2f G/s' % (d['value'] / 1e9), {k: round(v['ms'], 2) for k, v in d['kernels'].items()})" | tee gpurun_out/quick.log
