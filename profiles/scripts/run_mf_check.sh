python -m pytest tests/test_gpu_parity.py tests/test_facade_cpp.py -x -q 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value %.2f G/s'%(d['value']/1e9)); [print(k, '%.1f ms %.1f G/s'%(v['ms'],v['gevals_per_s'])) for k,v in d['kernels'].items()]"
