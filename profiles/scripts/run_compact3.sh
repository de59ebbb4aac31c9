mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "compaction or beckmann or lean_kernels" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('compact-2q', 'value %.2f G/s' % (d['value'] / 1e9), {k: round(v['ms'], 2) for k, v in d['kernels'].items()})" | tee gpurun_out/compact_2q.log
ncu --set full --clock-control none --import-source on -k regex:mf_beck_compact -c 2 -f -o gpurun_out/prof_r01_e_compact3 \
    python bench.py --steps 1 --warmup 0 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_compact2.log 2>&1
tail -1 gpurun_out/ncu_compact2.log
