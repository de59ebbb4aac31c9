# A/B of the warp-compacting Beckmann kernel: bit-equality test, then timing with compaction on / off
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "compaction or beckmann" 2>&1 | tail -5
for mode in on off; do
  if [ $mode = off ]; then export DJB200_MF_NOCOMPACT=1; else unset DJB200_MF_NOCOMPACT; fi
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$mode', 'value %.2f G/s' % (d['value'] / 1e9), {k: round(v['ms'], 2) for k, v in d['kernels'].items()})"
done | tee gpurun_out/compact_ab.log
