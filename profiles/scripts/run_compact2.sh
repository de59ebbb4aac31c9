# where does the warp-compacting Beckmann kernel spend its time? ncu of the kernel, then the occupancy variant
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mf_beck_compact -c 2 -f -o gpurun_out/prof_r01_e_compact \
    python bench.py --steps 1 --warmup 0 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_compact.log 2>&1
tail -1 gpurun_out/ncu_compact.log
for minb in 4 5; do
  DJB200_NVCC_EXTRA="-DDJB200_COMPACT_MINB=$minb" python -m dj_brdf_b200.build > /dev/null 2>&1
  grep -A3 "mf_beck_compact_kernelILi0ELi0" dj_brdf_b200/build/kernels_mf.ptxas.log | grep -E "Used|spill" | head -2
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('minb=$minb', 'value %.2f G/s' % (d['value'] / 1e9), {k: round(v['ms'], 2) for k, v in d['kernels'].items()})"
done | tee gpurun_out/compact_minb.log
