# round 2 (n): which of the lean double functions in the params construction costs the fused LEAN shading kernel time
for v in "" "-DDJB200_LIBM_PARAMS_A" "-DDJB200_LIBM_PARAMS_B" "-DDJB200_LIBM_PARAMS_C" "-DDJB200_LIBM_PARAMS_A -DDJB200_LIBM_PARAMS_B -DDJB200_LIBM_PARAMS_C" "-DDJB200_LEANSRC_MINB=1" "-DDJB200_LEANSRC_MINB=3"; do
  DJB200_NVCC_EXTRA="$v" python -m dj_brdf_b200.build --force > /dev/null 2>&1
  grep -A2 "mf_lean_kernelILi0ELi0ELi1ELi2ELb0" dj_brdf_b200/build/kernels_mf.ptxas.log | grep -E "Used" | sed 's/ptxas info    ://'
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --pairs 50000000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1])
print('[$v]', 'lean_shading', round(d['lean_shading']['ms'],3), 'two_pass', round(d['lean_shading']['two_pass_ms'],3), 'tab_eval', round(d['tabular_eval']['ms'],3), 'taniso_sample', round(d['tabular_anisotropic_sample']['ms'],3))"
done
