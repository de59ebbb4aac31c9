# round 2 (l): GGX r2 contracted in the 1e-5 tier (tests + timing), then the fast Beckmann sampling kernel at 4 resident CTAs
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "fast_tier" 2>&1 | tail -3
show() { python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1])
print('$1', d['ms_per_step'], {k: round(v['ms'],2) for k,v in d['kernels'].items()})
"; }
show base
DJB200_NVCC_EXTRA="-DDJB200_FSAMPLE_MINB=4" python -m dj_brdf_b200.build > /dev/null 2>&1
grep -A3 "mf_lean_kernelILi0ELi0ELi3ELi0ELb1E" dj_brdf_b200/build/kernels_mf.ptxas.log | grep -E "Used|spill" | head -2
show minb4
