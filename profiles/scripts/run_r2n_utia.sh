# round 2 (n): utia eval with 32-byte paired entries (one 256-bit load per two taps): parity, then resident-CTA variants
python -m pytest tests -m gpu -q --tb=short -x -k "utia or sgd or abc or widening" 2>&1 | tail -4
for v in "" "-DDJB200_UTIA_MINB=5" "-DDJB200_UTIA_MINB=6" "-DDJB200_UTIA_MINB=3"; do
  DJB200_NVCC_EXTRA="$v" python -m dj_brdf_b200.build --force > /dev/null 2>&1
  grep -A2 "utia_eval_kernel" dj_brdf_b200/build/kernels_tables.ptxas.log | grep -E "Used" | sed 's/ptxas info    ://'
  python profiles/scripts/utia_time.py "[$v]"
done
