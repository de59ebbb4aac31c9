#!/bin/bash
# round-1, session e (records r01_f and r01_g were taken with this script): full GPU suite, smoke, both bench arms, launch list, ncu of the compacting kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01_g_bench_n1_reference.json 2> gpurun_out/bench_g.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r01_g_bench_n1.json 2>> gpurun_out/bench_g.err
cat gpurun_out/r01_g_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r01_g_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mf_beck_compact|mf_lean_kernel' -c 6 -f \
    -o gpurun_out/prof_r01_g_mf python bench.py --steps 1 --warmup 0 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_mf_g.log 2>&1
tail -1 gpurun_out/ncu_mf_g.log
python /tmp/dm.py 2>/dev/null || true
tail -2 gpurun_out/bench_g.err
