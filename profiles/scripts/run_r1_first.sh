#!/bin/bash
# first GPU measurement of round 1: smoke, tests, bench, ncu launch list, ncu --set full of the microfacet kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
cat gpurun_out/bench_first.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_first_ref.json 2>> gpurun_out/bench_first.err
cat gpurun_out/bench_first_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_first.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mf_broadcast -s 6 -c 6 -f -o gpurun_out/prof_mf_first \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merl_eval|lean_kernel" -c 2 -f -o gpurun_out/prof_tables_first \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
