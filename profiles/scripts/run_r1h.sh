#!/bin/bash
# round-1 final record: both bench arms + launch list of the bench command (state after the last kernel change)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01_h_bench_n1_reference.json 2> gpurun_out/bench_h.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r01_h_bench_n1.json 2>> gpurun_out/bench_h.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_h_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_h.log 2>&1
tail -3 gpurun_out/bench_h.err
