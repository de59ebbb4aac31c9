#!/bin/bash
# round-1 final measurement on one B200: tests, smoke, both bench arms, ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2>> gpurun_out/bench_final.err
cat gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/bench_final.err
