# round 2 (q): ncu launch list of the six step kernels on the final code (shares to compare with the bench's kernel times)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_q_launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_r02_q_bench.log 2>&1
tail -1 gpurun_out/ncu_r02_q_bench.log | cut -c1-200
grep -c "gpu__time_duration" gpurun_out/r02_q_launches.csv
