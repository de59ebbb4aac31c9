ncu --set full --clock-control none --import-source on -k regex:mf_lean_sample -s 2 -c 2 -f -o gpurun_out/prof_mf_sample python bench.py --steps 1 --warmup 1 --pairs 20000000 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
