N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then python profiles/scripts/e2e_host_probe.py > gpurun_out/r02_e2e_probe_n$N.json 2>gpurun_out/r02_e2e_probe_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 profiles/scripts/e2e_host_probe.py > gpurun_out/r02_e2e_probe_n$N.json 2>gpurun_out/r02_e2e_probe_n$N.err; fi
tail -2 gpurun_out/r02_e2e_probe_n$N.err; cat gpurun_out/r02_e2e_probe_n$N.json | tail -1
lscpu | grep -E "Socket|NUMA|^CPU\(s\)|Model name" | head -8; free -g | head -2
