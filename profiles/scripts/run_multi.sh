N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_fit_check.py 2>&1 | grep -v -i "warn\|OMP_NUM" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 2 --warmup 3 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step","gpu_launches","clocks")}); print(d["e2e"]); print(d["roofline"])
PY
tail -3 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
