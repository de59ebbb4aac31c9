N=${1:-2}
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "Socket|NUMA|^CPU\(s\)|Model name" | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 2 --warmup 3 --no-extras 2>gpurun_out/bench_numa.err | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value']/1e9,'e2e',d['e2e'])"
tail -2 gpurun_out/bench_numa.err
