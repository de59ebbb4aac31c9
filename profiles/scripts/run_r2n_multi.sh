# round 2 (n): N GPUs -- the in-library NCCL path of the row-sharded anisotropic fit, then the bench at N (bash run_r2h_multi.sh N)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_fit_check.py 2>&1 | grep -v -i "warn\|OMP_NUM" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_n_bench_n$N.json 2> gpurun_out/r02_n_bench_n$N.err
tail -3 gpurun_out/r02_n_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_n_bench_n$N.json'))
print('value G/s', d['value']/1e9, 'ms/step', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value']/1e9)
for k in ('c1','merl','lean','fit','aniso_fit'):
    print(k, json.dumps(d.get(k))[:700])
PY
