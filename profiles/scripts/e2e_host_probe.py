"""Host ceiling of the e2e leg: every rank copies device <-> pinned host memory at the same time (what the DJB200_MEM_HOST path
does: 13.6 GB up, 89.6 GB down per rank and step), nothing else.  Reports per-rank and aggregate GB/s for

  * D2H into default pinned memory (cudaHostAlloc), one stream and two streams per rank
  * D2H into write-combined pinned memory (cudaHostAllocWriteCombined)
  * H2D, and both directions at once

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/e2e_host_probe.py

One JSON line on rank 0.  The aggregate D2H figure is the ceiling of bench.py's e2e value: 11.2 B of results per query.
"""
import ctypes as C
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = C.CDLL("libcudart.so.12")
GB = 1 << 30
NBYTES = 2 * GB
dev_a = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
    return p


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=4):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
D2H, H2D = 2, 1


def copy(dst, src, nbytes, kind, stream):
    assert rt.cudaMemcpyAsync(C.c_void_p(dst), C.c_void_p(src), C.c_size_t(nbytes), C.c_int(kind), C.c_void_p(stream.cuda_stream)) == 0


res = {"n_gpus": world, "bytes_per_copy": NBYTES}
for name, flags in (("pinned", 0), ("write_combined", 4)):
    h1, h2 = host_alloc(NBYTES, flags), host_alloc(NBYTES, flags)
    da, db = dev_a.data_ptr(), dev_b.data_ptr()
    dt = timed(lambda: copy(h1.value, da, NBYTES, D2H, s1))
    res[f"d2h_{name}_gbs_per_gpu"] = NBYTES / dt / 1e9
    dt = timed(lambda: (copy(h1.value, da, NBYTES // 2, D2H, s1), copy(h2.value, db, NBYTES // 2, D2H, s2)))
    res[f"d2h_{name}_two_streams_gbs_per_gpu"] = NBYTES / dt / 1e9
    dt = timed(lambda: copy(da, h1.value, NBYTES, H2D, s1))
    res[f"h2d_{name}_gbs_per_gpu"] = NBYTES / dt / 1e9
    dt = timed(lambda: (copy(h1.value, da, NBYTES, D2H, s1), copy(db, h2.value, NBYTES, H2D, s2)))
    res[f"bidir_{name}_gbs_per_gpu_each_way"] = NBYTES / dt / 1e9
    rt.cudaFreeHost(h1)
    rt.cudaFreeHost(h2)
for k in list(res):
    if k.endswith("_per_gpu") or k.endswith("each_way"):
        res[k.replace("_per_gpu", "_aggregate")] = res[k] * world
try:
    res["cpus_visible"] = len(os.sched_getaffinity(0))
except Exception:
    pass
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
