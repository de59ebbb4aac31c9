"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 paths (dj_brdf_b200/fit_sharded.py) -- row-block
partitioning, the all-gather of the iterate between power iterations, material round-robin and the residual gather.
The CUDA stages are replaced by a dense numpy matrix here; on the GPU box tests/test_gpu_fit.py runs the same
driver over the real stages."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dj_brdf_b200 import fit_sharded as fs
from dj_brdf_b200 import sharding


def test_shard_rows_partition():
    for n in (1, 7, 89, 8010):
        for world in (1, 2, 3, 8):
            blocks = [fs.shard_rows(n, world, r) for r in range(world)]
            covered = sum(b[1] - b[0] for b in blocks)
            assert covered == n
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            items = sorted(k for r in range(world) for k in fs.shard_items(n, world, r))
            assert items == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, iters, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    K = torch.from_numpy(rng.uniform(0.0, 1.0, (n, n)))

    def matvec(v_in, out, row0, row1):
        v = torch.ones(n, dtype=torch.float64) if v_in is None else v_in
        for r in range(row0, row1):  # sums in index order, like matrix::transform
            acc = 0.0
            for k in range(n):
                acc += float(K[r, k]) * float(v[k])
            out[r] = acc

    v, res = fs.power_iterations_sharded(matvec, n, iters, torch.device("cpu"), rank, world, None)
    srows = fs.sharded_rows_apply(lambda out, a, b: out[a:b].copy_(torch.arange(a, b, dtype=torch.float32)), n,
                                  torch.float32, torch.device("cpu"), rank, world, None)
    q.put((rank, v.numpy().copy(), res.numpy().copy(), srows.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [13, 40])
def test_row_sharded_power_iterations_gloo(n):
    world, iters = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    K = rng.uniform(0.0, 1.0, (n, n))
    v = np.ones(n)
    for _ in range(iters):
        nv = np.zeros(n)
        for r in range(n):
            acc = 0.0
            for k in range(n):
                acc += K[r, k] * v[k]
            nv[r] = acc
        v = nv
    for rank, gv, res, srows in got:
        assert np.array_equal(gv, v), f"rank {rank}: sharded iterate differs from the unsharded one"
        assert np.array_equal(srows, np.arange(n, dtype=np.float32))
        assert res.shape == (iters,) and np.isfinite(res).all() and res[-1] <= res[0] + 1e-6


# ---- LEAN map row bands: the CUDA entry is replaced by the oracle port here ---------------------------
def test_shard_range_partition():
    for n in (1, 5, 64, 8192, 8191):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _lean_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import api
    from tests import cases
    orc = api.PortOracle()
    nm = cases.synthetic_nmap(37, 24)  # 37 rows: uneven bands
    l1, l2 = sharding.nmap2leanmap_sharded(nm, 1e-5, 25.0, gather=True, kernel=orc.nmap2leanmap)
    r0, r1, b1, b2 = sharding.nmap2leanmap_sharded(nm, 1e-5, 25.0, gather=False, kernel=orc.nmap2leanmap)
    q.put((rank, l1.copy(), l2.copy(), r0, r1, np.asarray(b1).copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_lean_map_row_bands_gloo():
    from oracle import api
    from tests import cases
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lean_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w1, w2 = api.PortOracle().nmap2leanmap(cases.synthetic_nmap(37, 24), 1e-5, 25.0)
    for rank, l1, l2, r0, r1, b1 in got:
        assert np.array_equal(l1.view(np.uint32), w1.view(np.uint32)) and np.array_equal(l2.view(np.uint32), w2.view(np.uint32))
        assert (r0, r1) == sharding.shard_range(37, world, rank)
        assert np.array_equal(b1.view(np.uint32), w1[:, r0:r1].view(np.uint32))
