"""Multi-GPU check of the fits, run under torchrun on the GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_fit_check.py

* one anisotropic fit whose matrix rows span the GPUs (NCCL all-gather of the iterate every power iteration) must
  be bit-identical to the same fit on one GPU;
* config 4 of BASELINE.json: 128 synthetic MERL tables, 50 iterations, sharded by material, residuals gathered.
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import dj_brdf_b200 as djb  # noqa: E402
from dj_brdf_b200 import fit_sharded as fs  # noqa: E402
from tests import cases  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ut = cases.random_utia_table(12)
    src = djb.utia(ut)
    er, ar = 90, 90
    fs.tabular_anisotropic_sharded(src, 20, 20)  # warm-up (the library's NCCL communicator)
    fs.tabular_anisotropic_sharded(src, 20, 20, in_library=False)  # warm-up (torch's)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    driven = fs.tabular_anisotropic_sharded(src, er, ar, True, 4, in_library=False)  # stage API driven from Python + torch collectives
    torch.cuda.synchronize(); dist.barrier()
    t_py = time.perf_counter() - t0
    tm = {}
    t0 = time.perf_counter()
    sharded = fs.tabular_anisotropic_sharded(src, er, ar, True, 4, timing=tm)  # loop + ncclAllGather inside libdjb200.so
    torch.cuda.synchronize(); dist.barrier()
    t_sh = time.perf_counter() - t0
    assert np.array_equal(driven.m_p22.view(np.uint32), sharded.m_p22.view(np.uint32))
    assert np.array_equal(driven.m_sigma.view(np.uint32), sharded.m_sigma.view(np.uint32))
    t0 = time.perf_counter()
    single = djb.tabular_anisotropic(src, er, ar, True, 4)
    t_1 = time.perf_counter() - t0
    same = all(np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32)) for a, b in
               ((sharded.m_p22, single.m_p22), (sharded.m_sigma, single.m_sigma),
                (sharded.m_fresnel_points, single.m_fresnel_points), (sharded.beckmann, single.beckmann), (sharded.ggx, single.ggx)))
    ok = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"aniso 90x90 row-sharded over {world} GPUs: in-library {t_sh * 1e3:.1f} ms (device {tm['device_ms']:.2f} ms, exchanges "
              f"{tm['exchange_ms']:.3f} ms), Python-driven {t_py * 1e3:.1f} ms, single GPU {t_1 * 1e3:.1f} ms; "
              f"bit-identical on every rank: {bool(ok.item())}; beckmann {sharded.beckmann}", flush=True)
    assert ok.item() == 1

    tables = {}

    def make(k):
        if k % 8 not in tables:
            tables[k % 8] = djb.merl(cases.smooth_merl_table(100 + k % 8))
        return tables[k % 8]

    fs.tabular_fit_batch_sharded(make, world, 90, True, 4)  # warm-up
    torch.cuda.synchronize(); dist.barrier()
    for iters in (4, 50):
        t0 = time.perf_counter()
        fits, residuals = fs.tabular_fit_batch_sharded(make, 128, 90, True, iters)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        assert residuals.shape == (128, iters) and np.isfinite(residuals).all() and (residuals[:, -1] > 0).any() or iters == 50
        if rank == 0:
            print(f"config 4: 128 MERL fits x {iters} iterations over {world} GPUs: {dt * 1e3:.2f} ms -> {128 / dt:.0f} fits/s; "
                  f"alpha_ggx[0] = {fits[0].alpha_ggx:.6f}; max final residual {residuals[:, -1].max():.3e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
