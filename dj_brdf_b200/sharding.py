"""Multi-GPU plumbing of the per-pixel / per-pair maps: independent units, contiguous index ranges per
GPU, replicated parameters and tables, NO data-path collective.  One process per GPU; torch.distributed only carries the
optional gather of the results.

* (wi, wo) pairs of eval / pdf / sample / MERL lookups: `shard_range(n, world, rank)` of the pair arrays (bench.py shards this way);
* LEAN map (utils/nmap2leanmap.cpp:18-54): ROW BANDS of the image -- every texel depends on its own normal only, so a band needs
  no halo: `nmap2leanmap_row_band` converts this rank's band, `nmap2leanmap_sharded(..., gather=True)` also reassembles the two
  maps on every rank.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, world, rank):
    """Contiguous block of rank `rank` out of n units: (begin, end); blocks differ by at most one unit."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def nmap2leanmap_row_band(nmap, rank, world, base_roughness=1e-5, bias=0.0, kernel=None):
    """This rank's band of rows of the LEAN maps of a planar [3, H, W] uint8 normal map (host or device array).
    Returns (row0, row1, leanmap_1[4, rows, W], leanmap_2[4, rows, W]).  `kernel` (tests): stands in for the CUDA entry."""
    if kernel is None:
        from .brdf import nmap2leanmap as kernel
    h = nmap.shape[1]
    row0, row1 = shard_range(h, world, rank)
    band = nmap[:, row0:row1, :]
    band = band.contiguous() if hasattr(band, "contiguous") else np.ascontiguousarray(band)
    l1, l2 = kernel(band, base_roughness, bias)
    return row0, row1, l1, l2


def nmap2leanmap_sharded(nmap, base_roughness=1e-5, bias=0.0, group=None, gather=False, kernel=None):
    """nmap2leanmap with the rows of the image split over the ranks of `group`.  gather=False: (row0, row1, band_1, band_2) of this
    rank (what a renderer that shards its texture the same way keeps).  gather=True: the two complete [4, H, W] maps on every
    rank (bands all-gathered; the only communication, and not part of the conversion itself)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    row0, row1, l1, l2 = nmap2leanmap_row_band(nmap, rank, world, base_roughness, bias, kernel)
    if not gather or world == 1:
        return (row0, row1, l1, l2) if not gather else (l1, l2)
    h, w = nmap.shape[1], nmap.shape[2]
    t1 = l1 if torch.is_tensor(l1) else torch.from_numpy(np.ascontiguousarray(l1))
    t2 = l2 if torch.is_tensor(l2) else torch.from_numpy(np.ascontiguousarray(l2))
    rows_max = (h + world - 1) // world
    full = []
    for t in (t1, t2):
        pad = torch.zeros(4, rows_max, w, dtype=t.dtype, device=t.device)
        pad[:, : row1 - row0] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.empty(4, h, w, dtype=t.dtype, device=t.device)
        for r in range(world):
            a, b = shard_range(h, world, r)
            out[:, a:b] = parts[r][:, : b - a]
        full.append(out if torch.is_tensor(l1) else out.numpy())
    return full[0], full[1]
