"""Multi-GPU plumbing of the fits: one process per GPU, torch.distributed (NCCL over
NVLink on the GPUs, gloo in the CPU tests) for the exchange steps, the CUDA stages of include/djb200.h for the math.

* isotropic fits shard by MATERIAL: rank r fits materials r, r + world, ...; no data-path collective.  The
  per-iteration residual diagnostics (not in the reference; never fed back) are all-gathered once at the end.
* one anisotropic fit whose n = (elev_res - 1) * azim_res matrix rows span GPUs shards by ROW BLOCK: every power
  iteration each rank computes its rows of K v and the iterate is all-gathered (n doubles = 64 KB at 90 x 90:
  latency-bound on NVLink); the projected-area table is sharded and gathered the same way.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check


def shard_rows(n, world, rank):
    """Contiguous row block of rank `rank`: (row0, row1, chunk) with chunk = ceil(n / world)."""
    chunk = (n + world - 1) // world
    return min(rank * chunk, n), min((rank + 1) * chunk, n), chunk


def shard_items(n, world, rank):
    """Round-robin item indices of rank `rank` (materials of a batched isotropic fit)."""
    return list(range(rank, n, world))


def _all_gather_blocks(local_block, world, group):
    """[chunk] per rank -> [world * chunk] in rank order."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * local_block.numel(), dtype=local_block.dtype, device=local_block.device)
    try:
        dist.all_gather_into_tensor(out, local_block.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = list(out.chunk(world))
        dist.all_gather(parts, local_block.contiguous(), group=group)
    return out


def sharded_rows_apply(fn, n, dtype, device, rank, world, group):
    """Run fn(out_full, row0, row1), which fills out_full[row0:row1], on every rank's own row block and return the
    all-gathered [n] result (the exchange step of a row-sharded stage)."""
    import torch
    row0, row1, chunk = shard_rows(n, world, rank)
    full = torch.zeros(world * chunk, dtype=dtype, device=device)
    fn(full, row0, row1)
    if world == 1:
        return full[:n]
    gathered = _all_gather_blocks(full[rank * chunk:(rank + 1) * chunk], world, group)
    return gathered[:n]


def power_iterations_sharded(matvec, n, iterations, device, rank=0, world=1, group=None):
    """matrix::eigenvector (dj_brdf.h:2467-2480) with the rows of the matrix sharded over `world` ranks.
    matvec(v_in_or_None, out_full, row0, row1) must fill out_full[row0:row1] = (K v_in)[row0:row1]
    (v_in None = the all-ones start vector).  Returns (v, residuals[iterations])."""
    import torch
    v, res = None, []
    for _ in range(iterations):
        prev = v
        v = sharded_rows_apply(lambda out, a, b: matvec(prev, out, a, b), n, torch.float64, device, rank, world, group)
        p = torch.ones_like(v) if prev is None else prev
        c = torch.dot(p, v) / (torch.linalg.vector_norm(p) * torch.linalg.vector_norm(v))
        res.append(torch.sqrt(torch.clamp(2.0 - 2.0 * c, min=0.0)))
    return v, torch.stack(res).to(torch.float32)


_comms = {}  # (id(group), device) -> djb200_comm handle: created once per process group


def library_comm(group=None):
    """The library's own NCCL communicator over the ranks of `group` (djb200_comm_create): rank 0 makes the unique id, torch.distributed
    carries its 128 bytes to the other ranks -- plumbing only; the all-gathers of the fit then run inside libdjb200.so."""
    import torch
    import torch.distributed as dist
    lib = capi.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = (id(group), torch.cuda.current_device())
    if key in _comms:
        return _comms[key]
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        check(lib.djb200_comm_unique_id(buf))
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    dev_ident = ident.cuda() if dist.get_backend(group) == "nccl" else ident
    dist.broadcast(dev_ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(dev_ident.cpu().tolist())
    h = C.c_void_p()
    check(lib.djb200_comm_create(raw, C.c_int32(world), C.c_int32(rank), C.byref(h)))
    _comms[key] = h
    return h


def tabular_anisotropic_sharded(source, elevation_res, azimuthal_res, shadow=True, iterations=4, group=None, timing=None,
                                in_library=True):
    """djb::tabular_anisotropic (dj_brdf.h:2238-2273) with ONE material's matrix rows spanning the GPUs of `group`.
    Every rank passes its own device-resident copy of `source`; every rank returns the full result.
    in_library (default): the iteration loop and the NCCL all-gathers run inside libdjb200.so (djb200_aniso_fit_run); otherwise the
    stage API is driven from here with torch.distributed collectives (the path the CPU gloo tests exercise).
    timing: optional dict, receives device_ms / exchange_ms of the in-library run."""
    import torch
    import torch.distributed as dist
    from .brdf import _source_struct, tabular_anisotropic
    lib = capi.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    s = _source_struct(source)
    h = C.c_void_p()
    sp = capi.current_stream_ptr(capi.MEM_DEVICE)
    check(lib.djb200_aniso_fit_create(C.byref(s), C.c_int32(elevation_res), C.c_int32(azimuthal_res),
                                      C.c_int32(int(shadow)), sp, C.byref(h)))
    try:
        n = int(lib.djb200_aniso_fit_size(h))
        if in_library:
            comm = library_comm(group) if world > 1 else None
            res_host = np.zeros(max(1, iterations), np.float32)
            tm = (C.c_float * 2)()
            check(lib.djb200_aniso_fit_run(h, comm, C.c_int32(iterations), C.c_void_p(res_host.ctypes.data), tm, sp))
            if timing is not None:
                timing.update(device_ms=float(tm[0]), exchange_ms=float(tm[1]))
            residuals_np = res_host
        else:
            def matvec(v_in, out, row0, row1):
                check(lib.djb200_aniso_fit_matvec(h, C.c_void_p(v_in.data_ptr()) if v_in is not None else None,
                                                  C.c_void_p(out.data_ptr()), C.c_int64(row0), C.c_int64(row1), sp))

            v, residuals = power_iterations_sharded(matvec, n, iterations, dev, rank, world, group)
            v = v.contiguous()
            check(lib.djb200_aniso_fit_set_iterate(h, C.c_void_p(v.data_ptr()), sp))
            sigma_rows = sharded_rows_apply(
                lambda out, a, b: check(lib.djb200_aniso_fit_sigma(h, C.c_void_p(out.data_ptr()), C.c_int64(a), C.c_int64(b), sp)),
                n, torch.float32, dev, rank, world, group).contiguous()
            check(lib.djb200_aniso_fit_finish(h, C.c_void_p(sigma_rows.data_ptr()), sp))
            residuals_np = residuals.cpu().numpy()
        er, ar = elevation_res, azimuthal_res
        a = dict(m_p22=np.zeros(er * ar, np.float32), m_sigma=np.zeros(er * ar, np.float32),
                 m_fresnel_points=np.zeros((er, 3), np.float32), residuals=residuals_np)
        f = capi.TabularAnisotropicFit()
        f.elev_res, f.azim_res = er, ar
        f.p22, f.sigma, f.fresnel = a["m_p22"].ctypes.data, a["m_sigma"].ctypes.data, a["m_fresnel_points"].ctypes.data
        check(lib.djb200_aniso_fit_download(h, C.byref(f), sp))
        a["beckmann"] = np.array(list(f.beckmann), np.float32)
        a["ggx"] = np.array(list(f.ggx), np.float32)
    finally:
        lib.djb200_aniso_fit_destroy(h)
    return tabular_anisotropic(None, elevation_res, azimuthal_res, shadow, iterations, _result=a)


def tabular_fit_batch_sharded(make_source, n_materials, resolution=90, shadow=True, iterations=4, group=None, sources=None,
                              packed=False):
    """Batched isotropic fits sharded by material.  make_source(k) builds material k on THIS rank's GPU (only called for the
    materials this rank owns; `sources` = a prebuilt tabular.source_array of those materials skips that).  Returns
    (local fits {k: tabular}, residuals [n_materials, iterations] gathered from every rank); packed=True returns the local fits as the
    packed arrays of tabular.fit_packed instead ({"materials": [k...], "p22": [n_local, res], ...}): no per-material objects."""
    import torch
    import torch.distributed as dist
    from .brdf import tabular
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_items(n_materials, world, rank)
    per_rank = (n_materials + world - 1) // world
    local = torch.zeros(per_rank, iterations, dtype=torch.float32)
    r = None
    if mine:
        srcs = sources if sources is not None else tabular.source_array([make_source(k) for k in mine])
        r = tabular.fit_packed(srcs, resolution, shadow, iterations)
        local[:len(mine)] = torch.from_numpy(r["residuals"])
    if world > 1:
        backend = dist.get_backend(group)
        block = local.cuda() if backend == "nccl" else local
        gathered = _all_gather_blocks(block.reshape(-1), world, group).reshape(world, per_rank, iterations).cpu()
    else:
        gathered = local.reshape(1, per_rank, iterations)
    # rank r owns materials r, r + world, ...: [world, per_rank] -> material order
    residuals = gathered.permute(1, 0, 2).reshape(per_rank * world, iterations)[:n_materials].numpy()
    if packed:
        out = dict(r or {})
        out["materials"] = mine
        return out, residuals
    fits = {}
    if r is not None:
        for j, k in enumerate(mine):
            fits[k] = tabular(None, resolution, shadow, iterations, _result=dict(
                m_p22=r["p22"][j], m_sigma=r["sigma"][j], m_cdf=r["cdf"][j], m_qf=r["qf"][j], m_fresnel_points=r["fresnel"][j],
                residuals=r["residuals"][j], alpha_beckmann=float(r["alpha"][j, 0]), alpha_ggx=float(r["alpha"][j, 1])))
    return fits, residuals
