#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 BRDF engine (contract: task spec / DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): batched GGX + Beckmann eval / pdf / sample over 1e8 (wi, wo) pairs x 16
anisotropic materials on each GPU.  One "step" = the six kernels (2 NDF families x 3 queries), every pair under
every material: 6 x 16 x 1e8 = 9.6e9 BRDF queries per GPU per step.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm (CUDA, through the C-ABI)
    python bench.py --impl reference [...]                         # the reference's own CPU code, host cores
    torchrun --nproc-per-node N bench.py --gpus N ...              # N > 1: one rank per GPU, weak scaling

* `value`  : whole-job BRDF queries/s with inputs resident in HBM (device pointers through the C-ABI).
* `e2e`    : the same step through the C-ABI with HOST (pinned) buffers: H2D of the directions and D2H of every
             result inside the timed region.
* `roofline`: the dominant kernel of the step against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
* `cpu_baseline`: oracle/_ref (the unmodified reference compiled in place) or the oracle port, timed on the
             host cores on a bounded sample of the same workload (rank 0, N = 1 only).

The oracle is executed here only as the CPU baseline / reference arm, never as the thing measured for `value`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "brdf_evals_per_s"
UNIT = "evals/s"
OPS = ("eval", "pdf", "sample")
NDFS = ("ggx", "beckmann")
# algorithmic bytes per (pair, material) query at M materials:
#   eval:   24 B pair read once + 12 B result per material
#   pdf:    24 B pair read once +  4 B result per material
#   sample: 20 B (u, wo) read once + 12 B result per material


def algo_bytes(op, pairs, mats):
    per_pair_in = {"eval": 24, "pdf": 24, "sample": 20}[op]
    per_out = {"eval": 12, "pdf": 4, "sample": 12}[op]
    return pairs * per_pair_in + pairs * mats * per_out + 48 * mats


def materials(m=16, seed=1):
    """config 2: alpha1, alpha2 log-uniform in [0.02, 0.8], phi_a uniform in [0, pi)."""
    rng = np.random.default_rng(seed)
    a1 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    a2 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    ph = rng.uniform(0, np.pi, m).astype(np.float32)
    return a1, a2, ph


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region with NVML
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code (oracle/_ref) or the oracle port, all host threads
def cpu_arm(pairs, mats, steps, warmup, threads):
    from oracle import api
    if api.ref_available():
        orc, kind = api.RefOracle(), "reference"
    else:
        orc, kind = api.PortOracle(), "port"
    wi = api.directions(pairs, 0)
    wo = api.directions(pairs, 2)
    u = np.stack([api.uniforms(pairs, 4), api.uniforms(pairs, 5)], axis=1)
    a1, a2, ph = materials(mats)
    P = [orc.params_elliptic(float(a), float(b), float(c)) for a, b, c in zip(a1, a2, ph)]

    def step():
        for ndf in (api.NDF_GGX, api.NDF_BECKMANN):
            for p in P:
                orc.eval(ndf, p, wi, wo, nthreads=threads)
                orc.pdf(ndf, p, wi, wo, nthreads=threads)
                orc.sample(ndf, p, u, wo, nthreads=threads)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    q = 6.0 * pairs * mats * steps
    return {"value": q / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{pairs} pairs x {mats} materials x 6 queries per step, {steps} steps (same seeded generator)",
            "ms_per_step": 1e3 * dt / steps}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    pairs = args.cpu_pairs
    r = cpu_arm(pairs, args.materials, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 at the reference's double sub-expressions)",
        "data": "synthetic",
        "config": {"workload": "configs[1]: GGX+Beckmann eval/pdf/sample, pairs x 16 anisotropic materials",
                   "pairs_per_step": pairs, "materials": args.materials,
                   "note": "bounded sample of the 1e8-pair workload on the host cores"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pin this rank's threads (and hence its pinned host buffers, first-touched below) to the CPUs NVML reports as
    local to GPU `index`: with one process per GPU the host<->device copies of the e2e leg then stay on the GPU's
    own socket instead of crossing the inter-socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import dj_brdf_b200 as djb
    from dj_brdf_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- dj_brdf_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = capi.load()

    pairs, M = args.pairs, args.materials
    a1, a2, ph = materials(M)
    P = np.stack([djb.params.elliptic(float(a), float(b), float(c)) for a, b, c in zip(a1, a2, ph)])

    # synthetic directions, generated on the device (throughput set; the parity sets are host generated)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def dirs():
        z = 1.0 - 0.999 * torch.rand(pairs, device=dev, generator=g)
        phi = 6.283185307179586 * torch.rand(pairs, device=dev, generator=g)
        r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
        return torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1).contiguous()

    wi, wo = dirs(), dirs()
    u = torch.rand(pairs, 2, device=dev, generator=g).contiguous()
    out = torch.empty(M * pairs * 3, dtype=torch.float32, device=dev)  # reused by every query (19.2 GB at 1e8 x 16)
    brdfs = {"ggx": djb.ggx(), "beckmann": djb.beckmann()}
    descs = {k: b._desc() for k, b in brdfs.items()}
    fn = {"eval": lib.djb200_microfacet_eval, "pdf": lib.djb200_microfacet_pdf, "sample": lib.djb200_microfacet_sample}
    Pp = C.c_void_p(P.ctypes.data)

    def launch(ndf, op, a, b, n, o, mem, stream):
        capi.check(fn[op](C.byref(descs[ndf]), Pp, C.c_int64(M), C.c_int(capi.PARAMS_BROADCAST),
                          C.c_void_p(a), C.c_void_p(b), C.c_int64(n), C.c_void_p(o), C.c_int(mem), stream))

    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    ins = {"eval": wi, "pdf": wi, "sample": u}
    kernels = [(n, o) for n in NDFS for o in OPS]

    def step(events=None):
        for k, (ndf, op) in enumerate(kernels):
            if events is not None:
                events[k][0].record(stream)
            launch(ndf, op, ins[op].data_ptr(), wo.data_ptr(), pairs, out.data_ptr(), capi.MEM_DEVICE, sptr)
            if events is not None:
                events[k][1].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in kernels]
          for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = djb.kernel_launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_beg.record(stream)
        for s in range(args.steps):
            step(ev[s])
        t_end.record(stream)
        barrier()
    launches = djb.kernel_launch_count() - launches0
    ms_total = t_beg.elapsed_time(t_end)
    kern_ms = {f"{n}_{o}": float(np.mean([ev[s][k][0].elapsed_time(ev[s][k][1]) for s in range(args.steps)]))
               for k, (n, o) in enumerate(kernels)}

    if os.environ.get("DJB200_BENCH_TRACE"):
        for s in range(args.steps):
            print(f"[bench] rank {rank} step {s}: " + " ".join(f"{n}_{o}={ev[s][k][0].elapsed_time(ev[s][k][1]):.1f}ms"
                                                                 for k, (n, o) in enumerate(kernels)), file=sys.stderr)
    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    q_step = 6.0 * pairs * M * world
    value = q_step / (ms_step * 1e-3)

    # ---- end to end through the C-ABI with host buffers (pinned), in slabs of the same 1e8-pair workload ----
    slab = min(pairs, args.e2e_slab)
    n_slabs = (pairs + slab - 1) // slab
    h_wi = torch.empty(slab, 3, dtype=torch.float32).pin_memory()
    h_wo = torch.empty(slab, 3, dtype=torch.float32).pin_memory()
    h_u = torch.empty(slab, 2, dtype=torch.float32).pin_memory()
    h_out = torch.empty(M * slab * 3, dtype=torch.float32).pin_memory()
    h_wi.copy_(wi[:slab]); h_wo.copy_(wo[:slab]); h_u.copy_(u[:slab])
    torch.cuda.synchronize()
    h_ins = {"eval": h_wi, "pdf": h_wi, "sample": h_u}

    def e2e_step():
        for s in range(n_slabs):
            n = min(slab, pairs - s * slab)
            for ndf, op in kernels:
                launch(ndf, op, h_ins[op].data_ptr(), h_wo.data_ptr(), n, h_out.data_ptr(), capi.MEM_HOST, None)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.no_e2e:
        e2e_steps, n_slabs = 1, 0
    e2e_step() if args.warmup else None
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = q_step * e2e_steps / e2e_s
    h2d = sum({"eval": 24, "pdf": 24, "sample": 20}[o] for _, o in kernels) * pairs
    d2h = sum({"eval": 12, "pdf": 4, "sample": 12}[o] for _, o in kernels) * pairs * M

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = measured_peaks()
    dom = max(kern_ms, key=kern_ms.get)
    dom_op = dom.split("_")[1]
    ab = algo_bytes(dom_op, pairs, M)
    achieved = ab / (kern_ms[dom] * 1e-3) / 1e9
    # DRAM bytes of that kernel from the committed ncu --set full capture (profiles/dram_traffic.json), scaled to
    # this launch size (the kernels stream: bytes are linear in the number of pairs)
    traffic = None
    try:
        tj = json.loads((ROOT / "profiles" / "dram_traffic.json").read_text())["kernels"][dom]
        if tj.get("materials", M) == M:
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * pairs / tj["pairs"]
    except Exception:
        pass
    per_kernel = {k: {"ms": v, "gevals_per_s": pairs * M / (v * 1e-3) / 1e9,
                      "algo_gbs": algo_bytes(k.split("_")[1], pairs, M) / (v * 1e-3) / 1e9} for k, v in kern_ms.items()}

    extra = {}
    if args.extras and rank == 0:
        extra = run_extras(args, djb, capi, lib, torch, dev, wi, wo, out, stream, sptr, peak)
    if args.extras:
        extra.update(run_fit_extra(djb, torch, world, rank))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args.cpu_pairs, M, 1, 1, os.cpu_count() or 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 at the reference's double sub-expressions)", "data": "synthetic",
            "config": {"workload": "configs[1]: GGX+Beckmann eval/pdf/sample, 1e8 pairs x 16 anisotropic materials per GPU",
                       "pairs_per_gpu": pairs, "materials": M, "queries_per_step_per_gpu": 6 * pairs * M,
                       "l2": "inputs (2.4 GB) and outputs (19.2 GB) larger than L2; no flush needed",
                       "sharding": "pairs sharded across ranks, params replicated, no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ab,
                         "note": "issue-bound, not HBM-bound: reproducing the reference's rounded floats costs ~190 (GGX eval) to "
                                 "~1700 (Beckmann sample: 5 Newton steps of erfinv + exp per sample) warp instructions per "
                                 "result; DRAM traffic equals the algorithmic bytes (profiles/dram_traffic.json)"},
            "kernels": per_kernel,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "steps": e2e_steps, "slab_pairs": slab,
                    "note": "pinned host buffers through the C-ABI (DJB200_MEM_HOST), wall clock max over ranks",
                    "cpus_bound_to_gpu_numa_node": numa},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, djb, capi, lib, torch, dev, wi, wo, out, stream, sptr, peak):
    """Secondary lines of BASELINE.json (configs 3 and 5) on the same GPU: MERL lookups and the LEAN map."""
    res = {}
    n = wi.shape[0]
    # config 3: synthetic 90x90x180 table, n Rusinkiewicz lookups
    rng = np.random.default_rng(0)
    table = rng.uniform(0.0, 3.0, 3 * 90 * 90 * 180)
    m = djb.merl(table)

    def timed(f, reps=3):
        f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            f()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms = timed(lambda: capi.check(lib.djb200_merl_eval(m._h, C.c_void_p(wi.data_ptr()), C.c_void_p(wo.data_ptr()),
                                                       C.c_int64(n), C.c_void_p(out.data_ptr()), C.c_int(capi.MEM_DEVICE), sptr)))
    gbs = 36.0 * n / (ms * 1e-3) / 1e9
    res["merl"] = {"lookups_per_s": n / (ms * 1e-3), "ms": ms, "lookups": n,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "traffic": None}}
    # config 5: 8192^2 normal map -> two planar RGBA float maps
    W = H = args.lean_size
    nm = torch.randint(64, 192, (3, H, W), dtype=torch.uint8, device=dev)
    nm[2] = torch.randint(128, 256, (H, W), dtype=torch.uint8, device=dev)
    l1 = out[: 4 * H * W]
    l2 = out[4 * H * W: 8 * H * W]
    ms = timed(lambda: capi.check(lib.djb200_nmap_to_leanmap(C.c_void_p(nm.data_ptr()), C.c_int32(W), C.c_int32(H),
                                                             C.c_float(1e-5), C.c_float(0.0), C.c_void_p(l1.data_ptr()),
                                                             C.c_void_p(l2.data_ptr()), C.c_int(capi.MEM_DEVICE), sptr)))
    gbs = 35.0 * W * H / (ms * 1e-3) / 1e9
    res["lean"] = {"pixels_per_s": W * H / (ms * 1e-3), "ms": ms, "size": [W, H],
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                "traffic": None}}
    # section 8f rows, measured the same way.  LEAN-filtered Beckmann shading (mitsuba/dj_beckmannconductor.cpp:283-319):
    # fused params construction + evalp against the two-pass route (params blocks written to HBM, then a PER_PAIR query)
    ns = min(n, 50_000_000)
    g = torch.Generator(device=dev).manual_seed(5)
    sl = torch.randn(ns, 2, device=dev, generator=g) * 0.25
    var = torch.rand(ns, 3, device=dev, generator=g)
    vx, vy = 1e-5 + 0.08 * var[:, 0], 1e-5 + 0.08 * var[:, 1]
    E = torch.stack([sl[:, 0] + 25, sl[:, 1] + 25, sl[:, 0] ** 2 + vx, sl[:, 1] ** 2 + vy,
                     sl[:, 0] * sl[:, 1] + (1.4 * var[:, 2] - 0.7) * torch.sqrt(vx * vy) + 625], 1).contiguous()
    al = torch.rand(ns, 3, device=dev, generator=g)
    al[:, :2] = 0.03 + 0.47 * al[:, :2]
    al[:, 2] *= 3.14159
    del sl, var, vx, vy
    cfg = capi.LeanShading()
    cfg.bias, cfg.dmap_scale, cfg.lean_filtering, cfg.alpha_per_pair = 25.0, 1.0, 1, 1
    bk = djb.beckmann()
    d = bk._desc()
    P = out[: 12 * ns]
    res_rgb = out[12 * ns: 15 * ns]
    pv = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    ms_f = timed(lambda: capi.check(lib.djb200_lean_shading_evalp(C.byref(d), C.byref(cfg), pv(al), pv(E), pv(wi), pv(wo),
                                                                  C.c_int64(ns), pv(res_rgb), C.c_int(capi.MEM_DEVICE), sptr)))

    def two_pass():
        capi.check(lib.djb200_lean_shading_params(C.byref(cfg), pv(al), pv(E), C.c_int64(ns), pv(P), C.c_int(capi.MEM_DEVICE), sptr))
        capi.check(lib.djb200_microfacet_evalp(C.byref(d), pv(P), C.c_int64(ns), C.c_int(capi.PARAMS_PER_PAIR), pv(wi), pv(wo),
                                               C.c_int64(ns), pv(res_rgb), C.c_int(capi.MEM_DEVICE), sptr))
    ms_2 = timed(two_pass)
    gbs = 68.0 * ns / (ms_f * 1e-3) / 1e9  # 24 B directions + 20 B moments + 12 B roughness in, 12 B out
    res["lean_shading"] = {"evals_per_s": ns / (ms_f * 1e-3), "ms": ms_f, "pairs": ns, "two_pass_ms": ms_2,
                           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                        "traffic": None}}
    del E, al
    # djb::sgd / djb::abc eval (36 B per pair; double exp / pow / acos per channel: issue bound)
    na = min(n, 20_000_000)
    for kind, name in (("sgd", "gold-metallic-paint"), ("abc", "blue-metallic-paint")):
        mobj = getattr(djb, kind)(name)
        fn = getattr(lib, f"djb200_{kind}_eval")
        ms = timed(lambda: capi.check(fn(C.byref(mobj._data), pv(wi), pv(wo), C.c_int64(na), pv(res_rgb),
                                         C.c_int(capi.MEM_DEVICE), sptr)))
        gbs = 36.0 * na / (ms * 1e-3) / 1e9
        res[kind] = {"evals_per_s": na / (ms * 1e-3), "ms": ms, "pairs": na,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                  "traffic": None}}
    # djb::tabular / djb::tabular_anisotropic as BRDFs (fitted tables of an analytic GGX), djb::utia eval
    nt = min(n, 20_000_000)
    u_t = torch.rand(nt, 2, device=dev, generator=g).contiguous()
    tab = djb.tabular(djb.ggx(), 90)
    th, _keep = tab._first_arg()
    ms = timed(lambda: capi.check(lib.djb200_tabular_eval(th, None, C.c_int64(0), C.c_int(capi.PARAMS_BROADCAST), pv(wi), pv(wo),
                                                          C.c_int64(nt), pv(res_rgb), C.c_int(capi.MEM_DEVICE), sptr)))
    res["tabular_eval"] = {"evals_per_s": nt / (ms * 1e-3), "ms": ms, "pairs": nt,
                           "roofline": {"bound": "hbm", "achieved": 36.0 * nt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": 36.0 * nt / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
    ta = djb.tabular_anisotropic(djb.ggx(), 90, 90)
    tah, _keep2 = ta._first_arg()
    ms = timed(lambda: capi.check(lib.djb200_tabular_sample(tah, None, C.c_int64(0), C.c_int(capi.PARAMS_BROADCAST), pv(u_t), pv(wo),
                                                            C.c_int64(nt), pv(res_rgb), C.c_int(capi.MEM_DEVICE), sptr)))
    res["tabular_anisotropic_sample"] = {"samples_per_s": nt / (ms * 1e-3), "ms": ms, "pairs": nt,
                                         "roofline": {"bound": "hbm", "achieved": 32.0 * nt / (ms * 1e-3) / 1e9, "peak": peak,
                                                      "unit": "GB/s", "frac": 32.0 * nt / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
    ut = djb.utia(np.random.default_rng(3).uniform(0.0, 40.0, 3 * 6 * 48 * 6 * 48))
    ms = timed(lambda: capi.check(lib.djb200_utia_eval(ut._h, pv(wi), pv(wo), C.c_int64(nt), pv(res_rgb), C.c_int(capi.MEM_DEVICE), sptr)))
    res["utia"] = {"evals_per_s": nt / (ms * 1e-3), "ms": ms, "pairs": nt,
                   "roofline": {"bound": "hbm", "achieved": 36.0 * nt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": 36.0 * nt / (ms * 1e-3) / 1e9 / peak, "traffic": None}}
    return res


def smooth_merl_table(seed):
    """Synthetic MERL table (config 4): a radially decreasing lobe over the theta_h index plus a diffuse floor."""
    rng = np.random.default_rng(seed)
    width = float(rng.integers(4, 40))
    k = np.arange(90, dtype=np.float64)
    lobe = 1.0 / (1.0 + (k / width) ** 2) ** 2
    td = 1.0 + (np.arange(90, dtype=np.float64) / 89.0) ** 4
    base = lobe[:, None, None] * td[None, :, None] * np.ones((1, 1, 180))
    return np.concatenate([(tint * base * (1.0 + 0.01 * rng.random(base.shape)) + 30.0 * (c + 1)).reshape(-1)
                           for c, tint in enumerate((900.0, 700.0, 500.0))])


def run_fit_extra(djb, torch, world, rank):
    """BASELINE.json config 4: 128 synthetic MERL tables, 50 power iterations each, sharded by material across the ranks,
    residual diagnostics gathered over NCCL.  Timed through the public API (fit_sharded.tabular_fit_batch_sharded)."""
    import torch.distributed as dist
    from dj_brdf_b200 import fit_sharded as fs
    n_mat, iters = 128, 50
    tables = {}

    def make(k):  # 8 distinct tables uploaded per GPU, reused round-robin (a table is 23 MB on the device)
        if k % 8 not in tables:
            tables[k % 8] = djb.merl(smooth_merl_table(100 + k % 8))
        return tables[k % 8]

    for k in range(rank, n_mat, world):
        make(k)
    fs.tabular_fit_batch_sharded(make, n_mat, 90, True, iters)  # warm-up (workspaces, NCCL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        fits, residuals = fs.tabular_fit_batch_sharded(make, n_mat, 90, True, iters)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank != 0:
        return {}
    return {"fit": {"fits_per_s": n_mat / dt, "ms": dt * 1e3, "materials": n_mat, "iterations": iters, "res": 90,
                    "alpha_ggx_0": float(fits[0].alpha_ggx), "max_final_residual": float(residuals[:, -1].max()),
                    "note": "config 4: isotropic power-iteration fits sharded by material, residuals gathered; wall clock"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=100_000_000, help="(wi, wo) pairs per GPU")
    ap.add_argument("--materials", type=int, default=16)
    ap.add_argument("--cpu-pairs", type=int, default=1_000_000, help="pairs per step of the CPU arm / cpu_baseline")
    ap.add_argument("--e2e-slab", type=int, default=12_500_000, help="pairs per host-buffer call of the e2e leg")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--lean-size", type=int, default=8192)
    ap.add_argument("--no-extras", dest="extras", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"bench.py: note: warmup {args.warmup} < 3 is below the timing rules", file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
