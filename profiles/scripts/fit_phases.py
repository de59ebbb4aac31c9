"""Where the isotropic fit kernel spends its time: SM clocks at the phase boundaries of material 0 (128 fits x 50 iterations)."""
import ctypes as C, json, sys
import numpy as np, torch
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi, workloads
lib = capi.load()
specs = workloads.fit_table_specs(128)
srcs = [djb.merl(workloads.fit_table(s)) for s in specs]
names = ["rows", "matrix", "iterations", "normalise", "ndf grid", "sigma", "fresnel ratios", "fresnel sums + cdf", "qf + params"]
for n_mat in (128, 16, 1):
    for it in (50, 4):
        djb.tabular.fit_packed(srcs[:n_mat], 90, True, it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); djb.tabular.fit_packed(srcs[:n_mat], 90, True, it); e1.record(); torch.cuda.synchronize()
        c = (C.c_int64 * 10)()
        capi.check(lib.djb200_debug_fit_phase_clocks(c))
        c = np.array(list(c), np.int64)
        d = np.diff(c)
        print(json.dumps({"materials": n_mat, "iterations": it, "call_ms": e0.elapsed_time(e1), "kernel_clocks": int(c[-1] - c[0]),
                          "phases_kclk": {k: round(float(v) / 1e3, 1) for k, v in zip(names, d)}}), flush=True)
