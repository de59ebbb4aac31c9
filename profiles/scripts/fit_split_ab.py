"""Isotropic fit, single launch (one CTA per material) against the split mode (one launch per phase, several CTAs per material):
call time for 1 / 4 / 16 / 32 / 128 materials x 50 iterations."""
import ctypes as C, time, json
import numpy as np, torch
import dj_brdf_b200 as djb
from dj_brdf_b200 import capi, workloads
lib = capi.load()
specs = workloads.fit_table_specs(128)
tabs = [djb.merl(workloads.fit_table(s)) for s in specs[:32]]
for n_mat in (1, 4, 16, 32, 128):
    srcs = djb.tabular.source_array([tabs[k % len(tabs)] for k in range(n_mat)])
    row = {"materials": n_mat}
    for parts in (1, 3, 4, 6, 8, 0):
        capi.check(lib.djb200_debug_fit_parts(C.c_int(parts)))
        djb.tabular.fit_packed(srcs, 90, True, 50)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            djb.tabular.fit_packed(srcs, 90, True, 50)
        row[f"parts{parts}_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    print(json.dumps(row), flush=True)
capi.check(lib.djb200_debug_fit_parts(C.c_int(0)))
