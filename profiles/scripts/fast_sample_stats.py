"""1e-5 tier of sample against the exact tier on the device (2e7 pairs x 16 materials) and against the CPU oracle (2e5 pairs x 6
params sets): distribution of the largest component difference of the sampled direction, (0, 0, 1) pattern, kernel times."""
import json
import sys

import numpy as np
import torch

import dj_brdf_b200 as djb
from oracle import api
from tests import cases

out = {}
port = api.PortOracle()
# ---- against the oracle ----
wi, wo, u = cases.pairs(200_000)
for ndf, name, cls in ((api.NDF_GGX, "ggx", djb.ggx), (api.NDF_BECKMANN, "beckmann", djb.beckmann)):
    b = cls()
    for pname, P in cases.param_sets(port).items():
        want = port.sample(ndf, P, u, wo)
        djb.set_precision("1e-5")
        got = b.sample(u, wo, P)
        djb.set_precision("bits")
        exact = b.sample(u, wo, P)
        err = np.abs(got.astype(np.float64) - want).max(axis=1)
        fb_w = (want == np.array([0, 0, 1], np.float32)).all(axis=1)
        fb_g = (got == np.array([0, 0, 1], np.float32)).all(axis=1)
        q = np.quantile(err, [0.5, 0.99, 0.999, 0.9999])
        out[f"oracle/{name}/{pname}"] = dict(median=q[0], q99=q[1], q999=q[2], q9999=q[3], max=float(err.max()),
                                             frac_le_1e5=float((err <= 1e-5).mean()), frac_le_1e6=float((err <= 1e-6).mean()),
                                             fallback_pattern_equal=bool(np.array_equal(fb_w, fb_g)),
                                             exact_bits=float((exact.view(np.uint32) == want.view(np.uint32)).all(axis=1).mean()),
                                             nan=int(np.isnan(got).sum()))
# ---- at scale against the exact tier ----
n = 20_000_000
g = torch.Generator(device="cuda").manual_seed(7)
z = 1.0 - 0.999 * torch.rand(n, device="cuda", generator=g)
ph = 6.283185307179586 * torch.rand(n, device="cuda", generator=g)
r = torch.sqrt(torch.clamp(1 - z * z, min=0))
wo_d = torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], 1).contiguous()
u_d = torch.rand(n, 2, device="cuda", generator=g)
mats = cases.c2_materials(port)
mats[15] = djb.params.pdfparams(0.3, 0.2, 0.4, 0.1, -0.2)
for name, cls in (("ggx", djb.ggx), ("beckmann", djb.beckmann)):
    b = cls()
    times = {}
    res = {}
    for mode in ("1e-5", "bits"):
        djb.set_precision(mode)
        b.sample(u_d, wo_d, mats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res[mode] = b.sample(u_d, wo_d, mats)
        e1.record()
        torch.cuda.synchronize()
        times[mode] = e0.elapsed_time(e1)
    d = (res["1e-5"].double() - res["bits"].double()).abs().amax(dim=-1).reshape(-1)
    nan = int(torch.isnan(res["1e-5"]).sum().item()), int(torch.isnan(res["bits"]).sum().item())
    d = torch.nan_to_num(d, nan=0.0)
    srt = torch.sort(d).values
    N = srt.numel()
    out[f"scale/{name}"] = dict(ms_fast=times["1e-5"], ms_exact=times["bits"], n=N, nan_fast_exact=nan,
                                median=srt[N // 2].item(), q99=srt[int(N * 0.99)].item(), q999=srt[int(N * 0.999)].item(),
                                q9999=srt[int(N * 0.9999)].item(), q99999=srt[int(N * 0.99999)].item(), max=srt[-1].item(),
                                frac_le_1e5=(d <= 1e-5).double().mean().item(), frac_le_1e4=(d <= 1e-4).double().mean().item(),
                                mean=d.mean().item())
    per_mat = d.reshape(16, -1)
    out[f"scale/{name}"]["frac_le_1e5_per_material"] = [(per_mat[m] <= 1e-5).double().mean().item() for m in range(16)]
    del res, d, srt
djb.set_precision("bits")
json.dump(out, sys.stdout, indent=1)
