"""Synthetic inputs of the benchmark configurations, shared by bench.py and the tests.

Host-side numpy only: these are the *workloads* (what a caller would hand to the engine), not part of the engine.

* config 2: `materials`      -- 16 anisotropic parameter sets, alpha log-uniform in [0.02, 0.8], phi_a uniform in [0, pi)
* config 3: `MerlSynth.table` -- a 90 x 90 x 180 x 3 MERL table sampled from an analytic microfacet lobe + diffuse term at the
                                 cell centres, stored unscaled (divided by the MERL channel scales, dj_brdf.h:897-899), values
                                 exactly representable in fp32, cells whose directions fall below the horizon set to -1 (the
                                 reference's "negative => 0" branch, dj_brdf.h:1016-1021)
* config 4: `fit_tables`      -- 128 such tables, 64 GGX + 64 Beckmann, alpha in [0.05, 0.6], Schlick f0 colours: the
                                 ground-truth roughness of every table is known
* config 5: `synthetic_nmap`  -- planar uint8 normal map, r, g in [64, 191], b in [128, 255]
"""
from __future__ import annotations

import numpy as np

MERL_SCALES = (1.00 / 1500.0, 1.15 / 1500.0, 1.66 / 1500.0)


def materials(m=16, seed=1):
    rng = np.random.default_rng(seed)
    a1 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    a2 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    ph = rng.uniform(0, np.pi, m).astype(np.float32)
    return a1, a2, ph


class MerlSynth:
    """Generator of analytic MERL tables.  The geometry of the 1,458,000 cell centres (Rusinkiewicz half / difference angles
    -> i, o) does not depend on the material, so it is computed once; a table is then a few array products."""

    def __init__(self):
        th = ((np.arange(90) + 0.5) / 90.0) ** 2 * (np.pi / 2)
        td = (np.arange(90) + 0.5) / 90.0 * (np.pi / 2)
        pd = (np.arange(180) + 0.5) / 180.0 * np.pi
        TH, TD, PD = np.meshgrid(th, td, pd, indexing="ij")
        # half / diff -> i, o (h in the xz-plane at elevation TH; d rotated by TH about y)
        dx, dz = np.sin(TD) * np.cos(PD), np.cos(TD)
        ix = dx * np.cos(TH) + dz * np.sin(TH)
        iz = -dx * np.sin(TH) + dz * np.cos(TH)
        hx, hz = np.sin(TH), np.cos(TH)
        dot = ix * hx + iz * hz
        oz = 2 * dot * hz - iz
        self.below = ((iz <= 0) | (oz <= 0)).reshape(-1)
        self.schlick5 = ((1 - np.clip(dot, 0, 1)) ** 5).reshape(-1)
        self.inv_geo = (1.0 / np.maximum(4 * np.abs(iz * oz), 1e-3)).reshape(-1)
        self.tan2_h = np.tan(th) ** 2  # per theta_h index
        self.cos4_h = np.cos(th) ** 4

    def ndf(self, alpha, kind):
        if kind == "ggx":
            return alpha ** 2 / (np.pi * self.cos4_h * (alpha ** 2 + self.tan2_h) ** 2)
        return np.exp(-self.tan2_h / alpha ** 2) / (np.pi * alpha ** 2 * self.cos4_h)

    def table(self, alpha=0.15, kind="ggx", f0=(0.04, 0.04, 0.04), tint=(0.8, 0.6, 0.4), diffuse=(0.1, 0.2, 0.3)):
        D = np.repeat(self.ndf(alpha, kind), 90 * 180)
        planes = []
        for c in range(3):
            F = f0[c] + (1.0 - f0[c]) * self.schlick5
            v = (tint[c] * D * F * self.inv_geo + diffuse[c] / np.pi) / MERL_SCALES[c]
            v = v.astype(np.float32).astype(np.float64)  # exactly representable in fp32
            v[self.below] = -1.0
            planes.append(v)
        return np.concatenate(planes)


_synth = None


def merl_synth():
    global _synth
    if _synth is None:
        _synth = MerlSynth()
    return _synth


def synthetic_merl_table(alpha=0.15, kind="ggx"):
    """config 3 (also the table of tests/cases.py): GGX(alpha) or Beckmann(alpha) lobe + diffuse."""
    return merl_synth().table(alpha, kind)


def fit_table_specs(n=128, seed=4):
    """config 4: (kind, alpha, f0 rgb) of table k; even k GGX, odd k Beckmann, alpha uniform in [0.05, 0.6]."""
    rng = np.random.default_rng(seed)
    alpha = rng.uniform(0.05, 0.6, n)
    f0 = rng.uniform(0.02, 0.95, (n, 3))
    return [("ggx" if k % 2 == 0 else "beckmann", float(alpha[k]), tuple(float(x) for x in f0[k])) for k in range(n)]


def fit_table(spec):
    kind, alpha, f0 = spec
    # a purely specular lobe: the diffuse floor of config 3 would put a roughness-independent plateau under the fitted NDF
    return merl_synth().table(alpha, kind, f0=f0, tint=(1.0, 1.0, 1.0), diffuse=(0.0, 0.0, 0.0))


def synthetic_nmap(h, w, seed=12345):
    rng = np.random.default_rng(seed)
    r = rng.integers(64, 192, (h, w), dtype=np.uint8)
    g = rng.integers(64, 192, (h, w), dtype=np.uint8)
    b = rng.integers(128, 256, (h, w), dtype=np.uint8)
    return np.stack([r, g, b])
