"""Shared definitions of the parity cases (inputs are regenerated from seeds, never stored)."""
import hashlib

import numpy as np

from oracle import api

N_SMALL = 4096  # golden fixtures
N_PARITY = 200_000  # GPU-vs-oracle parity sweeps (the oracle finishes each in well under a second)


def pairs(n, stream=0):
    wi = api.directions(n, stream)
    wo = api.directions(n, stream + 2)
    u = np.stack([api.uniforms(n, stream + 4), api.uniforms(n, stream + 5)], axis=1)
    return wi, wo, u


def edge_pairs():
    """Hand-picked edge inputs the reference treats specially (SURVEY.md section 7 'edge cases')."""
    v = []
    z = np.float32
    def d(x, y, zz):
        a = np.array([x, y, zz], np.float64)
        return (a / np.linalg.norm(a)).astype(np.float32)
    v.append((d(0, 0, 1), d(0, 0, 1)))                   # normal incidence, h = z (pole guard)
    v.append((d(1e-4, 0, 1), d(-1e-4, 0, 1)))            # h within the 0.99999 pole guard
    v.append((d(1, 0, 1e-3), d(1, 0, 1e-3)))             # grazing retro-reflection
    v.append((d(1, 0, 1e-3), d(-1, 0, 1e-3)))            # grazing forward: h ~ z
    v.append((d(0.3, 0.4, -0.2), d(0.1, 0.2, 0.9)))      # i below the horizon (-> G1 gate)
    v.append((d(0.3, 0.4, 0.5), d(0.1, 0.2, -0.9)))      # o below the horizon
    v.append((d(0.6, 0, 0.8), d(-0.6, 0, 0.8)))          # mirror pair in the xz plane
    v.append((d(0, 0.6, 0.8), d(0, -0.6, 0.8)))          # mirror pair in the yz plane
    v.append((np.array([0.6, 0.0, 0.8], z), np.array([0.0, 0.0, 1.0], z)))
    v.append((np.array([0.0, 1.0, 0.0], z), np.array([0.0, 0.0, 1.0], z)))  # i.z == 0 -> inf/NaN path
    v.append((d(-0.5, -0.5, 0.7), d(-0.5, -0.5, 0.7)))
    v.append((d(1, 1, 1e-6), d(0, 0, 1)))
    wi = np.stack([a for a, _ in v]).astype(np.float32)
    wo = np.stack([b for _, b in v]).astype(np.float32)
    u = np.array([[0, 0], [1, 1], [0.5, 0.5], [1e-7, 0.999999], [0.25, 0.75], [0.3, 0.7], [0.9, 0.1],
                  [0.5, 0.0], [0.0, 0.5], [1.0, 0.5], [0.123, 0.456], [0.999, 0.001]], np.float32)
    return wi, wo, u


def param_sets(o):
    """Named params blocks built with oracle `o` (RefOracle or PortOracle)."""
    return {
        "iso0.1": o.params_elliptic(0.1, 0.1, 0.0),
        "iso0.5": o.params_elliptic(0.5, 0.5, 0.0),
        "aniso": o.params_elliptic(0.1, 0.4, 0.7),
        "aniso2": o.params_elliptic(0.6, 0.05, 2.5),
        "offcentre": o.params_pdfparams(0.3, 0.2, 0.4, 0.1, -0.2),
        "standard": o.params_elliptic(1.0, 1.0, 0.0),
    }


def c2_materials(o, m=16, seed=1):
    """config 2 of BASELINE.json: 16 anisotropic materials, alpha log-uniform [0.02, 0.8], phi_a in [0, pi)."""
    rng = np.random.default_rng(seed)
    a1 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    a2 = np.exp(rng.uniform(np.log(0.02), np.log(0.8), m)).astype(np.float32)
    ph = rng.uniform(0, np.pi, m).astype(np.float32)
    return np.stack([o.params_elliptic(float(a), float(b), float(c)) for a, b, c in zip(a1, a2, ph)])


def fresnels():
    return {
        "ideal": api.Fresnel.ideal(),
        "schlick": api.Fresnel.schlick([0.9, 0.5, 0.2]),
        "unpolarized": api.Fresnel.unpolarized([1.5, 1.8, 2.4]),
        "spline": api.Fresnel.spline(np.linspace(0.2, 1.0, 30, dtype=np.float32).reshape(10, 3)),
    }


def synthetic_merl_table(alpha=0.15, seed=7, kind="ggx"):
    """config 3: analytic microfacet lobe + diffuse sampled at MERL cell centres, stored unscaled (divided
    by the MERL channel scales); below-horizon cells are -1 (exercises dj_brdf.h:1016-1021).  The generator is the
    benchmark's (dj_brdf_b200/workloads.py): tests and bench.py use the same tables."""
    from dj_brdf_b200 import workloads
    return workloads.synthetic_merl_table(alpha, kind)


def synthetic_nmap(h, w, seed=12345):
    from dj_brdf_b200 import workloads
    return workloads.synthetic_nmap(h, w, seed)


# ---- tables made only of platform-independent arithmetic (PCG64 doubles, + - * /), for the golden files ----
def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def random_merl_table(seed):
    return np.random.default_rng(seed).uniform(-0.05, 3.0, 3 * api.MERL_CELLS)


def random_utia_table(seed):
    return np.random.default_rng(seed).uniform(-0.5, 60.0, 3 * 6 * 48 * 6 * 48)


def smooth_merl_table(seed):
    """A fit-able table made only of exactly reproducible arithmetic: a radially decreasing lobe in the
    theta_h index plus a diffuse floor, jittered by PCG64 (no libm calls)."""
    rng = np.random.default_rng(seed)
    width = float(rng.integers(4, 40))
    k = np.arange(90, dtype=np.float64)
    lobe = 1.0 / (1.0 + (k / width) ** 2) ** 2            # over theta_h index
    td = 1.0 + (np.arange(90, dtype=np.float64) / 89.0) ** 4  # mild growth towards grazing theta_d
    base = lobe[:, None, None] * td[None, :, None] * np.ones((1, 1, 180))
    planes = []
    for c, tint in enumerate((900.0, 700.0, 500.0)):
        jitter = 1.0 + 0.01 * rng.random(base.shape)
        planes.append((tint * base * jitter + 30.0 * (c + 1)).reshape(-1))
    return np.concatenate(planes)


def lean_texels(n, seed=7, bias=25.0):
    """n texels of a (biased) LEAN map pair as a renderer would fetch them: E1..E5 = mean slopes + bias, second moments
    of the unbiased slopes (E3, E4) and the biased cross moment (E5), utils/nmap2leanmap_biased.cpp:43-58 -- plus a
    per-texel base roughness (alpha1, alpha2, alphaAngle)."""
    rng = np.random.default_rng(seed)
    sx, sy = rng.normal(0, 0.25, n), rng.normal(0, 0.25, n)
    vx, vy = rng.uniform(1e-5, 0.08, n), rng.uniform(1e-5, 0.08, n)
    cxy = rng.uniform(-0.7, 0.7, n) * np.sqrt(vx * vy)
    E = np.stack([sx + bias, sy + bias, sx * sx + vx, sy * sy + vy, sx * sy + cxy + bias * bias], 1).astype(np.float32)
    alpha = np.stack([rng.uniform(0.03, 0.5, n), rng.uniform(0.03, 0.5, n), rng.uniform(0, np.pi, n)], 1).astype(np.float32)
    return E, alpha
