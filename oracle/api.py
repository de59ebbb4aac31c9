"""oracle/api.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the two CPU checkers:

* ``PortOracle``  -- oracle/libdjb_oracle.so, the C restatement (oracle/djb_oracle*.c).
* ``RefOracle``   -- oracle/_ref/libdjbref.so (+ libleanref*.so), the UNMODIFIED reference compiled
  in place from /root/reference by oracle/Makefile.  Exists only where it was built.

Both expose the same numpy-level methods so a parity test can be written once and pointed at
either.  Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("DJB_REFERENCE_ROOT", "/root/reference"))

NDF_BECKMANN, NDF_GGX = 0, 1
F_IDEAL, F_SCHLICK, F_UNPOLARIZED, F_SGD, F_SPLINE = 0, 1, 2, 3, 4
MERL_CELLS = 90 * 90 * 180

c_f32p = C.c_void_p
i64 = C.c_int64


def _ptr(a):
    return None if a is None else a.ctypes.data


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def build_port(quiet=True):
    subprocess.run(["make", "-C", str(HERE), "port"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def build_ref(quiet=True):
    """Compile the reference in place; only possible where /root/reference exists."""
    if not (REF_ROOT / "dj_brdf.h").exists():
        return False
    subprocess.run(["make", "-C", str(HERE), "ref", f"REF={REF_ROOT}"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.DEVNULL if quiet else None)
    return True


def build_plugins(quiet=True):
    """The six Mitsuba plugin sources of the reference, compiled in place against the mock Mitsuba API, once per backend
    (oracle/Makefile `plugins`); needs /root/reference and a built dj_brdf_b200/libdjb200.so."""
    if not (REF_ROOT / "mitsuba" / "dj_merl.cpp").exists():
        return False
    subprocess.run(["make", "-C", str(HERE), "plugins", f"REF={REF_ROOT}"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.DEVNULL if quiet else None)
    return True


def port_available():
    return (HERE / "libdjb_oracle.so").exists()


def ref_available():
    return (HERE / "_ref" / "libdjbref.so").exists()


class Fresnel:
    """kind + data, mirrors djb::fresnel::{ideal,schlick,unpolarized,sgd,spline}."""

    def __init__(self, kind=F_IDEAL, data=()):
        self.kind = kind
        self.data = _f32(np.asarray(data, dtype=np.float32).reshape(-1))

    @staticmethod
    def ideal():
        return Fresnel(F_IDEAL)

    @staticmethod
    def schlick(f0):
        return Fresnel(F_SCHLICK, f0)

    @staticmethod
    def unpolarized(ior):
        return Fresnel(F_UNPOLARIZED, ior)

    @staticmethod
    def spline(points):
        return Fresnel(F_SPLINE, points)


class _OrcFresnel(C.Structure):
    _fields_ = [("kind", C.c_int), ("v", C.c_float * 6), ("pts", C.c_void_p), ("npts", C.c_int)]


class _OrcSource(C.Structure):
    _fields_ = [("kind", C.c_int), ("ndf", C.c_int), ("F", _OrcFresnel), ("shadow", C.c_int),
                ("P", C.c_float * 12), ("table", C.c_void_p), ("sgd", C.c_void_p), ("abc", C.c_void_p)]


def sgd_field_major(coef_channel_major):
    """[3, 11] channel-major coefficients (the product's djb200_sgd_data) -> the 33 doubles of orc_sgd (field-major,
    the order of the reference's sgd::data, dj_brdf.h:482-497)."""
    return np.ascontiguousarray(np.asarray(coef_channel_major, np.float64).reshape(3, 11).T).reshape(-1)


class Source:
    """Fit input: an analytic microfacet BRDF, a MERL table (3*1458000 float64) or a raw UTIA table."""

    def __init__(self, kind, ndf=NDF_GGX, fresnel=None, shadow=True, params=None, table=None):
        self.kind, self.ndf, self.fresnel, self.shadow = kind, ndf, fresnel or Fresnel.ideal(), shadow
        self.params, self.table = params, table

    @staticmethod
    def microfacet(ndf, fresnel=None, shadow=True):
        return Source("microfacet", ndf=ndf, fresnel=fresnel, shadow=shadow)

    @staticmethod
    def merl(table):
        return Source("merl", table=np.ascontiguousarray(table, dtype=np.float64).reshape(-1))

    @staticmethod
    def utia(raw_table):
        return Source("utia", table=np.ascontiguousarray(raw_table, dtype=np.float64).reshape(-1))

    @staticmethod
    def sgd(name, coef_channel_major):
        """djb::sgd(name): the reference looks the name up itself, the port gets the 33 coefficients."""
        s = Source("sgd", table=sgd_field_major(coef_channel_major))
        s.name = name
        return s

    @staticmethod
    def abc(name, coef9):
        s = Source("abc", table=np.ascontiguousarray(coef9, dtype=np.float64).reshape(9))
        s.name = name
        return s


def write_merl_file(path, table):
    t = np.ascontiguousarray(table, dtype=np.float64).reshape(-1)
    assert t.size == 3 * MERL_CELLS
    with open(path, "wb") as f:
        f.write(np.array([90, 90, 180], dtype=np.int32).tobytes())
        f.write(t.tobytes())


# --------------------------------------------------------------------------------------------------
class PortOracle:
    name = "port"

    def __init__(self):
        if not port_available():
            build_port()
        self.lib = C.CDLL(str(HERE / "libdjb_oracle.so"))

    # -- helpers
    def _fres(self, fr):
        fr = fr or Fresnel.ideal()
        s = _OrcFresnel()
        s.kind = fr.kind
        if fr.kind == F_SPLINE:
            s.pts = fr.data.ctypes.data
            s.npts = fr.data.size // 3
        else:
            for k in range(min(6, fr.data.size)):
                s.v[k] = fr.data[k]
        return s

    def params_elliptic(self, a1, a2, phi_a=0.0):
        out = np.zeros(12, np.float32)
        self.lib.orc_params_elliptic(C.c_float(a1), C.c_float(a2), C.c_float(phi_a), c_f32p(out.ctypes.data))
        return out

    def params_pdfparams(self, ax, ay, rho=0.0, tx=0.0, ty=0.0):
        out = np.zeros(12, np.float32)
        self.lib.orc_params_pdfparams(C.c_float(ax), C.c_float(ay), C.c_float(rho), C.c_float(tx),
                                      C.c_float(ty), c_f32p(out.ctypes.data))
        return out

    def _mf(self, fn, ndf, fresnel, shadow, params, a, b, outs, nthreads):
        a, b = _f32(a), _f32(b)
        n = b.shape[0]
        fs = self._fres(fresnel)
        p = None if params is None else _f32(params)
        getattr(self.lib, fn)(C.c_int(ndf), C.byref(fs), C.c_int(int(shadow)), c_f32p(_ptr(p)),
                              c_f32p(a.ctypes.data), c_f32p(b.ctypes.data), i64(n),
                              *[c_f32p(o.ctypes.data) for o in outs], C.c_int(nthreads))

    def eval(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._mf("orc_microfacet_eval", ndf, fresnel, shadow, params, wi, wo, [out], nthreads)
        return out

    def evalp(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._mf("orc_microfacet_evalp", ndf, fresnel, shadow, params, wi, wo, [out], nthreads)
        return out

    def pdf(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty(len(wo), np.float32)
        self._mf("orc_microfacet_pdf", ndf, fresnel, shadow, params, wi, wo, [out], nthreads)
        return out

    def sample(self, ndf, params, u, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._mf("orc_microfacet_sample", ndf, fresnel, shadow, params, u, wo, [out], nthreads)
        return out

    def evalp_is(self, ndf, params, u, wo, fresnel=None, shadow=True, nthreads=1):
        w = np.empty((len(wo), 3), np.float32)
        i = np.empty((len(wo), 3), np.float32)
        pdf = np.empty(len(wo), np.float32)
        self._mf("orc_microfacet_evalp_is", ndf, fresnel, shadow, params, u, wo, [w, i, pdf], nthreads)
        return w, i, pdf

    COMPONENTS = {"ndf": 0, "gaf": 1, "g1": 2, "sigma": 3, "p22": 4, "vp22": 5, "vndf": 6, "fresnel": 7}

    def component(self, what, ndf, params, a, b=None, c=None, fresnel=None, shadow=True):
        """djb::microfacet's public component queries (dj_brdf.h:258-272) under one params block."""
        code = self.COMPONENTS[what]
        a = _f32(a)
        b = None if b is None else _f32(b)
        c = None if c is None else _f32(c)
        n = len(a)
        out = np.zeros((n, 3) if code == 7 else n, np.float32)
        fr = fresnel or Fresnel.ideal()
        fs = self._fres(fr)
        p = None if params is None else _f32(params)
        self.lib.orc_microfacet_component(C.c_int(ndf), C.byref(fs), C.c_int(int(shadow)), c_f32p(_ptr(p)), C.c_int(code),
                                          c_f32p(a.ctypes.data), c_f32p(_ptr(b)), c_f32p(_ptr(c)), i64(n), c_f32p(out.ctypes.data))
        return out

    def io_to_hd(self, wi, wo):
        wi, wo = _f32(wi), _f32(wo)
        h, d = np.empty_like(wi), np.empty_like(wi)
        self.lib.orc_io_to_hd(c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                              c_f32p(h.ctypes.data), c_f32p(d.ctypes.data))
        return h, d

    def hd_to_io(self, h, d):
        h, d = _f32(h), _f32(d)
        wi, wo = np.empty_like(h), np.empty_like(h)
        self.lib.orc_hd_to_io(c_f32p(h.ctypes.data), c_f32p(d.ctypes.data), i64(len(h)),
                              c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data))
        return wi, wo

    def merl_index(self, wi, wo, nthreads=1):
        wi, wo = _f32(wi), _f32(wo)
        idx = np.empty(len(wi), np.int32)
        self.lib.orc_merl_index(c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                                C.c_void_p(idx.ctypes.data), C.c_int(nthreads))
        return idx

    def merl_eval(self, table, wi, wo, nthreads=1):
        table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1)
        assert table.size == 3 * MERL_CELLS
        wi, wo = _f32(wi), _f32(wo)
        out = np.empty((len(wi), 3), np.float32)
        self.lib.orc_merl_eval(C.c_void_p(table.ctypes.data), c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data),
                               i64(len(wi)), c_f32p(out.ctypes.data), C.c_int(nthreads))
        return out

    def utia_eval(self, raw_table, wi, wo, nthreads=1):
        t = np.array(raw_table, dtype=np.float64).reshape(-1)
        self.lib.orc_utia_normalize(C.c_void_p(t.ctypes.data))
        wi, wo = _f32(wi), _f32(wo)
        out = np.empty((len(wi), 3), np.float32)
        self.lib.orc_utia_eval(C.c_void_p(t.ctypes.data), c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data),
                               i64(len(wi)), c_f32p(out.ctypes.data), C.c_int(nthreads))
        return out

    def lrep_to_params(self, E5):
        E5 = _f32(E5).reshape(-1, 5)
        out = np.empty((len(E5), 12), np.float32)
        self.lib.orc_lrep_to_params(c_f32p(E5.ctypes.data), i64(len(E5)), c_f32p(out.ctypes.data))
        return out

    def params_to_lrep(self, params):
        params = _f32(params).reshape(-1, 12)
        out = np.empty((len(params), 5), np.float32)
        self.lib.orc_params_to_lrep(c_f32p(params.ctypes.data), i64(len(params)), c_f32p(out.ctypes.data))
        return out

    def lean_shading_params(self, E5, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        """mitsuba/dj_beckmannconductor.cpp:283-314 per pair; alpha: 3 floats or [n, 3]."""
        E5, alpha = _f32(E5).reshape(-1, 5), _f32(alpha)
        out = np.zeros((len(E5), 12), np.float32)
        self.lib.orc_lean_shading_params(C.c_float(bias), C.c_float(dmap_scale), C.c_int(int(lean_filtering)),
                                         C.c_int(int(alpha.ndim == 2)), c_f32p(alpha.ctypes.data), c_f32p(E5.ctypes.data),
                                         i64(len(E5)), c_f32p(out.ctypes.data))
        return out

    def nmap2leanmap(self, nmap_planar, base_roughness=1e-5, bias=0.0):
        nmap = np.ascontiguousarray(nmap_planar, dtype=np.uint8)
        _, h, w = nmap.shape
        l1 = np.empty((4, h, w), np.float32)
        l2 = np.empty((4, h, w), np.float32)
        self.lib.orc_nmap2leanmap(C.c_void_p(nmap.ctypes.data), C.c_int(w), C.c_int(h),
                                  C.c_float(base_roughness), C.c_float(bias),
                                  c_f32p(l1.ctypes.data), c_f32p(l2.ctypes.data))
        return l1, l2

    def leanmap_half_mips(self, leanmap, levels=0):
        """The LEAN map as a half-float RGBA mip pyramid: level 0 = save_exr's (half)(float) of the four planes
        (utils/CImg.h:44940-44947), level L = 2 x 2 box filter of level L - 1 in float32, ((a + b) + (c + d)) * 0.25 with edge
        texels repeated, rounded to half (nearest even) once per level.  numpy restatement (IEEE float32 adds, no FMA)."""
        lm = np.asarray(leanmap, np.float32)
        cur = np.ascontiguousarray(np.moveaxis(lm, 0, -1))
        full = 1
        hh, ww = cur.shape[:2]
        while hh > 1 or ww > 1:
            hh, ww, full = max(1, hh // 2), max(1, ww // 2), full + 1
        n = full if levels <= 0 or levels > full else levels
        out = [cur.astype(np.float16)]
        while len(out) < n:
            sh, sw = cur.shape[:2]
            dh, dw = max(1, sh // 2), max(1, sw // 2)
            y0, y1 = np.minimum(2 * np.arange(dh), sh - 1), np.minimum(2 * np.arange(dh) + 1, sh - 1)
            x0, x1 = np.minimum(2 * np.arange(dw), sw - 1), np.minimum(2 * np.arange(dw) + 1, sw - 1)
            a, b, c, d = cur[y0][:, x0], cur[y0][:, x1], cur[y1][:, x0], cur[y1][:, x1]
            cur = ((a + b) + (c + d)) * np.float32(0.25)
            out.append(cur.astype(np.float16))
        return out

    def erf(self, x):
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        self.lib.orc_erf(c_f32p(x.ctypes.data), i64(len(x)), c_f32p(out.ctypes.data))
        return out

    def radial_query(self, what, x, ndf=None, fit=None):
        """djb::radial's p22_radial / sigma_std_radial / cdf_radial / qf_radial for an analytic family (ndf) or a
        fit_tabular() result (fit)."""
        code = {"p22": 0, "sigma_std": 1, "cdf": 2, "qf": 3}[what]
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        if fit is not None:
            t = [_f32(fit[k]) for k in ("p22", "sigma", "qf", "cdf")]
            self.lib.orc_radial_query(C.c_int(2), C.c_int(code), *[c_f32p(a.ctypes.data) for a in t], C.c_int(len(t[0])),
                                      c_f32p(x.ctypes.data), i64(len(x)), c_f32p(out.ctypes.data))
        else:
            self.lib.orc_radial_query(C.c_int(ndf), C.c_int(code), None, None, None, None, C.c_int(0),
                                      c_f32p(x.ctypes.data), i64(len(x)), c_f32p(out.ctypes.data))
        return out

    def dmap2nmap(self, dmap, scale=0.1):
        d = np.ascontiguousarray(dmap, np.uint8)
        h, w = d.shape
        out = np.zeros((3, h, w), np.uint8)
        self.lib.orc_dmap2nmap(C.c_void_p(d.ctypes.data), C.c_int(w), C.c_int(h), C.c_float(scale), C.c_void_p(out.ctypes.data))
        return out

    # -- fits
    def _source(self, src):
        s = _OrcSource()
        keep = []
        if src.kind == "microfacet":
            s.kind, s.ndf, s.shadow = 0, src.ndf, int(src.shadow)
            s.F = self._fres(src.fresnel)
            keep.append(src.fresnel)
            p = self.params_elliptic(1.0, 1.0, 0.0) if src.params is None else _f32(src.params)
            for k in range(12):
                s.P[k] = p[k]
        elif src.kind == "merl":
            s.kind, s.table = 1, src.table.ctypes.data
        elif src.kind == "utia":
            t = np.array(src.table, dtype=np.float64)
            self.lib.orc_utia_normalize(C.c_void_p(t.ctypes.data))
            keep.append(t)
            s.kind, s.table = 2, t.ctypes.data
        elif src.kind == "sgd":
            s.kind, s.sgd = 3, src.table.ctypes.data
        elif src.kind == "abc":
            s.kind, s.abc = 4, src.table.ctypes.data
        else:
            raise ValueError(src.kind)
        return s, keep

    def sgd_eval(self, coef_channel_major, wi, wo, nthreads=1):
        m = sgd_field_major(coef_channel_major)
        wi, wo = _f32(wi), _f32(wo)
        out = np.zeros((len(wi), 3), np.float32)
        self.lib.orc_sgd_eval(C.c_void_p(m.ctypes.data), c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                              c_f32p(out.ctypes.data), C.c_int(nthreads))
        return out

    def abc_eval(self, coef9, wi, wo, nthreads=1):
        m = np.ascontiguousarray(coef9, dtype=np.float64).reshape(9)
        wi, wo = _f32(wi), _f32(wo)
        out = np.zeros((len(wi), 3), np.float32)
        self.lib.orc_abc_eval(C.c_void_p(m.ctypes.data), c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                              c_f32p(out.ctypes.data), C.c_int(nthreads))
        return out

    def fit_tabular(self, src, res=90, shadow=True, iterations=4):
        s, keep = self._source(src)
        p22, sigma, cdf, qf = (np.zeros(res, np.float32) for _ in range(4))
        fres = np.zeros((res, 3), np.float32)
        alpha = np.zeros(2, np.float32)
        self.lib.orc_fit_tabular(C.byref(s), C.c_int(res), C.c_int(int(shadow)), C.c_int(iterations),
                                 *[c_f32p(a.ctypes.data) for a in (p22, sigma, cdf, qf, fres, alpha)])
        return dict(p22=p22, sigma=sigma, cdf=cdf, qf=qf, fresnel=fres, alpha=alpha)

    def tabular_query(self, op, fit, u_or_wi, wo, params=None, shadow=True, fresnel=None, nthreads=1):
        """djb::tabular as a BRDF on the tables of `fit` (a fit_tabular() result): op in eval/evalp/pdf/sample/evalp_is."""
        code = {"eval": 0, "evalp": 1, "pdf": 2, "sample": 3, "evalp_is": 4}[op]
        a, b = _f32(u_or_wi), _f32(wo)
        n = len(b)
        fr = fresnel or Fresnel.spline(fit["fresnel"])
        fs = self._fres(fr)
        p = None if params is None else _f32(params)
        o0 = np.zeros(n if code == 2 else (n, 3), np.float32)
        o1 = np.zeros((n, 3), np.float32)
        o2 = np.zeros(n, np.float32)
        res = len(fit["p22"])
        self.lib.orc_tabular_query(C.c_int(code), c_f32p(_f32(fit["p22"]).ctypes.data), c_f32p(_f32(fit["sigma"]).ctypes.data),
                                   c_f32p(_f32(fit["qf"]).ctypes.data), C.c_int(res), C.byref(fs), C.c_int(int(shadow)),
                                   c_f32p(_ptr(p)), c_f32p(a.ctypes.data), c_f32p(b.ctypes.data), i64(n),
                                   c_f32p(o0.ctypes.data), c_f32p(o1.ctypes.data), c_f32p(o2.ctypes.data), C.c_int(nthreads))
        return (o0, o1, o2) if code == 4 else o0

    def tabular_aniso_query(self, op, fit, er, ar, wi, wo, params=None, shadow=True, nthreads=1):
        """djb::tabular_anisotropic as a BRDF on the tables of a fit_tabular_anisotropic() result: eval / evalp / pdf."""
        code = {"eval": 0, "evalp": 1, "pdf": 2}[op]
        a, b = _f32(wi), _f32(wo)
        n = len(b)
        fs = self._fres(Fresnel.spline(fit["fresnel"]))
        p = None if params is None else _f32(params)
        o0 = np.zeros(n if code == 2 else (n, 3), np.float32)
        self.lib.orc_tabular_aniso_query(C.c_int(code), c_f32p(_f32(fit["p22"]).ctypes.data), c_f32p(_f32(fit["sigma"]).ctypes.data),
                                         C.c_int(er), C.c_int(ar), C.byref(fs), C.c_int(int(shadow)), c_f32p(_ptr(p)),
                                         c_f32p(a.ctypes.data), c_f32p(b.ctypes.data), i64(n), c_f32p(o0.ctypes.data), C.c_int(nthreads))
        return o0

    def aniso_sampling_tables(self, p22, er, ar):
        """tabular_anisotropic's sampling tables from its normalised p22 table (dj_brdf.h:2848-3103)."""
        p22 = _f32(p22).reshape(-1)
        one = [np.zeros(ar, np.float32) for _ in range(3)]
        two = [np.zeros(er * ar, np.float32) for _ in range(3)]
        counts = (C.c_int * 2)()
        self.lib.orc_aniso_sampling_tables(c_f32p(p22.ctypes.data), C.c_int(er), C.c_int(ar),
                                           *[c_f32p(a.ctypes.data) for a in (one[0], one[1], one[2], two[0], two[1], two[2])],
                                           counts)
        return dict(pdf1=one[0], cdf1=one[1], qf1=one[2], pdf2=two[0], cdf2=two[1], qf2=two[2],
                    n_qf1=counts[0], n_qf2=counts[1])

    def tabular_aniso_sample_query(self, op, fit, tabs, er, ar, u, wo, params=None, shadow=True, nthreads=1):
        """sample / evalp_is of djb::tabular_anisotropic on a fit_tabular_anisotropic() result + aniso_sampling_tables()."""
        code = {"sample": 3, "evalp_is": 4}[op]
        a, b = _f32(u), _f32(wo)
        n = len(b)
        fs = self._fres(Fresnel.spline(fit["fresnel"]))
        p = None if params is None else _f32(params)
        o0, o1, o2 = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        self.lib.orc_tabular_aniso_sample_query(
            C.c_int(code), c_f32p(_f32(fit["p22"]).ctypes.data), c_f32p(_f32(fit["sigma"]).ctypes.data),
            c_f32p(tabs["qf1"].ctypes.data), C.c_int(tabs["n_qf1"]), c_f32p(tabs["qf2"].ctypes.data), C.c_int(er), C.c_int(ar),
            C.byref(fs), C.c_int(int(shadow)), c_f32p(_ptr(p)), c_f32p(a.ctypes.data), c_f32p(b.ctypes.data), i64(n),
            c_f32p(o0.ctypes.data), c_f32p(o1.ctypes.data), c_f32p(o2.ctypes.data), C.c_int(nthreads))
        return (o0, o1, o2) if code == 4 else o0

    def fit_tabular_anisotropic(self, src, elev_res=90, azim_res=90, shadow=True, iterations=4, nthreads=8):
        s, keep = self._source(src)
        n = elev_res * azim_res
        p22, sigma = np.zeros(n, np.float32), np.zeros(n, np.float32)
        fres = np.zeros((elev_res, 3), np.float32)
        b5, g5 = np.zeros(5, np.float32), np.zeros(5, np.float32)
        self.lib.orc_fit_tabular_anisotropic(C.byref(s), C.c_int(elev_res), C.c_int(azim_res),
                                             C.c_int(int(shadow)), C.c_int(iterations),
                                             *[c_f32p(a.ctypes.data) for a in (p22, sigma, fres, b5, g5)],
                                             C.c_int(nthreads))
        return dict(p22=p22, sigma=sigma, fresnel=fres, beckmann=b5, ggx=g5)


# --------------------------------------------------------------------------------------------------
class RefOracle:
    """The unmodified reference, through oracle/ref_harness.cpp."""
    name = "reference"

    def __init__(self, opened=False):
        """opened=True loads libdjbref_open.so: the same harness + reference compiled with the access specifiers opened
        (oracle/ref_open.cpp), which additionally exports the reference's private sampling tables."""
        if not ref_available():
            raise RuntimeError("oracle/_ref/libdjbref.so not built (needs /root/reference; run `make -C oracle ref`)")
        self.opened = opened
        L = self.lib = C.CDLL(str(HERE / "_ref" / ("libdjbref_open.so" if opened else "libdjbref.so")))
        for fn in ("ref_microfacet_create", "ref_merl_open", "ref_utia_open", "ref_sgd_create",
                   "ref_abc_create", "ref_tabular_create", "ref_tabular_anisotropic_create"):
            getattr(L, fn).restype = C.c_void_p
        self._lean = {}
        self._tmp = tempfile.TemporaryDirectory(prefix="djbref_")
        self._handles = []
        assert L.ref_sizeof_params() == 48

    def _lean_lib(self, biased):
        name = "libleanref_biased.so" if biased else "libleanref.so"
        if name not in self._lean:
            self._lean[name] = C.CDLL(str(HERE / "_ref" / name))
        return self._lean[name]

    def params_elliptic(self, a1, a2, phi_a=0.0):
        out = np.zeros(12, np.float32)
        self.lib.ref_params_elliptic(C.c_float(a1), C.c_float(a2), C.c_float(phi_a), c_f32p(out.ctypes.data))
        return out

    def params_pdfparams(self, ax, ay, rho=0.0, tx=0.0, ty=0.0):
        out = np.zeros(12, np.float32)
        self.lib.ref_params_pdfparams(C.c_float(ax), C.c_float(ay), C.c_float(rho), C.c_float(tx),
                                      C.c_float(ty), c_f32p(out.ctypes.data))
        return out

    def microfacet(self, ndf, fresnel=None, shadow=True):
        fr = fresnel or Fresnel.ideal()
        h = self.lib.ref_microfacet_create(C.c_int(ndf), C.c_int(fr.kind), c_f32p(_ptr(fr.data) if fr.data.size else None),
                                           C.c_int(fr.data.size), C.c_int(int(shadow)))
        assert h
        return C.c_void_p(h)

    def destroy(self, h):
        self.lib.ref_brdf_destroy(h)

    def _q(self, fn, h, params, a, b, outs, nthreads):
        a, b = _f32(a), _f32(b)
        p = None if params is None else _f32(params)
        getattr(self.lib, fn)(h, c_f32p(_ptr(p)), c_f32p(a.ctypes.data), c_f32p(b.ctypes.data), i64(len(b)),
                              *[c_f32p(o.ctypes.data) for o in outs], C.c_int(nthreads))

    # generic-handle queries
    def brdf_eval(self, h, params, wi, wo, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._q("ref_brdf_eval", h, params, wi, wo, [out], nthreads)
        return out

    def _with_mf(self, ndf, fresnel, shadow, fn):
        h = self.microfacet(ndf, fresnel, shadow)
        try:
            return fn(h)
        finally:
            self.destroy(h)

    def eval(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        return self._with_mf(ndf, fresnel, shadow, lambda h: self.brdf_eval(h, params, wi, wo, nthreads))

    def evalp(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._with_mf(ndf, fresnel, shadow, lambda h: self._q("ref_brdf_evalp", h, params, wi, wo, [out], nthreads))
        return out

    def pdf(self, ndf, params, wi, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty(len(wo), np.float32)
        self._with_mf(ndf, fresnel, shadow, lambda h: self._q("ref_brdf_pdf", h, params, wi, wo, [out], nthreads))
        return out

    def sample(self, ndf, params, u, wo, fresnel=None, shadow=True, nthreads=1):
        out = np.empty((len(wo), 3), np.float32)
        self._with_mf(ndf, fresnel, shadow, lambda h: self._q("ref_brdf_sample", h, params, u, wo, [out], nthreads))
        return out

    def evalp_is(self, ndf, params, u, wo, fresnel=None, shadow=True, nthreads=1):
        w = np.empty((len(wo), 3), np.float32)
        i = np.empty((len(wo), 3), np.float32)
        pdf = np.empty(len(wo), np.float32)
        self._with_mf(ndf, fresnel, shadow,
                      lambda h: self._q("ref_brdf_evalp_is", h, params, u, wo, [w, i, pdf], nthreads))
        return w, i, pdf

    def component(self, what, ndf, params, a, b=None, c=None, fresnel=None, shadow=True):
        code = PortOracle.COMPONENTS[what]
        a = _f32(a)
        b = None if b is None else _f32(b)
        c = None if c is None else _f32(c)
        n = len(a)
        out = np.zeros((n, 3) if code == 7 else n, np.float32)
        p = None if params is None else _f32(params)
        h = self.microfacet(ndf, fresnel, shadow)
        try:
            rc = self.lib.ref_microfacet_component(h, c_f32p(_ptr(p)), C.c_int(code), c_f32p(a.ctypes.data), c_f32p(_ptr(b)),
                                                   c_f32p(_ptr(c)), i64(n), c_f32p(out.ctypes.data))
            assert rc == 0, rc
        finally:
            self.destroy(h)
        return out

    def io_to_hd(self, wi, wo):
        wi, wo = _f32(wi), _f32(wo)
        h, d = np.empty_like(wi), np.empty_like(wi)
        self.lib.ref_io_to_hd(c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                              c_f32p(h.ctypes.data), c_f32p(d.ctypes.data))
        return h, d

    def hd_to_io(self, h, d):
        h, d = _f32(h), _f32(d)
        wi, wo = np.empty_like(h), np.empty_like(h)
        self.lib.ref_hd_to_io(c_f32p(h.ctypes.data), c_f32p(d.ctypes.data), i64(len(h)),
                              c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data))
        return wi, wo

    def merl_index(self, wi, wo, nthreads=1):
        wi, wo = _f32(wi), _f32(wo)
        idx = np.empty(len(wi), np.int32)
        self.lib.ref_merl_index(c_f32p(wi.ctypes.data), c_f32p(wo.ctypes.data), i64(len(wi)),
                                C.c_void_p(idx.ctypes.data), C.c_int(nthreads))
        return idx

    def merl_open(self, path):
        h = self.lib.ref_merl_open(str(path).encode())
        assert h, path
        return C.c_void_p(h)

    def merl_from_table(self, table):
        path = os.path.join(self._tmp.name, f"merl_{len(os.listdir(self._tmp.name))}.binary")
        write_merl_file(path, table)
        h = self.merl_open(path)
        os.unlink(path)  # the reference copied the whole file into memory (dj_brdf.h:979-980)
        return h

    def utia_from_table(self, raw_table):
        path = os.path.join(self._tmp.name, f"utia_{len(os.listdir(self._tmp.name))}.bin")
        np.ascontiguousarray(raw_table, dtype=np.float64).tofile(path)
        h = self.lib.ref_utia_open(path.encode())
        assert h
        os.unlink(path)
        return C.c_void_p(h)

    def merl_eval(self, table, wi, wo, nthreads=1):
        h = self.merl_from_table(table)
        try:
            return self.brdf_eval(h, None, wi, wo, nthreads)
        finally:
            self.destroy(h)

    def utia_eval(self, raw_table, wi, wo, nthreads=1):
        h = self.utia_from_table(raw_table)
        try:
            return self.brdf_eval(h, None, wi, wo, nthreads)
        finally:
            self.destroy(h)

    def lrep_to_params(self, E5):
        E5 = _f32(E5).reshape(-1, 5)
        out = np.empty((len(E5), 12), np.float32)
        self.lib.ref_lrep_to_params(c_f32p(E5.ctypes.data), i64(len(E5)), c_f32p(out.ctypes.data))
        return out

    def params_to_lrep(self, params):
        params = _f32(params).reshape(-1, 12)
        out = np.empty((len(params), 5), np.float32)
        self.lib.ref_params_to_lrep(c_f32p(params.ctypes.data), i64(len(params)), c_f32p(out.ctypes.data))
        return out

    def lean_shading_params(self, E5, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        E5, alpha = _f32(E5).reshape(-1, 5), _f32(alpha)
        out = np.zeros((len(E5), 12), np.float32)
        self.lib.ref_lean_shading_params(C.c_float(bias), C.c_float(dmap_scale), C.c_int(int(lean_filtering)),
                                         C.c_int(int(alpha.ndim == 2)), c_f32p(alpha.ctypes.data), c_f32p(E5.ctypes.data),
                                         i64(len(E5)), c_f32p(out.ctypes.data))
        return out

    def nmap2leanmap(self, nmap_planar, base_roughness=1e-5, bias=0.0, run_check=False):
        assert bias in (0.0, 25.0)
        L = self._lean_lib(bias != 0.0)
        nmap = np.ascontiguousarray(nmap_planar, dtype=np.uint8)
        _, h, w = nmap.shape
        l1 = np.empty((4, h, w), np.float32)
        l2 = np.empty((4, h, w), np.float32)
        L.ref_nmap2leanmap(C.c_void_p(nmap.ctypes.data), C.c_int(w), C.c_int(h), C.c_float(base_roughness),
                           c_f32p(l1.ctypes.data), c_f32p(l2.ctypes.data), C.c_int(int(run_check)))
        return l1, l2

    def radial_query(self, what, x, ndf=None, src=None, res=90):
        """the same queries on the reference's objects: djb::ggx / djb::beckmann (ndf) or djb::tabular(src, res)"""
        code = {"p22": 0, "sigma_std": 1, "cdf": 2, "qf": 3}[what]
        x = _f32(x)
        out = np.zeros(len(x), np.float32)
        if src is not None:
            hs = self._source_handle(src)
            h = C.c_void_p(self.lib.ref_tabular_create(hs, C.c_int(res), C.c_int(1)))
        else:
            hs, h = None, self.microfacet(ndf)
        try:
            rc = self.lib.ref_radial_query(h, C.c_int(code), c_f32p(x.ctypes.data), C.c_int(len(x)), c_f32p(out.ctypes.data))
            assert rc == 0, rc
        finally:
            self.destroy(h)
            if hs is not None:
                self.destroy(hs)
        return out

    def dmap2nmap(self, dmap, scale=0.1):
        d = np.ascontiguousarray(dmap, np.uint8)
        h, w = d.shape
        out = np.zeros((3, h, w), np.uint8)
        lib = C.CDLL(str(HERE / "_ref" / "libdmapref.so"))
        lib.ref_dmap2nmap(C.c_void_p(d.ctypes.data), C.c_int(w), C.c_int(h), C.c_float(scale), C.c_void_p(out.ctypes.data))
        return out

    # -- fits
    def _source_handle(self, src):
        if src.kind == "microfacet":
            return self.microfacet(src.ndf, src.fresnel, src.shadow)
        if src.kind == "merl":
            return self.merl_from_table(src.table)
        if src.kind == "utia":
            return self.utia_from_table(src.table)
        if src.kind in ("sgd", "abc"):
            return self.analytic(src.kind, src.name)
        raise ValueError(src.kind)

    def analytic(self, kind, name):
        """djb::sgd(name) / djb::abc(name); None when the reference throws (unknown material)."""
        h = (self.lib.ref_sgd_create if kind == "sgd" else self.lib.ref_abc_create)(str(name).encode())
        return C.c_void_p(h) if h else None

    def sgd_eval(self, name, wi, wo, nthreads=1):
        h = self.analytic("sgd", name)
        try:
            return self.brdf_eval(h, None, wi, wo, nthreads)
        finally:
            self.destroy(h)

    def abc_eval(self, name, wi, wo, nthreads=1):
        h = self.analytic("abc", name)
        try:
            return self.brdf_eval(h, None, wi, wo, nthreads)
        finally:
            self.destroy(h)

    def fit_tabular(self, src, res=90, shadow=True, iterations=4):
        assert iterations == 4, "the reference hard-codes 4 power iterations (dj_brdf.h:2518)"
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_create(h, C.c_int(res), C.c_int(int(shadow))))
        p22, sigma, cdf, qf = (np.zeros(res, np.float32) for _ in range(4))
        fres = np.zeros((res, 3), np.float32)
        alpha = np.zeros(2, np.float32)
        n = self.lib.ref_tabular_get(t, *[c_f32p(a.ctypes.data) for a in (p22, sigma, cdf, qf, fres, alpha)])
        assert n == res, n
        self.destroy(t)
        self.destroy(h)
        return dict(p22=p22, sigma=sigma, cdf=cdf, qf=qf, fresnel=fres, alpha=alpha)

    def tabular_query(self, op, src, res, u_or_wi, wo, params=None, shadow=True, nthreads=1):
        """The reference's djb::tabular object built from `src`, queried through the virtual brdf interface."""
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_create(h, C.c_int(res), C.c_int(int(shadow))))
        n = len(wo)
        try:
            if op == "eval":
                return self.brdf_eval(t, params, u_or_wi, wo, nthreads)
            if op == "evalp_is":
                w, i, pdf = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty(n, np.float32)
                self._q("ref_brdf_evalp_is", t, params, u_or_wi, wo, [w, i, pdf], nthreads)
                return w, i, pdf
            out = np.empty(n if op == "pdf" else (n, 3), np.float32)
            self._q({"evalp": "ref_brdf_evalp", "pdf": "ref_brdf_pdf", "sample": "ref_brdf_sample"}[op], t, params, u_or_wi,
                    wo, [out], nthreads)
            return out
        finally:
            self.destroy(t)
            self.destroy(h)

    def tabular_aniso_query(self, op, src, er, ar, wi, wo, params=None, shadow=True, nthreads=1):
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_anisotropic_create(h, C.c_int(er), C.c_int(ar), C.c_int(int(shadow))))
        n = len(wo)
        try:
            if op == "eval":
                return self.brdf_eval(t, params, wi, wo, nthreads)
            if op == "evalp_is":
                w, i, pdf = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty(n, np.float32)
                self._q("ref_brdf_evalp_is", t, params, wi, wo, [w, i, pdf], nthreads)
                return w, i, pdf
            out = np.empty(n if op == "pdf" else (n, 3), np.float32)
            self._q({"evalp": "ref_brdf_evalp", "pdf": "ref_brdf_pdf", "sample": "ref_brdf_sample"}[op], t, params, wi, wo,
                    [out], nthreads)
            return out
        finally:
            self.destroy(t)
            self.destroy(h)

    def aniso_sampling_tables(self, src, er, ar, shadow=True):
        """The reference's private m_pdf1 / m_cdf1 / m_qf1 / m_pdf2 / m_cdf2 / m_qf2 (needs RefOracle(opened=True))
        plus its p22 table -> dict; `sizes` = the six vector sizes."""
        assert self.opened, "private tables need RefOracle(opened=True)"
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_anisotropic_create(h, C.c_int(er), C.c_int(ar), C.c_int(int(shadow))))
        try:
            one = [np.zeros(ar, np.float32) for _ in range(3)]
            two = [np.zeros(er * ar, np.float32) for _ in range(3)]
            sizes = (C.c_int * 6)()
            rc = self.lib.ref_tabular_anisotropic_sampling_tables(
                t, *[c_f32p(a.ctypes.data) for a in (one[0], one[1], one[2], two[0], two[1], two[2])], sizes)
            assert rc == 0, rc
            p22, sigma = np.zeros(er * ar, np.float32), np.zeros(er * ar, np.float32)
            self.lib.ref_tabular_anisotropic_get(t, c_f32p(p22.ctypes.data), c_f32p(sigma.ctypes.data), None, None, None)
            return dict(pdf1=one[0], cdf1=one[1], qf1=one[2], pdf2=two[0], cdf2=two[1], qf2=two[2], p22=p22,
                        sigma=sigma, sizes=list(sizes))
        finally:
            self.destroy(t)
            self.destroy(h)

    # ---- the remaining public scalar members (dj_brdf.h:366-369, 384-389, 450-455, 506-509, 531-533) -----------------
    MEMBER_CODES = {"qf1": 0, "qf2_radial": 1, "qf3_radial": 2, "ndf": 20, "gaf": 21, "g1": 22, "fresnel": 23}

    def member_query(self, what, a, b=None, c=None, ndf=None, sgd=None, abc=None):
        """beckmann / ggx ::qf1 / qf2_radial / qf3_radial (ndf=...), sgd ::ndf / gaf / g1 / fresnel (sgd=name), abc ::ndf / gaf /
        fresnel (abc=name) on the reference's own objects; vec3 arguments as [n, 3]."""
        code = self.MEMBER_CODES[what]
        h = self.microfacet(ndf) if ndf is not None else self.analytic("sgd" if sgd else "abc", sgd or abc)
        a = _f32(a)
        b, c = (None if x is None else _f32(x) for x in (b, c))
        vec_in = ndf is None and what != "fresnel"
        n = len(a) if not vec_in else a.reshape(-1, 3).shape[0]
        out = np.zeros((n, 3), np.float32)
        try:
            rc = self.lib.ref_member_query(h, C.c_int(code), c_f32p(a.ctypes.data), c_f32p(_ptr(b)), c_f32p(_ptr(c)), C.c_int(n),
                                           c_f32p(out.ctypes.data))
            assert rc in (1, 3), rc
        finally:
            self.destroy(h)
        return out.reshape(-1)[:n].copy() if rc == 1 else out

    def tabular_aniso_lookup(self, src, er, ar, what, a, b=None, shadow=True):
        """tabular_anisotropic::pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 of the reference's object fitted to `src` (needs opened=True:
        the wrapper lives in oracle/ref_open.cpp)."""
        assert self.opened, "ref_tabular_anisotropic_lookup is built into libdjbref_open.so"
        code = {"pdf1": 0, "cdf1": 1, "qf1": 2, "pdf2": 3, "cdf2": 4, "qf2": 5}[what]
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_anisotropic_create(h, C.c_int(er), C.c_int(ar), C.c_int(int(shadow))))
        try:
            return self.tabular_aniso_lookup_on(t, what, a, b)
        finally:
            self.destroy(t)
            self.destroy(h)

    def tabular_aniso_lookup_on(self, t, what, a, b=None):
        code = {"pdf1": 0, "cdf1": 1, "qf1": 2, "pdf2": 3, "cdf2": 4, "qf2": 5}[what]
        a = _f32(a)
        b = _f32(b) if b is not None else np.zeros_like(a)
        out = np.zeros(len(a), np.float32)
        rc = self.lib.ref_tabular_anisotropic_lookup(t, C.c_int(code), c_f32p(a.ctypes.data), c_f32p(b.ctypes.data),
                                                     C.c_int(len(a)), c_f32p(out.ctypes.data))
        assert rc == 0, rc
        return out

    def fit_tabular_anisotropic(self, src, elev_res=90, azim_res=90, shadow=True, iterations=4, nthreads=1):
        assert iterations == 4
        h = self._source_handle(src)
        t = C.c_void_p(self.lib.ref_tabular_anisotropic_create(h, C.c_int(elev_res), C.c_int(azim_res),
                                                               C.c_int(int(shadow))))
        n = elev_res * azim_res
        p22, sigma = np.zeros(n, np.float32), np.zeros(n, np.float32)
        fres = np.zeros((elev_res, 3), np.float32)
        b5, g5 = np.zeros(5, np.float32), np.zeros(5, np.float32)
        got = self.lib.ref_tabular_anisotropic_get(t, *[c_f32p(a.ctypes.data) for a in (p22, sigma, fres, b5, g5)])
        assert got == n, got
        self.destroy(t)
        self.destroy(h)
        return dict(p22=p22, sigma=sigma, fresnel=fres, beckmann=b5, ggx=g5)


# --------------------------------------------------------------------------------------------------
# Seeded synthetic inputs -- generated on the host with numpy so that the
# checker and the CUDA path see bit-identical floats.
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniforms(n, stream, seed=0x9E3779B97F4A7C15):
    """n float32 uniforms in [0,1) from a counter-based splitmix64 stream."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) * np.uint64(8) + np.uint64(stream)
        bits = _splitmix64(idx ^ np.uint64(seed))
    return ((bits >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)


def directions(n, stream, zmin=0.001):
    """Upper-hemisphere unit vectors: z = 1 - (1-zmin) u1, phi = 2 pi u2 (float32 AoS [n,3])."""
    u1 = uniforms(n, stream).astype(np.float64)
    u2 = uniforms(n, stream + 1).astype(np.float64)
    z = 1.0 - (1.0 - zmin) * u1
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * np.pi * u2
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
