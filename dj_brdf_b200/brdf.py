"""Host-side mirror of the reference's ``djb::`` interface for the accelerated paths, on top of the
C-ABI (``libdjb200.so``).  Same names, argument order and error behaviour as dj_brdf.h, but every
query is batched: ``i``/``o`` are ``[n, 3]`` float32 arrays (numpy = host memory, torch CUDA tensor
= device memory, results come back in the same space).

    ggx = djb.ggx(djb.fresnel.schlick([1.0, 0.8, 0.6]))
    fr  = ggx.eval(i, o, djb.params.elliptic(0.1, 0.4, 0.7))         # dj_brdf.h:1551
    fr16 = ggx.eval(i, o, [p0, ..., p15])                             # 16 materials x n pairs

Direction convention as in the reference (dj_brdf.h:23-26): i -> light, o -> viewer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import Buf, DjbError, check


# --------------------------------------------------------------------------------------------------
class fresnel:
    """djb::fresnel::* (dj_brdf.h:149-207)."""

    class impl:
        kind = capi.FRESNEL_IDEAL

        def __init__(self, data=(), points=None):
            self._v = np.zeros(6, np.float32)
            d = np.asarray(data, np.float32).reshape(-1)
            self._v[:d.size] = d
            self._points = None if points is None else np.ascontiguousarray(points, np.float32).reshape(-1, 3)

        def _c(self):
            f = capi.Fresnel()
            f.kind = self.kind
            for k in range(6):
                f.v[k] = float(self._v[k])
            if self._points is not None:
                f.points = self._points.ctypes.data
                f.n_points = len(self._points)
            return f

        def copy(self):
            c = type(self).__new__(type(self))
            c._v = self._v.copy()
            c._points = None if self._points is None else self._points.copy()
            return c

    class ideal(impl):
        kind = capi.FRESNEL_IDEAL

        def __init__(self):
            super().__init__()

    class schlick(impl):
        kind = capi.FRESNEL_SCHLICK

        def __init__(self, f0):
            super().__init__(np.broadcast_to(np.asarray(f0, np.float32), (3,)))

    class unpolarized(impl):
        kind = capi.FRESNEL_UNPOLARIZED

        def __init__(self, ior):
            super().__init__(np.broadcast_to(np.asarray(ior, np.float32), (3,)))

    class sgd(impl):
        kind = capi.FRESNEL_SGD

        def __init__(self, f0, f1):
            super().__init__(np.concatenate([np.asarray(f0, np.float32), np.asarray(f1, np.float32)]))

    class spline(impl):
        kind = capi.FRESNEL_SPLINE

        def __init__(self, points):
            super().__init__((), points)

        def get_points(self):
            return self._points

    @staticmethod
    def ior_to_f0(ior):  # dj_brdf.h:1255-1262
        ior = np.asarray(ior, np.float32)
        tmp = ((ior.astype(np.float64) - 1.0) / (ior.astype(np.float64) + 1.0)).astype(np.float32)
        return tmp * tmp


# --------------------------------------------------------------------------------------------------
class params:
    """djb::microfacet::params (dj_brdf.h:213-243): a 48-byte block, here a float32[12] ndarray with
    fields n[3], a1, a2, phi_a, ax, ay, rho, sqrt(1-rho^2), tx_n, ty_n."""

    @staticmethod
    def _make(fn, *args):
        out = np.zeros(12, np.float32)
        check(getattr(capi.load(), fn)(*[C.c_float(a) for a in args], C.c_void_p(out.ctypes.data)))
        return out

    @staticmethod
    def standard():
        return params._make("djb200_params_standard")

    @staticmethod
    def isotropic(a):
        return params._make("djb200_params_isotropic", a)

    @staticmethod
    def elliptic(a1, a2, phi_a=0.0):
        return params._make("djb200_params_elliptic", a1, a2, phi_a)

    @staticmethod
    def pdfparams(ax, ay, rho=0.0, tx_n=0.0, ty_n=0.0):
        return params._make("djb200_params_pdfparams", ax, ay, rho, tx_n, ty_n)

    @staticmethod
    def get_ellipse(p):
        return float(p[3]), float(p[4]), float(p[5])

    @staticmethod
    def get_pdfparams(p):
        return float(p[6]), float(p[7]), float(p[8]), float(p[10]), float(p[11])


def _n_pairs(x, width):
    n = Buf(x, np.float32).n
    if n % width:
        raise ValueError(f"array size {n} is not a multiple of {width}")
    return n // width


# --------------------------------------------------------------------------------------------------
class brdf:
    """djb::brdf (dj_brdf.h:74-109), batched."""

    def eval(self, i, o, user_param=None):
        raise NotImplementedError

    def evalp(self, i, o, user_param=None):
        # brdf::evalp: eval * cos(theta_i), dj_brdf.h:803-806
        fr = self.eval(i, o, user_param)
        iz = i[:, 2:3] if not capi._is_torch(i) else i[:, 2:3]
        return fr * iz

    @staticmethod
    def io_to_hd(i, o):
        bi, bo = Buf(i, np.float32), Buf(o, np.float32)
        mem = capi.same_space(bi, bo)
        n = bi.n // 3
        h = capi.empty_like_space(bi.keep, (n, 3), np.float32)
        d = capi.empty_like_space(bi.keep, (n, 3), np.float32)
        bh, bd = Buf(h, np.float32, True), Buf(d, np.float32, True)
        check(capi.load().djb200_io_to_hd(bi.ptr, bo.ptr, C.c_int64(n), bh.ptr, bd.ptr, mem,
                                          capi.current_stream_ptr(mem)))
        return h, d

    @staticmethod
    def hd_to_io(h, d):
        bh, bd = Buf(h, np.float32), Buf(d, np.float32)
        mem = capi.same_space(bh, bd)
        n = bh.n // 3
        i = capi.empty_like_space(bh.keep, (n, 3), np.float32)
        o = capi.empty_like_space(bh.keep, (n, 3), np.float32)
        bi, bo = Buf(i, np.float32, True), Buf(o, np.float32, True)
        check(capi.load().djb200_hd_to_io(bh.ptr, bd.ptr, C.c_int64(n), bi.ptr, bo.ptr, mem,
                                          capi.current_stream_ptr(mem)))
        return i, o


class microfacet(brdf):
    """djb::microfacet (dj_brdf.h:210-298) restricted to the radial families the kernels implement."""
    _ndf = None

    def __init__(self, fresnel_impl=None, shadow=True):
        self.m_fresnel = (fresnel_impl or fresnel.ideal()).copy()  # deep copy, dj_brdf.h:1514
        self.m_shadow = bool(shadow)

    # mutators / accessors, dj_brdf.h:278-282
    def set_shadow(self, shadow):
        self.m_shadow = bool(shadow)

    def set_fresnel(self, f):
        self.m_fresnel = f.copy()

    def get_shadow(self):
        return self.m_shadow

    def get_fresnel(self):
        return self.m_fresnel

    def supports_smith_vndf_sampling(self):
        return True

    def _desc(self):
        m = capi.Microfacet()
        m.ndf = self._ndf
        m.shadow = int(self.m_shadow)
        m.fresnel = self.m_fresnel._c()
        return m

    _prefix = "djb200_microfacet_"

    def _first_arg(self):
        """First argument of the C-ABI query: the construction-state descriptor (-> pointer, keepalive)."""
        d = self._desc()
        return C.byref(d), d

    def _params(self, user_param, n, mem):
        """-> (pointer, n_params, layout, keepalive, broadcast_count or None)"""
        if user_param is None:
            return None, 0, capi.PARAMS_BROADCAST, None, None
        if capi._is_torch(user_param):  # per-pair params living on the device
            b = Buf(user_param, np.float32)
            if b.n != 12 * n:
                raise ValueError("device params must be [n, 12] (PER_PAIR layout)")
            return b.ptr, n, capi.PARAMS_PER_PAIR, b, None
        p = np.ascontiguousarray(user_param, np.float32)
        if p.ndim == 1:
            p = p.reshape(1, 12)
            return C.c_void_p(p.ctypes.data), 1, capi.PARAMS_BROADCAST, p, None
        return C.c_void_p(p.ctypes.data), len(p), capi.PARAMS_BROADCAST, p, len(p)

    def _query(self, fn, a, b, a_width, out_widths, user_param, per_pair=False):
        ba, bb = Buf(a, np.float32), Buf(b, np.float32)
        mem = capi.same_space(ba, bb)
        n = bb.n // 3
        if ba.n != n * a_width:
            raise ValueError("input arrays disagree on the number of pairs")
        if per_pair and user_param is not None and not capi._is_torch(user_param):
            p = np.ascontiguousarray(user_param, np.float32).reshape(-1, 12)
            if len(p) != n:
                raise ValueError("per_pair params must be [n, 12]")
            if mem != capi.MEM_HOST:
                raise ValueError("host per-pair params need host direction arrays")
            pptr, npar, layout, keep, M = C.c_void_p(p.ctypes.data), n, capi.PARAMS_PER_PAIR, p, None
        else:
            pptr, npar, layout, keep, M = self._params(user_param, n, mem)
        outs = []
        for w in out_widths:
            shape = ((n, w) if w > 1 else (n,)) if M is None else ((M, n, w) if w > 1 else (M, n))
            outs.append(capi.empty_like_space(bb.keep, shape, np.float32))
        bouts = [Buf(x, np.float32, True) for x in outs]
        first, keep_first = self._first_arg()
        check(getattr(capi.load(), fn)(first, pptr, C.c_int64(npar), C.c_int(layout), ba.ptr, bb.ptr,
                                       C.c_int64(n), *[x.ptr for x in bouts], C.c_int(mem),
                                       capi.current_stream_ptr(mem)))
        return outs

    # BRDF interface (dj_brdf.h:247-256).  user_param: None (standard), one params block, a list /
    # [M,12] array of blocks (every pair under every block -> leading dim M), or with
    # per_pair=True an [n,12] array (pair k under block k).
    def eval(self, i, o, user_param=None, per_pair=False):
        return self._query(self._prefix + "eval", i, o, 3, [3], user_param, per_pair)[0]

    def evalp(self, i, o, user_param=None, per_pair=False):
        return self._query(self._prefix + "evalp", i, o, 3, [3], user_param, per_pair)[0]

    def pdf(self, i, o, user_param=None, per_pair=False):
        return self._query(self._prefix + "pdf", i, o, 3, [1], user_param, per_pair)[0]

    def sample(self, u, o, user_param=None, per_pair=False):
        """u: [n, 2] uniforms (u1, u2) -- microfacet::sample(u1, u2, o, user_param)."""
        return self._query(self._prefix + "sample", u, o, 2, [3], user_param, per_pair)[0]

    def evalp_is(self, u, o, user_param=None, per_pair=False):
        """-> (weight rgb, i, pdf) -- microfacet::evalp_is(u1, u2, o, &i, &pdf, user_param)."""
        return tuple(self._query(self._prefix + "evalp_is", u, o, 2, [3, 3, 1], user_param, per_pair))


    # ---- the public component queries (dj_brdf.h:258-272), batched under one params block ---------------------------------
    def _component(self, what, a, b=None, c=None, user_param=None):
        bufs = [Buf(x, np.float32) for x in (a, b, c)]
        mem = capi.same_space(*[x for x in bufs if x.mem is not None])
        n = bufs[0].n // 3
        out = capi.empty_like_space(bufs[0].keep, (n, 3) if what == 7 else (n,), np.float32)
        bo = Buf(out, np.float32, True)
        p = None if user_param is None else np.ascontiguousarray(user_param, np.float32).reshape(12)
        d = self._desc()
        check(capi.load().djb200_microfacet_component(C.byref(d), None if p is None else C.c_void_p(p.ctypes.data), C.c_int(what),
                                                      bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, C.c_int64(n), bo.ptr, C.c_int(mem),
                                                      capi.current_stream_ptr(mem)))
        return out

    def ndf(self, h, user_param=None):
        return self._component(0, h, user_param=user_param)

    def gaf(self, h, i, o, user_param=None):
        return self._component(1, h, i, o, user_param)

    def g1(self, h, k, user_param=None):
        return self._component(2, h, k, user_param=user_param)

    def sigma(self, k, user_param=None):
        return self._component(3, k, user_param=user_param)

    def p22(self, xy, user_param=None):
        """xy: [n, 3] with the slopes in the first two columns."""
        return self._component(4, xy, user_param=user_param)

    def vp22(self, xy, k, user_param=None):
        return self._component(5, xy, k, user_param=user_param)

    def vndf(self, h, k, user_param=None):
        return self._component(6, h, k, user_param=user_param)

    def fresnel_term(self, cos_theta_d):
        """microfacet::fresnel(cos_theta_d), dj_brdf.h:258; cos_theta_d: [n, 3] with the cosine in the first column -> rgb."""
        return self._component(7, cos_theta_d)

    # ---- LEAN-filtered shading, fused (mitsuba/dj_beckmannconductor.cpp:283-319, 338-366, 379-410) ----------------
    @staticmethod
    def _lean_cfg(alpha, n, bias, dmap_scale, lean_filtering, like):
        cfg = capi.LeanShading()
        cfg.bias, cfg.dmap_scale, cfg.lean_filtering = bias, dmap_scale, int(bool(lean_filtering))
        if capi._is_torch(alpha) or np.ndim(alpha) == 2:
            ba = Buf(alpha, np.float32)
            if ba.n != 3 * n:
                raise ValueError("per-pair roughness must be [n, 3] (alpha1, alpha2, alphaAngle)")
            cfg.alpha_per_pair = 1
            return cfg, ba
        a = np.asarray(alpha, np.float32).reshape(3)
        cfg.alpha_per_pair = 0
        cfg.alpha[0], cfg.alpha[1], cfg.alpha[2] = float(a[0]), float(a[1]), float(a[2])
        return cfg, Buf(None, np.float32)

    @staticmethod
    def lean_shading_params(E, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        """The params block the plugin builds per shading point: E [n, 5] LEAN texels (E1..E5 as fetched, still biased),
        alpha = (alpha1, alpha2, alphaAngle) for all pairs or [n, 3] -> [n, 12]."""
        bE = Buf(E, np.float32)
        n = bE.n // 5
        cfg, ba = microfacet._lean_cfg(alpha, n, bias, dmap_scale, lean_filtering, bE.keep)
        mem = capi.same_space(bE, ba)
        out = capi.empty_like_space(bE.keep, (n, 12), np.float32)
        bo = Buf(out, np.float32, True)
        check(capi.load().djb200_lean_shading_params(C.byref(cfg), ba.ptr, bE.ptr, C.c_int64(n), bo.ptr, C.c_int(mem),
                                                     capi.current_stream_ptr(mem)))
        return out

    def _lean_query(self, fn, a, b, a_width, out_widths, E, alpha, bias, dmap_scale, lean_filtering):
        ba_, bb, bE = Buf(a, np.float32), Buf(b, np.float32), Buf(E, np.float32)
        n = bb.n // 3
        if ba_.n != n * a_width or bE.n != 5 * n:
            raise ValueError("input arrays disagree on the number of pairs")
        cfg, balpha = microfacet._lean_cfg(alpha, n, bias, dmap_scale, lean_filtering, bb.keep)
        mem = capi.same_space(ba_, bb, bE, balpha)
        outs = [capi.empty_like_space(bb.keep, (n, w) if w > 1 else (n,), np.float32) for w in out_widths]
        bouts = [Buf(x, np.float32, True) for x in outs]
        d = self._desc()
        check(getattr(capi.load(), fn)(C.byref(d), C.byref(cfg), balpha.ptr, bE.ptr, ba_.ptr, bb.ptr, C.c_int64(n),
                                       *[x.ptr for x in bouts], C.c_int(mem), capi.current_stream_ptr(mem)))
        return outs

    def evalp_lean(self, i, o, E, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        return self._lean_query("djb200_lean_shading_evalp", i, o, 3, [3], E, alpha, bias, dmap_scale, lean_filtering)[0]

    def pdf_lean(self, i, o, E, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        return self._lean_query("djb200_lean_shading_pdf", i, o, 3, [1], E, alpha, bias, dmap_scale, lean_filtering)[0]

    def evalp_is_lean(self, u, o, E, alpha, bias=25.0, dmap_scale=1.0, lean_filtering=True):
        return tuple(self._lean_query("djb200_lean_shading_evalp_is", u, o, 2, [3, 3, 1], E, alpha, bias, dmap_scale,
                                      lean_filtering))


    # ---- djb::radial's public scalar queries (dj_brdf.h:307-310), batched -----------------------------------------------
    def _radial(self, what, x):
        b = Buf(x, np.float32)
        out = capi.empty_like_space(b.keep, (b.n,), np.float32)
        bo = Buf(out, np.float32, True)
        if self._ndf is None:  # tabular: the handle carries the tables
            h, _keep = self._first_arg()
            if getattr(self, "m_azimuthal_res", 0):
                raise DjbError(1, "tabular_anisotropic is not a radial distribution")
            check(capi.load().djb200_radial_query(C.c_int(what), C.c_int(0), h, b.ptr, C.c_int64(b.n), bo.ptr, C.c_int(b.mem),
                                                  capi.current_stream_ptr(b.mem)))
        else:
            check(capi.load().djb200_radial_query(C.c_int(what), C.c_int(self._ndf), None, b.ptr, C.c_int64(b.n), bo.ptr,
                                                  C.c_int(b.mem), capi.current_stream_ptr(b.mem)))
        return out

    def p22_radial(self, r_sqr):
        return self._radial(0, r_sqr)

    def sigma_std_radial(self, cos_theta_k):
        return self._radial(1, cos_theta_k)

    def cdf_radial(self, r):
        return self._radial(2, r)

    def qf_radial(self, u):
        return self._radial(3, u)

    # microfacet::qf2 / qf3 and radial::qf2_radial / qf3_radial throw in the reference unless the family overrides them
    # (dj_brdf.h:1783-1791, 1848-1860); ggx / beckmann override the radial pair below
    def qf2(self, u, k):
        raise DjbError(1, "djb_error: Not Implemented")

    def qf3(self, u, k, qf2):
        raise DjbError(1, "djb_error: Not Implemented")

    def qf2_radial(self, u, cos_theta_k, sin_theta_k):
        raise DjbError(1, "djb_error: Not Implemented")

    def qf3_radial(self, u, qf2):
        raise DjbError(1, "djb_error: Not Implemented")


def _member_call(fn_name, first, what, args, widths, out_width):
    """One of the scalar-member entry points (include/djb200.h: djb200_quantile_query, djb200_tabular_anisotropic_query,
    djb200_sgd_member, djb200_abc_member): `args` are up to three argument arrays (None = unused) of `widths` floats per item."""
    bufs = [Buf(a, np.float32) if a is not None else None for a in args]
    lead = next(b for b in bufs if b is not None)
    mem = capi.same_space(*[b for b in bufs if b is not None])
    n = lead.n // widths[0]
    out = capi.empty_like_space(lead.keep, (n, out_width) if out_width > 1 else (n,), np.float32)
    bo = Buf(out, np.float32, True)
    ptrs = [b.ptr if b is not None else None for b in bufs]
    check(getattr(capi.load(), fn_name)(first, C.c_int32(what), *ptrs, C.c_int64(n), bo.ptr, C.c_int(mem),
                                        capi.current_stream_ptr(mem)))
    return out


class _quantile_members:
    """beckmann / ggx ::qf1, qf2_radial, qf3_radial (dj_brdf.h:366-369, 384-389), batched."""

    def qf1(self, u):
        return _member_call("djb200_quantile_query", C.c_int32(self._ndf), capi.MEMBER_QF1, [u, None, None], [1], 1)

    def qf2_radial(self, u, cos_theta_k, sin_theta_k):
        return _member_call("djb200_quantile_query", C.c_int32(self._ndf), capi.MEMBER_QF2_RADIAL, [u, cos_theta_k, sin_theta_k],
                            [1, 1, 1], 1)

    def qf3_radial(self, u, qf2):
        return _member_call("djb200_quantile_query", C.c_int32(self._ndf), capi.MEMBER_QF3_RADIAL, [u, qf2, None], [1, 1], 1)


class ggx(_quantile_members, microfacet):
    _ndf = capi.NDF_GGX


class beckmann(_quantile_members, microfacet):
    _ndf = capi.NDF_BECKMANN

    # LEAN algebra (dj_brdf.h:355-356)
    @staticmethod
    def lrep_to_params(E):
        b = Buf(E, np.float32)
        n = b.n // 5
        out = capi.empty_like_space(b.keep, (n, 12), np.float32)
        bo = Buf(out, np.float32, True)
        check(capi.load().djb200_lrep_to_params(b.ptr, C.c_int64(n), bo.ptr, b.mem, capi.current_stream_ptr(b.mem)))
        return out

    @staticmethod
    def params_to_lrep(p):
        b = Buf(p, np.float32)
        n = b.n // 12
        out = capi.empty_like_space(b.keep, (n, 5), np.float32)
        bo = Buf(out, np.float32, True)
        check(capi.load().djb200_params_to_lrep(b.ptr, C.c_int64(n), bo.ptr, b.mem, capi.current_stream_ptr(b.mem)))
        return out


# --------------------------------------------------------------------------------------------------
class _table_brdf(brdf):
    _destroy = None
    _eval = None

    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                getattr(capi.load(), self._destroy)(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def eval(self, i, o, user_param=None):
        bi, bo = Buf(i, np.float32), Buf(o, np.float32)
        mem = capi.same_space(bi, bo)
        n = bi.n // 3
        out = capi.empty_like_space(bi.keep, (n, 3), np.float32)
        bout = Buf(out, np.float32, True)
        check(getattr(capi.load(), self._eval)(self._h, bi.ptr, bo.ptr, C.c_int64(n), bout.ptr, C.c_int(mem),
                                               capi.current_stream_ptr(mem)))
        return out


class merl(_table_brdf):
    """djb::merl (dj_brdf.h:126-133): from a .binary file name or a [3*1458000] float64 sample array."""
    _destroy, _eval = "djb200_merl_destroy", "djb200_merl_eval"

    def __init__(self, filename_or_samples):
        super().__init__()
        if isinstance(filename_or_samples, (str, bytes)) or hasattr(filename_or_samples, "__fspath__"):
            check(capi.load().djb200_merl_load(str(filename_or_samples).encode(), C.byref(self._h)))
        else:
            s = np.ascontiguousarray(filename_or_samples, np.float64).reshape(-1)
            if s.size != 3 * 90 * 90 * 180:
                raise DjbError(1, "MERL sample array must hold 3*90*90*180 doubles")
            check(capi.load().djb200_merl_create(C.c_void_p(s.ctypes.data), C.byref(self._h)))

    @staticmethod
    def index(i, o):
        """The cell index merl::eval computes (dj_brdf.h:997-1002)."""
        bi, bo = Buf(i, np.float32), Buf(o, np.float32)
        mem = capi.same_space(bi, bo)
        n = bi.n // 3
        out = capi.empty_like_space(bi.keep, (n,), np.int32)
        bout = Buf(out, np.int32, True)
        check(capi.load().djb200_merl_index(bi.ptr, bo.ptr, C.c_int64(n), bout.ptr, C.c_int(mem),
                                            capi.current_stream_ptr(mem)))
        return out


def merl_filter_stats(i, o):
    """Property check of the filtered MERL lookup on device arrays (djb200_debug_merl_filter_stats):
    -> dict(rejected, certified_wrong, max_d_error, h_mismatch)."""
    bi, bo = Buf(i, np.float32), Buf(o, np.float32)
    if capi.same_space(bi, bo) != capi.MEM_DEVICE:
        raise ValueError("merl_filter_stats needs device arrays")
    st = (C.c_uint64 * 5)()
    check(capi.load().djb200_debug_merl_filter_stats(bi.ptr, bo.ptr, C.c_int64(bi.n // 3), st,
                                                     capi.current_stream_ptr(capi.MEM_DEVICE)))
    return dict(rejected=int(st[0]), certified_wrong=int(st[1]),
                max_d_error=float(np.array([st[2] & 0xFFFFFFFF], np.uint32).view(np.float32)[0]), h_mismatch=int(st[3]),
                max_acos_error=float(np.array([st[4] & 0xFFFFFFFF], np.uint32).view(np.float32)[0]))


class utia(_table_brdf):
    """djb::utia (dj_brdf.h:136-146): from a .bin file name or the raw [3*6*48*6*48] float64 samples."""
    _destroy, _eval = "djb200_utia_destroy", "djb200_utia_eval"

    def __init__(self, filename_or_samples):
        super().__init__()
        if isinstance(filename_or_samples, (str, bytes)) or hasattr(filename_or_samples, "__fspath__"):
            check(capi.load().djb200_utia_load(str(filename_or_samples).encode(), C.byref(self._h)))
        else:
            s = np.ascontiguousarray(filename_or_samples, np.float64).reshape(-1)
            if s.size != 3 * 6 * 48 * 6 * 48:
                raise DjbError(1, "UTIA sample array must hold 3*6*48*6*48 doubles")
            check(capi.load().djb200_utia_create(C.c_void_p(s.ctypes.data), C.byref(self._h)))


class _analytic_brdf(brdf):
    """Shared by djb::sgd and djb::abc: coefficients looked up by MERL material name, eval through the C-ABI."""
    _preset = _eval = _names = None
    _data_type = None

    def __init__(self, name):
        self._data = self._data_type()
        self.name = name
        check(getattr(capi.load(), self._preset)(str(name).encode(), C.byref(self._data)))

    @classmethod
    def names(cls):
        lib = capi.load()
        return [getattr(lib, cls._names)(C.c_int32(k)).decode() for k in range(lib.djb200_preset_count())]

    def eval(self, i, o, user_param=None):
        bi, bo = Buf(i, np.float32), Buf(o, np.float32)
        mem = capi.same_space(bi, bo)
        n = bi.n // 3
        out = capi.empty_like_space(bi.keep, (n, 3), np.float32)
        bout = Buf(out, np.float32, True)
        check(getattr(capi.load(), self._eval)(C.byref(self._data), bi.ptr, bo.ptr, C.c_int64(n), bout.ptr, C.c_int(mem),
                                               capi.current_stream_ptr(mem)))
        return out


class sgd(_analytic_brdf):
    """djb::sgd (dj_brdf.h:481-511): shifted-gamma-distribution BRDF of a MERL material, by name."""
    _preset, _eval, _names, _data_type = "djb200_sgd_preset", "djb200_sgd_eval", "djb200_sgd_preset_name", capi.SgdData

    def coefficients(self):
        """[3, 11] float64: rhoD rhoS alpha p f0 f1 kap lambda c k theta0 per colour channel."""
        return np.array([[self._data.ch[c][f] for f in range(11)] for c in range(3)], np.float64)

    def get_fresnel(self):
        c = self.coefficients()
        return fresnel.sgd(c[:, 4].astype(np.float32), c[:, 5].astype(np.float32))

    # the per-channel terms eval is made of (dj_brdf.h:506-509, 3471-3499): vec3 per argument
    def ndf(self, h):
        return _member_call("djb200_sgd_member", C.byref(self._data), capi.MEMBER_NDF, [h, None, None], [3], 3)

    def gaf(self, h, i, o):
        return _member_call("djb200_sgd_member", C.byref(self._data), capi.MEMBER_GAF, [h, i, o], [3, 3, 3], 3)

    def g1(self, k):
        return _member_call("djb200_sgd_member", C.byref(self._data), capi.MEMBER_G1, [k, None, None], [3], 3)

    def fresnel_term(self, cos_theta_d):
        return _member_call("djb200_sgd_member", C.byref(self._data), capi.MEMBER_FRESNEL, [cos_theta_d, None, None], [1], 3)


class abc(_analytic_brdf):
    """djb::abc (dj_brdf.h:514-535): ABC-distribution BRDF of a MERL material, by name."""
    _preset, _eval, _names, _data_type = "djb200_abc_preset", "djb200_abc_eval", "djb200_abc_preset_name", capi.AbcData

    def coefficients(self):
        """[9] float64: kD[3], A[3], B, C, ior."""
        d = self._data
        return np.array([*d.kD, *d.A, d.B, d.C, d.ior], np.float64)

    def get_fresnel(self):
        return fresnel.unpolarized([np.float32(self._data.ior)] * 3)

    # dj_brdf.h:531-533, 3649-3668: ndf -> vec3, gaf -> scalar, fresnel -> vec3
    def ndf(self, h):
        return _member_call("djb200_abc_member", C.byref(self._data), capi.MEMBER_NDF, [h, None, None], [3], 3)

    def gaf(self, h, i, o):
        return _member_call("djb200_abc_member", C.byref(self._data), capi.MEMBER_GAF, [h, i, o], [3, 3, 3], 1)

    def fresnel_term(self, cos_theta_d):
        return _member_call("djb200_abc_member", C.byref(self._data), capi.MEMBER_FRESNEL, [cos_theta_d, None, None], [1], 3)


# --------------------------------------------------------------------------------------------------
def nmap2leanmap(nmap, base_roughness=1e-5, bias=0.0):
    """utils/nmap2leanmap.cpp:18-54 (bias=0) / nmap2leanmap_biased.cpp:23-63 (bias=25).
    nmap: planar uint8 [3, h, w] -> (leanmap_1, leanmap_2), planar float32 [4, h, w]."""
    b = Buf(nmap, np.uint8)
    shape = tuple(b.keep.shape)
    if len(shape) != 3 or shape[0] != 3:
        raise ValueError("nmap must be planar [3, h, w]")
    _, h, w = shape
    l1 = capi.empty_like_space(b.keep, (4, h, w), np.float32)
    l2 = capi.empty_like_space(b.keep, (4, h, w), np.float32)
    b1, b2 = Buf(l1, np.float32, True), Buf(l2, np.float32, True)
    check(capi.load().djb200_nmap_to_leanmap(b.ptr, C.c_int32(w), C.c_int32(h), C.c_float(base_roughness),
                                             C.c_float(bias), b1.ptr, b2.ptr, C.c_int(b.mem),
                                             capi.current_stream_ptr(b.mem)))
    return l1, l2


def leanmap_half_mips(leanmap, levels=0):
    """A LEAN map as the renderer consumes it: half-float RGBA with a mip pyramid (djb200_leanmap_to_half_mips).  Level 0 = what
    the reference's save_exr writes (utils/CImg.h:44940-44947); level L = 2 x 2 box filter of level L - 1.
    leanmap: planar float32 [4, h, w] -> list of float16 arrays [h_L, w_L, 4], one per level (levels <= 0: down to 1 x 1)."""
    b = Buf(leanmap, np.float32)
    shape = tuple(b.keep.shape)
    if len(shape) != 3 or shape[0] != 4:
        raise ValueError("leanmap must be planar [4, h, w]")
    _, h, w = shape
    lib = capi.load()
    n_levels = int(lib.djb200_leanmap_mip_levels(C.c_int32(w), C.c_int32(h), C.c_int32(levels)))
    texels = int(lib.djb200_leanmap_mip_texels(C.c_int32(w), C.c_int32(h), C.c_int32(levels)))
    if capi._is_torch(b.keep):
        import torch
        out = torch.empty(texels * 4, dtype=torch.float16, device=b.keep.device)
    else:
        out = np.empty(texels * 4, np.float16)
    optr = C.c_void_p(out.data_ptr() if capi._is_torch(out) else out.ctypes.data)
    check(lib.djb200_leanmap_to_half_mips(b.ptr, C.c_int32(w), C.c_int32(h), C.c_int32(levels), optr, C.c_int(b.mem),
                                          capi.current_stream_ptr(b.mem)))
    res, off, lw, lh = [], 0, w, h
    for _ in range(n_levels):
        res.append(out[off * 4:(off + lw * lh) * 4].reshape(lh, lw, 4))
        off += lw * lh
        lw, lh = max(1, lw // 2), max(1, lh // 2)
    return res


def dmap2nmap(dmap, scale=0.01):
    """utils/dmap2nmap.cpp:13-44: uint8 displacement map [h, w] -> planar uint8 normal map [3, h, w]."""
    b = Buf(dmap, np.uint8)
    shape = tuple(b.keep.shape)
    if len(shape) != 2:
        raise ValueError("dmap must be [h, w]")
    h, w = shape
    if capi._is_torch(b.keep):
        import torch
        out = torch.empty((3, h, w), dtype=torch.uint8, device=b.keep.device)
    else:
        out = np.empty((3, h, w), np.uint8)
    bo = Buf(out, np.uint8, True)
    check(capi.load().djb200_dmap_to_nmap(b.ptr, C.c_int32(w), C.c_int32(h), C.c_float(scale), bo.ptr, C.c_int(b.mem),
                                          capi.current_stream_ptr(b.mem)))
    return out


def leanmap_to_params(leanmap_1, leanmap_2, bias=0.0):
    """check_lean_maps (utils/nmap2leanmap.cpp:57-76) as a producer: per-texel lrep_to_params -> [h*w, 12]."""
    b1, b2 = Buf(leanmap_1, np.float32), Buf(leanmap_2, np.float32)
    mem = capi.same_space(b1, b2)
    _, h, w = tuple(b1.keep.shape)
    out = capi.empty_like_space(b1.keep, (h * w, 12), np.float32)
    bo = Buf(out, np.float32, True)
    check(capi.load().djb200_leanmap_to_params(b1.ptr, b2.ptr, C.c_int32(w), C.c_int32(h), C.c_float(bias), bo.ptr,
                                               C.c_int(mem), capi.current_stream_ptr(mem)))
    return out


# --------------------------------------------------------------------------------------------------
def _source_struct(src):
    s = capi.Source()
    if isinstance(src, merl):
        s.kind, s.merl = capi.SOURCE_MERL, src._h
    elif isinstance(src, utia):
        s.kind, s.utia = capi.SOURCE_UTIA, src._h
    elif isinstance(src, microfacet):
        s.kind, s.microfacet = capi.SOURCE_MICROFACET, src._desc()
    elif isinstance(src, sgd):
        s.kind, s.sgd = capi.SOURCE_SGD, C.pointer(src._data)
    elif isinstance(src, abc):
        s.kind, s.abc = capi.SOURCE_ABC, C.pointer(src._data)
    else:
        raise DjbError(6, f"cannot fit from a {type(src).__name__}")
    return s


class tabular(microfacet):
    """djb::tabular (dj_brdf.h:394-425): the isotropic power-iteration fit of any source BRDF -- and, like in the
    reference, itself a microfacet BRDF: ``eval / evalp / pdf / sample / evalp_is`` run on the fitted tables
    (normal-map sampling, since tabular does not support Smith VNDF sampling, dj_brdf.h:413).

    ``tabular(brdf, res, shadow)`` fits one material; ``tabular.fit_batch(brdfs, ...)`` fits many in
    one device pass.  ``iterations`` is 4 in the reference (dj_brdf.h:2518)."""
    _prefix = "djb200_tabular_"

    def __init__(self, source, resolution, shadow=True, iterations=4, _result=None):
        r = _result or tabular._run([source], resolution, shadow, iterations)[0]
        self.__dict__.update(r)
        self.m_shadow = bool(shadow)
        self.m_fresnel = fresnel.spline(self.m_fresnel_points)
        self._h = None

    def supports_smith_vndf_sampling(self):
        return False

    def _first_arg(self):
        if self._h is None:  # upload the tables once
            f = capi.TabularFit()
            f.res = len(self.m_p22)
            f.p22, f.sigma, f.cdf, f.qf = (self.__dict__[x].ctypes.data for x in ("m_p22", "m_sigma", "m_cdf", "m_qf"))
            f.fresnel = self.m_fresnel_points.ctypes.data
            h = C.c_void_p()
            check(capi.load().djb200_tabular_create(C.byref(f), C.c_int32(int(self.m_shadow)), C.byref(h)))
            self._h = h
        return self._h, self

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                capi.load().djb200_tabular_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def source_array(sources):
        """The C descriptor array of a list of source BRDFs; build it once when the same materials are fitted repeatedly."""
        srcs = (capi.Source * len(sources))(*[_source_struct(s) for s in sources])
        srcs._keep = list(sources)  # the handles must outlive the array
        return srcs

    @staticmethod
    def fit_packed(sources, res=90, shadow=True, iterations=4):
        """All fits of a batch in ONE packed result (djb200_fit_tabular_packed): no per-material objects, one device -> host
        copy.  `sources`: a list of BRDFs or a `source_array`.  Returns a dict of arrays with a leading material axis:
        p22, sigma, cdf, qf [n, res]; fresnel [n, res, 3]; alpha [n, 2] = (beckmann, ggx); residuals [n, iterations]."""
        if res <= 2:
            raise DjbError(1, "Invalid Resolution")  # DJB_ASSERT, dj_brdf.h:2218
        srcs = sources if isinstance(sources, C.Array) else tabular.source_array(list(sources))
        n = len(srcs)
        lib = capi.load()
        out = np.empty(n * (7 * res + 2), np.float32)
        resid = np.empty((n, max(1, iterations)), np.float32)
        check(lib.djb200_fit_tabular_packed(srcs, C.c_int32(n), C.c_int32(res), C.c_int32(int(shadow)), C.c_int32(iterations),
                                            C.c_void_p(out.ctypes.data), C.c_void_p(resid.ctypes.data), C.c_int(capi.MEM_HOST), None))
        pk = n * res
        return dict(p22=out[:pk].reshape(n, res), sigma=out[pk:2 * pk].reshape(n, res), cdf=out[2 * pk:3 * pk].reshape(n, res),
                    qf=out[3 * pk:4 * pk].reshape(n, res), fresnel=out[4 * pk:7 * pk].reshape(n, res, 3),
                    alpha=out[7 * pk:].reshape(n, 2), residuals=resid)

    @staticmethod
    def _run(sources, res, shadow, iterations):
        r = tabular.fit_packed(sources, res, shadow, iterations)
        return [dict(m_p22=r["p22"][k], m_sigma=r["sigma"][k], m_cdf=r["cdf"][k], m_qf=r["qf"][k],
                     m_fresnel_points=r["fresnel"][k], residuals=r["residuals"][k],
                     alpha_beckmann=float(r["alpha"][k, 0]), alpha_ggx=float(r["alpha"][k, 1]))
                for k in range(len(r["alpha"]))]

    @staticmethod
    def fit_batch(sources, resolution=90, shadow=True, iterations=4):
        return [tabular(None, resolution, shadow, iterations, _result=r)
                for r in tabular._run(list(sources), resolution, shadow, iterations)]

    # accessors, dj_brdf.h:404-407
    def get_p22v(self):
        return self.m_p22

    def get_sigmav(self):
        return self.m_sigma

    def get_cdfv(self):
        return self.m_cdf

    def get_qfv(self):
        return self.m_qf

    def get_fresnel(self):
        return fresnel.spline(self.m_fresnel_points)

    # dj_brdf.h:402-403
    @staticmethod
    def fit_beckmann_parameters(tab):
        return params.isotropic(tab.alpha_beckmann)

    @staticmethod
    def fit_ggx_parameters(tab):
        return params.isotropic(tab.alpha_ggx)


class tabular_anisotropic(microfacet):
    """djb::tabular_anisotropic (dj_brdf.h:428-478): eval tables + parameter fits, and an evaluable / samplable BRDF:
    ``eval / evalp / pdf`` on the elevation x azimuth tables, ``sample / evalp_is`` by normal-map sampling through the
    marginal / conditional quantile tables (built on the device when the handle is created, dj_brdf.h:2848-3103)."""
    _prefix = "djb200_tabular_"

    def __init__(self, source, elevation_res, azimuthal_res, shadow=True, iterations=4, _result=None):
        r = _result or tabular_anisotropic._run([source], elevation_res, azimuthal_res, shadow, iterations)[0]
        self.__dict__.update(r)
        self.m_elevation_res, self.m_azimuthal_res = elevation_res, azimuthal_res
        self.m_shadow = bool(shadow)
        self.m_fresnel = fresnel.spline(self.m_fresnel_points)
        self._h = None

    def supports_smith_vndf_sampling(self):
        return False

    def _first_arg(self):
        if self._h is None:
            f = capi.TabularAnisotropicFit()
            f.elev_res, f.azim_res = self.m_elevation_res, self.m_azimuthal_res
            f.p22, f.sigma, f.fresnel = self.m_p22.ctypes.data, self.m_sigma.ctypes.data, self.m_fresnel_points.ctypes.data
            h = C.c_void_p()
            check(capi.load().djb200_tabular_anisotropic_create(C.byref(f), C.c_int32(int(self.m_shadow)), C.byref(h)))
            self._h = h
        return self._h, self

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                capi.load().djb200_tabular_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def _run(sources, er, ar, shadow, iterations):
        if er <= 1 or ar <= 1:
            raise DjbError(1, "Invalid Resolution")  # dj_brdf.h:2244
        n = len(sources)
        srcs = (capi.Source * n)(*[_source_struct(s) for s in sources])
        fits = (capi.TabularAnisotropicFit * n)()
        arrays = []
        for k in range(n):
            a = dict(m_p22=np.zeros(er * ar, np.float32), m_sigma=np.zeros(er * ar, np.float32),
                     m_fresnel_points=np.zeros((er, 3), np.float32),
                     residuals=np.zeros(max(1, iterations), np.float32))
            arrays.append(a)
            f = fits[k]
            f.elev_res, f.azim_res = er, ar
            f.p22, f.sigma = a["m_p22"].ctypes.data, a["m_sigma"].ctypes.data
            f.fresnel = a["m_fresnel_points"].ctypes.data
            f.residuals = a["residuals"].ctypes.data
        check(capi.load().djb200_fit_tabular_anisotropic(srcs, C.c_int32(n), C.c_int32(er), C.c_int32(ar),
                                                         C.c_int32(int(shadow)), C.c_int32(iterations), fits, None))
        for k in range(n):
            arrays[k]["beckmann"] = np.array(list(fits[k].beckmann), np.float32)
            arrays[k]["ggx"] = np.array(list(fits[k].ggx), np.float32)
        return arrays

    def get_p22v(self):
        return self.m_p22, self.m_elevation_res, self.m_azimuthal_res

    def get_sigmav(self):
        return self.m_sigma, self.m_elevation_res, self.m_azimuthal_res

    def sampling_tables(self):
        """The tables behind pdf1 / cdf1 / qf1 / pdf2 / cdf2 / qf2 (dj_brdf.h:2766-2824) -> dict of float32 arrays
        (+ n_qf1 / n_qf2: the entries the reference's inversion searches produce)."""
        h, _ = self._first_arg()
        er, ar = self.m_elevation_res, self.m_azimuthal_res
        one = [np.zeros(ar, np.float32) for _ in range(3)]
        two = [np.zeros(er * ar, np.float32) for _ in range(3)]
        counts = (C.c_int32 * 2)()
        check(capi.load().djb200_tabular_anisotropic_sampling_tables(
            h, *[C.c_void_p(a.ctypes.data) for a in (one[0], one[1], one[2], two[0], two[1], two[2])], counts))
        return dict(pdf1=one[0], cdf1=one[1], qf1=one[2], pdf2=two[0], cdf2=two[1], qf2=two[2], n_qf1=counts[0],
                    n_qf2=counts[1])

    # the public table queries (dj_brdf.h:450-455, 2766-2824), batched
    def _table_member(self, what, a, b=None):
        h, _ = self._first_arg()
        bufs = [Buf(a, np.float32), Buf(b, np.float32) if b is not None else None]
        mem = capi.same_space(*[x for x in bufs if x is not None])
        out = capi.empty_like_space(bufs[0].keep, (bufs[0].n,), np.float32)
        bo = Buf(out, np.float32, True)
        check(capi.load().djb200_tabular_anisotropic_query(h, C.c_int32(what), bufs[0].ptr, bufs[1].ptr if bufs[1] else None,
                                                           C.c_int64(bufs[0].n), bo.ptr, C.c_int(mem),
                                                           capi.current_stream_ptr(mem)))
        return out

    def pdf1(self, phi):
        return self._table_member(capi.MEMBER_PDF1, phi)

    def cdf1(self, phi):
        return self._table_member(capi.MEMBER_CDF1, phi)

    def qf1(self, u1):
        return self._table_member(capi.MEMBER_TQF1, u1)

    def pdf2(self, theta, phi):
        return self._table_member(capi.MEMBER_PDF2, theta, phi)

    def cdf2(self, theta, phi):
        return self._table_member(capi.MEMBER_CDF2, theta, phi)

    def qf2(self, u, phi):
        return self._table_member(capi.MEMBER_TQF2, u, phi)

    @staticmethod
    def fit_beckmann_parameters(tab):
        return params.pdfparams(*[float(x) for x in tab.beckmann])

    @staticmethod
    def fit_ggx_parameters(tab):
        return params.pdfparams(*[float(x) for x in tab.ggx])
