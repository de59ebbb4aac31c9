"""ctypes binding of libdjb200.so (include/djb200.h) -- the same C-ABI a cgo/JNI/C++ host binds.

The library is built in-tree by ``dj_brdf_b200.build``.  Importing this module never falls back to
any CPU implementation: if the shared library is missing, ``load()`` raises; if there is no CUDA
device, every compute entry point returns DJB200_ERR_NO_DEVICE and ``check()`` raises ``DjbError``
(the Python face of the reference's ``djb::exc``, dj_brdf.h:54-59).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libdjb200.so"

MEM_HOST, MEM_DEVICE = 0, 1
NDF_BECKMANN, NDF_GGX = 0, 1
FRESNEL_IDEAL, FRESNEL_SCHLICK, FRESNEL_UNPOLARIZED, FRESNEL_SGD, FRESNEL_SPLINE = range(5)
PARAMS_BROADCAST, PARAMS_PER_PAIR = 0, 1
SOURCE_MERL, SOURCE_UTIA, SOURCE_MICROFACET, SOURCE_SGD, SOURCE_ABC = 0, 1, 2, 3, 4

STATUS_NAMES = {0: "OK", 1: "INVALID_ARGUMENT", 2: "NO_DEVICE", 3: "CUDA", 4: "OUT_OF_MEMORY", 5: "IO",
                6: "UNSUPPORTED"}

# every symbol include/djb200.h declares; tests check that the library exports all of them
EXPORTED_SYMBOLS = [
    "djb200_last_error", "djb200_version", "djb200_device_count", "djb200_set_device",
    "djb200_kernel_launch_count", "djb200_release_cache", "djb200_set_precision", "djb200_get_precision", "djb200_debug_force_generic", "djb200_debug_beckmann_compaction",
    "djb200_params_standard", "djb200_params_isotropic", "djb200_params_elliptic", "djb200_params_pdfparams",
    "djb200_microfacet_eval", "djb200_microfacet_evalp", "djb200_microfacet_pdf", "djb200_microfacet_sample",
    "djb200_microfacet_evalp_is", "djb200_microfacet_component",
    "djb200_io_to_hd", "djb200_hd_to_io",
    "djb200_merl_create", "djb200_merl_load", "djb200_merl_destroy", "djb200_merl_eval", "djb200_merl_index",
    "djb200_debug_merl_filter_stats",
    "djb200_utia_create", "djb200_utia_load", "djb200_utia_destroy", "djb200_utia_eval",
    "djb200_sgd_preset", "djb200_abc_preset", "djb200_preset_count", "djb200_sgd_preset_name", "djb200_abc_preset_name",
    "djb200_sgd_eval", "djb200_abc_eval",
    "djb200_nmap_to_leanmap", "djb200_leanmap_to_half_mips", "djb200_leanmap_mip_levels", "djb200_leanmap_mip_texels", "djb200_dmap_to_nmap", "djb200_lrep_to_params", "djb200_params_to_lrep", "djb200_leanmap_to_params",
    "djb200_lean_shading_params", "djb200_lean_shading_evalp", "djb200_lean_shading_pdf", "djb200_lean_shading_evalp_is",
    "djb200_debug_fit_phase_clocks", "djb200_debug_fit_parts", "djb200_debug_dmath", "djb200_fit_tabular", "djb200_fit_tabular_packed", "djb200_fit_tabular_packed_floats", "djb200_fit_tabular_anisotropic",
    "djb200_radial_query", "djb200_quantile_query", "djb200_tabular_anisotropic_query", "djb200_sgd_member", "djb200_abc_member", "djb200_tabular_create", "djb200_tabular_anisotropic_create", "djb200_tabular_anisotropic_sampling_tables", "djb200_tabular_destroy", "djb200_tabular_eval", "djb200_tabular_evalp", "djb200_tabular_pdf",
    "djb200_tabular_sample", "djb200_tabular_evalp_is",
    "djb200_aniso_fit_create", "djb200_aniso_fit_destroy", "djb200_aniso_fit_size", "djb200_aniso_fit_matvec",
    "djb200_aniso_fit_set_iterate", "djb200_aniso_fit_sigma", "djb200_aniso_fit_finish", "djb200_aniso_fit_download",
    "djb200_aniso_fit_run", "djb200_comm_unique_id", "djb200_comm_create", "djb200_comm_destroy",
]


class DjbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"djb200 error {STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Fresnel(C.Structure):
    _fields_ = [("kind", C.c_int32), ("v", C.c_float * 6), ("points", C.c_void_p), ("n_points", C.c_int32)]


class Microfacet(C.Structure):
    _fields_ = [("ndf", C.c_int32), ("shadow", C.c_int32), ("fresnel", Fresnel)]


class SgdData(C.Structure):
    _fields_ = [("ch", (C.c_double * 11) * 3)]


class AbcData(C.Structure):
    _fields_ = [("kD", C.c_double * 3), ("A", C.c_double * 3), ("B", C.c_double), ("C", C.c_double),
                ("ior", C.c_double)]


class LeanShading(C.Structure):
    _fields_ = [("bias", C.c_float), ("dmap_scale", C.c_float), ("lean_filtering", C.c_int32),
                ("alpha_per_pair", C.c_int32), ("alpha", C.c_float * 3)]


class Source(C.Structure):
    _fields_ = [("kind", C.c_int32), ("merl", C.c_void_p), ("utia", C.c_void_p), ("microfacet", Microfacet),
                ("sgd", C.POINTER(SgdData)), ("abc", C.POINTER(AbcData))]


class TabularFit(C.Structure):
    _fields_ = [("res", C.c_int32), ("p22", C.c_void_p), ("sigma", C.c_void_p), ("cdf", C.c_void_p),
                ("qf", C.c_void_p), ("fresnel", C.c_void_p), ("alpha_beckmann", C.c_float),
                ("alpha_ggx", C.c_float), ("residuals", C.c_void_p)]


class TabularAnisotropicFit(C.Structure):
    _fields_ = [("elev_res", C.c_int32), ("azim_res", C.c_int32), ("p22", C.c_void_p), ("sigma", C.c_void_p),
                ("fresnel", C.c_void_p), ("beckmann", C.c_float * 5), ("ggx", C.c_float * 5),
                ("residuals", C.c_void_p)]


_lib = None


def load():
    """Load libdjb200.so (raises if it has not been built -- there is no other implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -m dj_brdf_b200.build` (needs nvcc). "
                          "dj_brdf_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.djb200_last_error.restype = C.c_char_p
    lib.djb200_version.restype = C.c_char_p
    lib.djb200_kernel_launch_count.restype = C.c_uint64
    lib.djb200_aniso_fit_size.restype = C.c_int64
    lib.djb200_fit_tabular_packed_floats.restype = C.c_int64
    lib.djb200_leanmap_mip_texels.restype = C.c_int64
    lib.djb200_sgd_preset_name.restype = C.c_char_p
    lib.djb200_abc_preset_name.restype = C.c_char_p
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise DjbError(status, load().djb200_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int(0)
    st = load().djb200_device_count(C.byref(n))
    return n.value if st == 0 else 0


def kernel_launch_count():
    return int(load().djb200_kernel_launch_count())


PRECISION_REFERENCE_BITS, PRECISION_1E5 = 0, 1
# djb200.h: DJB200_MEMBER_*
MEMBER_QF1, MEMBER_QF2_RADIAL, MEMBER_QF3_RADIAL = 0, 1, 2
MEMBER_PDF1, MEMBER_CDF1, MEMBER_TQF1, MEMBER_PDF2, MEMBER_CDF2, MEMBER_TQF2 = 10, 11, 12, 13, 14, 15
MEMBER_NDF, MEMBER_GAF, MEMBER_G1, MEMBER_FRESNEL = 20, 21, 22, 23


def set_precision(mode):
    """Precision of microfacet eval / evalp / pdf / sample (djb200_set_precision): "1e-5" (default: eval / pdf within 1e-5 relative of
    the reference's floats with the identical zero pattern; sampled directions within 1e-5 except where Beckmann's quantile search
    stops one trip apart from the reference's -- include/djb200.h has the measured figures; ~2x faster) or "bits" (the reference's
    rounded floats)."""
    code = {"bits": PRECISION_REFERENCE_BITS, "1e-5": PRECISION_1E5, 0: 0, 1: 1}[mode]
    check(load().djb200_set_precision(C.c_int(code)))


def get_precision():
    return "1e-5" if int(load().djb200_get_precision()) == PRECISION_1E5 else "bits"


# ---- buffer plumbing: numpy arrays are host memory, torch CUDA tensors are device memory --------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Buf:
    """A bulk array handed to the C-ABI: pointer + memory space (+ the object keeping it alive)."""

    def __init__(self, obj, dtype, writable=False):
        if obj is None:
            self.ptr, self.mem, self.keep, self.n = None, None, None, 0
            return
        if _is_torch(obj):
            import torch
            want = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8,
                    np.float64: torch.float64}[dtype]
            t = obj
            if t.dtype != want or not t.is_contiguous():
                if writable:
                    raise ValueError("output tensor must be contiguous and of the right dtype")
                t = t.to(want).contiguous()
            self.keep = t
            self.ptr = C.c_void_p(t.data_ptr())
            self.mem = MEM_DEVICE if t.is_cuda else MEM_HOST
            self.n = t.numel()
        else:
            a = obj
            if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous):
                if writable:
                    raise ValueError("output array must be a C-contiguous ndarray of the right dtype")
                a = np.ascontiguousarray(a, dtype=dtype)
            self.keep = a
            self.ptr = C.c_void_p(a.ctypes.data)
            self.mem = MEM_HOST
            self.n = a.size


def current_stream_ptr(mem):
    if mem != MEM_DEVICE:
        return None
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def same_space(*bufs):
    mems = {b.mem for b in bufs if b.mem is not None}
    if len(mems) != 1:
        raise ValueError("all bulk arrays of one call must live in the same memory space")
    return mems.pop()


def empty_like_space(ref_obj, shape, dtype):
    """Allocate an output next to the inputs (numpy -> numpy, cuda tensor -> cuda tensor)."""
    if _is_torch(ref_obj):
        import torch
        tdt = {np.float32: torch.float32, np.int32: torch.int32}[dtype]
        return torch.empty(shape, dtype=tdt, device=ref_obj.device)
    return np.empty(shape, dtype=dtype)
