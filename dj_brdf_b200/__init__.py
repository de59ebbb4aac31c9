"""dj_brdf_b200 -- B200-native (sm_100a) microfacet BRDF evaluation and fitting engine.

Drop-in for the numeric hot paths of jdupuy/dj_brdf: batched GGX/Beckmann eval / pdf / sample /
evalp_is, MERL and UTIA table lookups, the power-iteration fits, and normal-map -> LEAN-map
conversion.  The product is ``libdjb200.so`` (hand-written CUDA behind the C-ABI declared in
``include/djb200.h``); this package is the Python host mirror of the reference's ``djb::`` interface.
There is no CPU fallback.
"""
from .capi import DjbError, device_count, get_precision, kernel_launch_count, load, set_precision  # noqa: F401
from .brdf import (abc, beckmann, brdf, dmap2nmap, fresnel, ggx, leanmap_half_mips, leanmap_to_params, merl, merl_filter_stats, microfacet, nmap2leanmap,  # noqa: F401
                   params, sgd, tabular, tabular_anisotropic, utia)

__all__ = ["DjbError", "device_count", "get_precision", "set_precision", "kernel_launch_count", "load", "abc", "sgd", "dmap2nmap", "beckmann", "brdf", "fresnel", "ggx",
           "leanmap_half_mips", "leanmap_to_params", "merl", "merl_filter_stats", "microfacet", "nmap2leanmap", "params", "tabular", "tabular_anisotropic",
           "utia"]
