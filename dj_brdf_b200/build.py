"""Build libdjb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m dj_brdf_b200.build [--force]

Flags that are part of the numerical contract (DESIGN.md "Numerics"):
  -fmad=false                 no FMA contraction on the device (the reference has none)
  -Xcompiler -ffp-contract=off   same for the host-side params factories
  (never -use_fast_math; IEEE -prec-div / -prec-sqrt are nvcc's defaults)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "build"
LIB = PKG / "libdjb200.so"
SOURCES = ["capi.cu", "capi_fit.cu", "kernels_mf.cu", "kernels_tables.cu", "kernels_merl.cu", "kernels_fit.cu", "kernels_tabular.cu",
           "kernels_analytic.cu", "presets.cu"]
HEADERS = ["djb_presets.inc", "djb_device.cuh", "djb_lean.cuh", "djb_fit.cuh", "djb_internal.h", "djb_dmath.cuh", "djb_dmath_tables.inc",
           "djb_glibcf.h", "../../include/djb200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
    *os.environ.get("DJB200_NVCC_EXTRA", "").split(),  # experiments only (e.g. -DDJB200_COMPACT_MINB=4)
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    hdrs = [(CSRC / h).resolve() for h in HEADERS]
    BUILD.mkdir(exist_ok=True)
    objs = []
    jobs = []
    for s in srcs:
        o = BUILD / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s, *hdrs, Path(__file__)]):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(compile_one, jobs))
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
