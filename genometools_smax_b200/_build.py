"""In-tree build of libsmax.so and the `smax` executable for sm_100a.

Explicit nvcc/gcc commands (no JIT cache): the built files land in
``genometools_smax_b200/lib/`` and travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
              "-I", INCLUDE, "-I", CSRC]
# experiments: extra nvcc definitions, e.g. SMAX_NVCC_DEFS="-DSMAX_X=1" (default: none)
NVCC_FLAGS += [d for d in os.environ.get("SMAX_NVCC_DEFS", "").split() if d.startswith("-D")]
GCC_FLAGS = ["-O2", "-std=gnu99", "-fPIC", "-Wall", "-Wextra", "-I", INCLUDE, "-I", CSRC]

CU_SOURCES = ["smax_scan.cu", "smax_ring.cu", "smax_device.cu", "smax_format.cu"]
C_SOURCES = ["smax_index.c", "smax_run.c", "smax_emit.c", "smax_tool.c", "smax_stream.c"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA/C source and link lib/libsmax.so + lib/smax."""
    os.makedirs(LIB, exist_ok=True)
    headers = [os.path.join(INCLUDE, "smax.h"), os.path.join(CSRC, "smax_host.h"),
               os.path.join(CSRC, "smax_kernels.cuh"), os.path.join(CSRC, "smax_ring.cuh"),
               os.path.join(CSRC, "smax_dec.h"),
               os.path.join(CSRC, "smax_swar.h")]
    nvcc = _nvcc()
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIB, src + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.check_call(cmd)
        objs.append(o)
    for src in C_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIB, src + ".o")
        if force or _stale(o, [s] + headers):
            subprocess.check_call(["gcc"] + GCC_FLAGS + ["-c", s, "-o", o])
        objs.append(o)
    so = os.path.join(LIB, "libsmax.so")
    if force or _stale(so, objs):
        subprocess.check_call([nvcc, "-shared", "-o", so] + objs +
                              ["-cudart", "static", "-lpthread", "-ldl", "-lrt"])
    exe = os.path.join(LIB, "smax")
    main_c = os.path.join(CSRC, "smax_main.c")
    if force or _stale(exe, [so, main_c]):
        subprocess.check_call(["gcc"] + GCC_FLAGS + [main_c, "-o", exe, "-L", LIB, "-lsmax",
                                                     "-Wl,-rpath,$ORIGIN"])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
