"""Tuning tool: build libsmax variants with different -D switches of smax_kernels.cu
into genometools_smax_b200/lib/variants/ (tools/probe_scan.py loads one via SMAX_LIB)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genometools_smax_b200 import _build as B

VARIANTS = {
    "base": [],
    "nocs": ["-DSMAX_NO_COMPACT_SMALL"],
    "cl": ["-DSMAX_COMPACT_LLV"],
    "outl": ["-DSMAX_OUTLINE_PASS"],
}

def main():
    B.build()
    out = os.path.join(B.LIB, "variants")
    os.makedirs(out, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    for name in names:
        o = os.path.join(out, "smax_kernels_%s.o" % name)
        subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + VARIANTS[name] +
                              ["-c", os.path.join(B.CSRC, "smax_kernels.cu"), "-o", o])
        objs = [o] + [os.path.join(B.LIB, s + ".o") for s in ["smax_device.cu"] + B.C_SOURCES]
        so = os.path.join(out, "libsmax_%s.so" % name)
        subprocess.check_call([B._nvcc(), "-shared", "-o", so] + objs +
                              ["-cudart", "static", "-lpthread", "-ldl", "-lrt"])
        print(so)

if __name__ == "__main__":
    main()
