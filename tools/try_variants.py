"""Build the experimental kernel variants next to the default library and compare them in ONE
gpurun call (tuning tool, not a benchmark).

    python tools/try_variants.py build          # here: default + one library per switch -> tools/_bin/
    python tools/try_variants.py run            # on the GPU box: parity subset + C2 probe per library

The switches are compile-time (csrc/smax_kernels.cuh, csrc/smax_scan.cu: SMAX_MINBLOCKS resident CTAs per SM,
SMAX_UNROLL_A unroll factor of the filter loop, SMAX_STATIC_EIGHTHS share of the units dealt round-robin).  tests and tools pick the library up from $SMAX_LIB."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "_bin")
VARIANTS = {
    "mb6": "-DSMAX_MINBLOCKS=6",
    "mb8": "-DSMAX_MINBLOCKS=8",
    "u2": "-DSMAX_UNROLL_A=2",
    "s3": "-DSMAX_STATIC_EIGHTHS=3",
    "s7": "-DSMAX_STATIC_EIGHTHS=7",
}
PARITY = "(golden_device_scan or few_ctas or fuzzed or uint32 or sharded or idempotent or window) and units"
PROBES = os.environ.get("SMAX_PROBES", "full,no-write,stream only")


def build():
    os.makedirs(BIN, exist_ok=True)
    lib = os.path.join(ROOT, "genometools_smax_b200", "lib", "libsmax.so")
    for name, defs in VARIANTS.items():
        env = dict(os.environ, SMAX_NVCC_DEFS=defs)
        subprocess.run([sys.executable, "-m", "genometools_smax_b200._build", "--force"], cwd=ROOT, env=env,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        shutil.copy(lib, os.path.join(BIN, "libsmax_%s.so" % name))
        print("built", name, defs)
    subprocess.run([sys.executable, "-m", "genometools_smax_b200._build", "--force"], cwd=ROOT, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)      # leave the default in place


def run(names):
    for name in names or VARIANTS:
        env = dict(os.environ, SMAX_LIB=os.path.join(BIN, "libsmax_%s.so" % name), SMAX_KERNEL="units")
        print("==", name, flush=True)
        p = subprocess.run(["timeout", "-s", "KILL", "400", sys.executable, "-m", "pytest",
                            os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-k", PARITY],
                           cwd=ROOT, env=env, capture_output=True, text=True)
        print(p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "no output (killed?)", flush=True)
        if p.returncode != 0:
            print(p.stdout[-1500:])
            continue
        p = subprocess.run(["timeout", "-s", "KILL", "240", sys.executable, os.path.join(ROOT, "tools", "probe_scan.py"),
                            "100000000", "c2", PROBES], cwd=ROOT, env=env,
                           capture_output=True, text=True)
        print("\n".join(p.stdout.strip().splitlines()[-(len(PROBES.split(",")) + 0):]), flush=True)
        if p.returncode != 0:
            print(p.stderr[-800:])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2:])
