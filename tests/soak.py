"""Randomised soak of the CUDA path against the C oracle (run on a B200):
random table kinds / sizes / grid limits / shard counts / minimum lengths / policies.
    python tests/soak.py [seconds] [seed]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from genometools_smax_b200 import capi
from oracle import smax_oracle as O
from util import fuzz_tables

O.build_c_oracle()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
kinds = ["dense", "alternating", "plateaus", "large", "huge", "widerun", "sparse"]
t0, cases, scans = time.time(), 0, 0
while time.time() - t0 < budget:
    kind = kinds[int(rng.integers(len(kinds)))]
    n = int(rng.choice([int(rng.integers(1, 40000)), int(rng.integers(40000, 400000)),
                        int(rng.integers(400000, 3000000))]))
    lcp, llv, bwt = fuzz_tables(rng, n, kind)
    if rng.random() < 0.5 and n > 100000:            # quiet stretches between busy ones
        lcp = lcp.copy()
        period = int(rng.integers(60000, 400000))
        a, b = sorted(rng.integers(0, period, 2))
        pos = np.arange(n) % period
        lcp[(pos >= a) & (pos < b)] = 0
        llv = llv[~((llv["position"] % period >= a) & (llv["position"] % period < b))]
    suf = rng.permutation(n).astype(np.uint64)
    nshards = int(rng.choice([1, 1, 2, 3, 5]))
    kernel = ["ring", "units", ""][int(rng.integers(3))]       # "": the device manager picks
    if kernel:
        os.environ["SMAX_KERNEL"] = kernel
    else:
        os.environ.pop("SMAX_KERNEL", None)
    idx = capi.Index.from_arrays(lcp, bwt, llv, suf)
    cuts = [((n // nshards) * g) & ~15 for g in range(nshards)] + [n]
    devs = [capi.Device(0) for _ in range(nshards)]
    try:
        views, limits = [], []
        for g, d in enumerate(devs):
            d.upload(idx, cuts[g], cuts[g + 1], True)
            if g:
                d.set_left_views(views[:g])
            views.append(d.view())
            limits.append(int(rng.choice([0, 1, 2, 3, 7, 40])))
            d.set_grid_limit(limits[-1])
        for m in rng.choice([1, 2, 3, 5, 8, 13, 20, 254, 255, 256, 300, 1 << 33], 3, replace=False):
            policy = int(rng.integers(2))
            for d in devs:
                d.scan(int(m), policy, True)
            parts = [d.fetch() for d in devs]
            recs = np.concatenate([p[0] for p in parts]); pos = np.concatenate([p[1] for p in parts])
            want = O.smax_c(lcp, llv, bwt, int(m), policy)
            if not (np.array_equal(recs, want) and np.array_equal(pos, O.positions_c(suf, want))):
                print("MISMATCH kernel=%s kind=%s n=%d shards=%d m=%d policy=%d seed=%d case=%d got=%d want=%d limits=%s cuts=%s"
                      % (kernel or "auto", kind, n, nshards, m, policy, seed, cases, len(recs), len(want), limits, cuts), flush=True)
                if len(recs) == len(want):
                    bad = np.flatnonzero((recs["lb"] != want["lb"]) | (recs["len"] != want["len"]) |
                                         (recs["width"] != want["width"]))
                    print("  record mismatches:", len(bad), "first", bad[:3], "lb", want["lb"][bad[:3]],
                          "got", recs[bad[:3]].tolist(), "want", want[bad[:3]].tolist())
                    print("  positions equal:", np.array_equal(pos, O.positions_c(suf, want)))
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                np.savez_compressed(os.path.join(ROOT, "gpurun_out", "soak_fail_%d.npz" % seed), lcp=lcp, bwt=bwt,
                                    llv=llv, recs=recs, m=int(m), policy=policy, cuts=np.array(cuts),
                                    limits=np.array(limits))
                sys.exit(1)
            scans += 1
    finally:
        for d in devs:
            d.close()
        idx.close()
    cases += 1
print("soak ok: %d tables, %d scans in %.0f s (seed %d)" % (cases, scans, time.time() - t0, seed))
