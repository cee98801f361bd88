"""Debug helper: one fuzz table, grid limited to a few CTAs, first mismatch vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from genometools_smax_b200 import capi
from oracle import smax_oracle as O
from util import fuzz_tables
O.build_c_oracle()
kind, ctas = sys.argv[1], int(sys.argv[2])
rng = np.random.default_rng(1000 * ctas + len(kind))
n = 1_000_003
lcp, llv, bwt = fuzz_tables(rng, n, kind)
if kind in ("plateaus", "sparse"):
    lcp = lcp.copy()
    for lo in range(0, n, 200_000):
        lcp[lo + 50_000: lo + 120_000] = 0
    keep = (llv["position"] % 200_000 < 50_000) | (llv["position"] % 200_000 >= 120_000)
    llv = llv[keep]
suf = rng.permutation(n).astype(np.uint64)
idx = capi.Index.from_arrays(lcp, bwt, llv, suf)
d = capi.Device(0)
d.upload(idx, 0, None, True)
d.set_grid_limit(ctas)
d.set_stats(True)
for m in (1, 4, 12, 255, 300):
    d.scan(m, 0, True)
    recs, pos = d.fetch()
    st = d.stats()
    want = O.smax_c(lcp, llv, bwt, m)
    ok = np.array_equal(recs, want)
    print("m", m, "ok" if ok else "MISMATCH", len(recs), len(want), "slow", st["slow_tiles"], "flushes", st["flushes"])
    if not ok and len(recs) == len(want):
        bad = np.flatnonzero((recs["lb"] != want["lb"]) | (recs["len"] != want["len"]) | (recs["width"] != want["width"]))
        print("  mismatches", len(bad), "first at", bad[:5], "tiles", (want["lb"][bad[:5]] // 16384))
        i = int(bad[0])
        print("  got ", recs[i - 2:i + 4].tolist())
        print("  want", want[i - 2:i + 4].tolist())
        # is it a permutation?
        print("  same multiset:", np.array_equal(np.sort(recs, order=["lb"]), want))
