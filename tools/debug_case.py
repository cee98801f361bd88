"""Re-run a saved soak failure (gpurun_out/soak_fail_<seed>.npz) in several set-ups (debug tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from genometools_smax_b200 import capi
from oracle import smax_oracle as O
O.build_c_oracle()
z = np.load(sys.argv[1])
lcp, bwt, llv, m, policy = z["lcp"], z["bwt"], z["llv"], int(z["m"]), int(z["policy"])
cuts0, limits0 = [int(c) for c in z["cuts"]], [int(c) for c in z["limits"]]
n = len(lcp)
suf = np.arange(n, dtype=np.uint64)
want = O.smax_c(lcp, llv, bwt, m, policy)
idx = capi.Index.from_arrays(lcp, bwt, llv, suf)
for kernel in ("units", "ring"):
    os.environ["SMAX_KERNEL"] = kernel
    for cuts, limits in ((cuts0, limits0), (cuts0, [0] * len(limits0)), ([0, n], [0]), ([0, n], [1])):
        devs = [capi.Device(0) for _ in range(len(cuts) - 1)]
        views = []
        for g, d in enumerate(devs):
            d.upload(idx, cuts[g], cuts[g + 1], True)
            if g:
                d.set_left_views(views[:g])
            views.append(d.view())
            d.set_grid_limit(limits[g])
        for d in devs:
            d.scan(m, policy, True)
        recs = np.concatenate([d.fetch()[0] for d in devs])
        ws = set(map(tuple, want.tolist())); gs = set(map(tuple, recs.tolist()))
        miss = sorted(ws - gs, key=lambda r: r[1]); extra = sorted(gs - ws, key=lambda r: r[1])
        ends = np.array([r[1] + r[2] - 1 for r in miss], dtype=np.int64)
        print(kernel, "cuts", cuts, "limits", limits, "got", len(recs), "want", len(want), "missing", len(miss),
              "extra", len(extra), "missing ends", (ends[:3].tolist(), ends[-3:].tolist()) if len(ends) else "", flush=True)
        for d in devs:
            d.close()
