import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from conftest import Golden
from genometools_smax_b200 import capi
from oracle import smax_oracle as O
O.build_c_oracle()
dev = capi.Device(0)
for name in ("atinsert", "random", "wide"):
    g = Golden(name); t = g.tables()
    for m in g.minlengths:
        for policy in (0, 1):
            idx = capi.Index.from_arrays(t.lcp, t.bwt, t.llv, t.suf)
            dev.upload(idx, 0, None, True); dev.scan(m, policy, True); recs, pos = dev.fetch(); idx.close()
            want = O.smax_c(t.lcp, t.llv, t.bwt, m, policy)
            ok = np.array_equal(recs, want)
            print(name, m, policy, "ok" if ok else "MISMATCH got %d want %d" % (len(recs), len(want)))
            if not ok:
                gs = set(map(tuple, recs.tolist())); ws = set(map(tuple, want.tolist()))
                print("  missing", sorted(ws - gs)[:6]); print("  extra", sorted(gs - ws)[:6])
                for (l, lb, w) in sorted(ws - gs)[:3] + sorted(gs - ws)[:3]:
                    print("   lcp", t.lcp[max(0,lb-2):lb+w+2].tolist(), "bwt", t.bwt[max(0,lb-2):lb+w+2].tolist(), "lb%16", lb % 16)
