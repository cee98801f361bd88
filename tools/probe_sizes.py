"""Tuning probe: scan time vs table size (synthetic byte tables, no ESA build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genometools_smax_b200 import capi
dev_t = torch.device("cuda", 0)
dev = capi.Device(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev_t)
for n in (16384, 1_000_000, 10_000_000, 100_000_000, 400_000_000, 1_600_000_000):
    # sparse table: lcp bytes like random DNA (mean 13), no large values
    g = torch.Generator(device=dev_t); g.manual_seed(1)
    lcp = (torch.rand(n + 64, device=dev_t, generator=g) * 6 + 10).to(torch.uint8)
    lcp[0] = 0; lcp[n:] = 0
    bwt = torch.randint(0, 4, (n + 64,), device=dev_t, dtype=torch.uint8, generator=g)
    suf = torch.arange(n + 8, device=dev_t, dtype=torch.int64)
    dev.adopt(lcp.data_ptr(), bwt.data_ptr(), 0, 0, suf.data_ptr(), 8, 0, n, 0, n, n, keep=(lcp, bwt, suf))
    for name, flags in (("full", 0), ("stream-only", 3)):
        dev.set_debug(flags)
        ts = []
        for k in range(13):
            flush.fill_(k)
            dev.scan(20, 0, True, 0)
            ts.append(dev.elapsed_ms()[0])
        ts = sorted(ts[3:])
        med = ts[len(ts) // 2]
        print("n=%-11d %-12s median %8.1f us  min %8.1f us  -> %6.0f GB/s" % (n, name, med * 1e3, ts[0] * 1e3, n / (med * 1e-3) / 1e9), flush=True)
    dev.set_debug(0)
    del lcp, bwt, suf
