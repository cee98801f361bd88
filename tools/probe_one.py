"""Tuning probe: a few launches of one scan variant (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genometools_smax_b200 import capi
from tools import synth
from tools.esa_build_torch import build_esa
import bench
length, kind, m, flags = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
dev_t = torch.device("cuda", 0)
seq = (np.random.Generator(np.random.PCG64(1)).integers(0, 4, length, dtype=np.uint8)
       if kind == "uniform" else synth.dna_c2(length, 20001))
esa = build_esa(torch.from_numpy(seq).to(dev_t), keep_on_device=True)
n = esa["n"]
lcp, bwt, suf, llv = bench.host_window(esa, 0, n)
del esa; torch.cuda.empty_cache()
idx = bench.index_from_host(capi, lcp, bwt, suf, llv, 0, n)
dev = capi.Device(0)
dev.upload(idx, 0, n, True)
dev.set_debug(flags)
for k in range(4):
    dev.scan(m, 0, True, 0)
    print(dev.elapsed_ms(), flush=True)
    try:
        print(dev.counts() if flags == 0 else "", flush=True)
    except Exception as ex:
        print("ERR", ex, flush=True)
