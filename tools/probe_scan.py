"""Tuning probe (not a benchmark): time the scan kernel with parts switched off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genometools_smax_b200 import capi
if os.environ.get("SMAX_LIB"):
    capi.LIB_PATH = os.environ["SMAX_LIB"]
from tools import synth
from tools.esa_build_torch import build_esa
import bench

length = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "c2"
dev_t = torch.device("cuda", 0)
MINLEN = 20
if kind == "uniform":
    seq = np.random.Generator(np.random.PCG64(1)).integers(0, 4, length, dtype=np.uint8)
    codes = torch.from_numpy(seq).to(dev_t)
elif kind.upper() in synth.WORKLOADS:
    cfg = synth.WORKLOADS[kind.upper()]
    MINLEN = cfg["minlength"]
    codes = torch.from_numpy(cfg["gen"](length, cfg["seed"])).to(dev_t)
    if cfg["mirrored"]:
        from tools.esa_build_torch import mirror_codes
        codes = mirror_codes(codes)
else:
    codes = torch.from_numpy(synth.dna_c2(length, 20001)).to(dev_t)
esa = build_esa(codes, keep_on_device=True)
del codes
n = esa["n"]
lcp, bwt, suf, llv = bench.host_window(esa, 0, n)
if esa["llv_pos"].numel():
    per_tile = torch.bincount((esa["llv_pos"] // 16384).to(torch.int64))
    q = torch.quantile(per_tile.float(), torch.tensor([0.5, 0.9, 0.99, 0.999], device=per_tile.device))
    print(".llv records per 16 KiB tile: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  p99.9 %.0f  max %d  tiles>1024: %.1f%%" % (
        per_tile.float().mean(), q[0], q[1], q[2], q[3], int(per_tile.max()),
        100.0 * float((per_tile > 1024).float().mean())), flush=True)
del esa; torch.cuda.empty_cache()
idx = bench.index_from_host(capi, lcp, bwt, suf, llv, 0, n)
dev = capi.Device(0)
dev.upload(idx, 0, n, True)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev_t)
flush2 = torch.zeros(512 << 20, dtype=torch.uint8, device=dev_t)
M = MINLEN
variants = (("full", 0, M), ("no-write", 16, M), ("no-arena-store", 32, M), ("no-entry-loop", 64, M), ("no-emit", 128, M), ("no-biglist", 256 | 32, M), ("no-width", 512 | 32, M), ("no-both", 256 | 512 | 32, M), ("stream only", 2 | 4 | 16, M),
            ("filter only", 4 | 8 | 16, M), ("small only", 4 | 16, M),
            ("large only", 2 | 16, M), ("no-large", 4, M), ("no-small", 2, M),
            ("full m=255", 0, 255), ("full m=14", 0, 14), ("full m=8", 0, 8))
if len(sys.argv) > 3:
    variants = [v for v in variants if v[0] in sys.argv[3].split(",")]
for name, flags, m in variants:
    dev.set_debug(flags)
    ts = []
    for k in range(13):
        flush.fill_(k)
        if os.environ.get("PROBE_FLUSH", "write") == "read":
            flush2.view(torch.int64).sum()      # evict the dirty lines with clean ones
        dev.scan(m, 0, True, 0)
        ms, _, _ = dev.elapsed_ms()
        ts.append(ms)
    ts = sorted(ts[3:])
    st = dev.stats()
    print("%-14s n=%d m=%d  median %.1f us  min %.1f us  -> %.0f GB/s lcp-only" % (
        name, n, m, 1e3 * ts[len(ts) // 2], 1e3 * ts[0], n / (ts[len(ts) // 2] * 1e-3) / 1e9),
        flush=True)
dev.set_debug(0)
