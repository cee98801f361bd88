"""Tuning probe (not a benchmark): how long does each unit of the unit kernel take, and which
ones are the slow ones?  (debug flag 1024: every warp stamps its units with %globaltimer)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genometools_smax_b200 import capi
from tools import synth
from tools.esa_build_torch import build_esa
import bench

length = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "c2"
dev_t = torch.device("cuda", 0)
cfg = synth.WORKLOADS[kind.upper()]
M = cfg["minlength"]
codes = torch.from_numpy(cfg["gen"](length, cfg["seed"])).to(dev_t)
if cfg["mirrored"]:
    from tools.esa_build_torch import mirror_codes
    codes = mirror_codes(codes)
esa = build_esa(codes, keep_on_device=True)
del codes
n = esa["n"]
lcp, bwt, suf, llv = bench.host_window(esa, 0, n)
del esa; torch.cuda.empty_cache()
idx = bench.index_from_host(capi, lcp, bwt, suf, llv, 0, n)
dev = capi.Device(0)
dev.upload(idx, 0, n, True)
L = capi.lib()
L.smax_device_debug_meta.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
nunits = (n + 4095) // 4096
meta = np.zeros(nunits, dtype=np.dtype([("count", "<u4"), ("ns", "<u4"), ("wsum", "<u8"), ("t0", "<u8")]))
dev.set_debug(1024 | 128)
for k in range(4):
    dev.scan(M, 0, True, 0)
    ms, _, _ = dev.elapsed_ms()
assert L.smax_device_debug_meta(dev.handle, meta.ctypes.data_as(ctypes.c_void_p), nunits) == 0
dev.set_debug(0)
ns = meta["ns"].astype(np.int64)
t0 = meta["t0"].astype(np.int64)
start = t0.min()
end = (t0 + ns).max()
print("scan %.1f us by events; units: %d  span %.1f us  mean %.2f us  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f" % (
    ms * 1e3, nunits, (end - start) / 1e3, ns.mean() / 1e3, *(np.percentile(ns, [50, 90, 99]) / 1e3), ns.max() / 1e3))
nllv = np.bincount((llv.numpy()[:, 0] // 4096).astype(np.int64), minlength=nunits)[:nunits]
lcp_u = lcp.numpy()
order = np.argsort(-ns)[:12]
print("slowest units:   unit      ns   begins_at_us  ends_at_us  records  llv  bytes>=m   us by phase: large filter K2large phaseB K3head K3entries")
for u in order:
    seg = lcp_u[u * 4096:(u + 1) * 4096]
    ph = int(meta["wsum"][u])
    print("              %7d %7d %10.1f %10.1f %7d %5d %6d      %s" % (u, ns[u], (t0[u] - start) / 1e3, (t0[u] + ns[u] - start) / 1e3,
                                                      meta["count"][u], nllv[u], int((seg >= M).sum()),
                                                      " ".join("%3d" % ((ph >> (8 * i)) & 255) for i in range(6))))
fin = np.sort(t0 + ns - start) / 1e3
print("finish times of the last units (us):", np.round(fin[-8:], 1), " 99%% of the units are done by %.1f us" % fin[int(0.99 * nunits)])
res = (ctypes.c_uint64 * 20)()
L.smax_device_debug_result.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
if L.smax_device_debug_result(dev.handle, res, 20) == 20:
    ph = np.array(list(res)[12:20], dtype=np.float64)
    names = ["large pass (+feed, dir, ticket)", "wait for the lcp bytes", "phase A (filter)", "K2 of large values",
             "phase B (+ENDs)", "K3 head (meta, sums)", "K3 entries / no survivors", "-"]
    print("warp time by phase (sum over all warps, %% of %.1f ms):" % (ph.sum() / 1e6))
    for nm, v in zip(names, ph):
        print("   %-34s %6.1f %%   %.2f us per unit" % (nm, 100 * v / ph.sum(), v / nunits / 1e3))
# least squares: what a unit costs by its contents
pad = np.zeros(nunits * 4096, dtype=np.uint8); pad[:n] = lcp_u[:n]
big = (pad.reshape(nunits, 4096) >= min(M, 255)).sum(1)
A = np.stack([np.ones(nunits), nllv, big, meta["count"]], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, ns.astype(np.float64), rcond=None)
print("ns per unit ~ %.0f + %.2f * llv + %.2f * bytes>=m + %.1f * records" % tuple(coef))
