"""Tuning probe (not a benchmark): where does the time of the plug-in call go when one process
drives several GPUs?  Builds one C2 index, keeps it in pinned host memory, times
smax_run + host emitter and smax_run_records for ngpus = 1, 2, ... and prints the per-shard
time line of one call each (SMAX_TRACE)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genometools_smax_b200 import capi
from tools import synth
from tools.esa_build_torch import build_esa
import bench

length = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
cfg = synth.WORKLOADS[(sys.argv[2] if len(sys.argv) > 2 else "C2").upper()]
codes = torch.from_numpy(cfg["gen"](length, cfg["seed"])).to("cuda:0")
esa = build_esa(codes, keep_on_device=True)
del codes
n = esa["n"]
full = bench.host_window(esa, 0, n)
del esa
torch.cuda.empty_cache()
idx = bench.index_from_host(capi, *full, 0, n)
M = cfg["minlength"]
for ngpus in [g for g in (1, 2, 4, 8) if g <= torch.cuda.device_count()]:
    for name, fn in (("records only", lambda: idx.run_records(M, ngpus=ngpus)),
                     ("host emitter", lambda: idx.run_emit_text(M, ngpus=ngpus, discard=True))):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(8):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        print("ngpus %d  %-13s  %.2f ms  (min %.2f)  -> %.1f G suffixes/s" % (
            ngpus, name, 1e3 * sum(ts) / len(ts), 1e3 * min(ts), n / (sum(ts) / len(ts)) / 1e9), flush=True)
    os.environ["SMAX_TRACE"] = "1"
    idx.run_emit_text(M, ngpus=ngpus, discard=True)
    del os.environ["SMAX_TRACE"]
