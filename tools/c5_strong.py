"""STRONG scaling of the scan on ONE fixed index (config C5 of BASELINE.json: synthetic
human-genome-scale DNA, 64-bit suffix table, SA range sharded over 2 / 4 / 8 B200 with P2P
halo views), with the parity check, in ONE process and ONE gpurun call: the index is built once
(tools/esa_build_torch.py on cuda:0), kept in pinned host memory, and scanned over N = 1, 2, 4,
8 devices through the same C-ABI device calls bench.py and smax_run use (one capi.Device per
shard, cost-balanced cuts, left views through peer access, one-sided count exchange).

    python tools/c5_strong.py [--length 3000000000] [--ns 1,2,4,8] [--steps 10] [--workload C5]

Prints one JSON line per N (the shape of bench.py's line; `scaling` = "strong").  Timing: CUDA
events on every device's launching stream (inside the C ABI), max over the devices per step,
L2 flushed between steps; wall clock of the step next to it.  Parity: records + positions of all
shards, concatenated, against the C oracle run once over the whole index.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C5")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--ns", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    from genometools_smax_b200 import capi
    from genometools_smax_b200.shard import balanced_cuts
    from oracle import smax_oracle as O
    from tools import synth
    from tools.esa_build_torch import build_esa, mirror_codes

    cfg = synth.WORKLOADS[args.workload]
    length = args.length or cfg["length"]
    minlength = cfg["minlength"]
    ngpu = torch.cuda.device_count()
    ns = [int(x) for x in args.ns.split(",") if int(x) <= ngpu]
    t0 = time.perf_counter()
    seq = cfg["gen"](length, cfg["seed"])
    t_gen = time.perf_counter() - t0
    dev0 = torch.device("cuda", 0)
    t0 = time.perf_counter()
    codes = torch.from_numpy(seq).to(dev0)
    if cfg["mirrored"]:
        codes = mirror_codes(codes)
    esa = build_esa(codes, keep_on_device=True, verbose=False)
    torch.cuda.synchronize(dev0)
    del codes, seq
    t_build = time.perf_counter() - t0
    n = esa["n"]
    nllv = int(esa["llv_pos"].shape[0])
    print("# index: %d suffixes, %d large values, maxlcp %d; generated in %.1f s, built in %.1f s"
          % (n, nllv, esa["maxlcp"], t_gen, t_build), file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    lcp_h, bwt_h, suf_h, llv_h = bench.host_window(esa, 0, n)
    llv_pos_h = llv_h[:, 0].contiguous()
    del esa
    torch.cuda.empty_cache()
    idx = bench.index_from_host(capi, lcp_h, bwt_h, suf_h, llv_h, 0, n)
    print("# tables in pinned host memory after %.1f s" % (time.perf_counter() - t0), file=sys.stderr, flush=True)

    want = wpos = None
    if not args.no_check:
        t0 = time.perf_counter()
        llv = np.zeros(nllv, dtype=O.LLV_DTYPE)
        llv["position"] = llv_h[:, 0].numpy()
        llv["value"] = llv_h[:, 1].numpy()
        want = O.smax_c(lcp_h.numpy(), llv, bwt_h.numpy(), minlength)
        wpos = O.positions_c(suf_h.numpy().view(np.uint64), want)
        del llv
        print("# oracle: %d records, %d positions in %.1f s" % (len(want), len(wpos), time.perf_counter() - t0),
              file=sys.stderr, flush=True)

    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    for N in ns:
        cuts = balanced_cuts(n, N, llv_pos_h)
        devs = [capi.Device(g) for g in range(N)]
        flush = [torch.empty(512 << 20, dtype=torch.uint8, device=torch.device("cuda", g)) for g in range(N)]
        try:
            views = []
            t0 = time.perf_counter()
            for g, d in enumerate(devs):
                d.upload(idx, cuts[g], cuts[g + 1], True)
                d.set_left_views(views[max(0, g - 8):g])
                views.append(d.view())
            t_up = time.perf_counter() - t0
            if N > 1:
                ptrs = [d.counts_export(N)[1] for d in devs]
                for g, d in enumerate(devs):
                    d.counts_connect(g, N, ptrs=ptrs)
            # algorithmic bytes per shard (stats build, untimed)
            alg = []
            for g, d in enumerate(devs):
                d.set_stats(True)
                if N > 1:
                    d.set_exchange_tag(1)
                d.scan(minlength, 0, True)
                st = d.stats()
                d.set_stats(False)
                nsh = cuts[g + 1] - cuts[g]
                k0, k1 = np.searchsorted(llv_pos_h.numpy(), [cuts[g], cuts[g + 1]])
                alg.append(nsh + st["candidate_width"] + 16 * min(st["llv_inspected"], int(k1 - k0))
                           + 16 * st["survivor_width"] + 24 * st["survivors"])
            ms_dev, ms_wall, launches = [], [], 1
            for step in range(args.warmup + args.steps):
                for g in range(N):
                    flush[g].fill_(step & 0xff)
                for g in range(N):
                    torch.cuda.synchronize(torch.device("cuda", g))
                t0 = time.perf_counter()
                for g, d in enumerate(devs):
                    if N > 1:
                        d.set_exchange_tag(step + 2)
                    d.scan(minlength, 0, True)
                per = []
                for d in devs:
                    ms, _, launches = d.elapsed_ms()          # waits for the device
                    per.append(ms)
                if step >= args.warmup:
                    ms_wall.append((time.perf_counter() - t0) * 1e3)
                    ms_dev.append(per)
            ms_dev = np.array(ms_dev)
            per_dev = ms_dev.mean(0)
            ms_step = float(ms_dev.max(1).mean())
            parity = "not checked"
            nrec = 0
            parts = [d.fetch() for d in devs]
            nrec = sum(len(p[0]) for p in parts)
            if want is not None:
                recs = np.concatenate([p[0] for p in parts])
                pos = np.concatenate([p[1] for p in parts])
                parity = "ok" if (np.array_equal(recs, want) and np.array_equal(pos, wpos)) else "FAIL"
                del recs, pos
            if N > 1:
                counts = devs[-1].peer_counts(args.warmup + args.steps + 1, N)
                if counts != [len(p[0]) for p in parts]:
                    parity = "FAIL (count exchange)"
            del parts
            slow = int(per_dev.argmax())
            line = {"metric": "suffixes scanned/sec", "value": n / (ms_step * 1e-3) / 1e9,
                    "unit": "G suffixes/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                    "config": {"workload": "%s: %s, %d suffixes total, ONE index, minlength=%d, suftab 64-bit, "
                                           "policy gt" % (args.workload, cfg["gen"].__name__, n, minlength),
                               "driver": "tools/c5_strong.py: one process, one capi.Device per shard, "
                                         "cost-balanced cuts, P2P left views, one-sided count exchange",
                               "l2": "flushed between steps", "cuts": [int(c) for c in cuts],
                               "largelcpvalues": nllv, "records": int(nrec)},
                    "parity": parity,
                    "parity_how": "records + positions of all shards, concatenated, == C oracle over the whole index",
                    "per_device": [{"device": g, "suffixes": int(cuts[g + 1] - cuts[g]), "kernel_ms": float(per_dev[g]),
                                    "algorithmic_bytes": int(alg[g]),
                                    "gbs": alg[g] / (per_dev[g] * 1e-3) / 1e9} for g in range(N)],
                    "roofline": {"bound": "hbm", "achieved": alg[slow] / (per_dev[slow] * 1e-3) / 1e9, "peak": peak,
                                 "unit": "GB/s", "frac": alg[slow] / (per_dev[slow] * 1e-3) / 1e9 / peak,
                                 "traffic": None, "note": "slowest shard (device %d)" % slow},
                    "gpu_launches": args.steps * int(launches) * N,
                    "timing": {"wall_ms_per_step": float(np.mean(ms_wall)), "upload_s": t_up,
                               "index_build_s": t_build, "sequence_gen_s": t_gen}}
            print(json.dumps(line), flush=True)
        finally:
            for d in devs:
                d.close()
            del flush
            torch.cuda.empty_cache()
    idx.close()


if __name__ == "__main__":
    main()
