#!/usr/bin/env python
"""Benchmark of the supermaximal-repeat scan (BASELINE.json metric:
G suffixes scanned / s and achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm
    python bench.py --impl reference [...]                         # reference CPU arm
    torchrun --nproc-per-node N bench.py --gpus N ...              # N > 1

A step = one pass of the hot path over the whole resident index: the fused
scan kernel (plateau detection + .llv resolution + left-distinctness +
ordered compaction with the position gather fused into its ordered write), plus -- for N > 1
-- the NCCL exchange of the shard record counts.  Workload at N = 1: config C2
of BASELINE.json (synthetic DNA 100 Mbp with injected tandem / interspersed
repeats, minlength 20); at N > 1 the index grows with N (weak scaling: one
C2-sized shard per GPU of ONE index over N x 100 Mbp).  The index is built on
the box by tools/esa_build_torch.py (bit-identical to the reference
suffixerator's tables on every golden fixture; construction is out of scope
and timed separately).

`value`  : resident tables, CUDA events on the launching stream, L2 flushed
           between steps, max over ranks.
`e2e`    : same metric through the C-ABI upload+scan+fetch calls with HOST
           (pinned) tables: H2D of lcp/bwt/llv, scan, D2H of the records, host
           gather of the positions from the host suffix table; wall clock.
`roofline`: k_scan alone, algorithmic bytes / its event duration.
`cpu_baseline`: reference code (oracle/_ref/gtref smax-lin, 1 thread) on a
           bounded sample of the same workload, or the C port if absent.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smax", choices=["smax", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--length", type=int, default=0, help="per-GPU sequence length (0 = config)")
    ap.add_argument("--minlength", type=int, default=0)
    ap.add_argument("--sample", type=int, default=10_000_000,
                    help="sequence length of the CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-emit", action="store_true", help="skip the device-side text formatting figures")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare with the C oracle (slow)")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions, polled through NVML
    every millisecond (nvidia-smi -lms cannot resolve a timed region of a few ms)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []          # (t, sm_mhz, reasons bitmask)
        self.windows = []          # [t0, t1] of the timed regions
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.perf_counter(), float(mhz), int(reasons)))
            except Exception:
                pass
            time.sleep(0.001)

    def open_window(self):
        self.windows.append([time.perf_counter(), None])

    def close_window(self):
        self.windows[-1][1] = time.perf_counter()

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nvml
        inside = [s for s in self.samples
                  if any(w[0] <= s[0] <= (w[1] or s[0]) for w in self.windows)]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in inside))
        return {"sm_mhz": statistics.median(s[1] for s in inside) if inside else None,
                "sm_max_mhz": self.max_mhz, "samples": len(inside),
                "samples_total": len(self.samples), "reasons": reasons,
                "how": "NVML polled every ms; median over the samples inside the timed regions "
                       "(resident steps + end-to-end steps)"}


# ---------------------------------------------------------------- workload
def make_sequence(args, world):
    from tools import synth
    cfg = synth.WORKLOADS[args.workload]
    per_gpu = args.length or cfg["length"]
    if args.workload in ("C3", "C5") and not args.length:
        per_gpu = cfg["length"]      # these configs name their own total size
        total = per_gpu
    else:
        total = per_gpu * world
    t0 = time.perf_counter()
    seq = cfg["gen"](total, cfg["seed"])
    return cfg, seq, time.perf_counter() - t0


def build_tables(cfg, seq, device):
    import torch
    from tools.esa_build_torch import build_esa, mirror_codes
    t0 = time.perf_counter()
    codes = torch.from_numpy(seq).to(device)
    if cfg["mirrored"]:
        codes = mirror_codes(codes)
    esa = build_esa(codes, keep_on_device=True)
    torch.cuda.synchronize(device)
    del codes
    return esa, time.perf_counter() - t0


def host_window(esa, lo, hi, pin=True):
    """Pinned host copies of the tables restricted to lcp indices [lo, hi)."""
    import torch
    lcp = esa["lcp"][lo:hi].cpu()
    bwt = esa["bwt"][lo:hi].cpu()
    suf = esa["suf"][lo:hi].cpu()
    sel = (esa["llv_pos"] >= lo) & (esa["llv_pos"] < hi)
    llv = torch.stack([esa["llv_pos"][sel], esa["llv_val"][sel]], dim=1).contiguous().cpu()
    if pin:
        lcp, bwt, suf, llv = (t.pin_memory() for t in (lcp, bwt, suf, llv))
    return lcp, bwt, suf, llv


def index_from_host(capi, lcp, bwt, suf, llv, base, n_total):
    return capi.Index.from_pointers(lcp.data_ptr(), bwt.data_ptr(), llv.data_ptr(),
                                    int(llv.shape[0]), suf.data_ptr(), 8, int(lcp.shape[0]),
                                    keep=(lcp, bwt, suf, llv), base=base, n_total=n_total)


# ------------------------------------------------------------ CPU baseline
def cpu_baseline(args, cfg, seq, what="first %d bp of the workload sequence"):
    """Reference code on the box's host cores, bounded sample of the workload."""
    from tools import synth
    sample_len = min(args.sample, seq.shape[0])
    what = what % sample_len
    minlength = args.minlength or cfg["minlength"]
    cores = 1   # the reference ESA path is single-threaded
    if os.path.exists(GTREF) and not cfg["mirrored"]:
        with tempfile.TemporaryDirectory() as tmp:
            fasta = os.path.join(tmp, "sample.fa")
            synth.to_fasta(seq[:sample_len], fasta, cfg["alphabet"], cfg["wildcard"])
            t0 = time.perf_counter()
            subprocess.run([GTREF, "suffixerator", "-db", fasta, "-suf", "-lcp", "-bwt", "-tis",
                            "-indexname", os.path.join(tmp, "s")] + cfg["flags"],
                           check=True, capture_output=True)
            t_build = time.perf_counter() - t0
            res = {}
            for tool in ("smax-lin", "smax-bu"):
                best = None
                for _ in range(3):
                    p = subprocess.run([GTREF, tool, os.path.join(tmp, "s"), str(minlength)],
                                       check=True, capture_output=True, text=True)
                    t = float(p.stderr.split("t_scan_s=")[1].split()[0])
                    best = t if best is None else min(best, t)
                res[tool] = best
            n = sample_len + 1
            return {"value": n / res["smax-lin"] / 1e9, "unit": "G suffixes/s", "cores": cores,
                    "kind": "reference",
                    "sample": "%s; index built by the reference suffixerator in %.1f s (not "
                              "counted); reference reader macros + linear plateau scan "
                              "(stand-in for esa_linsmax), 1 thread, best of 3" % (what, t_build),
                    "bottomup_value": n / res["smax-bu"] / 1e9,
                    "bottomup_note": "same sample through the reference's gt_esa_bottomup sweep "
                                     "(stand-in for esa-smax)"}
    # C port of the oracle on tables built here (no reference binary on this box)
    import torch
    from oracle import smax_oracle as O
    from tools.esa_build_torch import build_esa, llv_records, mirror_codes
    codes = torch.from_numpy(seq[:sample_len])
    if cfg["mirrored"]:
        codes = mirror_codes(codes)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    esa = build_esa(codes.to(dev))
    llv = llv_records(esa["llv_pos"], esa["llv_val"])
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        O.smax_c(esa["lcp"], llv, esa["bwt"], minlength, 0, "linear")
        t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    return {"value": esa["n"] / best / 1e9, "unit": "G suffixes/s", "cores": cores, "kind": "port",
            "sample": "%s, C restatement (oracle/smax_oracle.c) on in-memory tables, best of 3"
                      % what}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU code on this box (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools import synth
    cfg = synth.WORKLOADS[args.workload]
    minlength = args.minlength or cfg["minlength"]
    sample_len = args.sample
    seq = cfg["gen"](sample_len, cfg["seed"])
    base = cpu_baseline(args, cfg, seq, "the workload generator at %d bp (same seed)")
    # K timed steps of the bounded sample (cpu_baseline already took best-of-3
    # per tool; here every step is one full reference scan of the sample)
    steps, warm = max(1, min(args.steps, 5)), min(args.warmup, 1)
    n = sample_len + 1
    ms = n / (base["value"] * 1e9) * 1e3
    line = {"impl": "reference", "metric": "suffixes scanned/sec", "value": base["value"],
            "unit": "G suffixes/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s sample (%d bp) minlength=%d, reference CPU code via "
                                   "oracle/_ref/gtref" % (args.workload, sample_len, minlength)},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "G suffixes/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit_line(line)


# ------------------------------------------------------------------- main
_REAL_STDOUT = None


def emit_line(line: dict):
    """The ONE JSON line of the contract, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL's version banner, torchrun notices) write to fd 1: keep stdout
    # clean for the JSON line by pointing fd 1 at stderr for everything else
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from genometools_smax_b200 import capi
    from genometools_smax_b200.shard import ShardedScan, shard_cuts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- workload: every rank builds the same index deterministically and
    # keeps only its shard (no table scatter, no host round trip of 8n bytes)
    cfg, seq, t_gen = make_sequence(args, world)
    minlength = args.minlength or cfg["minlength"]
    esa, t_build = build_tables(cfg, seq, device)
    n = esa["n"]
    cuts = shard_cuts(n, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    w_lo = max(0, lo - 256) & ~15
    w_hi = min(n, hi + 16)
    lcp_h, bwt_h, suf_h, llv_h = host_window(esa, w_lo, w_hi)
    nllv_total = int(esa["llv_pos"].shape[0])
    maxlcp = esa["maxlcp"]
    if not args.check:
        del esa
        torch.cuda.empty_cache()
    idx = index_from_host(capi, lcp_h, bwt_h, suf_h, llv_h, w_lo, n)

    dev = capi.Device(local_rank)
    scan = ShardedScan(dev, rank, world)
    h2d_resident = scan.load(idx, n, with_suf=True)
    stream = torch.cuda.current_stream(device).cuda_stream

    # ---- algorithmic bytes of one scan (stats build of the kernel, untimed)
    dev.set_stats(True)
    scan.launch(minlength, capi.POLICY_GT, True, stream)
    st = dev.stats()
    recs0, pos0 = scan.fetch()
    dev.set_stats(False)
    n_shard = hi - lo
    # n lcp bytes + bwt bytes of candidate plateaus + 16 B per inspected .llv
    # record (each record counted once) + suf entries read and positions written for
    # survivors + 24 B per record written
    nllv_shard = int(llv_h.shape[0])
    alg_scan = (n_shard + st["candidate_width"] + 16 * min(st["llv_inspected"], nllv_shard)
                + (8 + 8) * st["survivor_width"] + 24 * st["survivors"])
    alg_step = alg_scan

    if args.check:
        from oracle import smax_oracle as O
        from tools.esa_build_torch import llv_records
        lcp_all = esa["lcp"].cpu().numpy(); bwt_all = esa["bwt"].cpu().numpy()
        llv_all = llv_records(esa["llv_pos"].cpu().numpy(), esa["llv_val"].cpu().numpy())
        want = O.smax_c(lcp_all, llv_all, bwt_all, minlength)
        ends = want["lb"] + want["width"] - 1
        mine = want[(ends >= lo) & (ends < hi)]
        assert np.array_equal(recs0, mine), "records differ from the oracle"
        assert np.array_equal(pos0, O.positions_c(esa["suf"].cpu().numpy().astype(np.uint64), mine))
        del esa
        torch.cuda.empty_cache()
        if rank == 0:
            print("# check ok: %d records on rank 0 match the oracle" % len(mine), file=sys.stderr)

    # ---- resident timing
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for _ in range(args.warmup):
        flush.fill_(1)
        scan.launch(minlength, capi.POLICY_GT, True, stream)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    scan_ms, total_ms = [], []
    t_wall0 = time.perf_counter()
    barrier()
    sampler.open_window()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        scan.launch(minlength, capi.POLICY_GT, True, stream)
        ev[k][1].record()
        # per-kernel events recorded inside the C-ABI launch (same stream)
        ms_all, ms_scan, launches = dev.elapsed_ms()
        scan_ms.append(ms_scan)
    barrier()
    sampler.close_window()
    t_wall = time.perf_counter() - t_wall0
    total_ms = [a.elapsed_time(b) for a, b in ev]
    ms_step = sum(total_ms) / len(total_ms)
    off, total_recs = scan.offsets()
    t = torch.tensor([ms_step, sum(scan_ms) / len(scan_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_scan_k = float(t[0]), float(t[1])

    # ---- emit path, outside the timed region: the records + positions
    # of the last scan rendered as text in HBM, next to the host emitter on the same records
    emit = None
    if not args.no_emit:
        nbytes = dev.format_text(capi.FORMAT_SMAX, False, fetch=False)
        fms = []
        for _ in range(10):
            flush.fill_(3)
            dev.format_text(capi.FORMAT_SMAX, False, fetch=False)
            fms.append(dev.format_elapsed_ms())
        t0 = time.perf_counter()
        text = dev.format_text(capi.FORMAT_SMAX, False)
        t_fetch = time.perf_counter() - t0
        recs_l, pos_l = scan.fetch()
        t0 = time.perf_counter()
        idx_e = capi.Index.from_arrays(np.zeros(1, np.uint8), np.zeros(1, np.uint8))
        idx_e.emit_text(recs_l, pos_l, capi.FORMAT_SMAX, False, discard=True)
        t_host = time.perf_counter() - t0
        if args.check:
            assert text == idx_e.emit_text(recs_l, pos_l, capi.FORMAT_SMAX, False), "device text differs"
            if rank == 0:
                print("# check ok: %d bytes of device-rendered text match the host emitter" % nbytes,
                      file=sys.stderr)
        idx_e.close()
        fmed = sorted(fms)[len(fms) // 2]
        emit = {"format": "smax, absolute positions", "records": int(len(recs_l)),
                "positions": int(len(pos_l)), "text_bytes": int(nbytes),
                "device_format_ms": fmed, "device_text_gbs": nbytes / (fmed * 1e-3) / 1e9,
                "device_format_plus_d2h_ms": t_fetch * 1e3,
                "host_emitter_ms": t_host * 1e3,
                "note": "rank 0; smax_scan_format = 4 launches (item sizes reduced per block, scan of the "
                        "block sums, offsets applied, items written), CUDA events; host = "
                        "smax_emitter_emit_records into /dev/null, 1 thread; not part of the timed steps"}
        del text

    # ---- end to end through the C ABI with host tables
    e2e = None
    if not args.no_e2e:
        h2d = d2h = 0
        dev2 = dev   # same device context: allocations are reused across steps
        times = []
        for k in range(args.warmup + args.steps):
            barrier()
            if k == args.warmup:
                sampler.open_window()
            t0 = time.perf_counter()
            h2d = dev2.upload(idx, lo, hi, with_suf=False)
            dev2.scan(minlength, capi.POLICY_GT, False, 0)
            recs, _ = dev2.fetch()
            # occurrence positions from the HOST suffix table (C-ABI ragged gather)
            posn = idx.gather_positions(recs)
            barrier()
            if k >= args.warmup:
                times.append(time.perf_counter() - t0)
            d2h = recs.nbytes + 64
        sampler.close_window()
        t_e2e = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": n / float(t_e2e[0]) / 1e9, "unit": "G suffixes/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(t_e2e[0]) * 1e3,
               "note": "smax_device_upload(lcp,bwt,llv from pinned host) + scan + record fetch "
                       "+ smax_index_gather_positions from the host suftab; wall clock"}

    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (k_scan)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_scan / (ms_scan_k * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("k_scan_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_scan", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_scan),
                "bytes_per_suffix": alg_scan / n_shard, "kernel_ms": ms_scan_k,
                "step_algorithmic_bytes": int(alg_step)}

    if rank == 0:
        line = {
            "metric": "suffixes scanned/sec", "value": n / (ms_step * 1e-3) / 1e9,
            "unit": "G suffixes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s: %s, %d suffixes total (%d per GPU), minlength=%d, "
                                   "suftab 64-bit, policy gt" % (
                                       args.workload, cfg["gen"].__name__, n, n // world, minlength),
                       "l2": "flushed between steps (512 MiB fill outside the timed events)",
                       "sharding": "SA range cut into %d shards; P2P left views; record counts "
                                   "exchanged by P2P stores of the scan kernel (no collective "
                                   "per step)" % world if world > 1 else "single shard",
                       "largelcpvalues": nllv_total, "maxbranchdepth": maxlcp,
                       "records": int(total_recs), "positions_rank0": int(st["positions"]),
                       "candidates_rank0": int(st["candidates"])},
            "roofline": roofline,
            "gpu_launches": args.steps * 1,
            "clocks": clocks,
            "timing": {"step_ms_min": min(total_ms), "step_ms_max": max(total_ms),
                       "wall_s_timed_region": t_wall, "index_build_s": t_build,
                       "sequence_gen_s": t_gen, "resident_upload_bytes": int(h2d_resident)},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if emit is not None:
            line["emit"] = emit
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args, cfg, seq)
        emit_line(line)
    barrier()
    dev.close()
    idx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
