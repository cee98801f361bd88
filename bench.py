#!/usr/bin/env python
"""Benchmark of the supermaximal-repeat scan (BASELINE.json metric:
G suffixes scanned / s and achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm
    python bench.py --impl reference [...]                         # reference CPU arm
    torchrun --nproc-per-node N bench.py --gpus N ...              # N > 1

A step = one pass of the hot path over the whole resident index: plateau
detection + .llv resolution + left-distinctness + ordered compaction + position
gather (the two launches of the unit kernel, k_scan + k_emit; one launch of the
ring kernel under SMAX_KERNEL=ring), plus -- for N > 1 -- the one-sided exchange
of the shard record counts.

Workloads (tools/synth.py):
  C2 (default)  synthetic DNA with tandem / interspersed repeats, minlength 20.
                N = 1: 100 Mbp (BASELINE.json configs[1]).  N > 1: WEAK scaling,
                ONE index over N independent 100 Mbp C2 blocks, SA range cut
                into N shards of equal cost.
  C4            protein 200 M residues, minlength 8 (weak, per-GPU length).
  C3, C5        500 Mbp -mirrored / 3 Gbp: ONE fixed index, STRONG scaling
                (tools/c5_strong.py is the one-process driver that builds the
                index once and runs N = 1, 2, 4, 8 on it).
The index is built on the box by tools/esa_build_torch.py (bit-identical to the
reference suffixerator's tables: golden fixtures in tests/, and the CPU-baseline
sample of every run); construction is out of scope and timed separately.

`value`   : resident tables, CUDA events on the launching stream, L2 flushed
            between steps, max over ranks.
`parity`  : EVERY run, outside the timed region: each rank compares the records
            and positions of its shard with the C oracle (oracle/smax_oracle.c)
            run over the same shard of the same tables.
`e2e`     : the plug-in call.  Rank 0 calls smax_run (host tables in pinned
            memory -> upload to all N GPUs -> scan -> records -> host emitter ->
            last byte of text) once per step, wall clock; the variants (text
            rendered on the devices, mmapped index files, -scan) sit next to it.
`roofline`: the scan kernel alone, algorithmic bytes / its event
            duration, per rank.
`cpu_baseline` / --impl reference: reference code (oracle/_ref/gtref, 1 thread:
            the reference ESA path is single-threaded) -- the reference arm on
            the SAME index as this arm (torch-built tables + the reference's own
            .esq), every step one full run; cpu_baseline on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
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
STRONG = ("C3", "C5")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smax", choices=["smax", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--length", type=int, default=0,
                    help="sequence length: per GPU for C2/C4 (weak), total for C3/C5 (strong); 0 = config")
    ap.add_argument("--minlength", type=int, default=0)
    ap.add_argument("--sample", type=int, default=10_000_000,
                    help="sequence length of the CPU-baseline sample")
    ap.add_argument("--equal-cuts", action="store_true", help="cut the SA range by length, not by cost")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-emit", action="store_true", help="skip the device-side text formatting figures")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the per-rank oracle comparison")
    ap.add_argument("--check", action="store_true", help="(kept for old command lines: the check is on by default)")
    ap.add_argument("--ref-seconds", type=float, default=150.0,
                    help="--impl reference: time budget of the timed steps")
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
def workload_shape(args, world):
    """(config, total sequence length, scaling)"""
    from tools import synth
    cfg = synth.WORKLOADS[args.workload]
    if args.workload in STRONG:
        return cfg, args.length or cfg["length"], "strong"
    return cfg, (args.length or cfg["length"]) * world, "weak"


def workload_name(args, cfg, n, world, minlength):
    return "%s: %s, %d suffixes total (%d per GPU), minlength=%d, suftab 64-bit, policy gt" % (
        args.workload, cfg["gen"].__name__, n, n // world, minlength)


def make_sequence(args, world):
    cfg, total, scaling = workload_shape(args, world)
    t0 = time.perf_counter()
    seq = cfg["gen"](total, cfg["seed"])
    return cfg, seq, scaling, time.perf_counter() - t0


def build_tables(cfg, seq, device):
    import torch
    from tools.esa_build_torch import build_esa, mirror_codes
    t0 = time.perf_counter()
    codes = torch.from_numpy(seq).to(device)
    if cfg["mirrored"]:
        codes = mirror_codes(codes)
    esa = build_esa(codes, keep_on_device=True)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    del codes
    return esa, time.perf_counter() - t0


def host_window(esa, lo, hi, pin=True):
    """Pinned host copies of the tables restricted to lcp indices [lo, hi)."""
    import torch
    lcp = esa["lcp"][lo:hi].cpu()
    bwt = esa["bwt"][lo:hi].cpu()
    suf = esa["suf"][lo:hi].cpu()
    k0, k1 = (int(x) for x in torch.searchsorted(
        esa["llv_pos"], torch.tensor([lo, hi], dtype=torch.int64, device=esa["llv_pos"].device)))
    llv = torch.stack([esa["llv_pos"][k0:k1], esa["llv_val"][k0:k1]], dim=1).contiguous().cpu()
    if pin:
        lcp, bwt, suf, llv = (t.pin_memory() for t in (lcp, bwt, suf, llv))
    return lcp, bwt, suf, llv


def index_from_host(capi, lcp, bwt, suf, llv, base, n_total):
    return capi.Index.from_pointers(lcp.data_ptr(), bwt.data_ptr(), llv.data_ptr(),
                                    int(llv.shape[0]), suf.data_ptr(), 8, int(lcp.shape[0]),
                                    keep=(lcp, bwt, suf, llv), base=base, n_total=n_total)


# ------------------------------------------------------------------ parity
def shard_oracle(esa, lo, hi, minlength, policy=0):
    """Records (global coordinates) and positions the C oracle finds for the plateaus that END
    in [lo, hi): it runs over the window [q, hi], q = the last index below lo whose lcp value
    is smaller than the minimum length -- no reported plateau can reach q, so the window holds
    every one of them whole, together with the entries either side that decide them."""
    import torch
    from oracle import smax_oracle as O
    n = esa["n"]
    mb = min(minlength, 255)
    q, w = 0, 1 << 16
    while lo > 0:
        s = max(0, lo - w)
        small = torch.nonzero(esa["lcp"][s:lo] < mb).flatten()
        if small.numel():
            q = s + int(small[-1])
            break
        if s == 0:
            break
        w *= 8
    e = min(n, hi + 1)
    lcp = esa["lcp"][q:e].cpu().numpy()
    bwt = esa["bwt"][q:e].cpu().numpy()
    k0, k1 = (int(x) for x in torch.searchsorted(
        esa["llv_pos"], torch.tensor([q, e], dtype=torch.int64, device=esa["llv_pos"].device)))
    llv = np.zeros(k1 - k0, dtype=O.LLV_DTYPE)
    llv["position"] = (esa["llv_pos"][k0:k1] - q).cpu().numpy()
    llv["value"] = esa["llv_val"][k0:k1].cpu().numpy()
    t0 = time.perf_counter()
    recs = O.smax_c(lcp, llv, bwt, minlength, policy)
    t_oracle = time.perf_counter() - t0
    ends = recs["lb"] + recs["width"] - 1 + q
    recs = recs[(ends >= lo) & (ends < hi)].copy()
    pos = O.positions_c(esa["suf"][q:e].cpu().numpy().astype(np.uint64), recs)
    recs["lb"] += q
    return recs, pos, t_oracle, e - q


def check_shard(esa, lo, hi, minlength, recs, pos):
    want, want_pos, t_oracle, span = shard_oracle(esa, lo, hi, minlength)
    ok = bool(np.array_equal(recs, want) and np.array_equal(pos, want_pos))
    return {"ok": ok, "records": int(len(want)), "positions": int(len(want_pos)),
            "oracle_s": round(t_oracle, 3), "oracle_span": int(span)}


# ------------------------------------------------------------ CPU baseline
def run_gtref(tool, base, minlength, quiet):
    """One run of the reference code; returns (t_scan_s measured inside gtref, repeats)."""
    cmd = [GTREF, tool, base, str(minlength)] + (["quiet"] if quiet else [])
    with open(os.devnull, "w") as null:
        p = subprocess.run(cmd, check=True, stdout=null, stderr=subprocess.PIPE, text=True)
    tail = p.stderr.split("t_scan_s=")[1].split()
    return float(tail[0]), int(tail[1].split("=")[1])


def cpu_baseline(args, cfg, seq, what="first %d bp of the workload sequence", use_ref=True):
    """Reference code on the box's host cores, bounded sample of the workload; on the way the
    torch builder's tables are compared with the reference suffixerator's for the sample."""
    from tools import synth
    sample_len = min(args.sample, seq.shape[0])
    what = what % sample_len
    minlength = args.minlength or cfg["minlength"]
    cores = 1   # the reference ESA path is single-threaded
    if use_ref and os.path.exists(GTREF):
        with tempfile.TemporaryDirectory() as tmp:
            fasta = os.path.join(tmp, "sample.fa")
            base = os.path.join(tmp, "s")
            synth.to_fasta(seq[:sample_len], fasta, cfg["alphabet"], cfg["wildcard"])
            t0 = time.perf_counter()
            subprocess.run([GTREF, "suffixerator", "-db", fasta, "-suf", "-lcp", "-bwt", "-tis",
                            "-indexname", base] + cfg["flags"], check=True, capture_output=True)
            t_build = time.perf_counter() - t0
            builder_check = None
            try:
                import torch
                from tools.esa_build_torch import build_esa, mirror_codes
                dev = "cuda" if torch.cuda.is_available() else "cpu"
                codes = torch.from_numpy(seq[:sample_len]).to(dev)
                if cfg["mirrored"]:
                    codes = mirror_codes(codes)
                esa = build_esa(codes)
                same = (np.array_equal(np.fromfile(base + ".suf", "<u8"), esa["suf"].astype(np.uint64))
                        and np.array_equal(np.fromfile(base + ".lcp", np.uint8), esa["lcp"])
                        and np.array_equal(np.fromfile(base + ".bwt", np.uint8), esa["bwt"])
                        and np.array_equal(np.fromfile(base + ".llv", "<u8").reshape(-1, 2)[:, 0],
                                           esa["llv_pos"].astype(np.uint64))
                        and np.array_equal(np.fromfile(base + ".llv", "<u8").reshape(-1, 2)[:, 1],
                                           esa["llv_val"].astype(np.uint64)))
                builder_check = ("tables of the torch builder identical to the reference suffixerator's "
                                 "(.suf .lcp .bwt .llv, %d suffixes)" % esa["n"]) if same else "TABLES DIFFER"
                del esa, codes
            except Exception as exc:          # the baseline itself does not depend on it
                builder_check = "not run: %s" % exc
            res = {}
            for tool in ("smax-lin", "smax-bu"):
                res[tool] = min(run_gtref(tool, base, minlength, True)[0] for _ in range(3))
            n = sample_len * (2 if cfg["mirrored"] else 1) + (2 if cfg["mirrored"] else 1)
            return {"value": n / res["smax-lin"] / 1e9, "unit": "G suffixes/s", "cores": cores,
                    "kind": "reference",
                    "sample": "%s; index built by the reference suffixerator in %.1f s (not "
                              "counted); reference reader macros + linear plateau scan "
                              "(stand-in for esa_linsmax), 1 thread, repeats counted not printed, "
                              "best of 3" % (what, t_build),
                    "bottomup_value": n / res["smax-bu"] / 1e9,
                    "bottomup_note": "same sample through the reference's gt_esa_bottomup sweep "
                                     "(stand-in for esa-smax)",
                    "builder_check": builder_check}
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


def scratch_dir(need_bytes):
    """A directory for index files: /dev/shm when it has the room (page-cache-warm either way)."""
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(cand).free > need_bytes * 1.2 + (1 << 30):
                return tempfile.mkdtemp(prefix="smaxbench_", dir=cand)
        except OSError:
            pass
    return tempfile.mkdtemp(prefix="smaxbench_")


def write_index(esa, cfg, seq, base, with_reference_esq):
    """The index of this run as GenomeTools files (tools/esa_files.py): torch-built tables,
    .esq / sequence keys by the reference encoder when gtref is there."""
    from tools import esa_files, synth
    template = None
    if with_reference_esq and os.path.exists(GTREF):
        fasta = base + ".fa"
        synth.to_fasta(seq, fasta, cfg["alphabet"], cfg["wildcard"])
        esa_files.encode_with_reference(GTREF, base, fasta, cfg["flags"])
        os.unlink(fasta)
        template = base + ".prj"
    esa_files.write_tables(base, esa)
    logical = seq
    if cfg["mirrored"]:
        logical = None      # the counts of the reference encoder (template) or none at all
    esa_files.write_prj(base, esa, logical, cfg["mirrored"], template=template)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU code on this box (rank 0 only), on the SAME
    index as the GPU arm at this N: torch-built tables (byte-identical to the reference
    suffixerator's, checked on the sample of every run) + the .esq of the reference encoder.
    Every step is one full run of `gtref smax-lin` (scan + one printed line per repeat into
    /dev/null, like the tool); as many of the requested steps as fit --ref-seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    world = args.gpus
    cfg, seq, scaling, _ = make_sequence(args, world)
    minlength = args.minlength or cfg["minlength"]
    have_gpu = torch.cuda.is_available()
    if not os.path.exists(GTREF) or not have_gpu:
        # no reference binary (or no GPU to build the full index): bounded sample
        base = cpu_baseline(args, cfg, seq)
        n = args.sample + 1
        ms = n / (base["value"] * 1e9) * 1e3
        emit_line({"impl": "reference", "metric": "suffixes scanned/sec", "value": base["value"],
                   "unit": "G suffixes/s", "n_gpus": args.gpus, "steps": 3, "warmup": 0,
                   "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                   "dtype": "u8", "data": "synthetic",
                   "config": {"workload": "%s sample (%d bp) minlength=%d" % (args.workload, args.sample, minlength)},
                   "cpu_baseline": base,
                   "e2e": {"value": base["value"], "unit": "G suffixes/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}})
        return
    device = torch.device("cuda", 0)
    esa, t_build = build_tables(cfg, seq, device)
    n = esa["n"]
    tmp = scratch_dir(n * 10 + 16 * int(esa["llv_pos"].shape[0]) + seq.shape[0] * 2)
    try:
        base = os.path.join(tmp, "idx")
        t0 = time.perf_counter()
        write_index(esa, cfg, seq, base, True)
        t_write = time.perf_counter() - t0
        del esa
        torch.cuda.empty_cache()
        warm = min(args.warmup, 1)
        for _ in range(warm):
            run_gtref("smax-lin", base, minlength, False)
        times, repeats, t_start = [], 0, time.perf_counter()
        while len(times) < max(1, args.steps) and (not times or time.perf_counter() - t_start < args.ref_seconds):
            t, repeats = run_gtref("smax-lin", base, minlength, False)
            times.append(t)
        t_quiet, _ = run_gtref("smax-lin", base, minlength, True)
        t_bu, _ = run_gtref("smax-bu", base, minlength, True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    ms = sum(times) / len(times) * 1e3
    value = n / (ms * 1e-3) / 1e9
    base_info = {"value": value, "unit": "G suffixes/s", "cores": 1, "kind": "reference",
                 "sample": "the whole workload index (%d suffixes): tables by the torch builder in %.1f s, "
                           ".esq by the reference encoder, files written in %.1f s (not counted); "
                           "`gtref smax-lin` = reference reader macros + linear plateau scan, 1 thread, "
                           "one line per repeat printed into /dev/null; mean of %d runs"
                           % (n, t_build, t_write, len(times)),
                 "scan_only_value": n / t_quiet / 1e9,
                 "scan_only_note": "same run with the repeats counted instead of printed",
                 "bottomup_value": n / t_bu / 1e9,
                 "bottomup_note": "the reference's gt_esa_bottomup sweep (stand-in for esa-smax), not printed",
                 "repeats": repeats}
    emit_line({"impl": "reference", "metric": "suffixes scanned/sec", "value": value,
               "unit": "G suffixes/s", "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
               "steps_requested": args.steps,
               "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
               "dtype": "u8", "data": "synthetic",
               "config": {"workload": workload_name(args, cfg, n, world, minlength)},
               "cpu_baseline": base_info,
               "e2e": {"value": value, "unit": "G suffixes/s", "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": 0}})


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


def measured_traffic(workload, n, kernel):
    """DRAM bytes per launch of the scan kernel from an ncu capture of EXACTLY this workload,
    size and kernel (profiles/traffic.json), else None: the bench cannot measure it itself."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    for ent in json.load(open(path)).get("captures", []):
        if ent.get("workload") == workload and ent.get("n") == n and ent.get("kernel") == kernel:
            return ent.get("dram_bytes_per_launch")
    return None


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
    from genometools_smax_b200.shard import ShardedScan, balanced_cuts, shard_cuts

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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")      # one node: the container hostname may not resolve
        try:
            cpu_group = dist.new_group(backend="gloo")
        except Exception as exc:                               # (then the waits below are NCCL barriers)
            print("# no gloo group: %s" % str(exc)[:200], file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def cpu_barrier():
        # a wait that leaves the GPUs idle: an NCCL barrier is a kernel that spins on the waiting
        # ranks' GPUs, and a GPU time-slices between the contexts of two processes -- rank 0's
        # end-to-end call on those GPUs then runs at half speed (21.5 instead of 9.7 ms at N = 2)
        torch.cuda.synchronize(device)
        if world > 1 and cpu_group is not None:
            dist.barrier(group=cpu_group)

    def gather_obj(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    # ---- workload: every rank builds the same index deterministically and
    # keeps only its shard (no table scatter, no host round trip of 8n bytes)
    cfg, seq, scaling, t_gen = make_sequence(args, world)
    minlength = args.minlength or cfg["minlength"]
    esa, t_build = build_tables(cfg, seq, device)
    n = esa["n"]
    cuts = shard_cuts(n, world) if args.equal_cuts else balanced_cuts(n, world, esa["llv_pos"])
    lo, hi = cuts[rank], cuts[rank + 1]
    w_lo = max(0, lo - 256) & ~15
    w_hi = min(n, hi + 16)
    lcp_h, bwt_h, suf_h, llv_h = host_window(esa, w_lo, w_hi)
    nllv_total = int(esa["llv_pos"].shape[0])
    maxlcp = esa["maxlcp"]
    idx = index_from_host(capi, lcp_h, bwt_h, suf_h, llv_h, w_lo, n)

    dev = capi.Device(local_rank)
    scan = ShardedScan(dev, rank, world)
    h2d_resident = scan.load(idx, n, with_suf=True, cuts=cuts)
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

    # ---- parity, every run: this rank's records + positions against the C oracle
    parity_mine = {"ok": None, "skipped": True}
    if not args.no_check:
        parity_mine = check_shard(esa, lo, hi, minlength, recs0, pos0)
        if not parity_mine["ok"]:
            print("# PARITY FAIL on rank %d: %d records here, %d in the oracle" % (
                rank, len(recs0), parity_mine["records"]), file=sys.stderr)
    keep_full = (rank == 0 and not args.no_e2e)
    if not keep_full:
        del esa
        torch.cuda.empty_cache()

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
    scan_ms, launches = [], 1
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
    ms_step_mine = sum(total_ms) / len(total_ms)
    ms_scan_mine = sum(scan_ms) / len(scan_ms)
    off, total_recs = scan.offsets()
    t = torch.tensor([ms_step_mine, ms_scan_mine], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_scan_k = float(t[0]), float(t[1])
    # the records of the timed scans are those of the checked one
    recs_t, pos_t = scan.fetch()
    if parity_mine.get("ok") is not None:
        parity_mine["ok"] = bool(parity_mine["ok"] and np.array_equal(recs_t, recs0)
                                 and np.array_equal(pos_t, pos0))
    per_rank = gather_obj({"rank": rank, "suffixes": int(n_shard), "largelcpvalues": nllv_shard,
                           "kernel_ms": ms_scan_mine, "step_ms": ms_step_mine,
                           "algorithmic_bytes": int(alg_scan),
                           "bytes_per_suffix": alg_scan / max(n_shard, 1),
                           "gbs": alg_scan / (ms_scan_mine * 1e-3) / 1e9,
                           "records": int(len(recs0)), "parity": parity_mine,
                           "kernel": "ring" if int(launches) == 1 else "units"})

    # ---- emit path, outside the timed region: the records + positions
    # of the last scan rendered as text in HBM, next to the host emitter on the same records
    emit = None
    if not args.no_emit and rank == 0:
        nbytes = dev.format_text(capi.FORMAT_SMAX, False, fetch=False)
        fms = []
        for _ in range(10):
            flush.fill_(3)
            dev.format_text(capi.FORMAT_SMAX, False, fetch=False)
            fms.append(dev.format_elapsed_ms())
        t0 = time.perf_counter()
        text = dev.format_text(capi.FORMAT_SMAX, False)
        t_fetch = time.perf_counter() - t0
        t0 = time.perf_counter()
        idx_e = capi.Index.from_arrays(np.zeros(1, np.uint8), np.zeros(1, np.uint8))
        idx_e.emit_text(recs_t, pos_t, capi.FORMAT_SMAX, False, discard=True)
        t_host = time.perf_counter() - t0
        text_ok = None
        if not args.no_check:
            text_ok = bool(text == idx_e.emit_text(recs_t, pos_t, capi.FORMAT_SMAX, False))
        idx_e.close()
        fmed = sorted(fms)[len(fms) // 2]
        emit = {"format": "smax, absolute positions", "records": int(len(recs_t)),
                "positions": int(len(pos_t)), "text_bytes": int(nbytes),
                "device_format_ms": fmed, "device_text_gbs": nbytes / (fmed * 1e-3) / 1e9,
                "device_format_plus_d2h_ms": t_fetch * 1e3,
                "host_emitter_ms": t_host * 1e3, "device_text_equals_host_text": text_ok,
                "note": "rank 0; smax_scan_format = 4 launches (item sizes reduced per block, scan of the "
                        "block sums, offsets applied, items written), CUDA events; host = "
                        "smax_emitter_emit_records into /dev/null, 1 thread; not part of the timed steps"}
        del text
    del flush

    # ---- end to end: the plug-in call (rank 0 drives all `world` GPUs through smax_run, the
    # way `gt smax -gpus N` does; the other ranks wait)
    e2e = None
    barrier()
    cpu_barrier()
    if not args.no_e2e:
        dev.close()                      # the library call owns the devices now
        dev = None
        if rank == 0:
            full = host_window(esa, 0, n)
            idx_full = index_from_host(capi, *full, 0, n)
            del esa
            torch.cuda.empty_cache()
            nshards = world
            h2d_host = 2 * (n + 272 * nshards) + 16 * nllv_total
            e2e_steps = max(1, min(args.steps, 20))
            e2e_warm = max(1, min(args.warmup, 3))

            h2d_seen = {}

            def timed(fn, key=None):
                ts = []
                for k in range(e2e_warm + e2e_steps):
                    if k == e2e_warm:
                        b0 = int(capi.lib().smax_h2d_bytes_total())
                    t0 = time.perf_counter()
                    fn()
                    if k >= e2e_warm:
                        ts.append(time.perf_counter() - t0)
                if key:
                    # bytes that really crossed the link per call (counted inside libsmax: the .llv
                    # records go as 4-byte values when they fit)
                    h2d_seen[key] = (int(capi.lib().smax_h2d_bytes_total()) - b0) // e2e_steps
                return sum(ts) / len(ts)

            sampler.open_window()
            t_host_emit = timed(lambda: idx_full.run_emit_text(minlength, ngpus=world, discard=True), "host")
            sampler.close_window()
            t_dev_emit = timed(lambda: idx_full.run_text(minlength, ngpus=world, discard=True), "device")
            text_bytes = getattr(idx_full, "last_text_bytes", 0)
            e2e_text_ok = None
            text_ref = None
            if not args.no_check:
                import hashlib
                a = idx_full.run_emit_text(minlength, ngpus=world)
                b = idx_full.run_text(minlength, ngpus=world)
                e2e_text_ok = bool(a == b and len(a) == text_bytes)
                text_ref = (len(a), hashlib.sha256(a).hexdigest())
                del a, b
            e2e = {"value": n / t_host_emit / 1e9, "unit": "G suffixes/s",
                   "h2d_bytes_per_step": int(h2d_seen.get("host", h2d_host)),
                   "h2d_bytes_per_step_full_records": int(h2d_host),
                   "d2h_bytes_per_step": int(24 * total_recs + 64 * nshards),
                   "ms_per_step": t_host_emit * 1e3, "steps": e2e_steps, "warmup": e2e_warm,
                   "call": "smax_run(idx, {minlength, ngpus=%d}, smax_emitter_emit) -- the tool's default "
                           "path (-emit host): pinned host tables -> upload of lcp/bwt/llv to every shard's "
                           "GPU (concurrent) -> scan -> records -> positions from the host suffix table -> "
                           "one line per repeat into /dev/null; wall clock per call, device handles and "
                           "allocations cached inside libsmax between calls" % world,
                   "device_emit": {"value": n / t_dev_emit / 1e9, "ms_per_step": t_dev_emit * 1e3,
                                   "h2d_bytes_per_step": int(h2d_seen.get("device", h2d_host + 8 * (n + 272 * nshards))),
                                   "d2h_bytes_per_step": int(text_bytes),
                                   "call": "smax_run_text (-emit device): the suffix table is uploaded too, "
                                           "positions gathered and text rendered in HBM, text bytes copied back"},
                   "text_identical_both_paths": e2e_text_ok}
            if world == 1:
                # the same call on index FILES (page-cache-warm mmap, pageable source -> staged copies),
                # and the -scan mode on the same files
                tmp = scratch_dir(n * 10 + 16 * nllv_total)
                try:
                    base = os.path.join(tmp, "idx")
                    tables = {"n": n, "maxlcp": maxlcp, "suf": full[2], "lcp": full[0], "bwt": full[1],
                              "llv_pos": full[3][:, 0], "llv_val": full[3][:, 1]}
                    write_index(tables, cfg, seq, base, False)
                    with capi.Index.open(base, capi.TAB_SUF | capi.TAB_LCP | capi.TAB_BWT) as idx_m:
                        t_mmap = timed(lambda: idx_m.run_emit_text(minlength, ngpus=1, discard=True), "mmap")
                        mmap_ok = None
                        if text_ref is not None:
                            # (the mapped tables take the stripped .llv upload: same text?)
                            c = idx_m.run_emit_text(minlength, ngpus=1)
                            mmap_ok = (len(c), hashlib.sha256(c).hexdigest()) == text_ref
                            del c
                    e2e["mmap_files"] = {"value": n / t_mmap / 1e9, "ms_per_step": t_mmap * 1e3,
                                         "h2d_bytes_per_step": h2d_seen.get("mmap"),
                                         "text_identical_to_pinned_path": mmap_ok,
                                         "call": "smax_index_open (mmap, page cache warm) once, then smax_run + host "
                                                 "emitter per step: what `gt smax -ii idx` does"}
                    with capi.Index.open(base, 0) as idx_s:
                        idx_s.run_stream_text(minlength)
                        t0 = time.perf_counter()
                        idx_s.run_stream_text(minlength)
                        t_scanmode = time.perf_counter() - t0
                    e2e["scan_mode"] = {"value": n / t_scanmode / 1e9, "ms_per_step": t_scanmode * 1e3,
                                        "call": "smax_run_stream (gt smax -scan): chunks read from the files, "
                                                "one run after one warm-up run"}
                except Exception as exc:
                    e2e["mmap_files"] = {"error": str(exc)[:200]}
                finally:
                    shutil.rmtree(tmp, ignore_errors=True)
            idx_full.close()
            capi.lib().smax_release_devices()
        cpu_barrier()
        barrier()

    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (k_scan)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if rank == 0:
        mine = per_rank[0]
        slow = max(per_rank, key=lambda r: r["kernel_ms"])
        achieved = mine["gbs"]
        roofline = {"bound": "hbm", "kernel": "k_scan (%s)" % (mine["kernel"] or "?"),
                    "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak,
                    "traffic": measured_traffic(args.workload, n, mine["kernel"]) if world == 1 else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": mine["algorithmic_bytes"],
                    "bytes_per_suffix": mine["bytes_per_suffix"], "kernel_ms": mine["kernel_ms"],
                    "note": "rank 0's shard; per_rank lists every shard (slowest: rank %d, %.4f ms, "
                            "%.0f GB/s)" % (slow["rank"], slow["kernel_ms"], slow["gbs"])}
        all_ok = all(r["parity"].get("ok") for r in per_rank) if not args.no_check else None
        line = {
            "metric": "suffixes scanned/sec", "value": n / (ms_step * 1e-3) / 1e9,
            "unit": "G suffixes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, cfg, n, world, minlength),
                       "l2": "flushed between steps (512 MiB fill outside the timed events)",
                       "sharding": ("SA range cut into %d shards of equal %s; P2P left views; record counts "
                                    "exchanged by P2P stores of the scan kernel (no collective per step)"
                                    % (world, "length" if args.equal_cuts else "cost (n + 16 per large value)"))
                                   if world > 1 else "single shard",
                       "cuts": [int(c) for c in cuts],
                       "largelcpvalues": nllv_total, "maxbranchdepth": maxlcp,
                       "records": int(total_recs), "positions_rank0": int(st["positions"]),
                       "candidates_rank0": int(st["candidates"])},
            "parity": ("ok" if all_ok else "FAIL") if all_ok is not None else "not checked",
            "parity_how": "every rank: records + positions of its shard == C oracle "
                          "(oracle/smax_oracle.c) over the same shard of the same tables, before and "
                          "after the timed steps; outside the timed region",
            "roofline": roofline,
            "per_rank": per_rank,
            "gpu_launches": args.steps * int(launches),
            "clocks": clocks,
            "timing": {"step_ms_min": min(total_ms), "step_ms_max": max(total_ms),
                       "wall_s_timed_region": t_wall, "index_build_s": t_build,
                       "sequence_gen_s": t_gen, "resident_upload_bytes": int(h2d_resident)},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if emit is not None:
            line["emit"] = emit
        if not args.no_cpu and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args, cfg, seq)
            except Exception as exc:             # the reference build failed on the sample: the port
                line["cpu_baseline"] = cpu_baseline(args, cfg, seq, use_ref=False)
                line["cpu_baseline"]["note"] = "reference run failed (%s): C port instead" % str(exc)[:160]
        emit_line(line)
    barrier()
    if dev is not None:
        dev.close()
    idx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
