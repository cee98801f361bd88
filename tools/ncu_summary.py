"""Turn an .ncu-rep of k_scan into the small tracked summaries under profiles/
(the .ncu-rep itself stays in gpurun_out/, which is scratch).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_k_scan   [--traffic C2 100000001 ring]

writes  <prefix>_summary.json   duration, DRAM bytes, throughputs, occupancy, stall mix
        <prefix>_lines.txt      per-source-line hot spots (tools/ncu_lines.py)
and, with --traffic, profiles/traffic.json (k_scan DRAM bytes per launch, read by bench.py).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
}

UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
              "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, val = rows[0], rows[1], rows[2]
    summ = {"report": os.path.basename(rep), "kernel": val[hdr.index("Kernel Name")]}
    stalls = {}
    for h, u, v in zip(hdr, units, val):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        if h in KEEP:
            summ[KEEP[h]] = x * UNIT_SCALE.get(u, 1)
            if u in UNIT_SCALE and UNIT_SCALE[u] != 1 or u in ("byte", "second"):
                summ[KEEP[h] + "_unit"] = "bytes" if "byte" in u else "seconds"
        elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(x, 3)
    summ["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    if "dram_bytes_read" in summ:
        summ["dram_bytes_total"] = summ["dram_bytes_read"] + summ.get("dram_bytes_write", 0)
    summ["note"] = ("ncu --set full --clock-control none, one launch, cold cache, serialised replays: "
                    "durations are NOT benchmark numbers")
    json.dump(summ, open(prefix + "_summary.json", "w"), indent=1)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"],
                           capture_output=True, text=True).stdout
    open(prefix + "_lines.txt", "w").write(lines)
    if "--traffic" in sys.argv:
        # --traffic WORKLOAD N KERNEL: bench.py reports `roofline.traffic` only for a run of exactly
        # this workload, size and kernel
        k = sys.argv.index("--traffic")
        workload, n, kernel = sys.argv[k + 1], int(sys.argv[k + 2]), sys.argv[k + 3]
        path = os.path.join(ROOT, "profiles", "traffic.json")
        doc = json.load(open(path)) if os.path.exists(path) else {}
        caps = [c for c in doc.get("captures", [])
                if not (c["workload"] == workload and c["n"] == n and c["kernel"] == kernel)]
        caps.append({"workload": workload, "n": n, "kernel": kernel,
                     "dram_bytes_per_launch": int(summ["dram_bytes_total"]),
                     "source": os.path.basename(prefix) + "_summary.json"})
        json.dump({"how": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of the scan kernel "
                          "(ncu --set full) per workload / size / kernel",
                   "captures": caps}, open(path, "w"), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main()
