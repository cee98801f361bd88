"""Stall samples / executed instructions of k_scan summed over regions of smax_scan.cu (tuning tool).

    python tools/ncu_regions.py gpurun_out/x.ncu-rep
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:k_scan"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, L = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        L.append((cur.split('/')[-1], int(r[0]), num(d["# Samples"]), num(d["Instructions Executed"]),
                  num(d["Thread Instructions Executed"])))
    src = open(os.path.join(ROOT, "genometools_smax_b200", "csrc", "smax_scan.cu")).read().split("\n")

    def find(s):
        for i, l in enumerate(src):
            if s in l:
                return i + 1
        return None
    marks = [("prologue (tickets of the first two units)", find("k_scan(const __grid_constant__")),
             ("unit head (bulk copy, pipeline words)", find("while (unit < nunits)")),
             ("K1a: large values in record space", find("K1a: large values")),
             ("wait for the lcp bytes + phase A filter", find("K1b + K2: small values")),
             ("pipeline step 2 + K2 of large values", find("pipeline: the ticket has arrived")),
             ("phase B (chunks -> ENDs -> K1/K2)", find("phase B, first level")),
             ("K3 first half (bitmaps -> arena)", find("K3, first half: the warp turns")),
             ("kernel tail", find("k_emit may be put in place")),
             ("k_emit", find("K3, second half"))]
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for f, n, s, i, t in L:
        if f == "smax_scan.cu":
            reg = "helpers (walks, list passes, accessors, mbarrier)"
            for name, ln in marks:
                if ln and n >= ln:
                    reg = name
        elif f == "smax_swar.h":
            reg = "smax_swar.h (bit-parallel K1/K2)"
        else:
            reg = "CUDA headers (shuffles, ballots, atomics)"
        a = agg[reg]
        a[0] += s
        a[1] += i
        a[2] += t
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print("%d stall samples, %d executed warp instructions" % (ts, ti))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-42s %5.1f%% samples %5.1f%% instructions  %4.1f lanes" % (k, 100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(a[1], 1)))


if __name__ == "__main__":
    main()
