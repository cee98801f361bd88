"""Per-source-line view of an .ncu-rep (tuning tool): stall samples and executed warp
instructions per line of the CUDA source, from ncu's own source page.

    python tools/ncu_src.py gpurun_out/x.ncu-rep [top] [file-substring]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        try:
            samples = int(d["# Samples"] or 0)
            inst = int(d["Instructions Executed"] or 0)
            thr = int(d["Thread Instructions Executed"] or 0)
        except ValueError:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items()
                  if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
        lines.append((cur_file, int(r[0]), r[1].strip(), samples, inst, thr, stalls))
    lines = [l for l in lines if want in (l[0] or "")]
    ts = sum(l[3] for l in lines) or 1
    ti = sum(l[4] for l in lines) or 1
    print("total samples %d, executed warp instructions %d, avg lanes %.1f" % (
        ts, ti, sum(l[5] for l in lines) / ti))
    print("--- by stall samples")
    for f, n, src, s, i, t, st in sorted(lines, key=lambda l: -l[3])[:top]:
        main_st = ",".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s   [%s]" % (100.0 * s / ts, 100.0 * i / ti,
              (f or "?").split("/")[-1], n, src[:90], main_st))
    print("--- by executed instructions")
    for f, n, src, s, i, t, st in sorted(lines, key=lambda l: -l[4])[:top]:
        print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / ts,
              t / max(i, 1), (f or "?").split("/")[-1], n, src[:90]))


if __name__ == "__main__":
    main()
