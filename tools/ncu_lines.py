"""Per-source-line hot spots of k_scan from an .ncu-rep (tuning tool).

ncu's CSV source page only lists SASS; this joins its per-instruction warp-stall
samples with the line table nvdisasm prints for the same cubin.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]

The line of a sample is that of the innermost inlined function nvdisasm names for the
instruction; waits show up on the branch / convergence instruction (BRA, BSSY, BSYNC) of
the spin loop or barrier, whose line can be that of a neighbouring statement -- the SASS
of the hottest instruction is printed in brackets to tell them apart.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_substr, obj=None):
    obj = obj or os.path.join(ROOT, "genometools_smax_b200", "lib", "smax_scan.cu.o")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], check=True,
                             capture_output=True, text=True).stdout
    table, on, cur = {}, False, None
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = kernel_substr in ln
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    obj = sys.argv[3] if len(sys.argv) > 3 else None        # the object file the profiled library was linked from
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:k_scan"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kname = rows[0][1]
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    # k_scan<(bool)S, (bool)W> -> the section of exactly that instantiation (_ZN4smax6k_scanILbSELbWEEE...)
    flags = re.findall(r"\(bool\)([01])", kname)
    table = line_table("k_scan" + "".join("%sLb%s" % ("I" if i == 0 else "E", f) for i, f in enumerate(flags)) + "EEE", obj)
    base = None
    per_line = collections.Counter()
    per_line_inst = collections.Counter()
    per_line_top = {}                       # line -> (samples, SASS text) of its hottest instruction
    src_cache = {}
    total = 0
    for r in rows[2:]:
        if len(r) <= isamp or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        base = addr if base is None else base
        loc, sass = table.get(addr - base, ((None, 0), ""))
        n = int(r[isamp] or 0)
        if n > per_line_top.get(loc, (0, ""))[0]:
            per_line_top[loc] = (n, sass)
        per_line[loc] += n
        per_line_inst[loc] += int(r[iex] or 0)
        total += n
    print("kernel %s, %d samples" % (kname, total))
    for loc, n in per_line.most_common(top):
        text = ""
        if loc and loc[0]:
            path = os.path.join(ROOT, "genometools_smax_b200", "csrc", loc[0])
            if path not in src_cache and os.path.exists(path):
                src_cache[path] = open(path).read().splitlines()
            if path in src_cache and 0 < loc[1] <= len(src_cache[path]):
                text = src_cache[path][loc[1] - 1].strip()[:90]
        print("%6d %5.1f%% inst %9d  %s:%s  %s   [%s]" % (
            n, 100.0 * n / max(total, 1), per_line_inst[loc], loc[0] if loc else "?",
            loc[1] if loc else 0, text, per_line_top.get(loc, (0, ""))[1].split(";")[0][:40]))


if __name__ == "__main__":
    main()
