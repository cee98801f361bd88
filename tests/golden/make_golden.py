#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Runs ONLY in the build container (needs /root/reference and the reference
binary oracle/_ref/gtref built by `make -f oracle/Makefile.ref`).  For every
fixture it

  1. builds the enhanced suffix array with the REFERENCE index builder
     (`gtref suffixerator`, i.e. gt_parseargsandcallsuffixerator,
     /root/reference/src/match/sfx-run.c:720), and
  2. records what REFERENCE CODE reports for the smax path on that index:
     `gtref smax-bu` (visitor on gt_esa_bottomup, esa-bottomup.c:116-273) and
     `gtref smax-lin` (linear scan on the reader macros, esa-seqread.h:96-215),
     in mapped and in -scan mode, which must all agree.

  3. records the same lines with RELATIVE positions ("<seqnum> <relpos>" per occurrence) as
     the reference's own gt_encseq_seqnum / gt_encseq_seqstartpos print them
     (`gtref smax-bu ... rel`, /root/reference/src/core/encseq.c:3815-3900, incl. the
     -mirrored arithmetic) on a second index of the same input built with -ssp
     (exp_rel_<m>): the pin of the tool's -rel output and of the device formatter.

The fixture is one compressed .npz holding the raw index files (so the tests
can re-materialise <idx>.prj/.esq/.suf/.lcp/.llv/.bwt byte for byte) and the
expected text for a list of minimum lengths.  The `plain`-policy expectation
(not a reference convention) is produced by the C restatement and marked so.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
    python tests/golden/make_golden.py u89959   # only the named fixtures
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import smax_oracle as O  # noqa: E402

GTREF = os.path.join(ROOT, "oracle", "_ref", "gtref")
TESTDATA = "/root/reference/testdata"
SUFFIXES = (".prj", ".esq", ".suf", ".lcp", ".llv", ".bwt")


def synth_llv(path):
    """2.3 kb DNA with a 600 bp exact repeat -> lcp values >= 255 (.llv)."""
    rng = np.random.Generator(np.random.PCG64(7))
    s = rng.integers(0, 4, 2300)
    s[1500:2100] = s[200:800]
    seq = "".join("acgt"[c] for c in s)
    with open(path, "w") as fh:
        fh.write(">llv\n" + seq + "\n")


def synth_wide(path):
    """Exact copies terminated by a wildcard: flat plateaus much wider than a
    warp, some with lcp >= 255, all-wildcard and mixed left contexts."""
    rng = np.random.Generator(np.random.PCG64(11))
    parts = []
    elem_a = "".join("acgt"[c] for c in rng.integers(0, 4, 40))
    elem_b = "".join("acgt"[c] for c in rng.integers(0, 4, 300))
    for k in range(700):
        # left context alternates over a,c,g,t,n -> duplicates after 4 copies
        parts.append("acgtn"[k % 5] + elem_a + "n")
        parts.append("".join("acgt"[c] for c in rng.integers(0, 4, int(rng.integers(0, 9)))))
    for k in range(40):
        parts.append("n" + elem_b + "n")          # left context always a wildcard
    for k in range(3):
        parts.append("acg"[k] + elem_b[:280] + "n")  # 3 distinct left chars, lcp 280
    seq = "".join(parts)
    with open(path, "w") as fh:
        fh.write(">wide\n" + seq + "\n")


def synth_multi(path):
    """Several sequences sharing repeats across separators."""
    rng = np.random.Generator(np.random.PCG64(13))
    rep = "".join("acgt"[c] for c in rng.integers(0, 4, 25))
    with open(path, "w") as fh:
        for k in range(9):
            body = "".join("acgt"[c] for c in rng.integers(0, 4, 120))
            if k % 2 == 0:
                body = rep + body          # repeat right after a separator
            if k % 3 == 0:
                body = body + rep          # ... and right before one
            fh.write(">s%d\n%s\n" % (k, body))


FIXTURES = [
    # name, source (path or generator), suffixerator flags, minimum lengths
    ("random", TESTDATA + "/Random.fna", ["-dna"], [1, 2, 8, 10, 13, 14, 20]),
    ("random_uint", TESTDATA + "/Random.fna", ["-dna", "-suftabuint"], [10]),
    ("random_parts3", TESTDATA + "/Random.fna", ["-dna", "-parts", "3"], [10]),
    ("randomn", TESTDATA + "/RandomN.fna", ["-dna"], [1, 8, 10]),
    ("random_small", TESTDATA + "/Random-Small.fna", ["-dna"], [1, 2]),
    ("random159", TESTDATA + "/Random159.fna", ["-dna"], [1, 2, 3]),
    ("random160", TESTDATA + "/Random160.fna", ["-dna"], [1, 2, 3]),
    ("atinsert", TESTDATA + "/Atinsert.fna", ["-dna"], [1, 8, 10, 14, 20, 50]),
    ("atinsert_mirrored", TESTDATA + "/Atinsert.fna", ["-dna", "-mirrored"], [8, 14, 20]),
    ("duplicate", TESTDATA + "/Duplicate.fna", ["-dna"], [1, 5, 20, 100]),
    ("ttt_small", TESTDATA + "/TTT-small.fna", ["-dna"], [1, 2]),
    ("repfind_example", TESTDATA + "/Repfind-example.fna", ["-dna"], [1, 2, 4]),
    ("trna", TESTDATA + "/trna_glutamine.fna", ["-dna"], [1, 3]),
    ("sw100k1", TESTDATA + "/sw100K1.fsa", ["-protein"], [1, 2, 3, 5, 8]),
    ("sw100k2", TESTDATA + "/sw100K2.fsa", ["-protein"], [2, 4]),
    ("llv", synth_llv, ["-dna"], [1, 20, 100, 254, 255, 256, 300, 599, 600, 601]),
    ("wide", synth_wide, ["-dna"], [1, 10, 39, 40, 41, 255, 279, 280, 281, 300, 301]),
    ("multi", synth_multi, ["-dna"], [1, 5, 20, 25, 26]),
    # real genomic DNA (108 kbp, wildcards, real repeat structure): 7 tiles of the scan kernel
    ("u89959", TESTDATA + "/U89959_genomic.fas", ["-dna"], [8, 12, 16, 20, 30]),
]


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, **kw)


def make_fixture(name, source, flags, minlengths, outdir):
    with tempfile.TemporaryDirectory() as tmp:
        if callable(source):
            fasta = os.path.join(tmp, name + ".fna")
            source(fasta)
        else:
            fasta = source
        idx = os.path.join(tmp, "idx")
        run([GTREF, "suffixerator", "-db", fasta, "-suf", "-lcp", "-bwt", "-tis",
             "-indexname", idx] + flags)
        # the same index with the sequence separator table, for the reference's seqnum/relpos
        idx_ssp = os.path.join(tmp, "idx_ssp")
        run([GTREF, "suffixerator", "-db", fasta, "-suf", "-lcp", "-bwt", "-tis", "-ssp",
             "-indexname", idx_ssp] + flags)
        files = {}
        for sfx in SUFFIXES:
            with open(idx + sfx, "rb") as fh:
                files["file" + sfx.replace(".", "_")] = np.frombuffer(fh.read(), dtype=np.uint8)
        # the absolute path of the input is baked into .prj ("dbfile=") and .esq; keep as is
        tabs = O.load_esa(idx, mmap=False)
        expected = {}
        uint_suftab = "-suftabuint" in flags
        for m in minlengths:
            outs = []
            # the reference's mapped loader rejects 4-byte suffix tables
            # (esa-map.c:351-355): those are only readable with -scan
            modes = [("smax-bu", ["scan"]), ("smax-lin", ["scan"])]
            if not uint_suftab:
                modes += [("smax-bu", []), ("smax-lin", [])]
            for tool, extra in modes:
                outs.append(run([GTREF, tool, idx, str(m)] + extra).stdout)
            assert all(o == outs[0] for o in outs), (name, m, "reference variants disagree")
            # cross-check the restatements before committing anything
            for algo in ("linear", "stack"):
                recs = O.smax_c(tabs.lcp, tabs.llv, tabs.bwt, m, 0, algo)
                txt = O.format_abs(recs, O.positions_c(tabs.suf, recs))
                assert txt == outs[0], (name, m, algo, "C restatement != reference run")
            recs = O.smax_numpy(tabs.lcp, tabs.llv, tabs.bwt, m, 0)
            assert O.format_abs(recs, O.gather_positions(tabs.suf, recs)) == outs[0], (name, m)
            expected["exp_gt_%d" % m] = np.frombuffer(outs[0], dtype=np.uint8)
            rel_extra = ["scan"] if uint_suftab else []
            assert run([GTREF, "smax-bu", idx_ssp, str(m)] + rel_extra).stdout == outs[0], (name, m, "-ssp index differs")
            rel = run([GTREF, "smax-bu", idx_ssp, str(m), "rel"] + rel_extra).stdout
            assert run([GTREF, "smax-lin", idx_ssp, str(m), "rel"] + rel_extra).stdout == rel, (name, m)
            expected["exp_rel_%d" % m] = np.frombuffer(rel, dtype=np.uint8)
            precs = O.smax_c(tabs.lcp, tabs.llv, tabs.bwt, m, 1, "linear")
            ptxt = O.format_abs(precs, O.positions_c(tabs.suf, precs))
            expected["exp_plain_%d" % m] = np.frombuffer(ptxt, dtype=np.uint8)
        np.savez_compressed(os.path.join(outdir, name + ".npz"),
                            minlengths=np.array(minlengths, dtype=np.int64),
                            flags=np.array(" ".join(flags)), **files, **expected)
        nrep = {m: expected["exp_gt_%d" % m].tobytes().count(b"\n") for m in minlengths}
        print("%-18s n=%-7d llv=%-5d repeats=%s" % (name, tabs.n, tabs.llv.shape[0], nrep))


def main():
    if not os.path.exists(GTREF):
        sys.exit("build the reference driver first: make -f oracle/Makefile.ref -j8")
    O.build_c_oracle()
    only = set(sys.argv[1:])
    for name, source, flags, minlengths in FIXTURES:
        if not only or name in only:
            make_fixture(name, source, flags, minlengths, HERE)


if __name__ == "__main__":
    main()
