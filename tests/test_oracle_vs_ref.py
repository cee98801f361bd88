"""Oracle vs REFERENCE CODE run live (CPU; needs oracle/_ref/gtref, skipped otherwise).

Random sequences -> reference suffixerator -> reference sweep (smax-bu) and
reference reader macros (smax-lin) vs the restatements; plus the necessary
condition against `repfind` (every pair of occurrences of a reported repeat
is a maximal pair of that length, esa-maxpairs.c).
"""
import itertools
import os
import subprocess

import numpy as np
import pytest

from conftest import GTREF

pytestmark = pytest.mark.skipif(not os.path.exists(GTREF), reason="oracle/_ref/gtref not built")


def write_fasta(path, rng, nseq, length, alphabet, wildcard, wild_rate, repeats):
    with open(path, "w") as fh:
        for s in range(nseq):
            seq = rng.choice(list(alphabet), size=length)
            for _ in range(repeats):
                L = int(rng.integers(5, 60))
                a, b = rng.integers(0, length - L, 2)
                seq[b:b + L] = seq[a:a + L]
            seq[rng.random(length) < wild_rate] = wildcard
            fh.write(">s%d\n%s\n" % (s, "".join(seq)))


@pytest.mark.parametrize("case", [
    dict(nseq=1, length=3000, alphabet="acgt", wildcard="n", wild_rate=0.0, repeats=20, flags=["-dna"]),
    dict(nseq=5, length=800, alphabet="acgt", wildcard="n", wild_rate=0.02, repeats=10, flags=["-dna"]),
    dict(nseq=3, length=900, alphabet="ac", wildcard="n", wild_rate=0.1, repeats=5, flags=["-dna", "-mirrored"]),
    dict(nseq=4, length=700, alphabet="LVIFKREDAGSTNQYWPHMC", wildcard="X", wild_rate=0.01, repeats=10,
         flags=["-protein"]),
])
def test_random_sequences(case, tmp_path, c_oracle):
    O = c_oracle
    rng = np.random.default_rng(1234)
    fasta = str(tmp_path / "in.fa")
    write_fasta(fasta, rng, case["nseq"], case["length"], case["alphabet"], case["wildcard"],
                case["wild_rate"], case["repeats"])
    idx = str(tmp_path / "idx")
    subprocess.run([GTREF, "suffixerator", "-db", fasta, "-suf", "-lcp", "-bwt", "-tis", "-ssp",
                    "-indexname", idx] + case["flags"], check=True, capture_output=True)
    t = O.load_esa(idx, mmap=False)
    for m in (1, 4, 8, 12, 30):
        ref_bu = subprocess.run([GTREF, "smax-bu", idx, str(m)], check=True, capture_output=True).stdout
        ref_lin = subprocess.run([GTREF, "smax-lin", idx, str(m), "scan"], check=True,
                                 capture_output=True).stdout
        assert ref_bu == ref_lin
        for algo in ("linear", "stack"):
            recs = O.smax_c(t.lcp, t.llv, t.bwt, m, 0, algo)
            assert O.format_abs(recs, O.positions_c(t.suf, recs)) == ref_bu, (m, algo)
    # necessary condition against the reference's maximal pairs
    m = 8
    pairs = set()
    out = subprocess.run([GTREF, "repfind", idx, str(m)], check=True, capture_output=True).stdout
    for line in out.decode().splitlines():
        l, p1, p2 = (int(x) for x in line.split())
        pairs.add((l, p1, p2))
    recs = O.smax_c(t.lcp, t.llv, t.bwt, m, 0)
    pos = O.positions_c(t.suf, recs)
    o = 0
    for r in recs:
        w = int(r["width"])
        for a, b in itertools.combinations(sorted(int(p) for p in pos[o:o + w]), 2):
            assert (int(r["len"]), a, b) in pairs, (r, a, b)
        o += w


@pytest.mark.skipif(not os.path.exists("/root/reference/testdata/at1MB"),
                    reason="the reference's testdata is not mounted")
def test_real_dna_one_megabase(tmp_path, c_oracle):
    """testdata/at1MB (1 Mbp of A. thaliana): reference suffixerator ->
    reference sweep and reader macros vs both restatements, mapped and -scan."""
    O = c_oracle
    idx = str(tmp_path / "at1mb")
    subprocess.run([GTREF, "suffixerator", "-db", "/root/reference/testdata/at1MB", "-dna", "-suf",
                    "-lcp", "-bwt", "-tis", "-indexname", idx], check=True, capture_output=True)
    t = O.load_esa(idx, mmap=False)
    for m in (12, 20, 40):
        ref_bu = subprocess.run([GTREF, "smax-bu", idx, str(m)], check=True, capture_output=True).stdout
        ref_lin = subprocess.run([GTREF, "smax-lin", idx, str(m), "scan"], check=True,
                                 capture_output=True).stdout
        assert ref_bu == ref_lin and ref_bu.count(b"\n") > 0
        for algo in ("linear", "stack"):
            recs = O.smax_c(t.lcp, t.llv, t.bwt, m, 0, algo)
            assert O.format_abs(recs, O.positions_c(t.suf, recs)) == ref_bu, (m, algo)
