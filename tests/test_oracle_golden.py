"""Oracle pinned against the committed golden vectors (CPU only).

The golden expectations were produced by REFERENCE code (oracle/_ref/gtref:
gt_esa_bottomup + reader macros of /root/reference) on indexes built by the
reference suffixerator -- see tests/golden/make_golden.py.
"""
import numpy as np
import pytest

from conftest import Golden, golden_names
from util import fuzz_tables

# the 19 lcp local maxima with l >= 10 of testdata/Random.fna
# (len, lb, rb, verdict under the GenomeTools convention / plain mask)
KNOWN_ANSWER = [
    (10, 2, 3, True, True), (10, 126, 127, True, True), (11, 363, 364, True, True),
    (10, 398, 399, True, True), (11, 483, 484, True, True), (10, 1147, 1150, False, False),
    (10, 1498, 1499, False, False), (10, 1568, 1569, True, True), (11, 1662, 1663, True, True),
    (11, 3713, 3714, False, False), (10, 3894, 3895, True, True), (10, 4575, 4576, True, True),
    (12, 4577, 4578, False, False), (10, 4580, 4581, True, True), (10, 4591, 4592, True, True),
    (10, 4629, 4631, True, False), (13, 4862, 4863, True, False), (10, 4962, 4963, False, False),
    (10, 4997, 4998, True, True),
]


@pytest.mark.parametrize("name", golden_names())
def test_restatements_match_golden(name, c_oracle):
    O = c_oracle
    g = Golden(name)
    t = g.tables()
    for m in g.minlengths:
        for policy, pname in ((0, "gt"), (1, "plain")):
            want = g.expected(m, pname)
            for algo in ("linear", "stack"):
                recs = O.smax_c(t.lcp, t.llv, t.bwt, m, policy, algo)
                assert O.format_abs(recs, O.positions_c(t.suf, recs)) == want, (name, m, algo)
            recs = O.smax_numpy(t.lcp, t.llv, t.bwt, m, policy)
            assert O.format_abs(recs, O.gather_positions(t.suf, recs)) == want, (name, m, "numpy")


def test_known_answer_vector_random_fna(c_oracle):
    O = c_oracle
    t = Golden("random").tables()
    for policy in (0, 1):
        recs = O.smax_c(t.lcp, t.llv, t.bwt, 10, policy)
        got = {(int(r["len"]), int(r["lb"]), int(r["lb"] + r["width"] - 1)) for r in recs}
        want = {(l, lb, rb) for l, lb, rb, gt, plain in KNOWN_ANSWER if (gt, plain)[policy]}
        assert got == want
    assert len(O.smax_c(t.lcp, t.llv, t.bwt, 10, 0)) == 14
    assert len(O.smax_c(t.lcp, t.llv, t.bwt, 10, 1)) == 12


@pytest.mark.parametrize("kind", ["dense", "alternating", "plateaus", "large", "huge",
                                  "widerun", "sparse"])
def test_restatements_agree_on_fuzzed_tables(kind, c_oracle):
    O = c_oracle
    rng = np.random.default_rng(hash(kind) % 2**32)
    for n in (1, 2, 17, 1000, 40000):
        lcp, llv, bwt = fuzz_tables(rng, n, kind)
        for m in (1, 3, 20, 255, 256):
            for policy in (0, 1):
                a = O.smax_c(lcp, llv, bwt, m, policy, "linear")
                b = O.smax_c(lcp, llv, bwt, m, policy, "stack")
                c = O.smax_numpy(lcp, llv, bwt, m, policy)
                assert np.array_equal(a, b) and np.array_equal(a, c), (kind, n, m, policy)


def test_properties(c_oracle):
    """Size-independent properties: disjoint, ascending, monotone in minlength."""
    O = c_oracle
    rng = np.random.default_rng(5)
    lcp, llv, bwt = fuzz_tables(rng, 200000, "plateaus")
    prev = None
    for m in (1, 5, 10, 20, 30):
        r = O.smax_c(lcp, llv, bwt, m)
        assert np.all(r["len"] >= m) and np.all(r["width"] >= 2)
        assert np.all(r["lb"][1:] >= r["lb"][:-1] + r["width"][:-1] - 0)  # disjoint, sorted
        if prev is not None:   # raising the threshold only removes repeats
            assert np.array_equal(r, prev[prev["len"] >= m])
        prev = r
