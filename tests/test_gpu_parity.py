"""Parity of the CUDA path against the oracle and the golden vectors (B200).

Everything goes through the C ABI of libsmax.so (genometools_smax_b200.capi is
a thin ctypes mirror).  Bit-exact: every length, left boundary, width and
position, and the emitted text byte for byte.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GTREF, Golden, golden_names
from util import fuzz_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["ring", "units"])
def scan_kernel(request, monkeypatch):
    """Every case of this file runs through BOTH scan kernels (the unit kernel is the one in
    production; SMAX_KERNEL forces it here).  The unit-kernel runs also take the stripped upload
    of the .llv records (values only, records rebuilt on the device from the 255 bytes of the lcp
    table) from the first record on; the ring-kernel runs upload the records whole."""
    monkeypatch.setenv("SMAX_KERNEL", request.param)
    monkeypatch.setenv("SMAX_LLV_STRIP", "1" if request.param == "units" else "0")
    return request.param


@pytest.fixture(scope="module")
def dev(libsmax):
    d = libsmax.Device(0)
    yield d
    d.close()


def scan(libsmax, dev, lcp, llv, bwt, suf, m, policy=0, lo=0, hi=None):
    idx = libsmax.Index.from_arrays(lcp, bwt, llv, suf)
    try:
        dev.upload(idx, lo, hi, with_suf=suf is not None)
        dev.scan(m, policy, gather=suf is not None)
        return dev.fetch()
    finally:
        idx.close()


@pytest.mark.parametrize("name", golden_names())
def test_golden_device_scan(name, dev, libsmax, c_oracle):
    O = c_oracle
    g = Golden(name)
    t = g.tables()
    for m in g.minlengths:
        for policy, pname in ((0, "gt"), (1, "plain")):
            recs, pos = scan(libsmax, dev, t.lcp, t.llv, t.bwt, t.suf, m, policy)
            want = O.smax_c(t.lcp, t.llv, t.bwt, m, policy)
            assert np.array_equal(recs, want), (name, m, pname)
            assert np.array_equal(pos, O.positions_c(t.suf, want)), (name, m, pname)
            assert O.format_abs(recs, pos) == g.expected(m, pname), (name, m, pname)


@pytest.mark.parametrize("name", golden_names())
def test_golden_tool_output(name, tmp_path, libsmax):
    """The `smax` tool on the re-materialised index files == reference-run text."""
    g = Golden(name)
    base = g.materialise(tmp_path)
    # every run is a process of its own (CUDA start-up dominates): two lengths per index
    for m in g.minlengths[:2]:
        for pname in ("gt", "plain"):
            p = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base, "-policy", pname],
                               capture_output=True)
            assert p.returncode == 0, p.stderr
            assert p.stdout == g.expected(m, pname), (name, m, pname)


def test_tool_formats_and_shards(tmp_path, libsmax, c_oracle):
    O = c_oracle
    g = Golden("atinsert")
    base = g.materialise(tmp_path)
    t = g.tables()
    recs = O.smax_c(t.lcp, t.llv, t.bwt, 10)
    itv = "".join("%d %d %d\n" % (r["len"], r["lb"], r["lb"] + r["width"] - 1) for r in recs)
    p = subprocess.run([libsmax.TOOL_PATH, "-l", "10", "-ii", base, "-format", "itv"],
                       capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == itv
    # relative positions against gt_encseq_seqnum semantics recomputed from the tables
    seps = np.sort(t.suf[t.bwt == 255].astype(np.int64) - 1)
    pos = O.positions_c(t.suf, recs)
    lines, o = [], 0
    for r in recs:
        w = int(r["width"])
        f = []
        for q in pos[o:o + w].astype(np.int64):
            k = int(np.searchsorted(seps, q, side="left"))
            f.append("%d %d" % (k, q - (seps[k - 1] + 1 if k else 0)))
        lines.append("%d %d %s\n" % (r["len"], w, " ".join(f)))
        o += w
    p = subprocess.run([libsmax.TOOL_PATH, "-l", "10", "-ii", base, "-rel"], capture_output=True,
                       text=True)
    assert p.returncode == 0 and p.stdout == "".join(lines)
    p = subprocess.run([libsmax.TOOL_PATH, "-l", "10", "-ii", base, "-v"], capture_output=True,
                       text=True)
    assert p.returncode == 0 and p.stdout.startswith("# indexname=")
    assert "".join(l + "\n" for l in p.stdout.splitlines() if not l.startswith("#")) \
        == g.expected(10).decode()


@pytest.mark.parametrize("kind", ["dense", "alternating", "plateaus", "large", "huge", "widerun",
                                  "sparse"])
def test_fuzzed_tables(kind, dev, libsmax, c_oracle):
    """Arbitrary tables: many plateaus per tile (stage overflow), runs wider than a
    tile, 255-runs, values beyond 32 bits, ragged sizes around the tile size."""
    O = c_oracle
    rng = np.random.default_rng(abs(hash(kind)) % 2**32)
    for n in (1, 2, 15, 16, 17, 4095, 16384, 16385, 70001, 300000):
        lcp, llv, bwt = fuzz_tables(rng, n, kind)
        suf = rng.permutation(n).astype(np.uint64)
        for m in (1, 3, 20, 255, 256, 1 << 33):
            for policy in (0, 1):
                recs, pos = scan(libsmax, dev, lcp, llv, bwt, suf, m, policy)
                want = O.smax_c(lcp, llv, bwt, m, policy)
                assert np.array_equal(recs, want), (kind, n, m, policy, len(recs), len(want))
                assert np.array_equal(pos, O.positions_c(suf, want)), (kind, n, m, policy)


@pytest.mark.parametrize("ctas", [1, 2, 5])
@pytest.mark.parametrize("kind", ["plateaus", "sparse", "large", "dense", "widerun"])
def test_few_ctas_walk_many_tiles(kind, ctas, libsmax, c_oracle):
    """With the grid limited to a few CTAs every CTA works through dozens of tiles: ring buffers and
    mbarrier phases wrap, dense / sparse regions alternate, the survivor log is flushed (and
    compacted) in mid-scan, tiles that overflow it are redone, generations are resolved in batches."""
    O = c_oracle
    rng = np.random.default_rng(1000 * ctas + len(kind))
    n = 1_000_003
    lcp, llv, bwt = fuzz_tables(rng, n, kind)
    if kind in ("plateaus", "sparse"):
        # stretches without any repeat next to busy ones
        lcp = lcp.copy()
        for lo in range(0, n, 200_000):
            lcp[lo + 50_000: lo + 120_000] = 0
        keep = (llv["position"] % 200_000 < 50_000) | (llv["position"] % 200_000 >= 120_000)
        llv = llv[keep]
    suf = rng.permutation(n).astype(np.uint64)
    idx = libsmax.Index.from_arrays(lcp, bwt, llv, suf)
    d = libsmax.Device(0)
    try:
        d.upload(idx, 0, None, True)
        d.set_grid_limit(ctas)
        for m in (1, 4, 12, 255, 300):
            d.scan(m, 0, True)
            recs, pos = d.fetch()
            want = O.smax_c(lcp, llv, bwt, m)
            assert np.array_equal(recs, want), (kind, ctas, m, len(recs), len(want))
            assert np.array_equal(pos, O.positions_c(suf, want)), (kind, ctas, m)
    finally:
        d.close()
        idx.close()


def test_suftab_uint32(dev, libsmax, c_oracle):
    O = c_oracle
    rng = np.random.default_rng(3)
    lcp, llv, bwt = fuzz_tables(rng, 50000, "plateaus")
    suf = rng.permutation(50000).astype(np.uint32)
    recs, pos = scan(libsmax, dev, lcp, llv, bwt, suf, 5)
    want = O.smax_c(lcp, llv, bwt, 5)
    assert np.array_equal(recs, want) and np.array_equal(pos, O.positions_c(suf, want))


def test_repeated_scans_are_idempotent(dev, libsmax, c_oracle):
    """Epoch-tagged look-back state and the ping-pong result block need no reset."""
    O = c_oracle
    rng = np.random.default_rng(4)
    lcp, llv, bwt = fuzz_tables(rng, 400000, "plateaus")
    suf = np.arange(400000, dtype=np.uint64)
    idx = libsmax.Index.from_arrays(lcp, bwt, llv, suf)
    dev.upload(idx, 0, None, True)
    for m in (4, 9, 4, 30, 4, 1, 1, 1):
        dev.scan(m, 0, True)
        recs, pos = dev.fetch()
        want = O.smax_c(lcp, llv, bwt, m)
        assert np.array_equal(recs, want) and np.array_equal(pos, O.positions_c(suf, want)), m
    idx.close()


@pytest.mark.parametrize("nshards", [2, 3, 8])
@pytest.mark.parametrize("kind", ["plateaus", "large", "widerun", "sparse"])
def test_sharded_scan_on_one_gpu(nshards, kind, libsmax, c_oracle):
    """N shards of the SA range as N device contexts on cuda:0, left views wired
    like the multi-GPU driver does: the concatenation must equal the 1-shard scan."""
    O = c_oracle
    rng = np.random.default_rng(nshards * 100 + len(kind))
    n = 200000 + int(rng.integers(0, 5000))
    lcp, llv, bwt = fuzz_tables(rng, n, kind)
    suf = rng.permutation(n).astype(np.uint64)
    idx = libsmax.Index.from_arrays(lcp, bwt, llv, suf)
    cuts = [((n // nshards) * g) & ~15 for g in range(nshards)] + [n]
    devs = [libsmax.Device(0) for _ in range(nshards)]
    try:
        views = []
        for g, d in enumerate(devs):
            d.upload(idx, cuts[g], cuts[g + 1], True)
            if g:
                d.set_left_views(views[:g])
            views.append(d.view())
        # one-sided count exchange: every shard's kernel stores its count into all arrays
        ptrs = [d.counts_export(nshards)[1] for d in devs]
        for g, d in enumerate(devs):
            d.counts_connect(g, nshards, ptrs=ptrs)
        for tag, m in enumerate((1, 7, 255, 300), start=1):
            for d in devs:
                d.set_exchange_tag(tag)
                d.scan(m, 0, True)
            parts = [d.fetch() for d in devs]
            for d in (devs[0], devs[-1]):
                assert d.peer_counts(tag, nshards) == [len(p[0]) for p in parts], (kind, nshards, m)
            recs = np.concatenate([p[0] for p in parts])
            pos = np.concatenate([p[1] for p in parts])
            want = O.smax_c(lcp, llv, bwt, m)
            assert np.array_equal(recs, want), (kind, nshards, m)
            assert np.array_equal(pos, O.positions_c(suf, want)), (kind, nshards, m)
    finally:
        for d in devs:
            d.close()
        idx.close()


@pytest.mark.parametrize("base", [3_000_000_000 & ~15, (1 << 32) + 4096, (1 << 33) + (1 << 31)])
@pytest.mark.parametrize("kind", ["plateaus", "large", "sparse"])
def test_window_beyond_32_bits(kind, base, libsmax, c_oracle):
    """A shard whose SA indices, .llv positions and suffix positions lie beyond 2^31 / 2^32 (the
    range of config C5): a window of a table with n_total = base + len entries
    (smax_index_from_memory_window), scanned as one shard and as three."""
    O = c_oracle
    rng = np.random.default_rng(base % 1000 + len(kind))
    n = 300000 + int(rng.integers(0, 999))
    lcp, llv, bwt = fuzz_tables(rng, n, kind)
    suf = (rng.permutation(n).astype(np.uint64) * np.uint64(9973) + np.uint64(base))
    llv_g = llv.copy()
    llv_g["position"] += np.uint64(base)
    idx = libsmax.Index.from_pointers(lcp.ctypes.data, bwt.ctypes.data, llv_g.ctypes.data, len(llv_g),
                                      suf.ctypes.data, 8, n, keep=(lcp, bwt, llv_g, suf),
                                      base=base, n_total=base + n)
    devs = [libsmax.Device(0) for _ in range(3)]
    try:
        for m in (1, 7, 255, 300):
            want = O.smax_c(lcp, llv, bwt, m)
            wpos = O.positions_c(suf, want)
            want = want.copy()
            want["lb"] += np.uint64(base)
            devs[0].upload(idx, base, base + n, True)
            devs[0].set_left_views([])
            devs[0].scan(m, 0, True)
            recs, pos = devs[0].fetch()
            assert np.array_equal(recs, want), (kind, base, m)
            assert np.array_equal(pos, wpos), (kind, base, m)
        cuts = [base, base + (n // 3 & ~15), base + (2 * n // 3 & ~15), base + n]
        views = []
        for g, d in enumerate(devs):
            d.upload(idx, cuts[g], cuts[g + 1], True)
            d.set_left_views(views[:g])
            views.append(d.view())
        for m in (3, 256):
            want = O.smax_c(lcp, llv, bwt, m)
            wpos = O.positions_c(suf, want)
            want = want.copy()
            want["lb"] += np.uint64(base)
            for d in devs:
                d.scan(m, 0, True)
            parts = [d.fetch() for d in devs]
            assert np.array_equal(np.concatenate([p[0] for p in parts]), want), (kind, base, m)
            assert np.array_equal(np.concatenate([p[1] for p in parts]), wpos), (kind, base, m)
    finally:
        for d in devs:
            d.close()
        idx.close()


def test_smax_run_callback_and_multi_gpu_arg(libsmax, c_oracle):
    O = c_oracle
    g = Golden("wide")
    t = g.tables()
    idx = libsmax.Index.from_arrays(t.lcp, t.bwt, t.llv, t.suf)
    res = idx.run(10)
    want = O.smax_c(t.lcp, t.llv, t.bwt, 10)
    wpos = O.positions_c(t.suf, want)
    assert [(r[0], r[1], r[2]) for r in res] == [(int(a), int(b), int(c)) for a, b, c in want]
    assert [p for r in res for p in r[3]] == [int(p) for p in wpos]
    with pytest.raises(libsmax.SmaxError, match="GPU"):
        idx.run_records(10, ngpus=libsmax.device_count() + 1)
    idx.close()


@pytest.mark.skipif(not os.path.exists(GTREF), reason="oracle/_ref/gtref not present")
def test_live_reference_index(tmp_path, libsmax):
    """FASTA -> reference suffixerator -> smax tool vs reference sweep, on the box."""
    rng = np.random.default_rng(77)
    seq = rng.choice(list("acgt"), size=200000)
    for _ in range(300):
        L = int(rng.integers(15, 400))
        a, b = rng.integers(0, 200000 - L, 2)
        seq[b:b + L] = seq[a:a + L]
    seq[rng.random(200000) < 0.001] = "n"
    fasta = str(tmp_path / "x.fa")
    open(fasta, "w").write(">x\n" + "".join(seq) + "\n")
    base = str(tmp_path / "x")
    subprocess.run([GTREF, "suffixerator", "-db", fasta, "-dna", "-suf", "-lcp", "-bwt", "-tis",
                    "-indexname", base], check=True, capture_output=True)
    for m in (12, 20, 100, 255, 300):
        ref = subprocess.run([GTREF, "smax-bu", base, str(m)], check=True, capture_output=True)
        got = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base], capture_output=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == ref.stdout, m


def test_randomised_soak_short():
    """tests/soak.py for 25 s: random table kinds / sizes / quiet stretches / grid limits / shard counts /
    minimum lengths / policies against the C oracle (seeds 1 and 2 found two log-flush bugs in round 1)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tests", "soak.py"), "25", "2"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "soak ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_default_kernel_is_the_unit_kernel(dev, libsmax, c_oracle, monkeypatch):
    """Without SMAX_KERNEL every index goes through the unit kernel (the faster one on every
    workload measured; the ring kernel stays behind SMAX_KERNEL=ring as the second implementation)."""
    monkeypatch.delenv("SMAX_KERNEL")
    O = c_oracle
    rng = np.random.default_rng(77)
    for kind, m, want_kernel in (("sparse", 20, "units"), ("dense", 1, "units"), ("large", 3, "units")):
        lcp, llv, bwt = fuzz_tables(rng, 200_000, kind)
        suf = rng.permutation(len(lcp)).astype(np.uint64)
        dev.set_stats(True)
        recs, pos = scan(libsmax, dev, lcp, llv, bwt, suf, m)
        st = dev.stats()
        dev.set_stats(False)
        want = O.smax_c(lcp, llv, bwt, m)
        assert np.array_equal(recs, want), (kind, m)
        assert st["kernel"] == want_kernel, (kind, m, st["kernel"])
