"""The emit path on the device (text formatting of records, relative
positions through a device-side separator table) against the host emitter of libsmax, a
plain-Python statement of the grammar, and the reference-run golden text.  Byte for byte,
through the C ABI (smax_scan_format / smax_scan_fetch_text / smax_device_*_separators /
smax_run_text / the tool's -emit device)."""
import subprocess

import numpy as np
import pytest

from conftest import Golden, golden_names
from util import fuzz_tables, render_text

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(libsmax):
    d = libsmax.Device(0)
    yield d
    d.close()


@pytest.mark.parametrize("name", golden_names())
def test_golden_device_text(name, tmp_path, dev, libsmax, c_oracle):
    O = c_oracle
    g = Golden(name)
    t = g.tables()
    base = g.materialise(tmp_path)
    idx = libsmax.Index.open(base)
    try:
        seps = idx.separators()
        want_seps = np.sort(t.suf[t.bwt == 255].astype(np.uint64) - np.uint64(1))
        assert np.array_equal(seps, want_seps)
        dev.upload(idx, with_suf=True)
        assert np.array_equal(dev.build_separators(), want_seps)
        for m in g.minlengths:
            dev.scan(m, 0, gather=True)
            recs, pos = dev.fetch()
            assert np.array_equal(recs, O.smax_c(t.lcp, t.llv, t.bwt, m, 0))
            text = dev.format_text(libsmax.FORMAT_SMAX, False)
            assert text == g.expected(m, "gt"), (name, m)
            assert text == idx.emit_text(recs, pos, libsmax.FORMAT_SMAX, False)
            rel = dev.format_text(libsmax.FORMAT_SMAX, True)
            assert rel == g.expected_rel(m), (name, m)           # the reference's own seqnum / relpos
            assert rel == idx.emit_text(recs, pos, libsmax.FORMAT_SMAX, True), (name, m)
            assert rel == render_text(recs, pos, "smax", want_seps), (name, m)
            itv = dev.format_text(libsmax.FORMAT_ITV, False)
            assert itv == idx.emit_text(recs, None, libsmax.FORMAT_ITV, False)
            assert itv == render_text(recs, None, "itv")
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["multi", "random_uint", "atinsert_mirrored"])
def test_tool_emit_device_matches_host(name, tmp_path, libsmax):
    g = Golden(name)
    base = g.materialise(tmp_path)
    m = g.minlengths[0]
    for extra in ([], ["-rel"], ["-format", "itv"], ["-policy", "plain"]):
        host = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base] + extra,
                              capture_output=True)
        devi = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base, "-emit", "device"]
                              + extra, capture_output=True)
        assert host.returncode == 0 and devi.returncode == 0, (host.stderr, devi.stderr)
        assert host.stdout == devi.stdout, (name, extra)
        if not extra:
            assert devi.stdout == g.expected(m, "gt")
        if extra == ["-rel"]:
            assert devi.stdout == g.expected_rel(m)      # the reference's own seqnum / relpos
    p = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base, "-emit", "device",
                        "-format", "pairs"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stderr.startswith("gt smax: error: ")


@pytest.mark.parametrize("kind", ["dense", "plateaus", "large", "huge", "widerun", "sparse"])
def test_fuzzed_tables_text(kind, dev, libsmax, c_oracle):
    """Many records, wide records, 20-digit values, empty results; separators supplied by
    the caller (ragged: none, one, many)."""
    O = c_oracle
    rng = np.random.default_rng(abs(hash("fmt" + kind)) % 2**32)
    for n in (1, 17, 4097, 70001, 300000):
        lcp, llv, bwt = fuzz_tables(rng, n, kind)
        suf = rng.permutation(n).astype(np.uint64)
        if kind == "huge":       # positions with up to 20 digits
            suf = suf + np.uint64(2**64 - 1 - n)
        idx = libsmax.Index.from_arrays(lcp, bwt, llv, suf)
        try:
            dev.upload(idx, with_suf=True)
            for m in (1, 3, 20, 256):
                dev.scan(m, 0, gather=True)
                recs, pos = dev.fetch()
                want = O.smax_c(lcp, llv, bwt, m, 0)
                assert np.array_equal(recs, want)
                assert dev.format_text(libsmax.FORMAT_SMAX, False) == O.format_abs(recs, pos)
                assert dev.format_text(libsmax.FORMAT_ITV, False) == render_text(recs, None, "itv")
                for nsep in (0, 1, max(1, n // 50)):
                    lo = int(suf.min())
                    seps = np.unique(rng.integers(0, max(n, 1), nsep).astype(np.uint64)) + np.uint64(lo)
                    dev.set_separators(seps)
                    got = dev.format_text(libsmax.FORMAT_SMAX, True)
                    assert got == render_text(recs, pos, "smax", seps), (kind, n, m, nsep)
                dev.set_separators(np.zeros(0, np.uint64))
        finally:
            idx.close()


def test_run_text_matches_run(tmp_path, libsmax):
    g = Golden("sw100k1")
    base = g.materialise(tmp_path)
    idx = libsmax.Index.open(base)
    try:
        for m in g.minlengths[:2]:
            assert idx.run_text(m) == g.expected(m, "gt")
            recs = idx.run_records(m)
            pos = idx.gather_positions(recs)
            assert idx.run_text(m, relative=True) == idx.emit_text(recs, pos, libsmax.FORMAT_SMAX, True)
            assert idx.run_text(m, fmt=libsmax.FORMAT_ITV) == render_text(recs, None, "itv")
        if libsmax.device_count() >= 2:
            assert idx.run_text(g.minlengths[0], ngpus=2) == g.expected(g.minlengths[0], "gt")
        with pytest.raises(libsmax.SmaxError):
            idx.run_text(g.minlengths[0], fmt=libsmax.FORMAT_PAIRS)
    finally:
        idx.close()


def test_format_errors(dev, libsmax):
    lcp = np.array([0, 3, 3, 1, 0], np.uint8)
    bwt = np.array([1, 2, 3, 1, 0], np.uint8)
    idx = libsmax.Index.from_arrays(lcp, bwt, None, None)
    try:
        dev.upload(idx, with_suf=False)
        dev.scan(1, 0, gather=False)
        with pytest.raises(libsmax.SmaxError):
            dev.format_text(libsmax.FORMAT_SMAX)          # no positions were gathered
        with pytest.raises(libsmax.SmaxError):
            dev.format_text(libsmax.FORMAT_PAIRS)
        with pytest.raises(libsmax.SmaxError):
            dev.build_separators()                        # no resident suffix table
        recs, _ = dev.fetch()
        assert dev.format_text(libsmax.FORMAT_ITV) == render_text(recs, None, "itv")
        with pytest.raises(libsmax.SmaxError):
            dev.set_separators(np.array([5, 5], np.uint64))   # not strictly ascending
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["wide", "atinsert_mirrored", "random_uint", "multi"])
def test_two_gpus_match_one(name, tmp_path, libsmax):
    """smax_run_records / smax_run_text with the SA range sharded over two GPUs (peer views
    over NVLink for plateaus that cross the cut) == one GPU.  Needs a box with two GPUs."""
    if libsmax.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = Golden(name)
    base = g.materialise(tmp_path)
    idx = libsmax.Index.open(base)
    try:
        for m in g.minlengths[:3]:
            one = idx.run_records(m, ngpus=1)
            two = idx.run_records(m, ngpus=2)
            assert np.array_equal(one, two), (name, m)
            assert idx.run_text(m, ngpus=2) == g.expected(m, "gt"), (name, m)
            assert idx.run_text(m, ngpus=2, relative=True) == idx.run_text(m, ngpus=1, relative=True)
            assert idx.run_text(m, ngpus=2, fmt=libsmax.FORMAT_ITV) == render_text(one, None, "itv")
        p = subprocess.run([libsmax.TOOL_PATH, "-l", str(g.minlengths[0]), "-ii", base, "-gpus", "2"],
                           capture_output=True)
        assert p.returncode == 0 and p.stdout == g.expected(g.minlengths[0], "gt")
    finally:
        idx.close()
