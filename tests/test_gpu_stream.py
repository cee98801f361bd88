"""The -scan mode (smax_run_stream): tables read chunk by chunk from the
index files, never mapped; three chunks resident on one GPU at a time.  Same bytes as the
mapped path and as the reference-run golden text, for every chunk size."""
import subprocess

import numpy as np
import pytest

from conftest import Golden, golden_names
from util import render_text, wide_plateau_tables, write_index_files

pytestmark = pytest.mark.gpu


def longest_run(lcp):
    """Length of the longest run of equal non-zero lcp bytes (an upper bound of the width of
    any plateau a scan with minlength >= 1 walks)."""
    if len(lcp) == 0:
        return 0
    cut = np.flatnonzero(np.diff(lcp.astype(np.int16)) != 0)
    edges = np.concatenate(([-1], cut, [len(lcp) - 1]))
    lens, vals = np.diff(edges), lcp[edges[1:]]
    return int(lens[vals > 0].max()) if (vals > 0).any() else 0


@pytest.mark.parametrize("name", golden_names())
def test_stream_matches_golden_for_every_chunk_size(name, tmp_path, libsmax):
    g = Golden(name)
    base = g.materialise(tmp_path)
    t = g.tables()
    run = longest_run(t.lcp)
    idx = libsmax.Index.open(base, libsmax.TAB_ESQ)        # no table is mapped
    try:
        # every index at the smallest chunk and at the default; the sizes in between on the
        # indexes whose plateaus are wider than a chunk
        chunks = (1024, 2048, 4096 + 16, 0) if run > 512 else (1024, 0)
        for chunk in chunks:
            for m in g.minlengths[:2]:
                # wider plateaus than two chunks + halo: the chunk is redone with a wider window
                got = idx.run_stream_text(m, chunk)
                assert got == g.expected(m, "gt"), (name, chunk, m)
        m = g.minlengths[0]
        assert idx.run_stream_text(m, 2048, policy=libsmax.POLICY_PLAIN) \
            == g.expected(m, "plain")
    finally:
        idx.close()


@pytest.mark.parametrize("name", ["multi", "random_uint", "atinsert_mirrored"])
def test_tool_scan_option(name, tmp_path, libsmax):
    g = Golden(name)
    base = g.materialise(tmp_path)
    m = g.minlengths[0]
    for extra in ([], ["-rel"], ["-format", "itv"], ["-format", "pairs"]):
        a = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base] + extra, capture_output=True)
        b = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base, "-scan"] + extra,
                           capture_output=True)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert a.stdout == b.stdout, (name, extra)
        if extra == ["-rel"]:
            assert b.stdout == g.expected_rel(m)         # the reference's own seqnum / relpos
    p = subprocess.run([libsmax.TOOL_PATH, "-l", str(m), "-ii", base, "-scan", "-emit", "device"],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "exclude each other" in p.stderr


def test_stream_relative_and_itv_small_chunks(tmp_path, libsmax, c_oracle):
    O = c_oracle
    g = Golden("atinsert_mirrored")
    base = g.materialise(tmp_path)
    t = g.tables()
    m = g.minlengths[0]
    recs = O.smax_c(t.lcp, t.llv, t.bwt, m, 0)
    pos = O.positions_c(t.suf, recs)
    seps = np.sort(t.suf[t.bwt == 255].astype(np.uint64) - np.uint64(1))
    idx = libsmax.Index.open(base, libsmax.TAB_ESQ)
    try:
        assert idx.run_stream_text(m, 1024, relative=True) == render_text(recs, pos, "smax", seps)
        assert idx.run_stream_text(m, 1024, fmt=libsmax.FORMAT_ITV) == render_text(recs, None, "itv")
    finally:
        idx.close()


def test_stream_redoes_chunks_with_plateaus_wider_than_the_resident_range(tmp_path, libsmax, c_oracle):
    """Plateaus of 6002 and 3001 entries against chunks of 1024 / 4096 suffixes: the chunk that
    ends such a plateau is redone with a wider window (smax_device_upload_halo)."""
    O = c_oracle
    rng = np.random.default_rng(7)
    lcp, llv, bwt = wide_plateau_tables(rng)
    suf = rng.permutation(len(lcp)).astype(np.uint64)
    base = str(tmp_path / "wideplateau")
    write_index_files(base, lcp, bwt, llv, suf)
    idx = libsmax.Index.open(base, 0)
    try:
        for m in (1, 1000):
            recs = O.smax_c(lcp, llv, bwt, m, 0)
            want = render_text(recs, O.positions_c(suf, recs))
            assert recs["width"].max() > 3000
            for chunk in (1024, 4096, 0):
                assert idx.run_stream_text(m, chunk) == want, (m, chunk)
    finally:
        idx.close()


def test_stream_missing_file(tmp_path, libsmax):
    import os
    g = Golden("random")
    base = g.materialise(tmp_path)
    idx = libsmax.Index.open(base, libsmax.TAB_ESQ)
    os.remove(base + ".bwt")
    try:
        with pytest.raises(libsmax.SmaxError, match="cannot open file"):
            idx.run_stream_text(g.minlengths[0], 0)
    finally:
        idx.close()
