"""C-ABI boundary on the CPU: the library loads, exports every symbol that
include/smax.h declares, the loader reproduces the reference loader's checks
and the tool parses options with GenomeTools' error texts.  No compute calls."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, Golden, golden_names


def header_symbols():
    text = open(os.path.join(ROOT, "include", "smax.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smax_[a-z0-9_]+)\s*\(", text)) - {"smax_emit_cb"})


def test_library_exports_every_declared_symbol(libsmax):
    lib = libsmax.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "libsmax.so does not export %s" % name
    # the ctypes mirror covers the header too
    assert set(declared) == set(libsmax.SIGNATURES)


def test_no_gpu_fails_loudly(libsmax):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(libsmax.SmaxError):
        libsmax.Device(0)
    g = Golden("random")
    t = g.tables()
    idx = libsmax.Index.from_arrays(t.lcp, t.bwt, t.llv, t.suf)
    with pytest.raises(libsmax.SmaxError, match="CUDA"):
        idx.run_records(10)      # no CPU fallback behind the boundary
    with pytest.raises(libsmax.SmaxError, match="CUDA"):
        idx.run_text(10)
    idx.close()


@pytest.mark.parametrize("name", golden_names())
def test_loader_matches_tables(name, tmp_path, libsmax):
    g = Golden(name)
    base = g.materialise(tmp_path)
    t = g.tables()
    with libsmax.Index.open(base) as idx:
        info = idx.info()
        n = t.n
        assert info.numberofallsortedsuffixes == n
        assert info.totallength == t.prj["totallength"]
        assert info.nonspecials == t.prj["totallength"] - t.prj["specialcharacters"]
        assert info.largelcpvalues == t.llv.shape[0]
        assert info.maxbranchdepth == t.prj["maxbranchdepth"]
        assert info.mirrored == t.prj["mirrored"]
        assert info.sufbytes == t.suf.dtype.itemsize
        assert info.numofchars == (20 if "-protein" in g.flags else 4)
        lib = libsmax.lib()
        for getter, arr in (("smax_index_lcptab", t.lcp), ("smax_index_bwttab", t.bwt),
                            ("smax_index_suftab", t.suf), ("smax_index_llvtab", t.llv)):
            p = getattr(lib, getter)(idx.handle)
            if arr.size == 0:
                continue
            raw = ctypes.string_at(p, arr.nbytes)
            assert raw == arr.tobytes(), getter


def test_loader_errors(tmp_path, libsmax):
    g = Golden("random")
    base = g.materialise(tmp_path)
    with pytest.raises(libsmax.SmaxError, match=r"cannot open file '.*nosuch\.esq'"):
        libsmax.Index.open(str(tmp_path / "nosuch"))
    # truncated .lcp -> the reference's size-check text (fa.c:703-722)
    with open(base + ".lcp", "r+b") as fh:
        fh.truncate(100)
    with pytest.raises(libsmax.SmaxError, match=r"mapping file .*\.lcp: number of mapped units "
                                                  r"\(of size 1\)  = 100 != 10004 = expected"):
        libsmax.Index.open(base)
    base = g.materialise(tmp_path)
    prj = open(base + ".prj").read()
    open(base + ".prj", "w").write(prj.replace("integersize=64", "integersize=32"))
    with pytest.raises(libsmax.SmaxError, match="index was generated for 32-bit integers"):
        libsmax.Index.open(base)
    open(base + ".prj", "w").write(prj.replace("readmode=0", "readmode=7"))
    with pytest.raises(libsmax.SmaxError, match="illegal readmode 7"):
        libsmax.Index.open(base)
    open(base + ".prj", "w").write("\n".join(l for l in prj.splitlines()
                                             if not l.startswith("longest=")) + "\n")
    with pytest.raises(libsmax.SmaxError, match="longest not defined"):
        libsmax.Index.open(base)
    open(base + ".prj", "w").write("\n".join(l for l in prj.splitlines()
                                             if not l.startswith("totallength=")) + "\n")
    with pytest.raises(libsmax.SmaxError, match=r'missing line beginning with "totallength="'):
        libsmax.Index.open(base)


def run_tool(libsmax, *args):
    p = subprocess.run([libsmax.TOOL_PATH] + list(args), capture_output=True, text=True)
    return p.returncode, p.stdout, p.stderr


def test_tool_option_errors(tmp_path, libsmax):
    """Messages of /root/reference/src/core/option.c and gt_repfind.c:521-525."""
    rc, out, err = run_tool(libsmax)
    assert rc == 1 and err == 'gt smax: error: option "-ii" is mandatory\n'
    rc, out, err = run_tool(libsmax, "-ii")
    assert rc == 1 and err == 'gt smax: error: missing argument to option "-ii"\n'
    rc, out, err = run_tool(libsmax, "-l", "0", "-ii", "x")
    assert rc == 1 and err == 'gt smax: error: argument to option "-l" must be an integer >= 1\n'
    rc, out, err = run_tool(libsmax, "-foo")
    assert rc == 1 and err == "gt smax: error: unknown option: -foo (-help shows possible options)\n"
    rc, out, err = run_tool(libsmax, "-abs", "-rel", "-ii", "x")
    assert rc == 1 and err == 'gt smax: error: option "-abs" and option "-rel" exclude each other\n'
    rc, out, err = run_tool(libsmax, "-l", "5", "-l", "6", "-ii", "x")
    assert rc == 1 and err == 'gt smax: error: option "l" already set\n'
    rc, out, err = run_tool(libsmax, "-ii", "x", "extra")
    assert rc == 1 and err == 'gt smax: error: superfluous arguments: "extra"\n'
    rc, out, err = run_tool(libsmax, "-ii", str(tmp_path / "missing"))
    assert rc == 1 and "cannot open file" in err and err.startswith("gt smax: error: fopen(): ")
    rc, out, err = run_tool(libsmax, "-policy", "weird", "-ii", "x")
    assert rc == 1 and 'argument to option "-policy" must be one of: gt, plain' in err
    rc, out, err = run_tool(libsmax, "-help")
    assert rc == 0 and out.startswith("Usage: gt smax [options] -ii indexname\n"
                                      "Compute supermaximal repeats.\n") and err == ""
    assert "-l " in out and "default: 20" in out
    rc, out, err = run_tool(libsmax, "--version")
    assert rc == 0 and "gt smax" in out


def test_emitter_formats(tmp_path, libsmax):
    """The text formats are a single switch; rel positions follow gt_encseq_seqnum."""
    g = Golden("multi")
    base = g.materialise(tmp_path)
    t = g.tables()
    lib = libsmax.lib()
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    # separator positions from the sequence lengths of the fixture generator
    seps = np.sort(t.suf[t.bwt == 255].astype(np.int64) - 1)

    def emit(fmt, relative, recs):
        path = str(tmp_path / ("out_%d_%d.txt" % (fmt, relative)))
        fp = libc.fopen(path.encode(), b"w")
        with libsmax.Index.open(base) as idx:
            opts = libsmax.Opts(minlength=1, relative=relative, ngpus=1, policy=0, format=fmt)
            em, err = ctypes.c_void_p(), ctypes.create_string_buffer(1024)
            assert lib.smax_emitter_new(idx.handle, ctypes.byref(opts), fp, ctypes.byref(em), err,
                                        1024) == 0
            for length, lb, pos in recs:
                arr = (ctypes.c_uint64 * len(pos))(*pos)
                assert lib.smax_emitter_emit(em, length, lb, len(pos), arr) == 0
            assert lib.smax_emitter_delete(em) == 0
        libc.fclose(fp)
        return open(path).read()

    recs = [(25, 3, [0, int(seps[1]) + 1, 700]), (7, 10, [5, 6])]
    assert emit(libsmax.FORMAT_SMAX, 0, recs) == "25 3 0 %d 700\n7 2 5 6\n" % (seps[1] + 1)
    assert emit(libsmax.FORMAT_ITV, 0, recs) == "25 3 5\n7 10 11\n"
    assert emit(libsmax.FORMAT_PAIRS, 0, recs[1:]) == "7 5 F 7 6\n"

    def rel(p):
        k = int(np.searchsorted(seps, p, side="left"))
        return k, p - (int(seps[k - 1]) + 1 if k else 0)

    want = "25 3 " + " ".join("%d %d" % rel(p) for p in recs[0][2]) + "\n"
    assert emit(libsmax.FORMAT_SMAX, 1, recs[:1]) == want
    assert rel(int(seps[1]) + 1) == (2, 0) and rel(0) == (0, 0)


def test_gather_positions_matches_oracle(libsmax, c_oracle):
    """smax_index_gather_positions (host side of the end-to-end path): suf[lb..lb+width) per record,
    8- and 4-byte suffix tables, and the range check."""
    O = c_oracle
    g = Golden("wide")
    t = g.tables()
    want = O.smax_c(t.lcp, t.llv, t.bwt, 10)
    for suf in (t.suf.astype(np.uint64), t.suf.astype(np.uint32)):
        idx = libsmax.Index.from_arrays(t.lcp, t.bwt, t.llv, suf)
        assert np.array_equal(idx.gather_positions(want), O.positions_c(t.suf, want))
        bad = want[:1].copy()
        bad["lb"] = len(t.lcp)
        with pytest.raises(libsmax.SmaxError, match="outside"):
            idx.gather_positions(bad)
        idx.close()


def test_emitter_batch_matches_python_grammar(libsmax):
    """smax_emitter_emit_records (the host reference of the device formatter) against a
    plain-Python statement of the line grammar, incl. relative positions on a real index."""
    from util import render_text
    rng = np.random.default_rng(5)
    idx = libsmax.Index.from_arrays(np.zeros(1, np.uint8), np.zeros(1, np.uint8))
    try:
        recs = np.zeros(500, libsmax.REC_DTYPE)
        recs["len"] = rng.integers(1, 2**63, 500, dtype=np.uint64) >> rng.integers(0, 63, 500).astype(np.uint64)
        recs["width"] = rng.integers(2, 40, 500)
        recs["lb"] = np.cumsum(recs["width"]) - recs["width"]
        pos = rng.integers(0, 2**63, int(recs["width"].sum()), dtype=np.uint64) >> \
            rng.integers(0, 63, int(recs["width"].sum())).astype(np.uint64)
        assert idx.emit_text(recs, pos) == render_text(recs, pos)
        assert idx.emit_text(recs, None, libsmax.FORMAT_ITV) == render_text(recs, None, "itv")
        assert idx.emit_text(recs[:0], pos[:0]) == b""
    finally:
        idx.close()


@pytest.mark.parametrize("name", golden_names())
def test_relative_positions_match_the_reference(name, tmp_path, libsmax, c_oracle):
    """-rel is pinned by reference code: the host emitter's '<seqnum> <relpos>' (separators
    recovered from the tables, smax_index_seqnum_relpos) against the lines the reference's own
    gt_encseq_seqnum / gt_encseq_seqstartpos print for the same repeats (tests/golden/*.npz
    exp_rel_<m>, /root/reference/src/core/encseq.c:3815-3900), multi-sequence, protein and
    -mirrored indexes included."""
    O = c_oracle
    g = Golden(name)
    t = g.tables()
    idx = libsmax.Index.open(g.materialise(tmp_path))
    try:
        for m in g.minlengths:
            recs = O.smax_c(t.lcp, t.llv, t.bwt, m, 0)
            pos = O.positions_c(t.suf, recs)
            assert idx.emit_text(recs, pos, libsmax.FORMAT_SMAX, True) == g.expected_rel(m), (name, m)
    finally:
        idx.close()


def test_emit_and_scan_options_without_gpu(tmp_path, libsmax):
    """Argument checks of -emit / -scan and the file errors of the streaming mode happen
    before any device call, so they can be pinned here."""
    g = Golden("random")
    base = g.materialise(tmp_path)
    rc, out, err = run_tool(libsmax, "-ii", base, "-emit", "gpu")
    assert rc == 1 and 'argument to option "-emit" must be one of: host, device' in err
    rc, out, err = run_tool(libsmax, "-ii", base, "-emit", "device", "-format", "pairs")
    assert rc == 1 and err.startswith('gt smax: error: option "-emit device" renders the formats smax and itv')
    rc, out, err = run_tool(libsmax, "-ii", base, "-scan", "-emit", "device")
    assert rc == 1 and err == 'gt smax: error: option "-scan" and option "-emit device" exclude each other\n'
    rc, out, err = run_tool(libsmax, "-help")
    assert "-emit " in out and "default: host" in out and "-scan " in out
    # streaming mode: tables are opened as files; a missing or truncated one is reported
    idx = libsmax.Index.open(base, libsmax.TAB_ESQ)
    try:
        os.rename(base + ".lcp", base + ".lcp.away")
        with pytest.raises(libsmax.SmaxError, match=r"fopen\(\): cannot open file '.*\.lcp'"):
            idx.run_stream_text(g.minlengths[0])
        with open(base + ".lcp", "wb") as fh:
            fh.write(b"\0" * 7)
        with pytest.raises(libsmax.SmaxError, match=r"number of units .* expected number of units"):
            idx.run_stream_text(g.minlengths[0])
        os.replace(base + ".lcp.away", base + ".lcp")
        # larger than the largest lcp value: empty answer without touching a table or a device
        assert idx.run_stream_text(10 ** 6) == b""
        import torch
        if not torch.cuda.is_available():           # the streaming mode has no CPU fallback either
            with pytest.raises(libsmax.SmaxError, match="CUDA"):
                idx.run_stream_text(g.minlengths[0])
        # pairs cannot be rendered on the device
        with pytest.raises(libsmax.SmaxError, match="host emitter"):
            idx.run_text(10 ** 6, fmt=libsmax.FORMAT_PAIRS)
    finally:
        idx.close()
