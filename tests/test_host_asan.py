"""The HOST side of libsmax (loader, smax_run, the -scan chunk driver smax_run_stream, the
emitter) end to end without a GPU: the C sources are compiled with AddressSanitizer + UBSan
and linked with tests/host_stub_device.c, a CPU stand-in for the device half of the C ABI
(test infrastructure; same shard / left-view / resident-range contract as the device manager).
Every golden index, mapped and streamed at several chunk sizes, must print the reference-run
text; any memory error aborts the run (cf. the reference's `testsuite.rb -memcheck` and
GT_MEM_BOOKKEEPING runs)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, Golden, golden_names
from util import render_text, wide_plateau_tables, write_index_files

CSRC = os.path.join(ROOT, "genometools_smax_b200", "csrc")
HOST_SOURCES = ["smax_index.c", "smax_run.c", "smax_emit.c", "smax_stream.c"]


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("asan") / "host_driver")
    # (-DTAB_READ_SLICE: the -scan driver cuts every table read of >= 8 KiB into slices read by
    # several threads -- in production that starts at 16 MiB)
    cmd = ["gcc", "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined",
           "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-Wall", "-Wextra", "-Werror",
           "-DTAB_READ_SLICE=4096", "-pthread",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += [os.path.join(CSRC, s) for s in HOST_SOURCES]
    cmd += [os.path.join(ROOT, "tests", "host_stub_device.c"),
            os.path.join(ROOT, "tests", "host_driver.c"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    """The `smax` tool itself (option parser + runner) over the stub device."""
    exe = str(tmp_path_factory.mktemp("asan_tool") / "smax")
    cmd = ["gcc", "-std=gnu99", "-g", "-O1", "-fsanitize=address,undefined",
           "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-Wall", "-Wextra", "-Werror",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += [os.path.join(CSRC, s) for s in HOST_SOURCES + ["smax_tool.c", "smax_main.c"]]
    cmd += [os.path.join(ROOT, "tests", "host_stub_device.c"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def run(driver, *args, devices=1, max_shard=0, pipeline=0):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1",
               SMAX_STUB_DEVICES=str(devices))
    if max_shard:
        env["SMAX_MAX_SHARD"] = str(max_shard)      # test hook of smax_run.c
    if pipeline:
        env["SMAX_PIPELINE"] = str(pipeline)        # shards per device of the upload / scan / emit pipeline
    return subprocess.run([driver] + [str(a) for a in args], capture_output=True, env=env)


def longest_run(lcp):
    cut = np.flatnonzero(np.diff(lcp.astype(np.int16)) != 0)
    edges = np.concatenate(([-1], cut, [len(lcp) - 1]))
    lens, vals = np.diff(edges), lcp[edges[1:]]
    return int(lens[vals > 0].max()) if (vals > 0).any() else 0


@pytest.mark.parametrize("name", golden_names())
def test_host_paths_under_sanitizers(name, tmp_path, driver):
    g = Golden(name)
    base = g.materialise(tmp_path)
    uint_suf = "-suftabuint" in g.flags
    for m in g.minlengths[:3]:
        want = g.expected(m, "gt")
        p = run(driver, base, m, "map", 0, "smax", 0)
        assert p.returncode == 0 and p.stdout == want, (name, m, p.stderr[-800:])
        for chunk in (1024, 2048, 5000, 0):
            # a plateau that reaches back over more than two chunks makes the driver redo the
            # chunk with a wider window: every chunk size gives the same bytes
            p = run(driver, base, m, "stream", chunk, "smax", 0)
            assert p.returncode == 0 and p.stdout == want, (name, m, chunk, p.stderr[-800:])
    m = g.minlengths[0]
    assert run(driver, base, m, "map", 0, "smax", 0, "plain").stdout == g.expected(m, "plain")
    assert run(driver, base, m, "stream", 1024, "smax", 0, "plain").stdout == g.expected(m, "plain")
    # the other renderings: mapped and streamed must agree byte for byte
    for fmt, rel in (("smax", 1), ("itv", 0), ("pairs", 0), ("pairs", 1)):
        a = run(driver, base, m, "map", 0, fmt, rel)
        b = run(driver, base, m, "stream", 0, fmt, rel)
        assert a.returncode == 0 and b.returncode == 0, (name, fmt, rel, a.stderr[-500:], b.stderr[-500:])
        assert a.stdout == b.stdout, (name, fmt, rel)
    assert uint_suf or run(driver, base, m, "map", 0, "itv", 0).stdout.count(b"\n") == \
        g.expected(m, "gt").count(b"\n")


@pytest.mark.parametrize("name", golden_names())
def test_shard_driver_on_stub_devices(name, tmp_path, driver):
    """smax_run with the SA range cut into 2, 3 and 8 shards (one stub device each, left views
    wired by the driver): the concatenation of the shards' records is the 1-shard answer."""
    g = Golden(name)
    base = g.materialise(tmp_path)
    for m in g.minlengths[:2]:
        for ngpus in (2, 3, 8):
            p = run(driver, base, m, "map", 0, "smax", 0, "gt", ngpus, devices=8)
            assert p.returncode == 0 and p.stdout == g.expected(m, "gt"), (name, m, ngpus, p.stderr[-500:])
    p = run(driver, base, g.minlengths[0], "map", 0, "smax", 0, "gt", 9, devices=8)
    assert p.returncode == 1 and b"9 GPU(s) requested" in p.stderr
    # an index larger than one shard may be (the kernel keeps 32-bit tile offsets; the limit is
    # lowered to 1024 suffixes here) is cut into several shards per device: up to 105 shards on
    # one device, or on three, each with its (at most 8) nearest left neighbours as views
    m = g.minlengths[0]
    n = len(g.tables().lcp)
    for ngpus in (1, 3):
        p = run(driver, base, m, "map", 0, "smax", 0, "gt", ngpus, devices=3, max_shard=1024)
        assert p.returncode == 0 and p.stdout == g.expected(m, "gt"), (name, ngpus, p.stderr[-500:])
    if n > 4096:
        p = run(driver, base, m, "map", 0, "smax", 0, "gt", 1, max_shard=n // 300 + 1024 if n > 300000 else 1024)
        assert p.returncode == 0


@pytest.mark.parametrize("name", ["atinsert", "u89959", "llv", "wide", "multi"])
def test_pipelined_shards(name, tmp_path, driver):
    """Several shards per device, uploaded by one thread per device while the calling thread
    launches and consumes them in order: same bytes for every format."""
    g = Golden(name)
    base = g.materialise(tmp_path)
    m = g.minlengths[0]
    for ngpus, pipeline in ((1, 2), (1, 5), (2, 3), (3, 4)):
        for fmt, rel in (("smax", 0), ("smax", 1), ("itv", 0)):
            a = run(driver, base, m, "map", 0, fmt, rel, "gt", 1)
            b = run(driver, base, m, "map", 0, fmt, rel, "gt", ngpus, devices=3, pipeline=pipeline)
            assert a.returncode == 0 and b.returncode == 0, (name, ngpus, pipeline, fmt, b.stderr[-500:])
            assert a.stdout == b.stdout, (name, ngpus, pipeline, fmt, rel)
    assert run(driver, base, m, "map", 0, "smax", 0, "gt", 1, pipeline=4).stdout == g.expected(m, "gt")


def test_shard_with_a_plateau_wider_than_its_views_is_redone(tmp_path, driver, c_oracle):
    """A plateau of 12001 entries against shards of 1024: more than the eight left neighbour views
    cover.  The shard that ends it is made resident again with a wider window (smax_run.c:
    redo_with_wider_halo) instead of failing."""
    O = c_oracle
    rng = np.random.default_rng(11)
    n = 40000
    L = rng.integers(0, 6, n).astype(np.uint64)
    L[5000:17001] = 9
    L[4999] = 2
    L[17001] = 3
    L[0] = 0
    bwt = rng.integers(0, 4, n).astype(np.uint8)
    bwt[4999:17001] = 255
    lcp, llv = np.minimum(L, 255).astype(np.uint8), np.zeros(0, O.LLV_DTYPE)
    suf = rng.permutation(n).astype(np.uint64)
    base = str(tmp_path / "verywide")
    write_index_files(base, lcp, bwt, llv, suf)
    recs = O.smax_c(lcp, llv, bwt, 4, 0)
    assert recs["width"].max() == 12002
    want = render_text(recs, O.positions_c(suf, recs))
    for ngpus in (1, 3):
        p = run(driver, base, 4, "map", 0, "smax", 0, "gt", ngpus, devices=3, max_shard=1024)
        assert p.returncode == 0 and p.stdout == want, (ngpus, p.stderr[-500:])


def test_stream_redoes_chunks_with_plateaus_wider_than_the_resident_range(tmp_path, driver, c_oracle):
    """Plateaus of 6002 and 3001 entries against chunks of 1024 ... 4096 suffixes: the chunk that
    ends such a plateau is redone with a wider window; same bytes as the mapped path."""
    O = c_oracle
    rng = np.random.default_rng(7)
    lcp, llv, bwt = wide_plateau_tables(rng)
    suf = rng.permutation(len(lcp)).astype(np.uint64)
    base = str(tmp_path / "wideplateau")
    write_index_files(base, lcp, bwt, llv, suf)
    for m in (1, 8, 1000):
        recs = O.smax_c(lcp, llv, bwt, m, 0)
        want = render_text(recs, O.positions_c(suf, recs))
        assert recs["width"].max() > 3000
        assert run(driver, base, m, "map", 0, "smax", 0).stdout == want
        for chunk in (1024, 2048, 4096, 0):
            p = run(driver, base, m, "stream", chunk, "smax", 0)
            assert p.returncode == 0 and p.stdout == want, (m, chunk, p.stderr[-500:])
    # with the plain policy the wide plateaus are no repeats (their left characters collide)
    recs = O.smax_c(lcp, llv, bwt, 1, 1)
    assert recs["width"].max() < 300
    p = run(driver, base, 1, "stream", 1024, "smax", 0, "plain")
    assert p.returncode == 0 and p.stdout == render_text(recs, O.positions_c(suf, recs))


def test_host_errors_under_sanitizers(tmp_path, driver):
    g = Golden("llv")
    base = g.materialise(tmp_path)
    p = run(driver, str(tmp_path / "missing"), 10, "map", 0, "smax", 0)
    assert p.returncode == 1 and b"cannot open file" in p.stderr
    # truncated tables are rejected by the size checks of both readers
    with open(base + ".llv", "r+b") as fh:
        fh.truncate(os.path.getsize(base + ".llv") - 16)
    for mode in ("map", "stream"):
        p = run(driver, base, 10, mode, 0, "smax", 0)
        assert p.returncode == 1 and b"expected number of" in p.stderr, (mode, p.stderr)
    g.materialise(tmp_path)
    with open(base + ".lcp", "r+b") as fh:       # a 255 entry without its .llv record
        data = bytearray(fh.read())
        data[5] = 255
        fh.seek(0)
        fh.write(data)
    for mode in ("map", "stream"):
        p = run(driver, base, 1, mode, 0, "smax", 0)
        assert p.returncode == 1 and b"inconsistent ESA tables" in p.stderr, (mode, p.stderr)


def test_tool_under_sanitizers(tmp_path, tool):
    g = Golden("atinsert")
    base = g.materialise(tmp_path)
    m = g.minlengths[1]
    want = g.expected(m, "gt")
    assert run(tool, "-l", m, "-ii", base).stdout == want
    assert run(tool, "-l", m, "-ii", base, "-scan").stdout == want
    assert run(tool, "-l", m, "-ii", base, "-gpus", 4, devices=4).stdout == want
    assert run(tool, "-l", m, "-ii", base, "-policy", "plain").stdout == g.expected(m, "plain")
    a = run(tool, "-l", m, "-ii", base, "-rel")
    b = run(tool, "-l", m, "-ii", base, "-rel", "-scan")
    assert a.returncode == 0 and a.stdout == b.stdout and a.stdout != want
    v = run(tool, "-l", m, "-ii", base, "-v")
    assert v.returncode == 0 and v.stdout.startswith(b"# indexname=")
    for args in (["-help"], ["--version"]):
        assert run(tool, *args).returncode == 0
    for args in ([], ["-ii"], ["-l", "0", "-ii", base], ["-foo"], ["-abs", "-rel", "-ii", base],
                 ["-ii", base, "extra"], ["-ii", str(tmp_path / "none")], ["-ii", base, "-emit", "device"],
                 ["-ii", base, "-gpus", 2]):
        p = run(tool, *args)
        assert p.returncode == 1 and p.stderr.startswith(b"gt smax: error: "), (args, p.stderr[-300:])


@pytest.fixture(scope="module")
def driver_tsan(tmp_path_factory):
    """The same host driver under ThreadSanitizer: smax_run's pipeline is threads (one uploader
    and up to four workers per device, the calling thread consumes), -scan reads ahead and cuts
    its table reads into slices."""
    exe = str(tmp_path_factory.mktemp("tsan") / "host_driver")
    cmd = ["gcc", "-std=gnu99", "-g", "-O1", "-fsanitize=thread", "-fno-omit-frame-pointer",
           "-DTAB_READ_SLICE=4096", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += [os.path.join(CSRC, s) for s in HOST_SOURCES]
    cmd += [os.path.join(ROOT, "tests", "host_stub_device.c"),
            os.path.join(ROOT, "tests", "host_driver.c"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


@pytest.mark.parametrize("name", ["u89959", "wide", "multi"])
def test_pipeline_threads_have_no_data_race(name, tmp_path, driver_tsan):
    g = Golden(name)
    base = g.materialise(tmp_path)
    m = g.minlengths[0]
    rel = "1" if name == "multi" else "0"
    want = None
    for ngpus, pipeline, max_shard in ((1, 1, 0), (1, 4, 0), (2, 3, 0), (3, 8, 0), (1, 0, 1024), (3, 0, 1024)):
        env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0", SMAX_STUB_DEVICES="3")
        if pipeline:
            env["SMAX_PIPELINE"] = str(pipeline)
        if max_shard:
            env["SMAX_MAX_SHARD"] = str(max_shard)
        for mode, chunk in (("map", "0"), ("scan", "2048")):
            p = subprocess.run([driver_tsan, base, str(m), mode, chunk, "smax", rel, "gt", str(ngpus)],
                               env=env, capture_output=True)
            want = p.stdout if want is None else want
            assert p.returncode == 0 and p.stdout == want, (ngpus, pipeline, max_shard, mode, p.stderr[-800:])
            assert b"ThreadSanitizer" not in p.stderr, (ngpus, pipeline, max_shard, mode, p.stderr[-1500:])
