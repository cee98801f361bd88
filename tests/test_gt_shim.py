"""The drop-in boundary proven with reference code: integration/gt_smax.c -- the file a
GenomeTools maintainer adds as src/tools/gt_smax.c -- is compiled against the reference
library (oracle/_ref/libgtref.a, unmodified reference sources), registered in a GtToolbox and
run through the reference's own gt_tool_run (/root/reference/src/core/tool.c:62-114) by
integration/gt_smax_harness.c.  oracle/Makefile.ref builds the binary oracle/_ref/gt_smax
(in the build container; it travels to the GPU box with the snapshot).
"""
import os
import subprocess

import pytest

from conftest import ROOT, Golden, golden_names

SHIM = os.path.join(ROOT, "oracle", "_ref", "gt_smax")
needs_shim = pytest.mark.skipif(not os.path.exists(SHIM), reason="oracle/_ref/gt_smax not built "
                                "(make -f oracle/Makefile.ref needs /root/reference)")


def shim(*args):
    return subprocess.run([SHIM, "smax"] + list(args), capture_output=True)


def tool(libsmax, *args):
    return subprocess.run([libsmax.TOOL_PATH] + list(args), capture_output=True)


@needs_shim
def test_parser_and_errors_are_the_reference_parsers(tmp_path, libsmax):
    """Option errors come from the reference's GtOptionParser here; the product tool
    (csrc/smax_tool.c) restates them -- both must say the same, with exit code 1."""
    base = Golden("random").materialise(tmp_path)
    cases = [[], ["-l", "0", "-ii", base], ["-l"], ["-abs", "-rel", "-ii", base], ["-ii", str(tmp_path / "nosuch")],
             ["-ii", base, "extra"], ["-ii", base, "-policy", "foo"], ["-ii", base, "-format", "x"],
             ["-ii", base, "-scan", "-emit", "device"], ["-ii", base, "-emit", "device", "-format", "pairs"],
             ["-nosuchoption"], ["-gpus", "0", "-ii", base]]
    for args in cases:
        a, b = shim(*args), tool(libsmax, *args)
        assert a.returncode == 1 and b.returncode == 1, (args, a.stderr, b.stderr)
        assert a.stderr == b.stderr, (args, a.stderr, b.stderr)
        assert a.stderr.startswith(b"gt smax: error: ") and a.stdout == b""
    # -help: same option table (the version banner differs)
    a, b = shim("-help"), tool(libsmax, "-help")
    assert a.returncode == 0 and b.returncode == 0
    assert a.stdout == b.stdout
    assert shim("-version").returncode == 0
    # an index that cannot be opened: the message of the loader, forwarded through GtError
    with open(base + ".lcp", "ab") as fh:
        fh.write(b"\0")
    a, b = shim("-ii", base), tool(libsmax, "-ii", base)
    assert a.returncode == 1 and a.stderr == b.stderr and b"number of mapped units" in a.stderr


@needs_shim
@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names())
def test_shim_output_is_the_reference_text(name, tmp_path):
    """gt_tool_run(gt_smax()) on the golden index files == the text the reference code
    printed for them (mapped, -scan, -rel and -emit device)."""
    g = Golden(name)
    base = g.materialise(tmp_path)
    for m in g.minlengths[:2]:
        p = shim("-l", str(m), "-ii", base)
        assert p.returncode == 0, p.stderr
        assert p.stdout == g.expected(m, "gt"), (name, m)
    m = g.minlengths[0]
    for extra, want in ((["-scan"], g.expected(m, "gt")), (["-rel"], g.expected_rel(m)),
                        (["-emit", "device"], g.expected(m, "gt")),
                        (["-emit", "device", "-rel"], g.expected_rel(m)),
                        (["-policy", "plain"], g.expected(m, "plain"))):
        p = shim("-l", str(m), "-ii", base, *extra)
        assert p.returncode == 0, (extra, p.stderr)
        assert p.stdout == want, (name, m, extra)
    p = shim("-v", "-l", str(m), "-ii", base)
    assert p.returncode == 0 and p.stdout.startswith(b"# indexname=")
