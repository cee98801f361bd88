"""The bench's torch ESA builder (tools/esa_build_torch.py) reproduces, bit for
bit, the tables the REFERENCE suffixerator wrote for every golden fixture
(sequence reconstructed from suf+bwt).  CPU torch here; the same code runs on
the GPU in bench.py."""
import numpy as np
import pytest
import torch

from conftest import Golden, golden_names
from tools.esa_build_torch import build_esa, llv_records, mirror_codes


def sequence_from_tables(t):
    n = t.n
    seq = np.zeros(n - 1, dtype=np.uint8)
    suf = t.suf.astype(np.int64)
    sel = suf >= 1
    seq[suf[sel] - 1] = t.bwt[sel]
    return seq


# (chunk, block): the defaults, and sizes small enough that every golden index is refined in many
# chunks of the previous order / compared in many blocks (the paths a 3e9-suffix index takes)
@pytest.mark.parametrize("sizes", [(1 << 30, 1 << 27, "levels"), (257, 1000, "levels"), (257, 1000, "compare")],
                         ids=["default", "chunked", "chunked-compare"])
@pytest.mark.parametrize("name", golden_names())
def test_builder_matches_reference_tables(name, sizes):
    t = Golden(name).tables()
    seq = sequence_from_tables(t)
    out = build_esa(torch.from_numpy(seq), chunk=sizes[0], block=sizes[1], lcp_method=sizes[2])
    assert np.array_equal(out["suf"].astype(np.uint64), t.suf.astype(np.uint64))
    assert np.array_equal(out["lcp"], t.lcp)
    assert np.array_equal(out["bwt"], t.bwt)
    assert np.array_equal(llv_records(out["llv_pos"], out["llv_val"]), t.llv)
    assert out["maxlcp"] == t.prj["maxbranchdepth"]


@pytest.mark.parametrize("name", ["random", "multi", "atinsert", "sw100k1", "llv"])
def test_builder_sliced_first_level(name, monkeypatch):
    """Above INT_MAX elements the first level is a counting sort and every pass over all n elements
    goes in slices (what an index of 3e9 suffixes takes): forced here with slices of 1000 elements."""
    import tools.esa_build_torch as B
    monkeypatch.setattr(B, "_SLICE", 1000)
    t = Golden(name).tables()
    seq = sequence_from_tables(t)
    out = B.build_esa(torch.from_numpy(seq), chunk=257, block=1000, lcp_method="compare")
    assert np.array_equal(out["suf"].astype(np.uint64), t.suf.astype(np.uint64))
    assert np.array_equal(out["lcp"], t.lcp)
    assert np.array_equal(out["bwt"], t.bwt)
    assert np.array_equal(llv_records(out["llv_pos"], out["llv_val"]), t.llv)


def test_mirror_matches_reference_mirrored_index():
    plain = Golden("atinsert").tables()
    mirrored = Golden("atinsert_mirrored").tables()
    seq = torch.from_numpy(sequence_from_tables(plain))
    out = build_esa(mirror_codes(seq))
    assert np.array_equal(out["suf"].astype(np.uint64), mirrored.suf)
    assert np.array_equal(out["lcp"], mirrored.lcp)
    assert np.array_equal(out["bwt"], mirrored.bwt)


def test_generators_give_indexable_sequences():
    """The workload generators of the bench (tools/synth.py) never put two separators next to
    each other or at either end: the reference suffixerator rejects a file with an empty
    sequence (the CPU baseline of `bench.py --workload C4` builds its sample with it)."""
    from tools import synth
    for name in ("C4", "C5"):
        cfg = synth.WORKLOADS[name]
        seq = cfg["gen"](2_000_000, cfg["seed"])
        seps = np.flatnonzero(seq == synth.SEPARATOR)
        assert seps.size > 0 and seps[0] > 0 and seps[-1] < seq.shape[0] - 1, name
        assert (np.diff(seps) > 1).all(), name
