"""The one-process-per-GPU driver (genometools_smax_b200/shard.py) on CPU: two
ranks over gloo, each with an oracle-backed stand-in for capi.Device.  Checks
the host-side logic of the N > 1 path -- shard cuts, ownership of plateaus by
their END, the count exchange (collective mode) and the global order of the
concatenated results -- without a GPU."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleDevice:
    """Same methods as capi.Device, results from the C oracle (test infrastructure)."""

    ordinal = 0

    def __init__(self, oracle):
        self.O = oracle
        self.uploads = 0

    def upload(self, index, lo, hi, with_suf=True):
        self.lcp, self.llv, self.bwt, self.suf = index
        self.lo, self.hi = lo, hi
        self.uploads += 1
        return int(hi - lo) * 2

    def ipc_export(self):
        return b"", b""

    def ipc_import(self, handles, view):
        return None

    def set_left_views(self, views):
        self.nleft = len(views)

    def scan(self, minlength, policy=0, gather=True, stream=0):
        want = self.O.smax_c(self.lcp, self.llv, self.bwt, minlength, policy)
        ends = want["lb"] + want["width"] - 1
        self.recs = want[(ends >= self.lo) & (ends < self.hi)]     # a shard owns the ENDS in its range
        self.pos = self.O.positions_c(self.suf, self.recs) if gather else np.zeros(0, np.uint64)

    def copy_count(self, d_dst, stream=0):
        ctypes.c_int64.from_address(d_dst).value = len(self.recs)

    def counts(self):
        return len(self.recs), len(self.pos)

    def fetch(self):
        return self.recs, self.pos


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import smax_oracle as O
        from util import fuzz_tables
        from genometools_smax_b200.shard import ShardedScan, gather_results, shard_cuts
        O.build_c_oracle()
        rng = np.random.default_rng(1234)                     # same tables on every rank
        n = 70001
        lcp, llv, bwt = fuzz_tables(rng, n, kind)
        suf = rng.permutation(n).astype(np.uint64)
        scan = ShardedScan(OracleDevice(O), rank, world, count_device=torch.device("cpu"))
        assert scan.exchange == "collective"
        scan.load((lcp, llv, bwt, suf), n, with_suf=True)
        assert scan.cuts == shard_cuts(n, world) and scan.cuts[0] == 0 and scan.cuts[-1] == n
        assert all(c % 16 == 0 for c in scan.cuts[:-1])
        for m in (1, 5, 300):
            scan.launch(m, 0, True, 0)
            off, total = scan.offsets()
            recs, pos = scan.fetch()
            want = O.smax_c(lcp, llv, bwt, m)
            assert total == len(want), (m, total, len(want))
            # this rank's records sit at [off, off + len) of the global order
            assert np.array_equal(recs, want[off:off + len(recs)]), (rank, m)
            allrecs, allpos = gather_results(scan, recs, pos)
            assert np.array_equal(allrecs, want), m
            assert np.array_equal(allpos, O.positions_c(suf, want)), m
        if rank == 0:
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["plateaus", "widerun"])
def test_sharded_driver_world2_gloo(kind, tmp_path):
    out = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, _free_port(), kind, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
