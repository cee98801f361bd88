"""One-process-per-GPU driver of the sharded scan (torchrun / torch.distributed).

The suffix-array index range is cut into `world` contiguous shards; rank r makes
shard r resident on its GPU (`capi.Device.upload`).  Plateaus that cross a cut
are resolved by the shard that owns their END: its kernel walks left into the
neighbours' tables through CUDA-IPC mapped peer pointers (P2P loads over
NVLink, no staging copy).  The only exchange is that of the per-shard record
counts (8 bytes per rank) that turns local output positions into global ones,
so the concatenation over ranks is in suffix-array order -- the order in which
the reference's sweep reports intervals
(/root/reference/src/match/esa-bottomup.c:160-170).  Two ways:

  "p2p"         (default on GPUs) the scan kernel itself stores its count, tagged
                with the step, into every shard's count array -- P2P stores over
                NVLink, no collective launch on the step's critical path;
  "collective"  one 8-byte all_gather per scan (NCCL on GPUs, gloo in the CPU
                tests).

torch.distributed is plumbing here: rendezvous, one object all_gather for the
IPC handles at load time, and the count all_gather of the collective mode.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

MAX_LEFT = 8


def shard_cuts(n: int, world: int) -> List[int]:
    """Cut points of the lcp index space [0, n): multiples of 16 (128-bit loads
    stay aligned), last cut = n.  Same rule as smax_run.c."""
    return [((n // world) * g) & ~15 for g in range(world)] + [n]


def balanced_cuts(n: int, world: int, llv_positions=None, max_shard: int = (1 << 32) - (1 << 20)) -> List[int]:
    """Cuts of [0, n) of equal COST instead of equal length -- the rule of
    smax_run.c (balanced_cuts): a shard reads one byte per entry and 16 bytes per
    large value, cost(x) = x + 16 * #{.llv positions < x}; `llv_positions` is the
    ascending position column of the .llv table (numpy or torch, any device).
    Falls back to equal lengths when a shard would exceed what a device holds."""
    if llv_positions is None or world == 1 or len(llv_positions) == 0:
        return shard_cuts(n, world)
    pos = llv_positions if isinstance(llv_positions, torch.Tensor) else torch.as_tensor(llv_positions)
    pos = pos.to(torch.int64)
    total = n + 16 * int(pos.shape[0])
    cuts = [0]
    for g in range(1, world):
        want = total * g // world
        lo, hi = 0, n
        while lo < hi:                          # smallest x with x + 16 * rank(x) >= want
            x = (lo + hi) // 2
            r = int(torch.searchsorted(pos, torch.tensor([x], dtype=torch.int64, device=pos.device))[0])
            if x + 16 * r < want:
                lo = x + 1
            else:
                hi = x
        cuts.append(max(lo & ~15, cuts[-1]))
    cuts.append(n)
    if any(b - a > max_shard for a, b in zip(cuts[:-1], cuts[1:])):
        return shard_cuts(n, world)
    return cuts


class ShardedScan:
    """Rank-local half of a scan over `world` shards.

    `device` is a capi.Device (or any object with the same upload / ipc_export /
    ipc_import / set_left_views / scan / copy_count / counts / fetch methods --
    the CPU tests plug in an oracle-backed stand-in to exercise this logic on
    gloo without a GPU)."""

    def __init__(self, device, rank: int, world: int, group=None, count_device=None,
                 exchange: str = "auto"):
        self.device = device
        self.rank = rank
        self.world = world
        self.group = group
        if exchange == "auto":
            exchange = "p2p" if hasattr(device, "counts_export") and count_device is None else "collective"
        self.exchange = exchange if world > 1 else "none"
        self.step = 0
        dev = count_device if count_device is not None else torch.device("cuda", device.ordinal)
        self._count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._all = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        self.cuts = None

    # ---------------------------------------------------------------- load
    def load(self, index, n_total: int, with_suf: bool = True, cuts=None) -> int:
        """Upload this rank's shard of `index` (whole table or a window that
        covers the shard) and wire the left neighbours.  `cuts`: world + 1 cut
        points (balanced_cuts); default equal lengths."""
        self.cuts = list(cuts) if cuts is not None else shard_cuts(n_total, self.world)
        nbytes = self.device.upload(index, self.cuts[self.rank], self.cuts[self.rank + 1], with_suf)
        self.connect()
        return nbytes

    def connect(self):
        """All-gather the IPC handles of every shard and map the (at most 8)
        shards to the left of this one."""
        if self.world == 1:
            self.device.set_left_views([])
            return
        mine = self.device.ipc_export()
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        left = [self.device.ipc_import(*everyone[r])
                for r in range(max(0, self.rank - MAX_LEFT), self.rank)]
        self.device.set_left_views(left)
        if self.exchange == "p2p":
            handle, _ = self.device.counts_export(self.world)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=self.group)
            self.device.counts_connect(self.rank, self.world, handles=handles)
        dist.barrier(group=self.group)

    # ---------------------------------------------------------------- scan
    def launch(self, minlength: int, policy: int = 0, gather: bool = True, stream: int = 0):
        """Scan kernel + count exchange, all enqueued on `stream`."""
        self.step += 1
        if self.exchange == "p2p":
            self.device.set_exchange_tag(self._tag())
        self.device.scan(minlength, policy, gather, stream)
        if self.exchange == "collective":
            self.device.copy_count(self._count.data_ptr(), stream)
            dist.all_gather(self._all, self._count, group=self.group)

    def _tag(self) -> int:
        return self.step % 0xfffffe + 1

    def offsets(self):
        """(global offset of this rank's first record, total records)."""
        if self.world == 1:
            nrec, _ = self.device.counts()
            return 0, nrec
        if self.exchange == "p2p":
            counts = self.device.peer_counts(self._tag(), self.world)
        else:
            counts = [int(t.item()) for t in self._all]
        return sum(counts[: self.rank]), sum(counts)

    def fetch(self):
        return self.device.fetch()


def gather_results(scan: ShardedScan, recs, pos):
    """Collect all shards' records/positions on rank 0 in SA order (for checks)."""
    parts: Sequence = [None] * scan.world
    if scan.world == 1:
        return recs, pos
    dist.all_gather_object(parts, (recs, pos), group=scan.group)
    import numpy as np
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
