"""Bench/test tooling (NOT the product path): build GenomeTools-format ESA tables
on the GPU with torch so that the benchmark can synthesise a 1e8..3e9-suffix
index on the box in seconds to minutes (the reference's `gt suffixerator` needs
~70 s per 100 Mbp single-threaded; index construction is out of scope of the hot
path and is timed separately).

Table semantics reproduced (validated bit-for-bit against indexes built by the
reference suffixerator in tests/test_esa_builder.py, and on the box against the
reference suffixerator's tables of the CPU-baseline sample in bench.py):
  * suffix order: letters by code; every special (wildcard 254 / separator 255 /
    the end of the text) is larger than all letters and different from every
    other special, earlier specials first -- hence suffixes that start with a
    special close the table in ascending text position and LCPs stop at
    specials (/root/reference/src/match/sfx-suffixer.c,
    sfx-lcpvalues.c:435-471)
  * lcp[i] = LCP(suf[i-1], suf[i]) capped at 255 with {position, value} overflow
    records (/root/reference/src/match/sfx-lcpvalues.c:371-433)
  * bwt[i] = code of text[suf[i]-1], 254 for suf[i]==0 (/root/reference/src/match/sfx-run.c:174-211)
  * -mirrored: text + separator + reverse complement (/root/reference/src/core/encseq.c:9449-9474)

Algorithm (n < 2^32; memory ~ 24 n bytes resident + 56 bytes per element of one
chunk, so 3e9 suffixes fit one 180 GB B200):
  * suffix array by prefix doubling.  The order of one level is refined in
    CHUNKS of the previous order (cut at group boundaries), so that the radix
    sort (torch.sort) never sees more than `chunk` keys; ranks are group starts,
    which makes "where does the group of this element start" a table look-up.
  * LCPs by comparing the text of neighbouring suffixes 8 characters per step
    (packed little-endian words; a special never equals anything, not even
    itself) over a shrinking set of still-equal pairs -- no rank levels kept.
"""
from __future__ import annotations

import numpy as np
import torch

WILDCARD, SEPARATOR = 254, 255
_OFF = 1 << 31
_L7 = 0x7F7F7F7F7F7F7F7F
_H7 = -0x7F7F7F7F7F7F7F80        # 0x8080808080808080 as int64
_C2 = 0x0202020202020202
_ONES = 0x0101010101010101


def mirror_codes(codes: torch.Tensor) -> torch.Tensor:
    """Logical sequence of a -mirrored index (DNA codes a,c,g,t = 0..3)."""
    rc = codes.flip(0)
    rc = torch.where(rc < 4, 3 - rc, rc)
    sep = torch.full((1,), SEPARATOR, dtype=codes.dtype, device=codes.device)
    return torch.cat([codes, sep, rc])


# torch.sort, nonzero, cumsum and the indexing kernels take at most INT_MAX elements per call:
# passes over all n elements go in slices of _SLICE (tests lower it to walk the sliced paths)
_SLICE = 1 << 30


def _nonzero_sliced(mask: torch.Tensor) -> torch.Tensor:
    n = mask.shape[0]
    if n <= _SLICE:
        return torch.nonzero(mask).flatten()
    parts = []
    for s in range(0, n, _SLICE):
        nz = torch.nonzero(mask[s:s + _SLICE]).flatten()
        if nz.numel():
            parts.append(nz.add_(s))
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=mask.device)


def _group_starts(newgroup: torch.Tensor, base: int) -> torch.Tensor:
    """index (+ base) of the first element of every element's group"""
    n = newgroup.shape[0]
    starts = _nonzero_sliced(newgroup)
    if n <= _SLICE:
        dense = torch.cumsum(newgroup, 0)
        dense.sub_(1)
        gs = starts[dense]
        del dense
    else:
        gs = torch.empty(n, dtype=torch.int64, device=newgroup.device)
        carry = 0
        for s in range(0, n, _SLICE):
            dense = torch.cumsum(newgroup[s:s + _SLICE], 0)
            dense.add_(carry - 1)
            carry = int(dense[-1]) + 1
            gs[s:s + _SLICE] = starts[dense]
            del dense
    del starts
    if base:
        gs.add_(base)
    return gs


def _first_level(key: torch.Tensor, n: int):
    """Suffixes ordered by their first letter (specials and the end of the text, key 255, each on
    its own, in text order): a stable sort of one byte key -- by counting when n exceeds what
    one torch.sort call takes.  Returns (sa, newgroup)."""
    dev = key.device
    if n <= min(_SLICE, (1 << 31) - 1):
        skey, sa = torch.sort(key, stable=True)
        ng = torch.ones(n, dtype=torch.bool, device=dev)
        if n > 1:
            ng[1:] = (skey[1:] != skey[:-1]) | (skey[1:] == 255)
        return sa, ng
    counts = torch.zeros(256, dtype=torch.int64, device=dev)
    for s in range(0, n, _SLICE):
        counts += torch.bincount(key[s:s + _SLICE].to(torch.int64), minlength=256)
    counts = counts.cpu().tolist()
    sa = torch.empty(n, dtype=torch.int64, device=dev)
    ng = torch.zeros(n, dtype=torch.bool, device=dev)
    at = 0
    for v in range(256):
        if counts[v] == 0:
            continue
        ng[at] = True
        if v == 255:
            ng[at:at + counts[v]] = True
        for s in range(0, n, _SLICE):
            nz = torch.nonzero(key[s:s + _SLICE] == v).flatten()
            if nz.numel():
                sa[at:at + nz.numel()] = nz.add_(s)
                at += nz.numel()
            del nz
    return sa, ng


def _suffix_array(codes: torch.Tensor, n: int, chunk: int, verbose: bool, levels=None):
    """levels: a list that receives the rank array (int32, group starts) of every level h = 1, 2,
    4, ... for the LCP descent, or None"""
    dev = codes.device
    T = n - 1
    # level h = 1: letters by code; specials and the end of the text are unique, in text order
    # (stable sort of one key for all of them)
    key = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    for s in range(0, T, _SLICE):
        c = codes[s:s + _SLICE]
        key[s:s + c.shape[0]] = torch.where(c >= 254, torch.full_like(c, 255), c)
        del c
    sa, ng = _first_level(key, n)
    del key
    done = bool(ng.all())
    rank = torch.empty(n, dtype=torch.int64, device=dev)
    gs = _group_starts(ng, 0)
    for s in range(0, n, _SLICE):
        rank[sa[s:s + _SLICE]] = gs[s:s + _SLICE]
    del ng, gs
    rank_new = torch.empty_like(rank) if not done else None
    h = 1
    chunk = min(chunk, _SLICE)
    while not done:
        if levels is not None:
            levels.append(rank.to(torch.int32))
        # chunk boundaries: multiples of `chunk` moved down to the start of their group
        cuts = [0]
        for c in range(chunk, n, chunk):
            b = int(rank[sa[c]])
            if b > cuts[-1]:
                cuts.append(b)
        cuts.append(n)
        done = True
        for c0, c1 in zip(cuts[:-1], cuts[1:]):
            idx = sa[c0:c1].clone()
            key = rank[idx]
            key.sub_(_OFF).bitwise_left_shift_(32)
            j = idx + h
            ok = j < n
            j.clamp_(max=n - 1)
            r2 = rank[j]
            r2.mul_(ok)
            del j, ok
            key.bitwise_or_(r2)
            del r2
            skey, perm = torch.sort(key)
            del key
            idx = idx[perm]
            del perm
            sa[c0:c1] = idx
            ngc = torch.ones(c1 - c0, dtype=torch.bool, device=dev)
            if c1 - c0 > 1:
                ngc[1:] = skey[1:] != skey[:-1]
            del skey
            done = done and bool(ngc.all())
            rank_new[idx] = _group_starts(ngc, c0)
            del idx, ngc
        rank, rank_new = rank_new, rank
        h *= 2
        if verbose:
            print("  doubling h=%d chunks=%d done=%s" % (h, len(cuts) - 1, done), flush=True)
    if levels is not None:
        levels.append(rank.to(torch.int32))          # the level at which all ranks are unique
    del rank, rank_new
    return sa


def _packed_words(codes: torch.Tensor, n: int, block: int) -> torch.Tensor:
    """W[i] = text[i .. i+8) as a little-endian word, the text padded with separators"""
    dev = codes.device
    pad = torch.cat([codes, torch.full((16,), SEPARATOR, dtype=torch.uint8, device=dev)])
    W = torch.empty(n, dtype=torch.int64, device=dev)
    for c0 in range(0, n, block):
        c1 = min(n, c0 + block)
        w = pad[c0:c1].to(torch.int64)
        for b in range(1, 8):
            w.bitwise_or_(pad[c0 + b:c1 + b].to(torch.int64).bitwise_left_shift_(8 * b))
        W[c0:c1] = w
        del w
    return W


def _special_ones(w: torch.Tensor) -> torch.Tensor:
    """0x01 in every byte of w that holds a special (>= 254)"""
    s = (w & _L7).add_(_C2).bitwise_and_(w).bitwise_and_(_H7)
    return s.bitwise_right_shift_(7).bitwise_and_(_ONES)


def _lcp_table(codes: torch.Tensor, suf: torch.Tensor, n: int, block: int):
    dev = codes.device
    W = _packed_words(codes, n, block)
    lcp = torch.zeros(n, dtype=torch.uint8, device=dev)
    llv_pos, llv_val = [], []
    maxlcp = 0
    for c0 in range(1, n, block):
        c1 = min(n, c0 + block)
        a = suf[c0 - 1:c1 - 1]
        b = suf[c0:c1]
        l = torch.zeros(c1 - c0, dtype=torch.int64, device=dev)
        act = torch.arange(c1 - c0, dtype=torch.int64, device=dev)
        while act.numel():
            la = l[act]
            wa = W[a[act] + la]
            wb = W[b[act] + la]
            del la
            # a special never matches: 254 on one side, 255 on the other
            wa.bitwise_and_(~_special_ones(wa))
            wb.bitwise_or_(_special_ones(wb))
            x = wa.bitwise_xor_(wb)
            del wb
            eq8 = x == 0
            l[act[eq8]] += 8
            ne = ~eq8
            xm = x[ne]
            del x
            if xm.numel():
                cnt = torch.zeros_like(xm)
                alive = torch.ones_like(xm, dtype=torch.bool)
                for byte in range(8):
                    alive &= (xm.bitwise_right_shift(8 * byte) & 0xFF) == 0
                    cnt += alive
                l[act[ne]] += cnt
                del cnt, alive
            del xm, ne
            act = act[eq8]
            del eq8
        lcp[c0:c1] = l.clamp(max=255).to(torch.uint8)
        big = torch.nonzero(l >= 255).flatten()
        if big.numel():
            llv_pos.append(big + c0)
            llv_val.append(l[big])
        if l.numel():
            maxlcp = max(maxlcp, int(l.max()))
        del l, a, b, big
    del W
    if llv_pos:
        return lcp, torch.cat(llv_pos), torch.cat(llv_val), maxlcp
    z = torch.zeros(0, dtype=torch.int64, device=dev)
    return lcp, z, z.clone(), maxlcp


def _lcp_by_levels(suf: torch.Tensor, levels, n: int, block: int):
    """LCPs by binary descent over the rank arrays of the doubling levels: ranks at level k are
    equal iff the first 2^k characters are (specials are unique from level 0 on, so a match never
    extends over one)."""
    dev = suf.device
    lcp = torch.zeros(n, dtype=torch.uint8, device=dev)
    llv_pos, llv_val = [], []
    maxlcp = 0
    for c0 in range(1, n, block):
        c1 = min(n, c0 + block)
        a = suf[c0 - 1:c1 - 1]
        b = suf[c0:c1]
        l = torch.zeros(c1 - c0, dtype=torch.int64, device=dev)
        for k in range(len(levels) - 1, -1, -1):
            lv = levels[k]
            ia = a + l
            ib = b + l
            ok = (ia < n) & (ib < n)
            ia.clamp_(max=n - 1)
            ib.clamp_(max=n - 1)
            eq = ok & (lv[ia] == lv[ib])
            l += eq.to(torch.int64) << k
            del ia, ib, ok, eq
        lcp[c0:c1] = l.clamp(max=255).to(torch.uint8)
        big = torch.nonzero(l >= 255).flatten()
        if big.numel():
            llv_pos.append(big + c0)
            llv_val.append(l[big])
        if l.numel():
            maxlcp = max(maxlcp, int(l.max()))
        del l, a, b, big
    if llv_pos:
        return lcp, torch.cat(llv_pos), torch.cat(llv_val), maxlcp
    z = torch.zeros(0, dtype=torch.int64, device=dev)
    return lcp, z, z.clone(), maxlcp


@torch.no_grad()
def build_esa(codes: torch.Tensor, keep_on_device: bool = False, verbose: bool = False,
              chunk: int = 1 << 30, block: int = 1 << 27, lcp_method: str = "auto"):
    """codes: uint8[T] (letters < 254, 254 wildcard, 255 separator).
    Returns dict(suf int64[n], lcp uint8[n], bwt uint8[n], llv_pos int64[L],
    llv_val int64[L]) with n = T + 1, as torch tensors on codes.device (or numpy).
    `chunk`: most keys one radix sort sees; `block`: elements per step of the
    chunked passes (both only bound the temporary memory).  lcp_method: "levels" keeps the
    rank array of every doubling level (4 n bytes each; time independent of the LCP values),
    "compare" compares text (no extra memory; time grows with the sum of the LCPs), "auto"
    takes levels up to 2^30 suffixes."""
    dev = codes.device
    T = int(codes.shape[0])
    n = T + 1
    if n >= (1 << 32):
        raise ValueError("builder handles n < 2^32")
    codes = codes.contiguous()
    if lcp_method == "auto":
        lcp_method = "levels" if n <= (1 << 30) + 8 else "compare"
    levels = [] if lcp_method == "levels" else None
    suf = _suffix_array(codes, n, chunk, verbose, levels)
    if levels is not None:
        lcp, llv_pos, llv_val, maxlcp = _lcp_by_levels(suf, levels, n, block)
        del levels
    else:
        lcp, llv_pos, llv_val, maxlcp = _lcp_table(codes, suf, n, block)
    # ---- BWT
    bwt = torch.empty(n, dtype=torch.uint8, device=dev)
    for c0 in range(0, n, block):
        c1 = min(n, c0 + block)
        s = suf[c0:c1]
        if T:
            bwt[c0:c1] = torch.where(s == 0, torch.full((1,), WILDCARD, dtype=torch.uint8, device=dev),
                                     codes[(s - 1).clamp_(min=0)])
        else:
            bwt[c0:c1] = WILDCARD
        del s
    out = dict(suf=suf, lcp=lcp, bwt=bwt, llv_pos=llv_pos, llv_val=llv_val, n=n, maxlcp=maxlcp)
    if keep_on_device:
        return out
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def llv_records(llv_pos, llv_val) -> np.ndarray:
    dt = np.dtype([("position", "<u8"), ("value", "<u8")])
    rec = np.zeros(len(llv_pos), dtype=dt)
    rec["position"] = np.asarray(llv_pos)
    rec["value"] = np.asarray(llv_val)
    return rec
