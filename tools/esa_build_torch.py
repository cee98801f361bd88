"""Bench/test tooling (NOT the product path): build GenomeTools-format ESA tables
on the GPU with torch so that the benchmark can synthesise a 1e8..1e9-suffix
index on the box in seconds (the reference's `gt suffixerator` needs ~70 s per
100 Mbp single-threaded; index construction is out of scope of the hot path and
is timed separately).

Table semantics reproduced (validated bit-for-bit against indexes built by the
reference suffixerator in tests/test_esa_builder.py):
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

Algorithm: prefix doubling with full radix sorts (torch.sort) keeping the rank
array of every level; LCPs by binary descent over the levels (no sequential
Kasai pass).
"""
from __future__ import annotations

import numpy as np
import torch

WILDCARD, SEPARATOR = 254, 255


def mirror_codes(codes: torch.Tensor) -> torch.Tensor:
    """Logical sequence of a -mirrored index (DNA codes a,c,g,t = 0..3)."""
    rc = codes.flip(0)
    rc = torch.where(rc < 4, 3 - rc, rc)
    sep = torch.full((1,), SEPARATOR, dtype=codes.dtype, device=codes.device)
    return torch.cat([codes, sep, rc])


@torch.no_grad()
def build_esa(codes: torch.Tensor, keep_on_device: bool = False, verbose: bool = False):
    """codes: uint8[T] (letters < 254, 254 wildcard, 255 separator).
    Returns dict(suf int64[n], lcp uint8[n], bwt uint8[n], llv_pos int64[L],
    llv_val int64[L]) with n = T + 1, as torch tensors on codes.device (or numpy)."""
    dev = codes.device
    T = int(codes.shape[0])
    n = T + 1
    if n >= (1 << 31):
        raise ValueError("builder handles n < 2^31 per call")
    special = codes >= 254
    sigma = 254
    # level 0 (h = 1): letters by code, specials unique in text order, sentinel last
    spec_rank = torch.cumsum(special.to(torch.int64), 0) - 1
    r = torch.where(special, sigma + spec_rank, codes.to(torch.int64))
    r = torch.cat([r, (sigma + spec_rank[-1:] + 1) if T else torch.full((1,), sigma, device=dev)])
    del spec_rank, special
    # densify so that keys stay below 2^31
    uniq, inv = torch.unique(r, sorted=True, return_inverse=True)
    rank = inv.to(torch.int64)
    del uniq, inv, r
    levels = [rank.to(torch.int32)]
    h = 1
    sa = None
    while True:
        r2 = torch.zeros(n, dtype=torch.int64, device=dev)
        if h < n:
            r2[: n - h] = rank[h:]
        key = (rank << 32) | r2
        del r2
        skey, sa = torch.sort(key)
        del key
        newgroup = torch.ones(n, dtype=torch.int64, device=dev)
        newgroup[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        del skey
        newrank_sorted = torch.cumsum(newgroup, 0) - 1
        del newgroup
        done = int(newrank_sorted[-1]) == n - 1
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[sa] = newrank_sorted
        del newrank_sorted
        h *= 2
        if verbose:
            print("  doubling h=%d done=%s" % (h, done), flush=True)
        if done:
            break
        levels.append(rank.to(torch.int32))
    del rank
    suf = sa  # int64[n]
    # ---- LCP by binary descent over the stored levels
    a = suf[:-1]
    b = suf[1:]
    l = torch.zeros(n - 1, dtype=torch.int64, device=dev)
    for k in range(len(levels) - 1, -1, -1):
        lv = levels[k]
        ia = a + l
        ib = b + l
        ok = (ia < n) & (ib < n)
        ia.clamp_(max=n - 1)
        ib.clamp_(max=n - 1)
        eq = ok & (lv[ia] == lv[ib])
        # level 0 ranks equal for two specials is impossible (unique), so eq
        # never extends a match across a special
        l += eq.to(torch.int64) << k
        del ia, ib, ok, eq
        levels[k] = None
    del levels, a, b
    lcpv = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), l])
    del l
    lcp = torch.clamp(lcpv, max=255).to(torch.uint8)
    llv_pos = torch.nonzero(lcpv >= 255).flatten()
    llv_val = lcpv[llv_pos]
    maxlcp = int(lcpv.max()) if n > 1 else 0
    del lcpv
    # ---- BWT
    prev = (suf - 1).clamp_(min=0)
    bwt = codes[prev.clamp(max=max(T - 1, 0))] if T else torch.zeros(n, dtype=torch.uint8, device=dev)
    bwt = torch.where(suf == 0, torch.full_like(bwt, WILDCARD), bwt)
    del prev
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
