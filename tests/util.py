"""Synthetic table generators for the parity tests (test infrastructure)."""
import numpy as np

from oracle import smax_oracle as O


def tables_from_values(L, bwt):
    """(lcp bytes, llv records) from resolved lcp values L (uint64, L[0] == 0)."""
    L = np.asarray(L, dtype=np.uint64)
    lcp = np.minimum(L, 255).astype(np.uint8)
    big = np.flatnonzero(L >= 255)
    llv = np.zeros(big.size, dtype=O.LLV_DTYPE)
    llv["position"] = big
    llv["value"] = L[big]
    return lcp, llv, np.asarray(bwt, dtype=np.uint8)


def fuzz_tables(rng, n, kind):
    """Arbitrary (not necessarily realisable) lcp/bwt tables that stress the scan."""
    if kind == "dense":          # local maxima at every other position, few symbols
        L = rng.integers(0, 6, n).astype(np.uint64)
        bwt = rng.integers(0, 4, n)
    elif kind == "alternating":  # 5,0,5,0: the maximum survivor density
        L = np.zeros(n, np.uint64)
        L[1::2] = 5
        bwt = np.arange(n) % 4
    elif kind == "plateaus":     # runs of random length, mixes widths 2..300
        vals, lens = rng.integers(0, 40, n), rng.geometric(0.08, n)
        L = np.repeat(vals, lens)[:n].astype(np.uint64)
        bwt = rng.integers(0, 20, n)
        bwt[rng.random(n) < 0.3] = 254
    elif kind == "large":        # values around the 255 overflow, long 255-runs
        vals, lens = rng.integers(250, 262, n), rng.geometric(0.3, n)
        L = np.repeat(vals, lens)[:n].astype(np.uint64)
        L[rng.random(n) < 0.2] = rng.integers(0, 5)
        bwt = rng.integers(0, 4, n)
        bwt[rng.random(n) < 0.5] = 255
    elif kind == "huge":         # values beyond 32 bits
        L = rng.integers(0, 3, n).astype(np.uint64)
        sel = rng.random(n) < 0.1
        L[sel] = (1 << 40) + rng.integers(0, 3, int(sel.sum())).astype(np.uint64)
        bwt = rng.integers(0, 253, n)
    elif kind == "widerun":      # a few runs much wider than a tile, specials on the left
        L = np.zeros(n, np.uint64)
        pos = 1
        while pos < n - 10:
            w = int(rng.integers(1, max(2, n // 3)))
            L[pos:pos + w] = rng.integers(1, 300)
            pos += w + int(rng.integers(1, 4))
        bwt = np.full(n, 254)
        bwt[rng.random(n) < 0.001] = 1
    elif kind == "sparse":       # like random DNA with minlength 20
        L = np.minimum(rng.geometric(0.75, n) + 10, 40).astype(np.uint64)
        bwt = rng.integers(0, 4, n)
    else:
        raise ValueError(kind)
    L = np.asarray(L, dtype=np.uint64)
    L[0] = 0
    return tables_from_values(L, bwt)


def render_text(recs, positions, fmt="smax", seps=None):
    """Plain-Python statement of the tool's line grammar (csrc/smax_emit.c):
    smax  '<len> <count> <pos>...'   (seps given: '<len> <count> <seq> <rel>...')
    itv   '<len> <lb> <rb>'."""
    out, o = [], 0
    if seps is not None:
        seps = np.asarray(seps, dtype=np.uint64)
    for r in recs:
        ln, lb, w = int(r["len"]), int(r["lb"]), int(r["width"])
        if fmt == "itv":
            out.append("%d %d %d\n" % (ln, lb, lb + w - 1))
            continue
        p = positions[o:o + w]
        o += w
        if seps is None:
            out.append("%d %d %s\n" % (ln, w, " ".join(str(int(x)) for x in p)))
        else:
            k = np.searchsorted(seps, p, side="left")
            f = []
            for x, kk in zip(p, k):
                start = int(seps[kk - 1]) + 1 if kk else 0
                f.append("%d %d" % (int(kk), int(x) - start))
            out.append("%d %d %s\n" % (ln, w, " ".join(f)))
    return "".join(out).encode()


def write_index_files(base, lcp, bwt, llv, suf):
    """Write <base>.prj/.lcp/.bwt/.llv/.suf for arbitrary tables in the byte layout of the
    reference's suffixerator (.prj keys as written by sfx-outprj.c:53-82)."""
    n = len(lcp)
    maxlcp = int(max(int(lcp.max()) if n else 0, int(llv["value"].max()) if len(llv) else 0))
    prj = {"totallength": n - 1, "specialcharacters": 0, "specialranges": 0, "realspecialranges": 0,
           "lengthofspecialprefix": 0, "lengthofspecialsuffix": 0, "wildcards": 0, "wildcardranges": 0,
           "realwildcardranges": 0, "lengthofwildcardprefix": 0, "lengthofwildcardsuffix": 0,
           "numofsequences": 1, "numofdbsequences": 1, "numofquerysequences": 0,
           "numberofallsortedsuffixes": n, "longest": 0, "prefixlength": 0,
           "largelcpvalues": len(llv), "averagelcp": "1.00", "maxbranchdepth": maxlcp,
           "integersize": 64, "littleendian": 1, "readmode": 0, "mirrored": 0}
    with open(base + ".prj", "w") as fh:
        for k, v in prj.items():
            fh.write("%s=%s\n" % (k, v))
    np.ascontiguousarray(lcp, np.uint8).tofile(base + ".lcp")
    np.ascontiguousarray(bwt, np.uint8).tofile(base + ".bwt")
    np.ascontiguousarray(llv).tofile(base + ".llv")
    np.ascontiguousarray(suf, np.uint64).tofile(base + ".suf")


def wide_plateau_tables(rng, n=40000):
    """Tables with plateaus far wider than a small chunk: a run of 6001 equal small values
    whose left characters are all specials (supermaximal under the gt policy), a run of 3000
    equal large values, and ordinary plateaus around them."""
    lcp, llv, bwt = fuzz_tables(rng, n, "plateaus")
    L = lcp.astype(np.uint64)
    if len(llv):
        L[llv["position"].astype(np.int64)] = llv["value"]
    L[9000:15001] = 7
    L[8999] = 2
    L[15001] = 3
    bwt = bwt.astype(np.uint8)
    bwt[8999:15001] = 255
    L[20000:23000] = 1000
    L[19999] = 999
    L[23000] = 5
    bwt[19999:23000] = np.where(np.arange(3001) % 2 == 0, 254, 255)
    L[0] = 0
    big = np.flatnonzero(L >= 255)
    llv2 = np.zeros(len(big), llv.dtype)
    llv2["position"] = big
    llv2["value"] = L[big]
    return np.minimum(L, 255).astype(np.uint8), llv2, bwt
