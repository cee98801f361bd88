"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the smax path + loaders.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product package never
does.  See the header of ``oracle/smax_oracle.c`` for the parity status (the
smax sources are absent from the mounted reference; the text format is
"parity unpinned"; tables, interval enumeration and the left-diversity
convention are pinned against reference code run through ``oracle/_ref/gtref``).

Reference anchors restated here:
  * ``.prj`` key=value file           src/match/sfx-outprj.c:37-82, src/match/esa-map.c:55-211
  * ``.lcp`` / ``.llv``               src/match/sfx-lcpvalues.c:371-433, src/match/lcpoverflow.h:24-30
  * ``.bwt``                          src/match/sfx-run.c:174-211
  * ``.suf`` 8- or 4-byte entries     src/match/sfx-suffixgetset.c:48-55,467-482
  * interval / left-diversity rules   src/match/esa-bottomup.c:116-273, src/match/esa-maxpairs.c:24-31
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LLV_DTYPE = np.dtype([("position", "<u8"), ("value", "<u8")])
REC_DTYPE = np.dtype([("len", "<u8"), ("lb", "<u8"), ("width", "<u8")])


# --------------------------------------------------------------------------
# table loading (independent of the product's C loader)
# --------------------------------------------------------------------------
@dataclass
class EsaTables:
    prj: dict
    lcp: np.ndarray   # uint8[n]
    bwt: np.ndarray   # uint8[n]
    llv: np.ndarray   # LLV_DTYPE[L]
    suf: np.ndarray   # uint64[n] or uint32[n]

    @property
    def n(self) -> int:
        return int(self.lcp.shape[0])


def read_prj(indexname: str) -> dict:
    prj = {}
    with open(indexname + ".prj") as fh:
        for line in fh:
            if line.startswith("dbfile=") or "=" not in line:
                continue
            k, v = line.rstrip("\n").split("=", 1)
            prj[k] = float(v) if k == "averagelcp" else int(v)
    return prj


def load_esa(indexname: str, mmap: bool = True) -> EsaTables:
    prj = read_prj(indexname)
    n = prj["numberofallsortedsuffixes"]
    rd = (lambda p, dt: np.memmap(p, dtype=dt, mode="r")) if mmap else np.fromfile
    lcp = rd(indexname + ".lcp", np.uint8)
    bwt = rd(indexname + ".bwt", np.uint8)
    sufbytes = os.path.getsize(indexname + ".suf") // n
    suf = rd(indexname + ".suf", np.uint64 if sufbytes == 8 else np.uint32)
    if os.path.getsize(indexname + ".llv") > 0:
        llv = np.fromfile(indexname + ".llv", dtype=LLV_DTYPE)
    else:
        llv = np.zeros(0, dtype=LLV_DTYPE)
    assert lcp.shape[0] == n and bwt.shape[0] == n and suf.shape[0] == n
    assert llv.shape[0] == prj["largelcpvalues"]
    return EsaTables(prj, lcp, bwt, llv, suf)


# --------------------------------------------------------------------------
# numpy restatement (vectorised; independent of the C restatement)
# --------------------------------------------------------------------------
def resolve_lcp(lcp: np.ndarray, llv: np.ndarray) -> np.ndarray:
    """Resolved lcp values: the k-th 255 byte takes llv[k].value."""
    L = lcp.astype(np.uint64)
    big = np.flatnonzero(lcp == 255)
    if big.size:
        assert big.size == llv.shape[0], "one .llv record per 255 byte"
        assert np.array_equal(big.astype(np.uint64), llv["position"])
        L[big] = llv["value"]
    return L


def smax_numpy(lcp, llv, bwt, minlength: int, policy: int = 0) -> np.ndarray:
    """Supermaximal repeats as REC_DTYPE records in ascending lb."""
    n = lcp.shape[0]
    minlength = max(1, int(minlength))
    L = np.concatenate([resolve_lcp(np.asarray(lcp), llv), np.zeros(1, np.uint64)])
    # run starts s in [1..n]: L[s] != L[s-1]
    diff = np.flatnonzero(L[1:] != L[:-1]) + 1          # indices s with a change
    if diff.size == 0:
        return np.zeros(0, dtype=REC_DTYPE)
    starts = diff[:-1]
    ends = diff[1:] - 1                                  # run [s..e]
    v = L[starts]
    up = L[starts - 1] < v
    down = L[ends + 1] < v
    cand = up & down & (v >= minlength)
    starts, ends, v = starts[cand], ends[cand], v[cand]
    lb = starts - 1
    width = ends - lb + 1
    keep = np.zeros(starts.shape[0], dtype=bool)
    bwt = np.asarray(bwt)
    # group by width so the distinctness test vectorises
    for w in np.unique(width):
        sel = np.flatnonzero(width == w)
        idx = lb[sel][:, None] + np.arange(int(w))[None, :]
        c = bwt[idx].astype(np.int64)
        if policy == 0:
            # specials are pairwise distinct: give each a unique negative code
            spec = c >= 254
            c = np.where(spec, -1 - np.arange(int(w))[None, :], c)
        cs = np.sort(c, axis=1)
        keep[sel] = ~np.any(cs[:, 1:] == cs[:, :-1], axis=1)
    out = np.zeros(int(keep.sum()), dtype=REC_DTYPE)
    out["len"] = v[keep]
    out["lb"] = lb[keep]
    out["width"] = width[keep]
    return out


def gather_positions(suf, recs: np.ndarray) -> np.ndarray:
    if recs.shape[0] == 0:
        return np.zeros(0, dtype=np.uint64)
    w = recs["width"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(w)[:-1]])
    idx = np.repeat(recs["lb"].astype(np.int64) - off, w) + np.arange(int(w.sum()))
    return np.asarray(suf)[idx].astype(np.uint64)


def format_abs(recs: np.ndarray, positions: np.ndarray) -> bytes:
    """Default text grammar of this project (format 'smax', absolute):
    ``<length> <count> <pos_1> ... <pos_count>\\n`` -- also what
    ``oracle/_ref/gtref smax-bu`` prints."""
    out = []
    o = 0
    for r in recs:
        w = int(r["width"])
        out.append("%d %d %s\n" % (int(r["len"]), w,
                                   " ".join(str(int(p)) for p in positions[o:o + w])))
        o += w
    return "".join(out).encode()


# --------------------------------------------------------------------------
# ctypes front-end of the C restatement (oracle/smax_oracle.c)
# --------------------------------------------------------------------------
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsmax_oracle.so")
    src = os.path.join(HERE, "smax_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-std=c99", "-shared",
                               "-fPIC", "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_build", "libsmax_oracle.so")
        if not os.path.exists(so):
            so = build_c_oracle()
        lib = ctypes.CDLL(so)
        for name in ("smax_oracle_linear", "smax_oracle_stack"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int64
            fn.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p,
                           ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                           ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.smax_oracle_free.argtypes = [ctypes.c_void_p]
        lib.smax_oracle_free.restype = None
        lib.smax_oracle_positions.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_uint64,
                                              ctypes.c_void_p]
        lib.smax_oracle_positions.restype = None
        _LIB = lib
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p) if a.size else None


def smax_c(lcp, llv, bwt, minlength: int, policy: int = 0, algo: str = "linear") -> np.ndarray:
    lib = _lib()
    lcp = np.ascontiguousarray(lcp, dtype=np.uint8)
    bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
    llv = np.ascontiguousarray(llv, dtype=LLV_DTYPE)
    out = ctypes.c_void_p()
    fn = lib.smax_oracle_linear if algo == "linear" else lib.smax_oracle_stack
    cnt = fn(_ptr(lcp), lcp.shape[0], _ptr(llv), llv.shape[0], _ptr(bwt),
             int(minlength), int(policy), ctypes.byref(out))
    if cnt < 0:
        raise RuntimeError("smax oracle failed: %d" % cnt)
    if cnt == 0:
        if out.value:
            lib.smax_oracle_free(out)
        return np.zeros(0, dtype=REC_DTYPE)
    buf = (ctypes.c_char * (cnt * REC_DTYPE.itemsize)).from_address(out.value)
    recs = np.frombuffer(buf, dtype=REC_DTYPE).copy()
    lib.smax_oracle_free(out)
    return recs


def positions_c(suf, recs: np.ndarray) -> np.ndarray:
    lib = _lib()
    total = int(recs["width"].sum()) if recs.shape[0] else 0
    pos = np.zeros(total, dtype=np.uint64)
    if total:
        suf = np.ascontiguousarray(suf)
        recs = np.ascontiguousarray(recs)
        lib.smax_oracle_positions(_ptr(suf), suf.dtype.itemsize, _ptr(recs),
                                  recs.shape[0], _ptr(pos))
    return pos
