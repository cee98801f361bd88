"""ctypes binding of libsmax.so (the C ABI declared in include/smax.h).

This is the Python-side mirror of the boundary: the same entry points a cgo /
JNI / GenomeTools-C caller would bind.  There is deliberately NO fallback: if
the shared library is missing, or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_float, c_int,
                    c_int32, c_size_t, c_uint, c_uint8, c_uint32, c_uint64, c_void_p)

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# SMAX_LIB: another build of the library (kernel experiments, tools/try_variants.py)
LIB_PATH = os.environ.get("SMAX_LIB") or os.path.join(PKG, "lib", "libsmax.so")
TOOL_PATH = os.path.join(PKG, "lib", "smax")

ERRLEN = 1024
TAB_ESQ, TAB_SUF, TAB_LCP, TAB_BWT = 1, 2, 4, 8
TAB_ALL = 15
POLICY_GT, POLICY_PLAIN = 0, 1
FORMAT_SMAX, FORMAT_ITV, FORMAT_PAIRS = 0, 1, 2
IPC_BYTES, IPC_TABLES = 64, 5

LLV_DTYPE = np.dtype([("position", "<u8"), ("value", "<u8")])
REC_DTYPE = np.dtype([("len", "<u8"), ("lb", "<u8"), ("width", "<u8")])


class SmaxError(RuntimeError):
    pass


class IndexInfo(Structure):
    _fields_ = [(n, c_uint64) for n in
                ("totallength", "specialcharacters", "numofsequences",
                 "numberofallsortedsuffixes", "nonspecials", "largelcpvalues",
                 "maxbranchdepth", "longest")] + \
               [(n, c_uint32) for n in
                ("integersize", "littleendian", "readmode", "mirrored", "sufbytes",
                 "alphatype", "numofchars", "reserved")]


class Opts(Structure):
    _fields_ = [("minlength", c_uint64), ("relative", c_int32), ("ngpus", c_int32),
                ("policy", c_int32), ("format", c_int32), ("first_device", c_int32),
                ("verbose", c_int32)]


class ShardView(Structure):
    _fields_ = [("a_lo", c_uint64), ("a_hi", c_uint64), ("d_lcp", c_uint64),
                ("d_bwt", c_uint64), ("d_llv", c_uint64), ("d_llvdir", c_uint64),
                ("d_suf", c_uint64), ("nllv", c_uint64), ("device", c_int32),
                ("sufbytes", c_uint32)]


EMIT_CB = ctypes.CFUNCTYPE(c_int, c_void_p, c_uint64, c_uint64, c_uint64, POINTER(c_uint64))

# name -> (restype, argtypes); also the list the CPU test checks against smax.h
SIGNATURES = {
    "smax_index_open": (c_int, [c_char_p, c_uint, POINTER(c_void_p), c_char_p, c_size_t]),
    "smax_index_from_memory": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_uint,
                                       c_uint64, POINTER(c_void_p), c_char_p, c_size_t]),
    "smax_index_from_memory_window": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p,
                                              c_uint, c_uint64, c_uint64, c_uint64,
                                              POINTER(c_void_p), c_char_p, c_size_t]),
    "smax_index_close": (None, [c_void_p]),
    "smax_index_info_get": (c_int, [c_void_p, POINTER(IndexInfo)]),
    "smax_index_lcptab": (c_void_p, [c_void_p]),
    "smax_index_bwttab": (c_void_p, [c_void_p]),
    "smax_index_llvtab": (c_void_p, [c_void_p]),
    "smax_index_suftab": (c_void_p, [c_void_p]),
    "smax_run": (c_int, [c_void_p, POINTER(Opts), EMIT_CB, c_void_p, c_char_p, c_size_t]),
    "smax_run_records": (c_int, [c_void_p, POINTER(Opts), POINTER(c_void_p), POINTER(c_uint64),
                                 c_char_p, c_size_t]),
    "smax_free": (None, [c_void_p]),
    "smax_emitter_new": (c_int, [c_void_p, POINTER(Opts), c_void_p, POINTER(c_void_p), c_char_p,
                                 c_size_t]),
    "smax_emitter_emit": (c_int, [c_void_p, c_uint64, c_uint64, c_uint64, POINTER(c_uint64)]),
    "smax_emitter_emit_records": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "smax_emitter_delete": (c_int, [c_void_p]),
    "smax_tool_main": (c_int, [c_int, POINTER(c_char_p)]),
    "smax_run_stream": (c_int, [c_void_p, POINTER(Opts), c_uint64, c_void_p, c_void_p, c_char_p,
                                c_size_t]),
    "smax_run_text": (c_int, [c_void_p, POINTER(Opts), c_void_p, POINTER(c_uint64), c_char_p,
                              c_size_t]),
    "smax_index_separators": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_uint64), c_char_p,
                                      c_size_t]),
    "smax_device_set_separators": (c_int, [c_void_p, c_void_p, c_uint64, c_char_p, c_size_t]),
    "smax_device_build_separators": (c_int, [c_void_p, POINTER(c_uint64), c_char_p, c_size_t]),
    "smax_device_fetch_separators": (c_int, [c_void_p, c_void_p, POINTER(c_uint64), c_char_p,
                                             c_size_t]),
    "smax_scan_format": (c_int, [c_void_p, c_int, c_int, POINTER(c_uint64), c_char_p, c_size_t]),
    "smax_scan_fetch_text": (c_int, [c_void_p, c_void_p, c_char_p, c_size_t]),
    "smax_scan_format_elapsed_ms": (c_int, [c_void_p, POINTER(c_float), c_char_p, c_size_t]),
    "smax_device_count": (c_int, [c_char_p, c_size_t]),
    "smax_device_create": (c_int, [c_int, POINTER(c_void_p), c_char_p, c_size_t]),
    "smax_device_destroy": (None, [c_void_p]),
    "smax_device_synchronize": (c_int, [c_void_p]),
    "smax_release_devices": (None, []),
    "smax_device_upload": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_int,
                                   POINTER(c_uint64), c_char_p, c_size_t]),
    "smax_device_upload_halo": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_uint64, c_int,
                                        POINTER(c_uint64), c_char_p, c_size_t]),
    "smax_device_adopt": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p,
                                  c_uint, c_uint64, c_uint64, c_uint64, c_uint64, c_uint64,
                                  c_char_p, c_size_t]),
    "smax_device_view": (c_int, [c_void_p, POINTER(ShardView)]),
    "smax_device_set_left_views": (c_int, [c_void_p, POINTER(ShardView), c_int, c_char_p,
                                           c_size_t]),
    "smax_device_ipc_export": (c_int, [c_void_p, c_void_p, POINTER(ShardView), c_char_p,
                                       c_size_t]),
    "smax_device_ipc_import": (c_int, [c_void_p, c_void_p, POINTER(ShardView), c_char_p,
                                       c_size_t]),
    "smax_index_gather_positions": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_char_p,
                                            c_size_t]),
    "smax_device_counts_export": (c_int, [c_void_p, c_int, c_void_p, POINTER(c_uint64), c_char_p,
                                          c_size_t]),
    "smax_device_counts_connect": (c_int, [c_void_p, c_int, c_int, c_void_p, POINTER(c_uint64),
                                           c_char_p, c_size_t]),
    "smax_device_set_exchange_tag": (c_int, [c_void_p, c_uint64]),
    "smax_device_set_grid_limit": (c_int, [c_void_p, c_int]),
    "smax_scan_peer_counts": (c_int, [c_void_p, c_uint64, POINTER(c_uint64), c_char_p, c_size_t]),
    "smax_scan_launch": (c_int, [c_void_p, c_uint64, c_int, c_int, c_void_p, c_char_p, c_size_t]),
    "smax_scan_counts": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64), c_char_p,
                                 c_size_t]),
    "smax_scan_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_char_p, c_size_t]),
    "smax_scan_elapsed_ms": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float),
                                     POINTER(c_int), c_char_p, c_size_t]),
    "smax_scan_copy_count": (c_int, [c_void_p, c_void_p, c_void_p, c_char_p, c_size_t]),
    "smax_scan_device_buffers": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64),
                                         POINTER(c_uint64)]),
    "smax_device_set_stats": (c_int, [c_void_p, c_int]),
    "smax_device_set_debug": (c_int, [c_void_p, c_int]),
    "smax_h2d_bytes_total": (c_uint64, []),
    "smax_scan_stats": (c_int, [c_void_p, POINTER(c_uint64), c_char_p, c_size_t]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libsmax.so (built in-tree by genometools_smax_b200._build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmaxError("libsmax.so is not built: run `python -m genometools_smax_b200._build`"
                            " (there is no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _err():
    return ctypes.create_string_buffer(ERRLEN)


def _check(rc, err):
    if rc != 0:
        raise SmaxError(err.value.decode(errors="replace") or "libsmax call failed (%d)" % rc)


def _np_ptr(a):
    return None if a is None or a.size == 0 else c_void_p(a.ctypes.data)


class Index:
    """Host-side ESA handle (mmapped files or caller-owned arrays)."""

    def __init__(self, handle, keep=()):
        self.handle = handle
        self._keep = keep      # arrays that must outlive the handle

    @classmethod
    def open(cls, indexname: str, demand: int = TAB_ALL) -> "Index":
        h, err = c_void_p(), _err()
        _check(lib().smax_index_open(indexname.encode(), demand, byref(h), err, ERRLEN), err)
        return cls(h)

    @classmethod
    def from_arrays(cls, lcp, bwt, llv=None, suf=None) -> "Index":
        lcp = np.ascontiguousarray(lcp, dtype=np.uint8)
        bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
        llv = np.zeros(0, LLV_DTYPE) if llv is None else np.ascontiguousarray(llv, dtype=LLV_DTYPE)
        sufbytes = 0
        if suf is not None:
            suf = np.ascontiguousarray(suf)
            if suf.dtype not in (np.dtype("<u8"), np.dtype("<u4")):
                raise SmaxError("suffix table must be uint64 or uint32")
            sufbytes = suf.dtype.itemsize
        h, err = c_void_p(), _err()
        _check(lib().smax_index_from_memory(_np_ptr(lcp), _np_ptr(bwt), _np_ptr(llv),
                                            llv.shape[0], _np_ptr(suf) if suf is not None else None,
                                            sufbytes, lcp.shape[0], byref(h), err, ERRLEN), err)
        return cls(h, keep=(lcp, bwt, llv, suf))

    @classmethod
    def from_pointers(cls, lcp_ptr, bwt_ptr, llv_ptr, nllv, suf_ptr, sufbytes, n, keep=(),
                      base=0, n_total=None) -> "Index":
        """Wrap raw host pointers (e.g. pinned torch tensors) holding the lcp
        indices [base, base+n) of a table with n_total entries."""
        h, err = c_void_p(), _err()
        _check(lib().smax_index_from_memory_window(
            c_void_p(lcp_ptr), c_void_p(bwt_ptr), c_void_p(llv_ptr) if nllv else None, nllv,
            c_void_p(suf_ptr) if suf_ptr else None, sufbytes, base, n,
            n if n_total is None else n_total, byref(h), err, ERRLEN), err)
        return cls(h, keep=keep)

    def info(self) -> IndexInfo:
        info = IndexInfo()
        lib().smax_index_info_get(self.handle, byref(info))
        return info

    @property
    def n(self) -> int:
        return int(self.info().numberofallsortedsuffixes)

    def close(self):
        if self.handle:
            lib().smax_index_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- whole path -----------------------------------------------------
    def run_records(self, minlength: int, ngpus: int = 1, policy: int = POLICY_GT,
                    first_device: int = 0) -> np.ndarray:
        opts = Opts(minlength=minlength, relative=0, ngpus=ngpus, policy=policy,
                    format=FORMAT_SMAX, first_device=first_device, verbose=0)
        out, cnt, err = c_void_p(), c_uint64(), _err()
        _check(lib().smax_run_records(self.handle, byref(opts), byref(out), byref(cnt), err,
                                      ERRLEN), err)
        n = cnt.value
        if n == 0:
            return np.zeros(0, REC_DTYPE)
        buf = (c_char * (n * REC_DTYPE.itemsize)).from_address(out.value)
        recs = np.frombuffer(buf, dtype=REC_DTYPE).copy()
        lib().smax_free(out)
        return recs

    def gather_positions(self, recs: np.ndarray) -> np.ndarray:
        """suf[lb .. lb+width) of every record, in record order (host suffix table)."""
        recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
        out = np.zeros(int(recs["width"].sum()), np.uint64)
        err = _err()
        _check(lib().smax_index_gather_positions(self.handle, _np_ptr(recs), len(recs),
                                                 _np_ptr(out), err, ERRLEN), err)
        return out

    def separators(self) -> np.ndarray:
        """Ascending absolute positions of the sequence separators (host tables)."""
        ptr, cnt, err = c_void_p(), c_uint64(), _err()
        _check(lib().smax_index_separators(self.handle, byref(ptr), byref(cnt), err, ERRLEN), err)
        if cnt.value == 0:
            return np.zeros(0, np.uint64)
        buf = (c_uint64 * cnt.value).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint64).copy()

    def _with_file(self, fn, discard: bool = False) -> bytes:
        """Run fn(FILE*) on a temporary C stream and return the bytes it wrote
        (discard: the stream is /dev/null and nothing is returned)."""
        import tempfile
        libc = ctypes.CDLL(None)
        libc.fopen.restype = c_void_p
        libc.fopen.argtypes = [c_char_p, c_char_p]
        libc.fclose.argtypes = [c_void_p]
        if discard:
            fp = libc.fopen(b"/dev/null", b"w")
            try:
                fn(c_void_p(fp))
            finally:
                libc.fclose(c_void_p(fp))
            return b""
        with tempfile.NamedTemporaryFile(suffix=".smax") as tmp:
            fp = libc.fopen(tmp.name.encode(), b"w")
            if not fp:
                raise SmaxError("cannot open a temporary file")
            try:
                fn(c_void_p(fp))
            finally:
                libc.fclose(c_void_p(fp))
            return open(tmp.name, "rb").read()

    def emit_text(self, recs: np.ndarray, positions, fmt: int = FORMAT_SMAX,
                  relative: bool = False, discard: bool = False) -> bytes:
        """The host emitter (smax_emitter_*) over records + their positions."""
        recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
        pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.uint64)
        opts = Opts(minlength=1, relative=int(relative), ngpus=1, policy=POLICY_GT, format=fmt,
                    first_device=0, verbose=0)

        def body(fp):
            em, err = c_void_p(), _err()
            _check(lib().smax_emitter_new(self.handle, byref(opts), fp, byref(em), err, ERRLEN),
                   err)
            try:
                if lib().smax_emitter_emit_records(em, _np_ptr(recs), len(recs), _np_ptr(pos)) != 0:
                    raise SmaxError("host emitter failed")
            finally:
                lib().smax_emitter_delete(em)

        return self._with_file(body, discard)

    def run_stream_text(self, minlength: int, chunk: int = 0, policy: int = POLICY_GT,
                        fmt: int = FORMAT_SMAX, relative: bool = False) -> bytes:
        """smax_run_stream (the -scan mode) with the host emitter as its callback."""
        opts = Opts(minlength=minlength, relative=int(relative), ngpus=1, policy=policy,
                    format=fmt, first_device=0, verbose=0)

        def body(fp):
            em, err = c_void_p(), _err()
            _check(lib().smax_emitter_new(self.handle, byref(opts), fp, byref(em), err, ERRLEN),
                   err)
            try:
                cb = ctypes.cast(lib().smax_emitter_emit, c_void_p)
                _check(lib().smax_run_stream(self.handle, byref(opts), chunk, cb, em, err,
                                             ERRLEN), err)
            finally:
                lib().smax_emitter_delete(em)

        return self._with_file(body)

    def run_text(self, minlength: int, ngpus: int = 1, policy: int = POLICY_GT,
                 fmt: int = FORMAT_SMAX, relative: bool = False, discard: bool = False) -> bytes:
        """smax_run_text: the result lines rendered on the devices (the tool's -emit device)."""
        opts = Opts(minlength=minlength, relative=int(relative), ngpus=ngpus, policy=policy,
                    format=fmt, first_device=0, verbose=0)

        def body(fp):
            err, nb = _err(), c_uint64()
            _check(lib().smax_run_text(self.handle, byref(opts), fp, byref(nb), err, ERRLEN), err)
            self.last_text_bytes = int(nb.value)

        return self._with_file(body, discard)

    def run_emit_text(self, minlength: int, ngpus: int = 1, policy: int = POLICY_GT,
                      fmt: int = FORMAT_SMAX, relative: bool = False, discard: bool = False) -> bytes:
        """smax_run with the host emitter as its callback -- what the tool does by default
        (-emit host): upload, scan on `ngpus` devices, records back, one line per repeat."""
        opts = Opts(minlength=minlength, relative=int(relative), ngpus=ngpus, policy=policy,
                    format=fmt, first_device=0, verbose=0)

        def body(fp):
            em, err = c_void_p(), _err()
            _check(lib().smax_emitter_new(self.handle, byref(opts), fp, byref(em), err, ERRLEN),
                   err)
            try:
                cb = ctypes.cast(lib().smax_emitter_emit, EMIT_CB)
                _check(lib().smax_run(self.handle, byref(opts), cb, em, err, ERRLEN), err)
            finally:
                lib().smax_emitter_delete(em)

        return self._with_file(body, discard)

    def run(self, minlength: int, ngpus: int = 1, policy: int = POLICY_GT):
        """smax_run with a Python callback; returns [(len, lb, width, [positions])]."""
        res = []

        def cb(_info, length, lb, width, pos):
            res.append((int(length), int(lb), int(width),
                        [int(pos[k]) for k in range(width)] if pos else None))
            return 0

        opts = Opts(minlength=minlength, relative=0, ngpus=ngpus, policy=policy,
                    format=FORMAT_SMAX, first_device=0, verbose=0)
        err = _err()
        _check(lib().smax_run(self.handle, byref(opts), EMIT_CB(cb), None, err, ERRLEN), err)
        return res


def device_count() -> int:
    err = _err()
    n = lib().smax_device_count(err, ERRLEN)
    if n < 0:
        raise SmaxError(err.value.decode(errors="replace"))
    return n


class Device:
    """One GPU holding a resident shard of the tables."""

    def __init__(self, ordinal: int = 0):
        h, err = c_void_p(), _err()
        _check(lib().smax_device_create(ordinal, byref(h), err, ERRLEN), err)
        self.handle = h
        self.ordinal = ordinal
        self._keep = ()

    def close(self):
        if self.handle:
            lib().smax_device_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, index: Index, lo: int = 0, hi: int | None = None, with_suf: bool = True,
               halo: int = 256) -> int:
        hi = index.n if hi is None else hi
        nbytes, err = c_uint64(0), _err()
        _check(lib().smax_device_upload_halo(self.handle, index.handle, lo, hi, halo, int(with_suf),
                                             byref(nbytes), err, ERRLEN), err)
        return nbytes.value

    def adopt(self, d_lcp: int, d_bwt: int, d_llv: int, nllv: int, d_suf: int, sufbytes: int,
              a_lo: int, a_hi: int, lo: int, hi: int, n_total: int, keep=()):
        err = _err()
        _check(lib().smax_device_adopt(self.handle, c_void_p(d_lcp), c_void_p(d_bwt),
                                       c_void_p(d_llv) if d_llv else None, nllv,
                                       c_void_p(d_suf) if d_suf else None, sufbytes, a_lo, a_hi,
                                       lo, hi, n_total, err, ERRLEN), err)
        self._keep = keep

    def view(self) -> ShardView:
        v = ShardView()
        lib().smax_device_view(self.handle, byref(v))
        return v

    def set_left_views(self, views):
        arr = (ShardView * max(1, len(views)))(*views)
        err = _err()
        _check(lib().smax_device_set_left_views(self.handle, arr, len(views), err, ERRLEN), err)

    def ipc_export(self):
        handles = (c_uint8 * (IPC_BYTES * IPC_TABLES))()
        v, err = ShardView(), _err()
        _check(lib().smax_device_ipc_export(self.handle, handles, byref(v), err, ERRLEN), err)
        return bytes(handles), bytes(v)

    def ipc_import(self, handles: bytes, view_bytes: bytes) -> ShardView:
        h = (c_uint8 * (IPC_BYTES * IPC_TABLES)).from_buffer_copy(handles)
        v = ShardView.from_buffer_copy(view_bytes)
        err = _err()
        _check(lib().smax_device_ipc_import(self.handle, h, byref(v), err, ERRLEN), err)
        return v

    # one-sided exchange of the shards' record counts (P2P stores by the kernel)
    def counts_export(self, world: int):
        """(IPC handle bytes, device address) of this shard's count array."""
        handle, ptr, err = (c_uint8 * IPC_BYTES)(), c_uint64(), _err()
        _check(lib().smax_device_counts_export(self.handle, world, handle, byref(ptr), err, ERRLEN),
               err)
        return bytes(handle), ptr.value

    def counts_connect(self, rank: int, world: int, handles=None, ptrs=None):
        """handles: IPC handles of all shards in rank order (other processes), or
        ptrs: device addresses (shards of this process)."""
        err = _err()
        h = None
        if handles is not None:
            h = (c_uint8 * (IPC_BYTES * world)).from_buffer_copy(b"".join(handles))
        p = (c_uint64 * world)(*(ptrs if ptrs is not None else [0] * world))
        _check(lib().smax_device_counts_connect(self.handle, rank, world, h, p, err, ERRLEN), err)

    def set_exchange_tag(self, tag: int):
        lib().smax_device_set_exchange_tag(self.handle, int(tag))

    def peer_counts(self, tag: int, world: int):
        out, err = (c_uint64 * world)(), _err()
        _check(lib().smax_scan_peer_counts(self.handle, int(tag), out, err, ERRLEN), err)
        return [int(x) for x in out]

    def set_grid_limit(self, max_ctas: int):
        lib().smax_device_set_grid_limit(self.handle, int(max_ctas))

    def set_debug(self, flags: int):
        lib().smax_device_set_debug(self.handle, int(flags))

    def set_stats(self, on: bool):
        lib().smax_device_set_stats(self.handle, int(on))

    def scan(self, minlength: int, policy: int = POLICY_GT, gather: bool = True, stream: int = 0):
        err = _err()
        _check(lib().smax_scan_launch(self.handle, minlength, policy, int(gather),
                                      c_void_p(stream) if stream else None, err, ERRLEN), err)

    def counts(self):
        nrec, npos, err = c_uint64(), c_uint64(), _err()
        _check(lib().smax_scan_counts(self.handle, byref(nrec), byref(npos), err, ERRLEN), err)
        return nrec.value, npos.value

    def fetch(self):
        nrec, npos = self.counts()
        recs = np.zeros(nrec, REC_DTYPE)
        pos = np.zeros(npos, np.uint64)
        err = _err()
        _check(lib().smax_scan_fetch(self.handle, _np_ptr(recs), _np_ptr(pos), err, ERRLEN), err)
        return recs, pos

    # ---- emit on the device -----------------------------------------------
    def set_separators(self, seps: np.ndarray):
        seps = np.ascontiguousarray(seps, dtype=np.uint64)
        err = _err()
        _check(lib().smax_device_set_separators(self.handle, _np_ptr(seps), len(seps), err,
                                                ERRLEN), err)

    def build_separators(self) -> np.ndarray:
        cnt, err = c_uint64(), _err()
        _check(lib().smax_device_build_separators(self.handle, byref(cnt), err, ERRLEN), err)
        out = np.zeros(cnt.value, np.uint64)
        _check(lib().smax_device_fetch_separators(self.handle, _np_ptr(out), byref(cnt), err,
                                                  ERRLEN), err)
        return out

    def format_text(self, fmt: int = FORMAT_SMAX, relative: bool = False, fetch: bool = True):
        """Render the last scan's records as text on the device; returns the bytes
        (or only their number with fetch=False)."""
        nb, err = c_uint64(), _err()
        _check(lib().smax_scan_format(self.handle, fmt, int(relative), byref(nb), err, ERRLEN), err)
        if not fetch:
            return nb.value
        buf = np.zeros(max(nb.value, 1), np.uint8)
        _check(lib().smax_scan_fetch_text(self.handle, c_void_p(buf.ctypes.data), err, ERRLEN), err)
        return buf[:nb.value].tobytes()

    def format_elapsed_ms(self) -> float:
        ms, err = c_float(), _err()
        _check(lib().smax_scan_format_elapsed_ms(self.handle, byref(ms), err, ERRLEN), err)
        return ms.value

    def elapsed_ms(self):
        """(ms of the whole launch sequence, ms of the scan kernel alone, launches)"""
        ms, ms_scan, launches, err = c_float(), c_float(), c_int(), _err()
        _check(lib().smax_scan_elapsed_ms(self.handle, byref(ms), byref(ms_scan), byref(launches),
                                          err, ERRLEN), err)
        return ms.value, ms_scan.value, launches.value

    def copy_count(self, d_dst: int, stream: int = 0):
        err = _err()
        _check(lib().smax_scan_copy_count(self.handle, c_void_p(d_dst),
                                          c_void_p(stream) if stream else None, err, ERRLEN), err)

    def stats(self):
        arr, err = (c_uint64 * 8)(), _err()
        _check(lib().smax_scan_stats(self.handle, arr, err, ERRLEN), err)
        keys = ("n", "candidates", "candidate_width", "llv_inspected", "survivors",
                "survivor_width", "positions")
        st = dict(zip(keys, [int(x) for x in arr]))
        st["kernel"] = "units" if int(arr[7]) >> 63 else "ring"      # which scan kernel ran
        st["walks"] = int(arr[7]) & ((1 << 63) - 1)
        return st


def tool_main(argv) -> int:
    """Run the `smax` tool in-process (argv[0] is the tool name)."""
    arr = (c_char_p * len(argv))(*[a.encode() for a in argv])
    return lib().smax_tool_main(len(argv), arr)
