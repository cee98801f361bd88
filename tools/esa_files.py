"""Bench/test tooling (NOT the product path): write ESA tables built by
tools/esa_build_torch.py as GenomeTools index files, so that the product's
mmap loader (smax_index_open) and the REFERENCE code (oracle/_ref/gtref) read
the very same index the GPU arm scans.

Byte layouts: SURVEY.md section A (.suf 8-byte entries, .lcp / .bwt one byte
per suffix, .llv {u64 position, u64 value} records, .prj key=value lines as
written by /root/reference/src/match/sfx-outprj.c:37-82).  The .esq (and the
.prj keys that describe the sequence) come from the reference's own encoder
when it is available: `gtref suffixerator -tis` only encodes (0.05 s per Mbp);
the keys that describe the tables are then replaced.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np

TABLE_KEYS = ("numberofallsortedsuffixes", "longest", "prefixlength", "largelcpvalues",
              "averagelcp", "maxbranchdepth", "integersize", "littleendian", "readmode", "mirrored")


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def write_tables(base: str, esa: dict, chunk: int = 1 << 26):
    """<base>.suf/.lcp/.bwt/.llv from the builder's dict (torch tensors on any device, or numpy)."""
    n = int(esa["n"])
    with open(base + ".suf", "wb") as fh:
        for c0 in range(0, n, chunk):
            _np(esa["suf"][c0:c0 + chunk]).astype("<u8").tofile(fh)
    _np(esa["lcp"]).astype(np.uint8).tofile(base + ".lcp")
    _np(esa["bwt"]).astype(np.uint8).tofile(base + ".bwt")
    rec = np.zeros(int(esa["llv_pos"].shape[0]), dtype=[("position", "<u8"), ("value", "<u8")])
    rec["position"] = _np(esa["llv_pos"])
    rec["value"] = _np(esa["llv_val"])
    rec.tofile(base + ".llv")


def table_keys(esa: dict, mirrored: bool) -> dict:
    suf = esa["suf"]
    longest = int((suf == 0).nonzero()[0][0]) if hasattr(suf, "nonzero") else 0
    return {"numberofallsortedsuffixes": int(esa["n"]), "longest": longest, "prefixlength": 0,
            "largelcpvalues": int(esa["llv_pos"].shape[0]), "averagelcp": "0.00",
            "maxbranchdepth": int(esa["maxlcp"]), "integersize": 64, "littleendian": 1,
            "readmode": 0, "mirrored": int(bool(mirrored))}


def write_prj(base: str, esa: dict, codes: np.ndarray | None, mirrored: bool, template: str | None = None):
    """<base>.prj: sequence keys from `template` (a .prj written by the reference encoder) or
    counted from `codes` (the LOGICAL sequence: already mirrored for a -mirrored index); table
    keys from the builder."""
    keys = {}
    if template is not None:
        for ln in open(template):
            k, _, v = ln.rstrip("\n").partition("=")
            if k and not k.startswith("dbfile") and k not in TABLE_KEYS:
                keys[k] = v
    else:
        special = int((codes >= 254).sum()) if codes is not None else 0
        wild = int((codes == 254).sum()) if codes is not None else 0
        nseq = (int((codes == 255).sum()) + 1) if codes is not None else 1
        total = int(esa["n"]) - 1
        keys = {"totallength": total, "specialcharacters": special,
                "specialranges": 0, "realspecialranges": 0, "lengthofspecialprefix": 0,
                "lengthofspecialsuffix": 0, "wildcards": wild, "wildcardranges": 0,
                "realwildcardranges": 0, "lengthofwildcardprefix": 0, "lengthofwildcardsuffix": 0,
                "numofsequences": nseq, "numofdbsequences": nseq, "numofquerysequences": 0}
    keys.update(table_keys(esa, mirrored))
    with open(base + ".prj", "w") as fh:
        for k, v in keys.items():
            fh.write("%s=%s\n" % (k, v))


def encode_with_reference(gtref: str, base: str, fasta: str, flags) -> bool:
    """<base>.esq/.prj/.des/.sds/.md5 by the reference encoder (no suffix sorting)."""
    if not os.path.exists(gtref):
        return False
    subprocess.run([gtref, "suffixerator", "-db", fasta, "-tis", "-indexname", base] + list(flags),
                   check=True, capture_output=True)
    return True
