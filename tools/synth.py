"""Deterministic synthetic sequences for the benchmark configs of BASELINE.json
(bench/test tooling).  Every generator is a pure function of (config, length,
seed); the seeds of SURVEY.md section 8d are the defaults.

Codes: DNA a,c,g,t = 0..3; protein = 0..19 in the order LVIFKREDAGSTNQYWPHMC
(/root/reference/src/core/alphabet.c:63-70); wildcard 254; separator 255.
"""
from __future__ import annotations

import numpy as np

WILDCARD, SEPARATOR = 254, 255
DNA = "acgt"
PROTEIN = "LVIFKREDAGSTNQYWPHMC"


def _mutate(rng, elem, rate, sigma):
    if rate <= 0:
        return elem
    out = elem.copy()
    hit = np.flatnonzero(rng.random(elem.shape[0]) < rate)
    if hit.size:
        out[hit] = (out[hit] + rng.integers(1, sigma, hit.size)) % sigma
    return out


def _place(rng, seq, elem):
    p = int(rng.integers(0, seq.shape[0] - elem.shape[0]))
    seq[p:p + elem.shape[0]] = elem


def dna_c2(length=100_000_000, seed=20001, budget=0.2):
    """C2: uniform DNA + tandem arrays + interspersed families with 0-2 %
    substitutions.  Injection stops when `budget` of the sequence
    has been overwritten (the literal parameter ranges would overwrite a
    100 Mbp sequence several times)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = rng.integers(0, 4, length, dtype=np.uint8)
    scale = length / 100_000_000
    used = 0
    limit = int(budget * length)
    for _ in range(max(1, int(2000 * scale))):          # tandem arrays
        unit = rng.integers(0, 4, int(rng.integers(2, 51)), dtype=np.uint8)
        arr = np.tile(unit, int(rng.integers(3, 201)))
        if arr.shape[0] >= length:
            continue
        _place(rng, seq, arr)
        used += arr.shape[0]
    for _ in range(max(1, int(500 * scale))):           # interspersed families
        if used >= limit:
            break
        elem = rng.integers(0, 4, int(rng.integers(100, 5001)), dtype=np.uint8)
        rate = float(rng.uniform(0.0, 0.02))
        for _ in range(int(rng.integers(2, 501))):
            if elem.shape[0] >= length:
                break
            _place(rng, seq, _mutate(rng, elem, rate, 4))
            used += elem.shape[0]
    return seq


def dna_c2_blocks(length=100_000_000, seed=20001, block=100_000_000):
    """Weak-scaling form of C2: independent C2 blocks of `block` bases, one seed each, joined
    into ONE sequence -- the repeat content per base (and with it the bytes the scan moves per
    suffix) does not depend on the number of blocks, which dna_c2(N * block) does not give
    (its families grow with the sequence).  One block is exactly dna_c2."""
    if length <= block:
        return dna_c2(length, seed)
    parts, done, b = [], 0, 0
    while done < length:
        part = min(block, length - done)
        parts.append(dna_c2(part, seed + 7919 * b))
        done += part
        b += 1
    return np.concatenate(parts)


def dna_c3(length=500_000_000, seed=30001):
    """C3 base sequence (index it with mirror_codes): C2 generator + long-plateau
    stress: families of EXACT copies, half of them terminated by a wildcard (so
    the lcp plateau is flat and as wide as the copy number, far beyond a warp or
    a tile, mostly with lcp >= 255) + homopolymer / low-complexity tracts."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = dna_c2(length, seed=seed + 1, budget=0.05)
    scale = length / 500_000_000
    used, limit = 0, int(0.2 * length)
    for fam in range(50):
        if used >= limit:
            break
        elen = int(rng.integers(300, 3001))
        copies = max(2, int(int(rng.integers(1000, 20001)) * scale))
        elem = rng.integers(0, 4, elen, dtype=np.uint8)
        if fam % 2 == 0:
            elem = np.concatenate([elem, np.array([WILDCARD], dtype=np.uint8)])
        for _ in range(copies):
            if used >= limit or elem.shape[0] >= length:
                break
            _place(rng, seq, elem)
            used += elem.shape[0]
    for _ in range(10):
        tl = max(50, int(int(rng.integers(100_000, 1_000_001)) * scale))
        if tl >= length:
            continue
        if rng.random() < 0.5:
            tract = np.full(tl, rng.integers(0, 4), dtype=np.uint8)
        else:
            tract = np.tile(rng.integers(0, 4, int(rng.integers(2, 7)), dtype=np.uint8), tl // 2)[:tl]
        _place(rng, seq, tract)
    return seq


def protein_c4(length=200_000_000, seed=40001):
    """C4: uniform 20-letter background, sequences of mean length 1000 joined by
    separators, 1000 domain families copied 2..2000x with 0-5 % substitutions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = rng.integers(0, 20, length, dtype=np.uint8)
    scale = length / 200_000_000
    used, limit = 0, int(0.2 * length)
    for _ in range(max(1, int(1000 * scale))):
        if used >= limit:
            break
        elem = rng.integers(0, 20, int(rng.integers(30, 400)), dtype=np.uint8)
        rate = float(rng.uniform(0.0, 0.05))
        for _ in range(int(rng.integers(2, 2001))):
            if elem.shape[0] >= length:
                break
            _place(rng, seq, _mutate(rng, elem, rate, 20))
            used += elem.shape[0]
    nsep = max(1, length // 1000) - 1
    if nsep > 0:
        seps = np.unique(rng.integers(1, length - 1, nsep))
        # (no two separators next to each other: the reference suffixerator rejects a file
        # with an empty sequence)
        seps = seps[np.concatenate([[True], np.diff(seps) > 1])]
        seq[seps] = SEPARATOR
    return seq


def dna_c5(length=3_000_000_000, seed=50001):
    """C5: C2 generator scaled, 24 'chromosomes' (23 separators), 0.5 % n runs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = dna_c2(length, seed=seed + 1)
    nruns = max(1, int(length * 0.005 / 1000))
    for _ in range(nruns):
        rl = int(rng.integers(100, 1901))
        if rl >= length:
            continue
        p = int(rng.integers(0, length - rl))
        seq[p:p + rl] = WILDCARD
    seps = np.unique(rng.integers(1, length - 1, 23))
    seq[seps] = SEPARATOR
    return seq


def to_fasta(codes: np.ndarray, path: str, alphabet: str = DNA, wildcard: str = "n"):
    """Write codes as FASTA (separators start a new record) for the reference
    suffixerator."""
    lut = np.zeros(256, dtype="S1")
    for k, ch in enumerate(alphabet):
        lut[k] = ch.encode()
    lut[WILDCARD] = wildcard.encode()
    cuts = np.flatnonzero(codes == SEPARATOR)
    start = 0
    with open(path, "wb") as fh:
        for k, c in enumerate(list(cuts) + [codes.shape[0]]):
            fh.write(b">s%d\n" % k)
            fh.write(lut[codes[start:c]].tobytes())
            fh.write(b"\n")
            start = c + 1


WORKLOADS = {
    "C2": dict(gen=dna_c2_blocks, length=100_000_000, seed=20001, minlength=20, mirrored=False,
               alphabet=DNA, wildcard="n", flags=["-dna"]),
    "C3": dict(gen=dna_c3, length=500_000_000, seed=30001, minlength=20, mirrored=True,
               alphabet=DNA, wildcard="n", flags=["-dna", "-mirrored"]),
    "C4": dict(gen=protein_c4, length=200_000_000, seed=40001, minlength=8, mirrored=False,
               alphabet=PROTEIN, wildcard="X", flags=["-protein"]),
    "C5": dict(gen=dna_c5, length=3_000_000_000, seed=50001, minlength=20, mirrored=False,
               alphabet=DNA, wildcard="n", flags=["-dna"]),
}
