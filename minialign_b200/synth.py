"""Deterministic synthetic genomes and PBSIM-CLR-like reads (no PBSIM / real genomes exist offline).

Shapes follow SURVEY.md §8(d): genome = i.i.d. uniform ACGT with planted repeat families; reads = uniform start,
strand 50/50, length ~N(20000, 2000) clipped to [100, 25000], per-read accuracy ~N(0.88, 0.07) clipped to
[0.75, 1.0], errors sub:ins:del = 10:60:30 (the read model minialign's README.md:53 quotes its numbers on).
Everything is numpy and seeded, so the same call gives the same bytes here and on the GPU box.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def make_genome(length: int, n_contigs: int = 1, seed: int = 1, repeats=((7, 5000), (40, 1500)), divergence: float = 0.02, weights=None):
    """Return a list of (name, uint8 ASCII array) contigs.  weights: relative contig sizes (default: equal); unequal contigs
    matter for parity because the reference tests a read's first seed against the PREVIOUS chain's reference length
    (minialign.c:3865 runs before 3873)."""
    rng = np.random.default_rng(seed)
    if weights is None:
        sizes = np.full(n_contigs, length // n_contigs, dtype=np.int64)
    else:
        w = np.asarray(weights, dtype=np.float64)
        assert w.size == n_contigs
        sizes = np.maximum(1, (length * w / w.sum()).astype(np.int64))
    sizes[-1] += length - sizes.sum()
    contigs = []
    for ci, sz in enumerate(sizes):
        g = _ACGT[rng.integers(0, 4, size=int(sz))]
        contigs.append(g)
    # planted repeat families: copies with small divergence, scattered over all contigs
    for fam, (copies, rlen) in enumerate(repeats):
        if rlen * 4 > sizes.min():
            continue
        unit = _ACGT[rng.integers(0, 4, size=rlen)]
        for _ in range(copies):
            c = int(rng.integers(0, n_contigs))
            pos = int(rng.integers(0, sizes[c] - rlen))
            cp = unit.copy()
            nmut = int(rlen * divergence)
            if nmut:
                idx = rng.integers(0, rlen, size=nmut)
                cp[idx] = _ACGT[rng.integers(0, 4, size=nmut)]
            if rng.integers(0, 2):
                cp = _COMP[cp[::-1]]
            contigs[c][pos:pos + rlen] = cp
    return [(f"chr{ci + 1}", g) for ci, g in enumerate(contigs)]


def _mutate(src: np.ndarray, acc: float, rng) -> np.ndarray:
    """Apply sub:ins:del = 10:60:30 errors at total rate (1-acc) per source base."""
    n = src.size
    err = 1.0 - acc
    r = rng.random(n)
    is_sub = r < err * 0.10
    is_ins = (r >= err * 0.10) & (r < err * 0.70)
    is_del = (r >= err * 0.70) & (r < err)
    out_len = np.ones(n, dtype=np.int64)
    out_len[is_ins] = 2
    out_len[is_del] = 0
    ofs = np.concatenate(([0], np.cumsum(out_len)))
    out = np.empty(int(ofs[-1]), dtype=np.uint8)
    keep = out_len > 0
    out[ofs[:-1][keep]] = src[keep]
    sub_idx = np.nonzero(is_sub)[0]
    if sub_idx.size:
        # substitute with one of the three other bases
        code = np.searchsorted(_ACGT, src[sub_idx])
        out[ofs[:-1][sub_idx]] = _ACGT[(code + rng.integers(1, 4, size=sub_idx.size)) & 3]
    ins_idx = np.nonzero(is_ins)[0]
    if ins_idx.size:
        out[ofs[:-1][ins_idx] + 1] = _ACGT[rng.integers(0, 4, size=ins_idx.size)]
    return out


def make_reads(contigs, total_bases: int, seed: int = 2, len_mean=20000, len_sd=2000, len_min=100, len_max=25000,
               acc_mean=0.88, acc_sd=0.07, acc_min=0.75, acc_max=1.0):
    """Return list of (name, uint8 ASCII array) reads whose lengths sum to ~total_bases."""
    rng = np.random.default_rng(seed)
    sizes = np.array([g.size for _, g in contigs], dtype=np.int64)
    cum = np.cumsum(sizes)
    reads = []
    acc_total = 0
    i = 0
    while acc_total < total_bases:
        L = int(np.clip(rng.normal(len_mean, len_sd), len_min, len_max))
        a = float(np.clip(rng.normal(acc_mean, acc_sd), acc_min, acc_max))
        g = int(rng.integers(0, cum[-1]))
        c = int(np.searchsorted(cum, g, side="right"))
        L = min(L, int(sizes[c]))
        pos = int(rng.integers(0, sizes[c] - L + 1))
        src = contigs[c][1][pos:pos + L]
        strand = int(rng.integers(0, 2))
        if strand:
            src = _COMP[src[::-1]]
        rd = _mutate(src, a, rng)
        reads.append((f"S1_{i}", rd))
        acc_total += rd.size
        i += 1
    return reads


def make_hard_reads(contigs, seed: int = 3, n: int = 60):
    """Edge-case reads: chimeras, <1 kb reads, N runs, junk inserts, unmappable, contig-end spanning, very short."""
    rng = np.random.default_rng(seed)
    base = make_reads(contigs, n * 8000, seed=seed + 100, len_mean=8000, len_sd=3000)
    out = []
    for i, (nm, rd) in enumerate(base):
        kind = i % 8
        rd = rd.copy()
        if kind == 0 and i + 1 < len(base):          # chimera of two reads
            rd = np.concatenate((rd[: rd.size // 2], base[i + 1][1][: 3000]))
        elif kind == 1:                               # short
            rd = rd[: int(rng.integers(20, 900))]
        elif kind == 2 and rd.size > 600:             # N run
            p = int(rng.integers(100, rd.size - 300))
            rd[p:p + int(rng.integers(1, 200))] = ord("N")
        elif kind == 3 and rd.size > 600:             # junk insert
            p = int(rng.integers(100, rd.size - 300))
            rd = np.concatenate((rd[:p], _ACGT[rng.integers(0, 4, size=int(rng.integers(50, 1500)))], rd[p:]))
        elif kind == 4:                               # unmappable
            rd = _ACGT[rng.integers(0, 4, size=int(rng.integers(200, 6000)))]
        elif kind == 5:                               # contig start/end spanning
            g = contigs[int(rng.integers(0, len(contigs)))][1]
            L = min(3000, g.size)
            junk = _ACGT[rng.integers(0, 4, size=1500)]
            rd = np.concatenate((junk, g[:L])) if rng.integers(0, 2) else np.concatenate((g[-L:], junk))
        elif kind == 6:                               # tiny (below k) / lowercase
            rd = rd[: int(rng.integers(1, 40))]
        out.append((f"H{kind}_{i}", rd))
    return out


def write_fasta(path, records, width: int = 0):
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            b = seq.tobytes()
            if width:
                for k in range(0, len(b), width):
                    f.write(b[k:k + width] + b"\n")
            else:
                f.write(b + b"\n")


def encode_2bit(seq: np.ndarray) -> np.ndarray:
    """ASCII -> minialign's 1 byte/base codes: the reader indexes a 16-entry table with the LOW NIBBLE of the character
    (encaf, minialign.c:214-232): A,a -> 0, C,c -> 1, G,g -> 2, T,t,U,u -> 3, N,n -> 4, everything else -> 0."""
    enc = np.zeros(16, dtype=np.uint8)
    for ch, v in ((ord("A"), 0), (ord("C"), 1), (ord("G"), 2), (ord("T"), 3), (ord("U"), 3), (ord("N"), 4)):
        enc[ch & 0x0F] = v
    return enc[seq & 0x0F]
