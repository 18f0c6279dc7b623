"""Parity proper: the CUDA path, called through the C ABI, against the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest

import ora
from conftest import build_emu, gold_pairs, unpack
from minialign_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(gold):
    m = api.Mapper(gold["blob"], "pacbio")
    yield m
    m.close()


def test_intrinsics_match_host_semantics(gpu, gold):
    """Every packed-SIMD / permute / warp primitive of the DP, device vs the CUDA-on-CPU shim, word for word."""
    e = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    a, b = gpu.selftest(), e.selftest()
    e.close()
    n = int(b[63, 0])
    assert n >= 40 and np.array_equal(a[:n], b[:n])


def test_sketch_seed_chain_golden(gpu, gold):
    st = gold["stage"]
    sk = unpack(st["sketch"], st["sketch_ofs"])
    for rnd in (0, 2):
        seeds, roots = unpack(st[f"seed{rnd}"], st[f"seedo{rnd}"]), unpack(st[f"root{rnd}"], st[f"rooto{rnd}"])
        for i, s in enumerate(gold["enc"]):
            if s.size < 15:
                continue
            if rnd == 0:
                w = gpu.sketch(s)
                assert len(w) == len(sk[i]) and np.array_equal(w[:-3], sk[i][:-3])
            ns, sd, rt = gpu.seed_chain(s, rnd)
            assert ns == st[f"ns{rnd}"][i] and np.array_equal(sd.reshape(-1), seeds[i]) and np.array_equal(rt.reshape(-1), roots[i])


@pytest.mark.parametrize("key,preset", [("pacbio", "pacbio"), ("ont", "ont.1dsq")])
def test_extend_pairs_golden(gold, key, preset):
    m = api.Mapper(gold["blob"], preset)
    pairs, res, alns = gold_pairs(gold["extend"], key)
    got = m.extend_pairs(pairs)
    m.close()
    for (r2, a2), r, a in zip(got, res, alns):
        assert np.array_equal(r, r2) and np.array_equal(a, a2)


@pytest.mark.parametrize("preset,prm", [("pacbio", ora.PACBIO), ("ont.1dsq", ora.ONT)] + [(None, p) for p, _ in ora.CUSTOM])
def test_extend_pairs_fuzz_vs_oracle(gold, preset, prm):
    m = api.Mapper(gold["blob"], preset or {k: prm[k] for k in ora.API_KEYS})
    o = ora.Oracle(prm)
    rng = np.random.default_rng(99)
    pairs = []
    while len(pairs) < 600:
        L = max(2, int(rng.choice([2, 5, 20, 40, 63, 64, 65, 100, 150, 300, 700, 2000, 9000])) + int(rng.integers(0, 30)))
        a = rng.integers(0, 4, size=L).astype(np.uint8)
        b = synth.encode_2bit(synth._mutate(np.frombuffer(b"ACGT", dtype=np.uint8)[a], float(rng.choice([1.0, 0.95, 0.88, 0.8, 0.7, 0.5])), rng))
        if b.size < 2:
            continue
        if rng.random() < 0.1:
            a[rng.integers(0, a.size, size=3)] = 4
        brev = int(rng.integers(0, 2))
        if brev:
            b = np.where(b[::-1] < 4, 3 - b[::-1], 4).astype(np.uint8)
        pairs.append((a, b, int(rng.integers(0, min(a.size, 60))), int(rng.integers(0, min(b.size, 60))), brev, int(rng.choice([0, 0, 0, 1, 2])), 0))
    got = m.extend_pairs(pairs)
    m.close()
    for p, (r2, a2) in zip(pairs, got):
        r1, a1 = o.extend(*p[:6], p[6])
        assert np.array_equal(r1, r2) and np.array_equal(a1, a2)


def test_map_batch_golden(gpu, gold):
    """All golden reads in file order == the reference's mm_align_seq results (-t1 order)."""
    got = gpu.map_batch(gold["enc"])
    assert sum(len(g) > 0 for g in got) > 50
    for e, g in zip(gold["align"], got):
        assert np.array_equal(e, g)


def test_map_batch_golden_unstaged_sortchain(gold, monkeypatch):
    """Same golden batch with the shared-memory staging of k_sortchain capped at 64 seeds: nearly every read takes the
    global-memory path of the exact sort / chaining, results must not change."""
    monkeypatch.setenv("MAB_SC_CAP", "64")
    m = api.Mapper(gold["blob"], "pacbio")
    got = m.map_batch(gold["enc"])
    m.close()
    for e, g in zip(gold["align"], got):
        assert np.array_equal(e, g)


def test_long_reads_under_a_small_arena_budget(tmp_path, monkeypatch):
    """Reads far longer than the benchmark's (~60-150 kb) with the DP-arena budget squeezed so that only a few warps stay
    resident: exact against the oracle."""
    import subprocess
    import refh
    from minialign_b200 import mai
    if not os.path.exists(refh.BIN):
        pytest.skip("oracle/_ref/minialign not built (needed to build the index)")
    monkeypatch.setenv("MAB_ARENA_BUDGET_MB", "512")
    g = synth.make_genome(1_500_000, 2, seed=41)
    fa, idx = str(tmp_path / "g.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    blob = mai.load_mai(idx)
    hdr = mai.parse_header(blob)
    reads = synth.make_reads(g, 1_200_000, seed=5, len_mean=100000, len_sd=30000, len_max=160000)
    enc = [synth.encode_2bit(r) for _, r in reads]
    assert max(e.size for e in enc) > 64000
    m = api.Mapper(blob, "pacbio")
    got = m.map_batch(enc)
    m.close()
    o = ora.Oracle(dict(ora.PACBIO, occ=hdr["occ"][:3]), blob)
    for s, x in zip(enc, got):
        assert np.array_equal(o.align(s), x)
    o.close()
    assert sum(len(x) > 0 for x in got) >= len(enc) // 2


def test_fill_peak_reports_a_ceiling(gpu):
    """The integer-roofline microbenchmark runs and the masked step is not faster than the unmasked one."""
    pm, pu = gpu.fill_peak(True, 500), gpu.fill_peak(False, 500)
    assert 0 < pm <= pu * 1.05


def test_map_batch_edge_cases(gold):
    m = api.Mapper(gold["blob"], "pacbio")
    assert m.map_batch([]) == []
    out = m.map_batch([np.zeros(3, dtype=np.uint8), np.full(500, 4, dtype=np.uint8), np.zeros(14, dtype=np.uint8)])   # < k, all-N, k-1
    assert all(len(o) == 0 for o in out)
    m.close()


def test_map_batch_full_size_vs_oracle_and_properties(gold):
    """20 kb reads (BASELINE read model): exact vs the oracle on a sample, plus size-independent properties on all reads:
    determinism across batch splits, path/segment consistency (#1 bits = blen, #0 bits = alen), score bounds."""
    g = [(n, s) for n, s in _contigs(gold)]
    reads = synth.make_reads(g, 1_200_000, seed=77)
    enc = [synth.encode_2bit(r) for _, r in reads]
    m = api.Mapper(gold["blob"], "pacbio")
    a = m.map_batch(enc)
    m.close()
    m2 = api.Mapper(gold["blob"], "pacbio")
    b = m2.map_batch(enc[: len(enc) // 2]) + m2.map_batch(enc[len(enc) // 2:])
    m2.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    o = ora.Oracle(dict(ora.PACBIO, occ=gold["hdr"]["occ"][:3]), gold["blob"])
    for s, x in list(zip(enc, a))[:12]:
        assert np.array_equal(o.align(s), x)
    o.close()
    n_aln = 0
    for s, w in zip(enc, a):
        if len(w) == 0:
            continue
        p = 2
        for _ in range(int(w[0])):
            slen, plen, npw = int(w[p + 7]), int(w[p + 8]), int(w[p + 9])
            segs = w[p + 16:p + 16 + 8 * slen].reshape(-1, 8)
            path = w[p + 16 + 8 * slen:p + 16 + 8 * slen + npw]
            bits = np.unpackbits(path.view(np.uint8), bitorder="little")[:plen]
            assert int(segs[:, 4].sum() + segs[:, 5].sum()) == plen
            assert int(bits.sum()) == int(segs[:, 5].sum()) and plen - int(bits.sum()) == int(segs[:, 4].sum())
            assert int(segs[:, 5].sum()) <= s.size
            score = int(w[p]) | (int(w[p + 1]) << 32)
            assert 0 < score <= 2 * s.size
            p += 16 + 8 * slen + npw
            n_aln += 1
    assert n_aln >= len(enc) // 2


def _contigs(gold):
    from minialign_b200 import mai
    raw = gold["blob"]
    dec = np.frombuffer(b"ACGTN", dtype=np.uint8)
    out = []
    for name, l_seq, ofs in mai.ref_seqs(raw):
        out.append((name, dec[raw[ofs:ofs + l_seq]]))
    return out
