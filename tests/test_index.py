"""Host-side index construction (`minialign-b200 -d`, minialign_b200/csrc/host/mab_index.cpp) against an index built by the
unmodified reference: same parameters, same occurrence thresholds, and for every minimizer the same occurrences in the same
order (the order is observable through the unstable seed sort)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT
from minialign_b200 import mai, synth
import refh

CLI = os.path.join(ROOT, "minialign_b200", "minialign-b200")


def build_cli():
    subprocess.check_call(["make", "-s", "-f", "minialign_b200/csrc/host/Makefile"], cwd=ROOT)


def tables(blob: np.ndarray):
    """{bucket: {key: [(pos, rid), ...]}} decoded from a relocatable index image."""
    h = mai.parse_header(blob)
    nb = h["mask"] + 1
    bk = np.frombuffer(blob, dtype=np.dtype([("mask", "<u4"), ("max", "<u4"), ("cnt", "<u4"), ("ub", "<u4"), ("a", "<u8"), ("p", "<u8")]), count=nb, offset=h["bkt"])
    out = {}
    for i in np.nonzero(bk["a"])[0]:
        a, p, n = int(bk["a"][i]), int(bk["p"][i]), int(bk["mask"][i]) + 1
        slots = np.frombuffer(blob, dtype="<u8", count=2 * n, offset=a).reshape(n, 2)
        used = slots[slots[:, 0] < np.uint64(0xFFFFFFFFFFFFFFFE)]
        np_words = int(np.frombuffer(blob, dtype="<u8", count=1, offset=p)[0])
        parr = np.frombuffer(blob, dtype="<u8", count=np_words + 1, offset=p)
        d = {}
        for key, val in used.tolist():
            if val >> 63:
                cnt, first = val & 0xFFFFFFFF, (val >> 32) & 0x7FFFFFFF
                d[key] = [(int(x) & 0xFFFFFFFF, int(x) >> 32) for x in parr[first:first + cnt]]
            else:
                d[key] = [(val & 0xFFFFFFFF, val >> 32)]
        out[int(i)] = d
    return h, out


def probe(blob: np.ndarray, h: dict, bucket: int, key: int):
    """the probe of mm_idx_get / kh_get_ptr (minialign.c:2727-2748, 634-643) on an image"""
    hmask, _, _, _, a, p = struct.unpack_from("<IIIIQQ", blob, h["bkt"] + 32 * bucket)
    if a == 0:
        return None
    pos = key & hmask
    while True:
        k, v = struct.unpack_from("<QQ", blob, a + 16 * pos)
        if k == key:
            return v
        if k == 0xFFFFFFFFFFFFFFFF:
            return None
        pos = (pos + 1) & hmask


def check_same(ours: np.ndarray, ref: np.ndarray):
    ho, to = tables(ours)
    hr, tr = tables(ref)
    for f in ("mask", "b", "w", "k", "n_occ", "n_seq"):
        assert ho[f] == hr[f], f
    assert ho["occ"][:hr["n_occ"]] == hr["occ"][:hr["n_occ"]]
    assert [(n, l) for n, l, _ in mai.ref_seqs(ours)] == [(n, l) for n, l, _ in mai.ref_seqs(ref)]
    for (_, l, o1), (_, _, o2) in zip(mai.ref_seqs(ours), mai.ref_seqs(ref)):
        assert np.array_equal(ours[o1:o1 + l], ref[o2:o2 + l])
    assert to.keys() == tr.keys()
    n_keys = 0
    for b in tr:
        assert to[b] == tr[b], b                      # same keys, same occurrences, same order
        for key in list(tr[b])[:3]:
            assert probe(ours, ho, b, key) is not None      # and our tables answer the reference's probe
        n_keys += len(tr[b])
    return n_keys


def test_index_of_golden_reference_matches(tmp_path):
    build_cli()
    out = str(tmp_path / "ours.mai")
    subprocess.check_call([CLI, "-xpacbio", "-d", out, os.path.join(GOLD, "small.fa")], stderr=subprocess.DEVNULL)
    n = check_same(mai.load_mai(out), mai.load_mai(os.path.join(GOLD, "small.mai")))
    assert n > 1000


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
@pytest.mark.parametrize("args", [["-xpacbio"], ["-k13", "-w7", "-B10", "-f0.1,0.02"], ["-k17", "-w12"]])
def test_index_matches_live_reference(tmp_path, args):
    """Repeat-rich multi-contig genome with N runs: buckets larger than the 64-element insertion-sort cutoff, minimizers above
    the occurrence threshold, k > 16 (the CRC term of the hash is live)."""
    build_cli()
    g = synth.make_genome(300_000, 5, seed=51, repeats=((20, 3000), (100, 800), (700, 120)))
    g = [(n, s.copy()) for n, s in g]
    g[1][1][5000:5400] = ord("N"); g[3][1][100:130] = ord("n")
    fa, a, b = str(tmp_path / "g.fa"), str(tmp_path / "ours.mai"), str(tmp_path / "ref.mai")
    synth.write_fasta(fa, g, 70)
    subprocess.check_call([CLI, *args, "-d", a, fa], stderr=subprocess.DEVNULL)
    subprocess.check_call([refh.BIN, *args, "-d", b, fa], stderr=subprocess.DEVNULL)
    assert check_same(mai.load_mai(a), mai.load_mai(b)) > 5000
