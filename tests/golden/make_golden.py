"""Generates the golden fixtures of tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile.ref).

Run here (needs /root/reference through oracle/_ref):  python tests/golden/make_golden.py
Outputs (committed):
  small.fa / small.mai     60 kb, 3 contigs of unequal length, planted repeats; index built by `minialign -xpacbio -d`
  reads.fa                 simulated + edge-case reads
  golden_align.npz         per read: mm_align_seq result in the flat layout (reference harness, -t1 order, one context)
  golden_stage.npz         per read: mm_sketch words, seed/root arrays after rounds 0 and 2
  golden_extend.npz        random sequence pairs with the reference's fill/search/trace results (pacbio + ont.1dsq scores)
  golden_pacbio.sam / golden_tags.sam   reference CLI output (-t1), without and with -TAS,XS,NM,MD,NH,IH
"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh
from minialign_b200 import synth

REF = os.path.join(ROOT, "oracle/_ref/minialign")


def genome():
    g = synth.make_genome(60_000, 3, seed=41, repeats=((6, 1500), (12, 400)), divergence=0.03)
    return [(g[0][0], g[0][1][:14000]), (g[1][0], g[1][1]), (g[2][0], g[2][1][:9000])]


def reads(g):
    r = synth.make_reads(g, 150_000, seed=42, len_mean=3000, len_sd=1500) + synth.make_hard_reads(g, seed=43, n=32)
    r += synth.make_reads(g, 45_000, seed=44, len_mean=15000, len_sd=3000)
    return r


def pairs(seed, n):
    rng = np.random.default_rng(seed); out = []
    while len(out) < n:
        L = max(2, int(rng.choice([5, 20, 40, 70, 100, 150, 300, 700, 1500])) + int(rng.integers(-3, 30)))
        a = rng.integers(0, 4, size=L).astype(np.uint8)
        acc = float(rng.choice([1.0, 0.95, 0.88, 0.8, 0.7, 0.5]))
        bsc = synth._mutate(np.frombuffer(b"ACGT", dtype=np.uint8)[a], acc, rng)
        if rng.random() < 0.3:
            bsc = np.concatenate((bsc, np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 200)))]))
        b = synth.encode_2bit(bsc)
        if b.size < 2: continue
        if rng.random() < 0.1: a[rng.integers(0, a.size, size=3)] = 4
        if rng.random() < 0.1: b[rng.integers(0, b.size, size=3)] = 4
        brev = int(rng.integers(0, 2))
        if brev: b = np.where(b[::-1] < 4, 3 - b[::-1], 4).astype(np.uint8)
        apos = int(rng.integers(0, max(1, min(a.size, 60)))); bpos = int(rng.integers(0, max(1, min(b.size, 60))))
        if rng.random() < 0.5: apos = bpos = min(apos, bpos)
        out.append((a, b, apos, bpos, brev, int(rng.choice([0, 0, 0, 1, 2]))))
    return out


def pack(arrs, dtype):
    ofs = np.zeros(len(arrs) + 1, dtype=np.int64)
    for i, a in enumerate(arrs): ofs[i + 1] = ofs[i] + len(a)
    return (np.concatenate([np.asarray(a, dtype=dtype) for a in arrs]) if arrs else np.zeros(0, dtype)), ofs


def main():
    g = genome(); rd = reads(g)
    synth.write_fasta(f"{HERE}/small.fa", g, 70); synth.write_fasta(f"{HERE}/reads.fa", rd)
    subprocess.check_call([REF, "-xpacbio", "-d", f"{HERE}/small.mai", f"{HERE}/small.fa"], stderr=subprocess.DEVNULL)
    with open(f"{HERE}/golden_pacbio.sam", "wb") as f:
        subprocess.check_call([REF, "-xpacbio", "-t1", f"{HERE}/small.mai", f"{HERE}/reads.fa"], stdout=f, stderr=subprocess.DEVNULL)
    with open(f"{HERE}/golden_tags.sam", "wb") as f:
        subprocess.check_call([REF, "-xpacbio", "-t1", "-TAS,XS,NM,MD,NH,IH", f"{HERE}/small.mai", f"{HERE}/reads.fa"], stdout=f, stderr=subprocess.DEVNULL)
    h = refh.RefHarness(f"{HERE}/small.mai")
    enc = [synth.encode_2bit(r) for _, r in rd]
    al, alo = pack([h.align(s) for s in enc], np.uint32)
    np.savez_compressed(f"{HERE}/golden_align.npz", words=al, ofs=alo)
    sk, sko = pack([h.sketch(s) if s.size >= 15 else np.zeros(0, np.uint64) for s in enc], np.uint64)
    st = {}
    for rnd in (0, 2):
        ns, sd, rt = [], [], []
        for s in enc:
            if s.size < 15: ns.append(0); sd.append(np.zeros(0, np.uint32)); rt.append(np.zeros(0, np.uint32)); continue
            n, a, b = h.seed_chain(s, rnd); ns.append(n); sd.append(a.reshape(-1)); rt.append(b.reshape(-1))
        st[f"ns{rnd}"] = np.array(ns, dtype=np.int64)
        st[f"seed{rnd}"], st[f"seedo{rnd}"] = pack(sd, np.uint32); st[f"root{rnd}"], st[f"rooto{rnd}"] = pack(rt, np.uint32)
    np.savez_compressed(f"{HERE}/golden_stage.npz", sketch=sk, sketch_ofs=sko, **st)
    ext = {}
    for preset, seed in (("pacbio", 51), ("ont.1dsq", 52)):
        hh = refh.RefHarness(f"{HERE}/small.mai", args=("-x" + preset,))
        ps = pairs(seed, 150)
        res, alns = [], []
        for a, b, apos, bpos, brev, narrow in ps:
            r, o = hh.extend(a, b, apos, bpos, brev, narrow, 0); res.append(r); alns.append(o)
        key = preset.split(".")[0]
        ext[f"{key}_a"], ext[f"{key}_ao"] = pack([p[0] for p in ps], np.uint8); ext[f"{key}_b"], ext[f"{key}_bo"] = pack([p[1] for p in ps], np.uint8)
        ext[f"{key}_args"] = np.array([p[2:] for p in ps], dtype=np.int64)
        ext[f"{key}_res"] = np.stack(res); ext[f"{key}_aln"], ext[f"{key}_alno"] = pack(alns, np.uint32)
        hh.close()
    np.savez_compressed(f"{HERE}/golden_extend.npz", **ext)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
