"""First on-GPU shake-out: stage parity (extend pairs, sketch, seed/chain) and end-to-end map_batch against the oracle."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ora
from minialign_b200 import synth, mai, api

out = os.path.join(ROOT, "gpurun_out"); os.makedirs(out, exist_ok=True)
work = "/tmp/mab_first"; os.makedirs(work, exist_ok=True)
g = synth.make_genome(1_000_000, 2, seed=1)
synth.write_fasta(f"{work}/g.fa", g, 80)
ref = os.path.join(ROOT, "oracle/_ref/minialign")
subprocess.check_call([ref, "-xpacbio", "-d", f"{work}/g.mai", f"{work}/g.fa"], stderr=subprocess.DEVNULL)
blob = mai.load_mai(f"{work}/g.mai")
hd = mai.parse_header(blob)
m = api.Mapper(blob, "pacbio")
o = ora.Oracle(dict(ora.PACBIO, occ=hd["occ"][:3]), blob)
rng = np.random.default_rng(7)
pairs = []
for it in range(400):
    L = max(2, int(rng.choice([5, 20, 40, 70, 100, 150, 300, 700, 2000, 6000])) + int(rng.integers(-3, 30)))
    a = rng.integers(0, 4, size=L).astype(np.uint8)
    acc = float(rng.choice([1.0, 0.95, 0.88, 0.8, 0.7, 0.5]))
    bsc = synth._mutate(np.frombuffer(b"ACGT", dtype=np.uint8)[a], acc, rng)
    b = synth.encode_2bit(bsc)
    if b.size < 2: continue
    brev = int(rng.integers(0, 2))
    if brev: b = np.where(b[::-1] < 4, 3 - b[::-1], 4).astype(np.uint8)
    apos = int(rng.integers(0, max(1, min(a.size, 60)))); bpos = int(rng.integers(0, max(1, min(b.size, 60))))
    pairs.append((a, b, apos, bpos, brev, int(rng.choice([0, 0, 0, 1, 2])), 0))
t = time.time(); got = m.extend_pairs(pairs); print("extend_pairs", len(pairs), "time", time.time() - t, flush=True)
bad = 0
for p, (r2, o2) in zip(pairs, got):
    r1, o1 = o.extend(*p[:6], p[6])
    bad += not (np.array_equal(r1, r2) and np.array_equal(o1, o2))
print("extend_pairs bad", bad, flush=True)
reads = synth.make_reads(g, 3_000_000, seed=2) + synth.make_hard_reads(g, seed=3)
enc = [synth.encode_2bit(r) for _, r in reads]
bad = 0
for s in enc[:20]:
    if s.size < 15: continue
    a = o.sketch(s); b = m.sketch(s)
    bad += not (len(a) == len(b) and np.array_equal(a[:-3], b[:-3]))
    for rnd in (0, 2):
        x = o.seed_chain(s, rnd); y = m.seed_chain(s, rnd)
        bad += not (x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]))
print("sketch/seed bad", bad, flush=True)
for rep in range(2):
    t = time.time(); res = m.map_batch(enc); dt = time.time() - t
    print("map_batch", len(enc), "reads", sum(e.size for e in enc), "bases wall", dt, m.stats(), flush=True)
bad = 0; nmap = 0
t = time.time()
for s, gr in zip(enc, res):
    exp = o.align(s); nmap += len(exp) > 0
    bad += not np.array_equal(exp, gr)
print("map_batch mapped", nmap, "bad", bad, "oracle time", time.time() - t, flush=True)
