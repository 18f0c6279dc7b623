"""INTEGRATION.md made executable: the reference binary with its batch worker (mm_align_worker, minialign.c:4589-4601) swapped
for mab_map_batch + a words -> mm_reg_t converter (oracle/ref_gpuworker.c).  The reference's own reader, thread pool, ordered
drain and SAM printer run unchanged, so the SAM must equal the stock binary's.  On the CPU the binding is linked against the
emulation build of the library; on the GPU box (`-m gpu`) against libminialign_b200.so (built by oracle/Makefile.ref)."""
import os
import subprocess

import pytest

from conftest import GOLD, ROOT, build_emu
import refh

REF_SRC = "/root/reference/minialign.c"
GW = os.path.join(ROOT, "oracle", "_ref", "minialign-gpuworker")


def sam(args, binary):
    p = subprocess.run([binary, *args], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()[-600:]
    return [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]


def small_reads(tmp_path, max_len, n):
    lines = open(os.path.join(GOLD, "reads.fa")).read().split("\n")
    recs = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 1, 2) if len(lines[i + 1]) <= max_len][:n]
    fa = str(tmp_path / "sub.fa")
    with open(fa, "w") as f:
        for h, s in recs:
            f.write(h + "\n" + s + "\n")
    return fa


@pytest.mark.skipif(not (os.path.exists(REF_SRC) and os.path.exists(refh.BIN)), reason="needs the reference source tree and oracle/_ref")
def test_reference_binary_with_the_emulated_worker(tmp_path):
    so = build_emu()
    exe = str(tmp_path / "minialign-gpuworker-emu")
    objs = [os.path.join(ROOT, "oracle", "_ref", f"gaba.{m}.{b}.pic.o") for m in ("linear", "affine", "combined") for b in (16, 32, 64)]
    subprocess.check_call(["gcc", "-o", exe, "-O2", "-std=c99", "-w", "-DMM_VERSION=\"minialign-0.6.0-devel\"", "-DUNITTEST=0", "-mavx2", "-mbmi", "-mbmi2", "-mlzcnt", "-mpopcnt",
                           "-I/root/reference", f"-DREF_SRC=\"{REF_SRC}\"", os.path.join(ROOT, "oracle", "ref_gpuworker.c"), *objs,
                           "-L" + os.path.dirname(so), "-lmab_emu", "-Wl,-rpath," + os.path.dirname(so), "-lm", "-lz", "-lpthread"])
    fa = small_reads(tmp_path, 3000, 36)
    idx = os.path.join(GOLD, "small.mai")
    for tags in ([], ["-TAS,XS,NM,MD,SA"]):
        exp = sam(["-xpacbio", "-t1", *tags, idx, fa], refh.BIN)
        env_small = dict(os.environ, GW_BATCH_KB="16")             # several batches: the worker's state carries over like the reference thread's
        p = subprocess.run([exe, "-xpacbio", "-t1", *tags, idx, fa], capture_output=True, env=env_small)
        assert p.returncode == 0, p.stderr.decode()[-600:]
        got = [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]
        assert got == exp and sum(1 for l in exp if l and not l.startswith("@") and l.split("\t")[1] != "4") > 10


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(GW) and os.path.exists(refh.BIN)), reason="oracle/_ref/minialign-gpuworker not built")
def test_reference_binary_with_the_gpu_worker(tmp_path):
    from minialign_b200 import synth
    g = synth.make_genome(2_000_000, 5, seed=111, weights=[6, 1, 3, 2, 4])
    reads = synth.make_reads(g, 8_000_000, seed=112) + synth.make_hard_reads(g, seed=113)
    fa, rd, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    for tags in ([], ["-TAS,XS,NM,MD,SA"]):
        exp = sam(["-xpacbio", "-t1", *tags, idx, rd], refh.BIN)
        p = subprocess.run([GW, "-xpacbio", "-t1", *tags, idx, rd], capture_output=True, env=dict(os.environ, GW_BATCH_KB="2048"))
        assert p.returncode == 0, p.stderr.decode()[-600:]
        got = [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]
        assert got == exp
