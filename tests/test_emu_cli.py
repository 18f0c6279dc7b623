"""The command line's host pipeline (chunking reader, several contexts per device, in-order rlen commit, writers, host fallback
for text the device reader does not take) run end to end on the emulation build: `minialign-emu` is mab_cli.cpp linked against
tests/emu/libmab_emu.so instead of the CUDA library.  Expected output: the reference CLI's SAM (golden, or live with -t1)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, build_emu_cli
import refh


def run(args, stdin=None, to_file=None):
    cli = build_emu_cli()
    if to_file:
        with open(to_file, "wb") as f:
            p = subprocess.run([cli, *args], stdout=f, stderr=subprocess.PIPE)
        out = open(to_file, "rb").read()
    else:
        p = subprocess.run([cli, *args], capture_output=True, input=stdin)
        out = p.stdout
    assert p.returncode == 0, p.stderr.decode()[-600:]
    return [l for l in out.decode().split("\n") if not l.startswith("@PG")], p.stderr.decode()


def golden(name):
    return [l for l in open(os.path.join(GOLD, name)).read().split("\n") if not l.startswith("@PG")]


def small_reads_fa(tmp_path, max_len=3000, n=40):
    """a prefix-closed subset would change the reference's order-dependent results, so the expected text is computed for the
    same subset file by the one-chunk, one-context run (itself compared with the golden SAM in test_emu_text.py)"""
    lines = open(os.path.join(GOLD, "reads.fa")).read().split("\n")
    recs = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 1, 2) if len(lines[i + 1]) <= max_len][:n]
    fa = str(tmp_path / "sub.fa")
    with open(fa, "w") as f:
        for h, s in recs:
            f.write(h + "\n" + s + "\n")
    return fa, recs


@pytest.mark.parametrize("tags", [[], ["-TAS,XS,NM,MD,NH,IH"]])
def test_emu_cli_many_chunks_three_contexts(tmp_path, tags):
    fa, recs = small_reads_fa(tmp_path)
    mai_path = os.path.join(GOLD, "small.mai")
    one, _ = run(["-xpacbio", *tags, "-c1", mai_path, fa])
    many, err = run(["-xpacbio", *tags, "-c3", "-N0.008", mai_path, fa])                 # ~8 KB chunks: a few reads each
    assert many == one and sum(1 for l in one if l and not l.startswith("@")) >= len(recs)
    two_dev, _ = run(["-xpacbio", *tags, "-c2", "-g0,0", "-N0.02", mai_path, fa], to_file=str(tmp_path / "o.sam"))   # regular file: parallel pwrite path
    assert two_dev == one


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
def test_emu_cli_fastq_gz_qualities_vs_live_reference(tmp_path):
    """FASTQ (plain and gzipped), with and without -Q, wrapped FASTQ (host fallback), multi-line FASTA; uneven contigs; the
    reference run with -t1 on the same files."""
    from minialign_b200 import synth
    g = synth.make_genome(80_000, 5, seed=81, repeats=((6, 500),), weights=[20, 2, 9, 3, 6])
    reads = synth.make_reads(g, 45_000, seed=82, len_mean=700, len_sd=300, len_min=40, len_max=1800)
    fa, idx = str(tmp_path / "g.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 60)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    rng = np.random.default_rng(5)
    fq, fqw, fam = str(tmp_path / "r.fq"), str(tmp_path / "rw.fq"), str(tmp_path / "rm.fa")
    with open(fq, "wb") as f, open(fqw, "wb") as fw, open(fam, "wb") as fm:
        for name, s in reads:
            q = bytes(rng.integers(33, 74, size=s.size).astype(np.uint8))
            b = s.tobytes()
            f.write(b"@" + name.encode() + b" desc\n" + b + b"\n+\n" + q + b"\n")
            fw.write(b"@" + name.encode() + b"\n" + b"\n".join(b[k:k + 70] for k in range(0, len(b), 70)) + b"\n+" + name.encode() + b"\n"
                     + b"\n".join(q[k:k + 70] for k in range(0, len(q), 70)) + b"\n")
            fm.write(b">" + name.encode() + b"\tx y\n" + b"\n".join(b[k:k + 50] for k in range(0, len(b), 50)) + b"\n\n")
    with open(fq, "rb") as f, gzip.open(fq + ".gz", "wb") as z:
        z.write(f.read())

    def ref(args, path):
        p = subprocess.run([refh.BIN, "-xpacbio", "-t1", *args, idx, path], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()[-300:]
        return [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]

    for args, path in (([], fq), (["-Q"], fq), (["-Q", "-TAS,NM,MD,SA"], fq + ".gz"), (["-Q"], fqw), ([], fam)):
        exp = ref(args, path)
        got, err = run(["-xpacbio", *args, "-c2", "-N0.03", idx, path])
        assert len(got) == len(exp), (args, path)
        bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
        assert not bad, (args, path, len(bad), got[bad[0]][:300], exp[bad[0]][:300])
        if path == fqw:
            assert "chunks parsed on the host: 0" not in err          # wrapped FASTQ goes through the host reader
