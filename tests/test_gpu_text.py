"""The text path on the GPU (device reader + post-processing + SAM printer behind mab_text_*): golden SAM of the reference CLI,
the host formatter on the record-level results, live reference runs (-t1) for index / scoring variants, several contexts and --
when two devices are visible -- two GPUs in one process and two ranks over NCCL."""
import ctypes as C
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLD, ROOT
from minialign_b200 import api, mai, synth
import refh

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "minialign_b200", "minialign-b200")


def nopg(text):
    return [l for l in text.split("\n") if not l.startswith("@PG")]


def run_cli(args, to_file=None):
    if to_file:
        with open(to_file, "wb") as f:
            p = subprocess.run([CLI, *args], stdout=f, stderr=subprocess.PIPE)
        out = open(to_file, "rb").read()
    else:
        p = subprocess.run([CLI, *args], capture_output=True)
        out = p.stdout
    assert p.returncode == 0, p.stderr.decode()[-800:]
    return nopg(out.decode()), p.stderr.decode()


def ref_cli(args):
    p = subprocess.run([refh.BIN, *args], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()[-400:]
    return nopg(p.stdout.decode())


@pytest.mark.parametrize("golden,taglist", [("golden_pacbio.sam", ""), ("golden_tags.sam", "AS,XS,NM,MD,NH,IH")])
def test_map_text_matches_golden_sam(gold, golden, taglist):
    m = api.Mapper(gold["blob"], "pacbio")
    got = m.map_text(open(os.path.join(GOLD, "reads.fa"), "rb").read(), api.parse_tags(taglist)).decode()
    st = m.stats()
    m.close()
    exp = [l for l in open(os.path.join(GOLD, golden)).read().split("\n") if not l.startswith("@")]
    assert got.split("\n") == exp
    assert st["n_launches"] >= 12 and st["n_failed"] == 0


def test_staged_set_up_and_announced_chunk_size(gold):
    """mab_load_begin / put / end (pieces out of order, every visible device) gives the contexts mab_init gives; with the chunk
    size announced a short chunk followed by the whole file maps like without"""
    import torch
    text = open(os.path.join(GOLD, "reads.fa"), "rb").read()
    short = text[: text.index(b">", 4000)]
    m0 = api.Mapper(gold["blob"], "pacbio")
    exp = [m0.map_text(short), m0.map_text(text)]
    m0.close()
    devs = tuple(range(min(2, torch.cuda.device_count())))
    ms = api.Mapper.staged(gold["blob"], "pacbio", devices=devs, piece=(1 << 20) + 4096, order=lambda st: st[::-1])
    for m in ms:
        m.text_reserve(len(text) + 1000)
        c = m.clone()
        assert [m.map_text(short), m.map_text(text)] == exp and c.map_text(text) == exp[1]
        c.close()
    for m in ms:
        m.close()


def test_map_text_device_resident_io(gold):
    """device-input + device-output mode (bench.py's kernel-side arm): same byte count as the host-buffer mode"""
    import torch
    text = open(os.path.join(GOLD, "reads.fa"), "rb").read()
    m = api.Mapper(gold["blob"], "pacbio")
    host = m.map_text(text)
    d = torch.frombuffer(bytearray(text + b"\n" * 64), dtype=torch.uint8).cuda()
    m.lib.mab_set_device_input(m.h, 1)
    m.text_begin(d.data_ptr(), len(text), api.TEXT_DEVICE_OUT, 0, True)
    info, ptr = m.text_finish()
    m.close()
    assert info.sam_bytes == len(host) and info.n_reads == host.count(b"\n") - sum(1 for l in host.split(b"\n") if l and int(l.split(b"\t")[1]) & 0x900)


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
@pytest.mark.parametrize("iargs", [["-k17", "-w12"], ["-k13", "-w7"]])
def test_cli_index_parameter_variants_vs_live_reference(tmp_path, iargs):
    """k > 16 (the CRC term of the minimizer hash is live on the device) and a window that is not 10, mapped -- not only indexed."""
    g = synth.make_genome(900_000, 3, seed=91, repeats=((10, 2000), (40, 500)), weights=[5, 1, 3])
    reads = synth.make_reads(g, 3_000_000, seed=92, len_mean=6000, len_sd=2500) + synth.make_hard_reads(g, seed=93, n=24)
    fa, rd, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-xpacbio", *iargs, "-d", idx, fa], stderr=subprocess.DEVNULL)
    exp = ref_cli(["-xpacbio", "-t1", "-TAS,XS,NM,MD,SA", idx, rd])
    got, _ = run_cli(["-xpacbio", "-TAS,XS,NM,MD,SA", "-c3", "-N0.4", idx, rd])
    assert len(got) == len(exp) and sum(1 for l in exp if l and not l.startswith("@") and l.split("\t")[1] != "4") > 100
    bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not bad, (len(bad), got[bad[0]][:200], exp[bad[0]][:200])
    # the FASTA reference on the command line with the same index parameters: built by mab_index.cpp
    got_fa, _ = run_cli(["-xpacbio", *iargs, "-TAS,XS,NM,MD,SA", fa, rd])
    assert got_fa == exp


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
def test_cli_three_contexts_uneven_contigs_fastq_gz_vs_live_reference(tmp_path):
    """sacCer3-like in small: 17 contigs of very different lengths (the reference thread's stale `rlen` flips the first seed test
    of many reads), 20 kb reads, three contexts with small chunks, FASTQ / gz / -Q, regular-file and pipe output."""
    g = synth.make_genome(5_000_000, 17, seed=95, repeats=((20, 3000), (80, 900)),
                          weights=[230, 813, 317, 1532, 577, 270, 1091, 563, 440, 746, 667, 1078, 924, 784, 1091, 948, 86])
    reads = synth.make_reads(g, 9_000_000, seed=96) + synth.make_hard_reads(g, seed=97)
    fa, rd, fq, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "r.fq"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    rng = np.random.default_rng(3)
    with open(fq, "wb") as f:
        for name, s in reads:
            f.write(b"@" + name.encode() + b" x\n" + s.tobytes() + b"\n+\n" + bytes(rng.integers(33, 74, size=s.size).astype(np.uint8)) + b"\n")
    with open(fq, "rb") as f, gzip.open(fq + ".gz", "wb", compresslevel=1) as z:
        z.write(f.read())
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    exp = ref_cli(["-xpacbio", "-t1", idx, rd])
    got, err = run_cli(["-xpacbio", "-c3", "-N2", idx, rd])
    assert got == exp, err[-400:]
    got1, _ = run_cli(["-xpacbio", "-c1", "-N1000", idx, rd], to_file=str(tmp_path / "o.sam"))
    assert got1 == exp
    expq = ref_cli(["-xpacbio", "-t1", "-Q", "-TAS,NM,MD,SA", idx, fq])
    gotq, _ = run_cli(["-xpacbio", "-Q", "-TAS,NM,MD,SA", "-c3", "-N3", idx, fq + ".gz"], to_file=str(tmp_path / "q.sam"))
    assert gotq == expq
    gotn, _ = run_cli(["-xpacbio", "-TAS,NM,MD,SA", "-c2", "-N5", idx, fq])
    assert gotn == ref_cli(["-xpacbio", "-t1", "-TAS,NM,MD,SA", idx, fq])


def _two_gpus():
    try:
        import torch
        return torch.cuda.device_count() >= 2
    except Exception:
        return False


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
@pytest.mark.skipif(not _two_gpus(), reason="needs two visible devices")
def test_two_gpus_one_merged_sam_vs_live_reference(tmp_path):
    """ONE read file across two GPUs, (a) in one process (`-g0,1`), (b) as two ranks over NCCL (mgpu.py): one SAM, equal to -t1."""
    g = synth.make_genome(3_000_000, 9, seed=101, weights=[9, 1, 5, 2, 7, 3, 4, 6, 1])
    reads = synth.make_reads(g, 12_000_000, seed=102)
    fa, rd, idx, out = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai"), str(tmp_path / "m.sam")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    exp = ref_cli(["-xpacbio", "-t1", "-TAS,NM", idx, rd])
    got, _ = run_cli(["-xpacbio", "-TAS,NM", "-g0,1", "-c2", "-N1.5", idx, rd])
    assert got == exp
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29571",
                        "-m", "minialign_b200.mgpu", "-xpacbio", "-TAS,NM", "-c2", "-N1.5", "-o", out, idx, rd], capture_output=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-1500:]
    assert nopg(open(out).read()) == exp
