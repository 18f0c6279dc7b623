"""Drop-in check at the command line: `minialign-b200 -xpacbio ref.mai reads.fa` must print the reference's SAM byte for byte
(every line except @PG)."""
import os
import subprocess

import pytest

from conftest import GOLD, ROOT
import refh

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "minialign_b200", "minialign-b200")


def run_cli(args):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-f", "minialign_b200/csrc/host/Makefile"], cwd=ROOT)
    p = subprocess.run([CLI, *args], capture_output=True)
    assert p.returncode == 0, p.stderr.decode()[-500:]
    return [l for l in p.stdout.decode().split("\n") if not l.startswith("@PG")]


@pytest.mark.parametrize("golden,extra", [("golden_pacbio.sam", []), ("golden_tags.sam", ["-TAS,XS,NM,MD,NH,IH"])])
def test_cli_matches_golden_sam(golden, extra):
    got = run_cli(["-xpacbio", *extra, os.path.join(GOLD, "small.mai"), os.path.join(GOLD, "reads.fa")])
    exp = [l for l in open(os.path.join(GOLD, golden)).read().split("\n") if not l.startswith("@PG")]
    assert got == exp


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
@pytest.mark.parametrize("preset", ["pacbio", "ont.1dsq"])
def test_cli_matches_live_reference_on_20kb_reads(tmp_path, preset):
    """BASELINE config 0 in small: E.coli-like reference, PBSIM-like 20 kb reads, reference run with -t1 on the box's CPU."""
    from minialign_b200 import synth
    g = synth.make_genome(1_000_000, 2, seed=21)
    reads = synth.make_reads(g, 6_000_000, seed=22) + synth.make_hard_reads(g, seed=23)
    fa, rd, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-x" + preset, "-d", idx, fa], stderr=subprocess.DEVNULL)
    ref = subprocess.run([refh.BIN, "-x" + preset, "-t1", "-TAS,XS,NM,MD,SA", idx, rd], capture_output=True)
    assert ref.returncode == 0
    exp = [l for l in ref.stdout.decode().split("\n") if not l.startswith("@PG")]
    got = run_cli(["-x" + preset, "-TAS,XS,NM,MD,SA", "-c2", "-N1.3", idx, rd])        # several chunks: state must carry across chunks and contexts
    assert len(got) == len(exp)
    bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not bad, (len(bad), got[bad[0]][:200], exp[bad[0]][:200])


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
@pytest.mark.parametrize("opts", [["-a1", "-b2", "-p2", "-q1", "-r2,2", "-Y30", "-s40", "-m0.2"],
                                  ["-a3", "-b5", "-p5", "-q3", "-r4,5", "-Y70", "-s200", "-m0.5", "-W3000", "-G2000"]])
def test_cli_matches_live_reference_custom_scoring(tmp_path, opts):
    """Non-preset scoring / filtering options (minialign.c:5950-6030): score matrix, gap model, X-drop, min score / ratio,
    chaining windows; reference run with -t1."""
    from minialign_b200 import synth
    g = synth.make_genome(800_000, 3, seed=41, repeats=((20, 3000), (60, 800)))
    reads = synth.make_reads(g, 3_000_000, seed=42) + synth.make_hard_reads(g, seed=43)
    fa, rd, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    ref = subprocess.run([refh.BIN, "-xpacbio", *opts, "-t1", "-TAS,XS,NM,MD,SA", idx, rd], capture_output=True)
    assert ref.returncode == 0
    exp = [l for l in ref.stdout.decode().split("\n") if not l.startswith("@PG")]
    got = run_cli(["-xpacbio", *opts, "-TAS,XS,NM,MD,SA", idx, rd])
    assert len(got) == len(exp) and sum(1 for l in exp if l and not l.startswith("@")) > 50
    bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not bad, (len(bad), got[bad[0]][:200], exp[bad[0]][:200])


@pytest.mark.skipif(not os.path.exists(refh.BIN), reason="oracle/_ref/minialign not built")
def test_cli_matches_live_reference_multi_contig_repeats(tmp_path):
    """BASELINE config 2 in small (sacCer3-like: 17 contigs) with heavier planted repeat families, so that the rescue rounds
    (occ thresholds), secondary / supplementary records and the seed-rich sort class are exercised; reference run with -t1."""
    from minialign_b200 import synth
    g = synth.make_genome(6_000_000, 17, seed=31, repeats=((30, 4000), (120, 1200), (400, 300)), weights=[7, 2, 9, 1, 5, 3, 8, 2, 6, 4, 1, 9, 3, 5, 2, 7, 1])
    reads = synth.make_reads(g, 8_000_000, seed=32) + synth.make_hard_reads(g, seed=33)
    fa, rd, idx = str(tmp_path / "g.fa"), str(tmp_path / "r.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80); synth.write_fasta(rd, reads)
    subprocess.check_call([refh.BIN, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    ref = subprocess.run([refh.BIN, "-xpacbio", "-t1", "-TAS,XS,NM,MD,NH,IH", idx, rd], capture_output=True)
    assert ref.returncode == 0
    exp = [l for l in ref.stdout.decode().split("\n") if not l.startswith("@PG")]
    got = run_cli(["-xpacbio", "-TAS,XS,NM,MD,NH,IH", idx, rd])
    assert len(got) == len(exp)
    bad = [i for i, (a, b) in enumerate(zip(got, exp)) if a != b]
    assert not bad, (len(bad), got[bad[0]][:200], exp[bad[0]][:200])
    assert sum(1 for l in exp if l and not l.startswith("@") and int(l.split("\t")[1]) & 0x900) > 0     # secondary / supplementary present
    # the same with the FASTA reference on the command line: the index is built by mab_index.cpp instead of the reference
    got_fa = run_cli(["-xpacbio", "-TAS,XS,NM,MD,NH,IH", fa, rd])
    assert got_fa == exp
