"""File-to-SAM wall time of the CLI against the reference CLI on the same files (E.coli-like genome, 8192 reads ~ 169 Mbases)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minialign_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = "/tmp/mab_cli_e2e"; os.makedirs(work, exist_ok=True)
g = synth.make_genome(4_640_000, 1, seed=1)
fa, rd, idx = f"{work}/g.fa", f"{work}/r.fa", f"{work}/g.mai"
synth.write_fasta(fa, g, 80)
N = int(os.environ.get("CLI_READS", "8192"))
PRESET = "-x" + os.environ.get("CLI_PRESET", "pacbio")
reads = synth.make_reads(g, N * 20_600, seed=1000)[:N]
synth.write_fasta(rd, reads)
bases = sum(r[1].size for r in reads)
REF, CLI = f"{ROOT}/oracle/_ref/minialign", f"{ROOT}/minialign_b200/minialign-b200"
subprocess.check_call([REF, PRESET, "-d", idx, fa], stderr=subprocess.DEVNULL)
for name, cmd in (("reference -t%d" % os.cpu_count(), [REF, PRESET, "-t%d" % min(os.cpu_count(), 127), idx, rd]), ("minialign-b200", [CLI, PRESET, idx, rd]), ("minialign-b200 (2nd run)", [CLI, PRESET, idx, rd])):
    t = time.time()
    p = subprocess.run(cmd, stdout=open(f"{work}/{name.split()[0]}.sam", "wb"), stderr=subprocess.PIPE, text=True)
    dt = time.time() - t
    tail = [l for l in p.stderr.split("\n") if "mapped" in l or "Real time" in l or "loaded" in l or "pipeline" in l]
    print(f"{name}: wall {dt:.2f} s -> {bases / 1e6 / dt:.0f} Mbases/s file-to-SAM;", " | ".join(tail), flush=True)
a = [l for l in open(f"{work}/reference.sam") if not l.startswith("@PG")]
b = [l for l in open(f"{work}/minialign-b200.sam") if not l.startswith("@PG")]
print(PRESET, N, "reads; SAM identical:", a == b, len(a), "lines")
