"""File-to-SAM of the CLI against the reference CLI on the same files (BASELINE configs[1] shape: E.coli-sized genome, 20 kb reads).

    CLI_READS=65536 CLI_REPEAT=3 python scripts/gpu_cli_e2e.py

Both programs get the same read file CLI_REPEAT times on the command line (so that the job is long compared with CUDA start-up
and the first chunk's buffer allocation).  Reported: wall time of the whole process and the mapping-phase throughput each program
prints; SAM written to a regular file and to /dev/null; the first reads compared byte for byte with the reference run with -t1."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minialign_b200 import synth

work = os.environ.get("MAB_BENCH_DIR", "/tmp/mab_cli_e2e"); os.makedirs(work, exist_ok=True)
N = int(os.environ.get("CLI_READS", "65536")); R = int(os.environ.get("CLI_REPEAT", "3"))
PRESET = "-x" + os.environ.get("CLI_PRESET", "pacbio")
g = synth.make_genome(4_640_000, 1, seed=1)
fa, rd, idx, srd = f"{work}/g.fa", f"{work}/r.fa", f"{work}/g.mai", f"{work}/sample.fa"
synth.write_fasta(fa, g, 80)
reads = []
for c in range((N + 16383) // 16384):
    reads += synth.make_reads(g, 16384 * 20_600, seed=1000 + c)[:16384]
reads = reads[:N]
synth.write_fasta(rd, reads); synth.write_fasta(srd, reads[:4096])
bases = sum(r[1].size for r in reads) * R
REF, CLI = f"{ROOT}/oracle/_ref/minialign", f"{ROOT}/minialign_b200/minialign-b200"
subprocess.check_call([REF, PRESET, "-d", idx, fa], stderr=subprocess.DEVNULL)
thr = min(os.cpu_count(), 127)
print(f"# {N} reads x {R} = {bases / 1e6:.0f} Mbases, {os.cpu_count()} host cores", flush=True)
runs = [(f"reference -t{thr} -> file", [REF, PRESET, f"-t{thr}", idx] + [rd] * R, f"{work}/ref.sam"),
        ("minialign-b200 -c4 -> file", [CLI, PRESET, "-c4", idx] + [rd] * R, f"{work}/ours.sam"),
        ("minialign-b200 -c4 -> /dev/null", [CLI, PRESET, "-c4", idx] + [rd] * R, os.devnull),
        ("minialign-b200 -c4 -> file (2nd run)", [CLI, PRESET, "-c4", idx] + [rd] * R, f"{work}/ours.sam")]
for name, cmd, out in runs:
    t = time.time()
    with open(out, "wb") as f:
        p = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, text=True)
    dt = time.time() - t
    tail = [l for l in p.stderr.split("\n") if "mapped" in l or "Real time" in l or "loaded" in l or "pipeline" in l]
    print(f"{name}: rc {p.returncode} wall {dt:.2f} s -> {bases / 1e6 / dt:.0f} Mbases/s file-to-SAM;", " | ".join(x.strip() for x in tail), flush=True)
p = subprocess.run([REF, PRESET, "-t1", idx, srd], capture_output=True)
exp = [l for l in p.stdout.split(b"\n") if l and not l.startswith(b"@")]
got = []
with open(f"{work}/ours.sam", "rb") as f:
    for l in f:
        if not l.startswith(b"@"):
            got.append(l.rstrip(b"\n"))
            if len(got) >= len(exp):
                break
print(PRESET, "first 4096 reads vs reference -t1: SAM identical:", got == exp, len(exp), "lines;",
      "sizes: ours", os.path.getsize(f"{work}/ours.sam"), "reference", os.path.getsize(f"{work}/ref.sam"))
