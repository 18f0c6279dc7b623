"""BASELINE configs[4] in shape: a human-sized synthetic genome (24 contigs in hg38's proportions, planted repeat families), the
index built by the reference (`oracle/_ref/minialign -d`), PBSIM-CLR-like 20 kb reads, mapped file-to-SAM by `minialign-b200`.

    python scripts/gpu_human_scale.py [genome_gb=3.1] [coverage=0.3] [out_dir=gpurun_out/human]

Writes <out_dir>/summary.json: sizes, set-up and mapping times of both programs, per-kernel device times (ncu launch list with a
single-pass metric set: time, instructions, DRAM bytes, L2 hit rate, long-scoreboard share) and the parity of a read sample
against the reference CLI (-t1).  Nothing here is part of the product path.
"""
import csv
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from minialign_b200 import synth

GB = float(sys.argv[1]) if len(sys.argv) > 1 else 3.1
COV = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "human")
WORK = os.environ.get("MAB_BENCH_DIR", "/tmp/mab_human")
REF = os.path.join(ROOT, "oracle", "_ref", "minialign")
CLI = os.path.join(ROOT, "minialign_b200", "minialign-b200")
HG38 = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]    # Mb, chr1..22, X, Y
os.makedirs(OUT, exist_ok=True); os.makedirs(WORK, exist_ok=True)
S = {"genome_gb": GB, "coverage": COV, "host_cores": os.cpu_count()}


def log(*a):
    print(f"[{time.strftime('%H:%M:%S')}]", *a, file=sys.stderr, flush=True)


def save():
    json.dump(S, open(os.path.join(OUT, "summary.json"), "w"), indent=1)


# host memory guard: the raw index (~9 B per reference base) exists twice on the host while a context is set up
try:
    avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    S["host_mem_available_gb"] = avail / 1e9
    fit = avail / 1e9 / 30.0
    if fit < GB:
        log(f"only {avail / 1e9:.0f} GB of host memory: genome scaled down from {GB} to {fit:.2f} Gb")
        GB = max(0.2, fit); S["genome_gb"] = GB
except Exception:
    pass
t0 = time.time()
bp = int(GB * 1e9)
scale = GB / 3.1
g = synth.make_genome(bp, len(HG38), seed=7, weights=HG38,
                      repeats=((int(3000 * scale), 6000), (int(30000 * scale), 1500), (int(300000 * scale), 300)), divergence=0.08)
S["t_genome_s"] = time.time() - t0; log("genome", S["t_genome_s"])
fa, idx, rd = os.path.join(WORK, "g.fa"), os.path.join(WORK, "g.mai"), os.path.join(WORK, "r.fa")
t0 = time.time(); synth.write_fasta(fa, g, 80); S["t_write_fa_s"] = time.time() - t0
t0 = time.time()
reads = synth.make_reads(g, int(bp * COV), seed=8)
synth.write_fasta(rd, reads)
S["n_reads"], S["read_bases"] = len(reads), int(sum(r.size for _, r in reads)); S["t_reads_s"] = time.time() - t0; log("reads", S["n_reads"], S["t_reads_s"])
sample = reads[:2000]
srd = os.path.join(WORK, "sample.fa"); synth.write_fasta(srd, sample)
nrd = os.path.join(WORK, "chunk.fa"); synth.write_fasta(nrd, reads[:16384])       # one full-size chunk for the per-kernel picture
del g, reads
thr = max(1, min(os.cpu_count() or 1, 127))
t0 = time.time()
p = subprocess.run([REF, "-xpacbio", f"-t{thr}", "-d", idx, fa], capture_output=True, text=True)
S["t_ref_index_s"] = time.time() - t0; S["mai_bytes"] = os.path.getsize(idx) if os.path.exists(idx) else None; log("index", S["t_ref_index_s"], p.stderr[-200:])
save()
if p.returncode != 0:
    raise SystemExit("index build failed: " + p.stderr[-500:])

if os.environ.get("HUMAN_GPUS"):               # the whole box: one process, all GPUs (BASELINE configs[4] in shape: REP x 0.3x coverage ~ 3x)
    ng = int(os.environ["HUMAN_GPUS"]); REP = int(os.environ.get("HUMAN_REPEAT", "10"))
    glist = ",".join(str(i) for i in range(ng))
    args = [CLI, "-xpacbio", "-g" + glist, "-c" + os.environ.get("HUMAN_CTX", "2"), "-N" + os.environ.get("HUMAN_CHUNK_MB", "160"), idx]
    for tag, files, out in (("x%d" % REP, [rd] * REP, os.devnull), ("x1_file", [rd], os.path.join(WORK, "ours8.sam"))):
        t1 = time.time()
        with open(out, "wb") as f:
            pc = subprocess.run(args + files, stdout=f, stderr=subprocess.PIPE, text=True)
        S["ours_%dgpu_%s_wall_s" % (ng, tag)] = time.time() - t1; S["ours_%dgpu_%s_stderr" % (ng, tag)] = pc.stderr[-1200:]
        mm = re.search(r"mapped (\d+) reads / ([0-9.]+) Mbases in ([0-9.]+) sec \(([0-9.]+) Mbases/s\)", pc.stderr)
        if mm:
            S["ours_%dgpu_%s_map_s" % (ng, tag)], S["ours_%dgpu_%s_mbases_per_s" % (ng, tag)] = float(mm.group(3)), float(mm.group(4))
        log("ours", ng, tag, S["ours_%dgpu_%s_wall_s" % (ng, tag)], pc.stderr[-400:])
        save()
    t1 = time.time()
    with open(os.devnull, "wb") as f:
        pr = subprocess.run([REF, "-xpacbio", f"-t{thr}", idx] + [rd] * REP, stdout=f, stderr=subprocess.PIPE, text=True)
    S["ref_x%d_wall_s" % REP] = time.time() - t1; S["ref_x%d_stderr" % REP] = pr.stderr[-600:]; S["ref_threads"] = thr
    log("reference", S["ref_x%d_wall_s" % REP], pr.stderr[-300:])
    save()
    ns = int(os.environ.get("HUMAN_PARITY_READS", "1000"))
    synth.write_fasta(srd, sample[:ns])
    pr1 = subprocess.run([REF, "-xpacbio", "-t1", idx, srd], capture_output=True)
    exp = [l for l in pr1.stdout.split(b"\n") if l and not l.startswith(b"@")]
    got = []
    with open(os.path.join(WORK, "ours8.sam"), "rb") as f:
        for l in f:
            if not l.startswith(b"@"):
                got.append(l.rstrip(b"\n"))
                if len(got) >= len(exp):
                    break
    S["parity_%dgpu" % ng] = {"reads": ns, "sam_lines": len(exp), "identical": got == exp, "against": "oracle/_ref/minialign -xpacbio -t1 on the first reads of the file; ours: ONE SAM written by %d GPUs" % ng}
    log("parity", S["parity_%dgpu" % ng])
    save()
    raise SystemExit(0)

if os.environ.get("HUMAN_TRACE"):             # stage timeline of the CLI (MAB_TRACE lines) for the listed context counts, nothing else
    REP = int(os.environ.get("HUMAN_REPEAT", "3"))
    for nc in [int(x) for x in os.environ["HUMAN_TRACE"].split(",")]:
        for ctas in os.environ.get("HUMAN_CTAS", "4").split(","):
            for var in os.environ.get("HUMAN_ENVS", "").split(";"):           # e.g. "MAB_SORT_WALK=0;MAB_SORT_WALK=1"
                extra = dict(kv.split("=") for kv in var.split(",") if kv)
                with open(os.devnull, "wb") as f:
                    pc = subprocess.run([CLI, "-xpacbio", f"-c{nc}", idx] + [rd] * REP, stdout=f, stderr=subprocess.PIPE, text=True, env=dict(os.environ, MAB_TRACE="1", MAB_EXT_CTAS=ctas, **extra))
                tag = f"c{nc}_ctas{ctas}" + ("_" + var.replace("=", "").replace(",", "_") if var else "")
                open(os.path.join(OUT, f"trace_{tag}.log"), "w").write(pc.stderr)
                log("trace", tag, pc.stderr[-300:])
    raise SystemExit(0)

# ---- our CLI, file to SAM ----
sam = os.path.join(WORK, "ours.sam")
t0 = time.time()
with open(sam, "wb") as f:
    p = subprocess.run([CLI, "-xpacbio", "-c3", idx, rd], stdout=f, stderr=subprocess.PIPE, text=True)
S["ours_wall_s"] = time.time() - t0; S["ours_stderr"] = p.stderr[-1500:]; log("ours", S["ours_wall_s"], p.stderr[-600:])
# steady state: the same read file eight times over (the first chunk of every context pays for its buffers: a three-chunk job is all warm-up)
REP = int(os.environ.get("HUMAN_REPEAT", "8"))
t1 = time.time()
with open(os.devnull, "wb") as f:
    p8 = subprocess.run([CLI, "-xpacbio", "-c4", idx] + [rd] * REP, stdout=f, stderr=subprocess.PIPE, text=True)
S["ours_x%d_wall_s" % REP] = time.time() - t1; S["ours_x%d_stderr" % REP] = p8.stderr[-900:]
m8 = re.search(r"mapped (\d+) reads / ([0-9.]+) Mbases in ([0-9.]+) sec \(([0-9.]+) Mbases/s\)", p8.stderr)
if m8:
    S["ours_x%d_map_s" % REP], S["ours_x%d_mbases_per_s" % REP] = float(m8.group(3)), float(m8.group(4))
log("ours x%d" % REP, S.get("ours_x%d_mbases_per_s" % REP))
S["contexts_sweep_mbases_per_s"] = {}
for nc in [int(x) for x in os.environ.get("HUMAN_CTX_SWEEP", "").split(",") if x]:
    for ctas in ("4", "6"):
        with open(os.devnull, "wb") as f:
            pc = subprocess.run([CLI, "-xpacbio", f"-c{nc}", idx] + [rd] * REP, stdout=f, stderr=subprocess.PIPE, text=True, env=dict(os.environ, MAB_EXT_CTAS=ctas))
        mc = re.search(r"in ([0-9.]+) sec \(([0-9.]+) Mbases/s\)", pc.stderr)
        S["contexts_sweep_mbases_per_s"][f"c{nc}_ctas{ctas}"] = float(mc.group(2)) if mc else pc.stderr[-200:]
        log("sweep", nc, ctas, S["contexts_sweep_mbases_per_s"][f"c{nc}_ctas{ctas}"])
save()
m = re.search(r"mapped (\d+) reads / ([0-9.]+) Mbases in ([0-9.]+) sec \(([0-9.]+) Mbases/s\)", p.stderr)
if m:
    S["ours_map_s"], S["ours_mbases_per_s"] = float(m.group(3)), float(m.group(4))
m = re.search(r"index file ([0-9.]+) s, .* ([0-9.]+) s\)", p.stderr)
if m:
    S["ours_index_load_s"], S["ours_device_setup_s"] = float(m.group(1)), float(m.group(2))
S["ours_rc"] = p.returncode; S["sam_bytes"] = os.path.getsize(sam)
save()

# ---- reference CLI on all host cores, same files ----
t0 = time.time()
with open(os.devnull, "wb") as f:
    p = subprocess.run([REF, "-xpacbio", f"-t{thr}", idx, rd], stdout=f, stderr=subprocess.PIPE, text=True)
S["ref_wall_s"] = time.time() - t0
m1 = re.search(r"main_align::([0-9.]+)\*[0-9.]+\] loaded/built index", p.stderr); m2 = re.search(r"Real time: ([0-9.]+) sec", p.stderr)
if m1 and m2:
    S["ref_index_load_s"] = float(m1.group(1)); S["ref_map_s"] = float(m2.group(1)) - float(m1.group(1)); S["ref_mbases_per_s"] = S["read_bases"] / 1e6 / S["ref_map_s"]; S["ref_threads"] = thr
log("reference", S.get("ref_map_s"), S.get("ref_mbases_per_s")); save()

# ---- parity: the first 2000 reads, reference -t1 (its results do not depend on what follows a read) ----
t0 = time.time()
p = subprocess.run([REF, "-xpacbio", "-t1", idx, srd], capture_output=True)
exp = b"".join(l for l in p.stdout.split(b"\n") if l and not l.startswith(b"@"))
n_lines = sum(1 for l in p.stdout.split(b"\n") if l and not l.startswith(b"@"))
got_lines = []
with open(sam, "rb") as f:
    for l in f:
        if l.startswith(b"@"):
            continue
        got_lines.append(l.rstrip(b"\n"))
        if len(got_lines) >= n_lines:
            break
S["parity"] = {"reads": len(sample), "sam_lines": n_lines, "identical": b"".join(got_lines) == exp, "t_s": time.time() - t0,
               "against": "oracle/_ref/minialign -xpacbio -t1 on the first 2000 reads of the file"}
log("parity", S["parity"]); save()

# ---- per-kernel device picture on this index: one-pass metrics only (replaying kernels would mean saving tens of GB of device memory) ----
mets = "gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"
lst = os.path.join(OUT, "human_launches.csv")
t0 = time.time()
with open(os.devnull, "wb") as f:
    p = subprocess.run(["ncu", "--metrics", mets, "--clock-control", "none", "--cache-control", "none", "-c", "60", "--csv", "--log-file", lst, CLI, "-xpacbio", "-c1", idx, nrd],
                       stdout=f, stderr=subprocess.PIPE, text=True)
S["ncu_wall_s"] = time.time() - t0
try:
    rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
    hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        k = r[ki].split("(")[0]; a = agg.setdefault(k, {})
        a.setdefault(r[mi], []).append(float(r[vi].replace(",", "")))
    S["kernels_one_chunk_16384_reads"] = {k: {"launches": len(v.get("gpu__time_duration.sum", [])), "ms": sum(v.get("gpu__time_duration.sum", [])) / 1e6,
                                          "ginst": sum(v.get("smsp__inst_executed.sum", [])) / 1e9, "dram_read_gb": sum(v.get("dram__bytes_read.sum", [])) / 1e9,
                                          "l2_hit_pct": float(np.mean(v.get("lts__t_sector_hit_rate.pct", [0]))),
                                          "long_scoreboard_stall_ratio": float(np.mean(v.get("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", [0])))} for k, v in agg.items()}
except Exception as e:
    S["kernels_error"] = repr(e) + " " + p.stderr[-300:]
save()
log("done")
print(json.dumps(S)[:3000])
