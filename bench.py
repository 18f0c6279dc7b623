#!/usr/bin/env python
"""bench.py -- Mbases aligned / s of the read->reference mapping path (BASELINE.json metric), FASTA text to SAM text.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation (oracle/_ref/minialign)

Workload (config.workload): N = 1: BASELINE.json configs[1] -- E.coli-MG1655-sized reference (4.64 Mb, synthetic: no genomes or
PBSIM exist offline) x100 coverage of PBSIM-CLR-like reads (20k +- 2k, accuracy 0.88 +- 0.07), -xpacbio.
N > 1: configs[2] -- sacCer3-sized reference (12.1 Mb, 17 contigs of the yeast chromosomes' sizes), same read model, ONE read
set cut into chunks that are dealt round-robin to the ranks (weak scaling: one chunk per rank and step).
A step = one chunk of `--batch-reads` reads per rank through the whole path: FASTA text -> parse -> seed -> sort/chain -> extend ->
post-processing -> SAM text.  Chunks cycle through a few distinct ones, each far larger than L2 (config.l2).
value   : kernel-side throughput, the FASTA text already resident in HBM when the timed region starts, SAM text left in HBM
e2e     : the same through the public text API (mab_text_begin / commit / finish) with HOST buffers: FASTA bytes in page-locked
          host memory in, SAM bytes in page-locked host memory out, both copies and (N > 1) the per-wave NCCL exchange of the rlen
          chain and the output offsets inside the timed region
roofline: the dominant kernel (k_extend, round 0) timed alone with CUDA events on its own stream inside the library, against the
          integer ceiling of its own DP step (k_fill_peak) -- the bound SURVEY.md section 8d names; HBM figures as secondary keys
After the timed runs the SAM text of a full-size chunk is checked against the reference CLI (-t1) on a sample of its reads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # hardware queues: several contexts x (stream + side streams) + NCCL; at the default 8 they share queues and wait for each other's kernels
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

GENOMES = {
    "ecoli": dict(bp=4_640_000, contigs=1, weights=None,
                  name="ecoli-like 4.64 Mb synthetic reference x100 PBSIM-CLR-like reads (20k+-2k, acc 0.88+-0.07), -xpacbio (BASELINE configs[1])"),
    "saccer3": dict(bp=12_100_000, contigs=17, weights=[230, 813, 317, 1532, 577, 270, 1091, 563, 440, 746, 667, 1078, 924, 784, 1091, 948, 86],
                    name="sacCer3-like 12.1 Mb / 17 contigs synthetic reference x100 PBSIM-CLR-like reads (20k+-2k, acc 0.88+-0.07), -xpacbio, one read set sharded by chunk (BASELINE configs[2])"),
    # configs[3]: D.melanogaster dm6-sized (143 Mb, ~1.9 k contigs: 7 chromosome arms holding ~96 % + a long tail of scaffolds), -xont.1dsq
    "dm6": dict(bp=143_000_000, contigs=1870, weights=[23500, 25300, 28100, 32100, 23500, 1350, 3670] + [max(1.0, 60.0 * 0.9985 ** i) for i in range(1863)], preset="ont.1dsq",
                name="dm6-like 143 Mb / 1870 contigs synthetic reference x20 PBSIM-CLR-like reads (20k+-2k, acc 0.88+-0.07), -xont.1dsq, one read set sharded by chunk (BASELINE configs[3])"),
}
BYTES_PER_BASE = 160.0          # algorithmic HBM bytes per read base
TRAFFIC_PER_BASE = 293.0        # dram__bytes_read.sum + dram__bytes_write.sum of k_extend per read base, ncu --set full capture
                                # (profiles/r02_ncu_k_extend_full.json: 47.0 + 52.0 GB for the 337.5 Mbase launch)
NCU_PIPE_ALU_PCT = 62.0         # sm__inst_executed_pipe_alu of k_extend in the same capture (issue slots busy 75.1 %)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minialign")
OUR_BIN = os.path.join(ROOT, "minialign_b200", "minialign-b200")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_genome(work: str, which: str, seed: int = 1):
    """Synthetic genome + index (built by the reference, like BASELINE.md section 3 step 3; by the product's own host builder when
    the reference binary is absent)."""
    from minialign_b200 import mai, synth
    G = GENOMES[which]
    os.makedirs(work, exist_ok=True)
    g = synth.make_genome(G["bp"], G["contigs"], seed=seed, weights=G["weights"])
    fa, idx = os.path.join(work, f"{which}_like.fa"), os.path.join(work, f"{which}_like.mai")
    if not os.path.exists(idx):
        tmp = idx + f".tmp{os.getpid()}.mai"
        synth.write_fasta(fa + f".{os.getpid()}", g, 80)
        builder = REF_BIN if os.path.exists(REF_BIN) else OUR_BIN
        subprocess.check_call([builder, "-x" + G.get("preset", "pacbio"), "-d", tmp, fa + f".{os.getpid()}"], stderr=subprocess.DEVNULL)
        os.replace(tmp, idx)
    return g, idx, mai.load_mai(idx)


def chunk_reads(g, chunk_id: int, batch_reads: int):
    """reads of chunk `chunk_id` of the (virtual) read file: seeded by the chunk id, so every rank can produce its own share"""
    from minialign_b200 import synth
    return synth.make_reads(g, batch_reads * 20_600, seed=1000 + chunk_id)[:batch_reads]


def fasta_bytes(reads, chunk_id: int = 0) -> bytes:
    return b"".join(b">c%d_" % chunk_id + n.encode() + b"\n" + s.tobytes() + b"\n" for n, s in reads)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, through NVML (nvidia-smi as a fallback: spawning it every 200 ms
    perturbs the run it is supposed to observe)."""

    def __init__(self, dev: int):
        super().__init__(daemon=True)
        self.dev, self.stop_flag, self.rows = dev, False, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = dev
            if vis:
                tok = vis.split(",")[dev].strip()
                if tok.isdigit():
                    idx = int(tok)
                else:
                    uuid = tok
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


PRESET = "pacbio"           # set from the workload in main()


def ref_run(idx: str, fasta: str, threads: int, out=None, extra=()):
    """One run of the reference CLI; returns mapping seconds = final Real time - index-load timestamp."""
    with open(out or os.devnull, "wb") as sink:
        p = subprocess.run([REF_BIN, "-x" + PRESET, f"-t{threads}", *extra, idx, fasta], stdout=sink, stderr=subprocess.PIPE, text=True)
    m1 = re.search(r"main_align::([0-9.]+)\*[0-9.]+\] loaded/built index", p.stderr)
    m2 = re.search(r"Real time: ([0-9.]+) sec", p.stderr)
    if p.returncode != 0 or not m1 or not m2:
        raise RuntimeError("reference run failed: " + p.stderr[-400:])
    return float(m2.group(1)) - float(m1.group(1))


def host_threads():
    return max(1, min(os.cpu_count() or 1, 127))          # MAX_THREADS requires -t < 128 (minialign.c:23, 5971)


def ref_hot_path(idx: str, reads, threads: int):
    """The reference's mm_align_seq alone (no file I/O, no SAM text) over all host cores, through oracle/_ref/libref_harness.so
    (refh_align_many): the like-for-like number for the C-ABI record-level call."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refh
    from minialign_b200 import api, synth
    if not refh.available():
        return None
    h = refh.RefHarness(idx, args=("-x" + PRESET, f"-t{threads}"))
    L = h.lib
    if not hasattr(L, "refh_align_many"):
        h.close()
        return None
    L.refh_align_many.restype = C.c_double
    L.refh_align_many.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    block, ofs, lens = api.pack_reads([synth.encode_2bit(s) for _, s in reads])
    nm = C.c_uint64(0)
    best = None
    for _ in range(2):
        secs = L.refh_align_many(h.h, block.ctypes.data, ofs.ctypes.data_as(C.POINTER(C.c_uint64)), lens.ctypes.data_as(C.POINTER(C.c_uint32)), len(lens), threads, C.byref(nm))
        best = secs if best is None else min(best, secs)
    h.close()
    return {"value": float(lens.sum()) / 1e6 / best, "unit": "Mbases/s", "cores": threads, "mapped": int(nm.value),
            "what": "mm_align_seq over all cores through libref_harness.so: reads pre-parsed in memory, results freed, no SAM text, best of 2"}


def cpu_baseline(idx, reads, work, budget_core_s=20.0):
    from minialign_b200 import synth
    thr = host_threads()
    # ~10 Mbases/s/core: bound the sample to roughly `budget_core_s` core-seconds
    sample, bases = [], 0
    for r in reads:
        sample.append(r); bases += r[1].size
        if bases >= budget_core_s * 10e6:
            break
    fa = os.path.join(work, f"cpu_sample.{os.getpid()}.fa")
    synth.write_fasta(fa, sample)
    secs = min(ref_run(idx, fa, thr) for _ in range(2))
    os.remove(fa)
    out = {"value": bases / 1e6 / secs, "unit": "Mbases/s", "cores": thr, "kind": "reference",
           "sample": f"{len(sample)} reads / {bases / 1e6:.1f} Mbases of the same workload, oracle/_ref/minialign -x{PRESET} -t{thr} FASTA file -> SAM text, best of 2, index load excluded"}
    try:
        hp = ref_hot_path(idx, sample, thr)
        if hp:
            out["hot_path"] = hp
    except Exception as e:
        out["hot_path"] = {"value": None, "unavailable": str(e)}
    return out


class _Solo:
    """rank-local stand-in for the wave exchange (the untimed parity pass runs on rank 0 alone)"""

    def __init__(self):
        self.rlen, self.out_base, self.n_collectives, self.world, self.rank = 0, 0, 0, 1, 0

    def begin_wave(self, valid, value):
        self._last = (valid, value)

    def settle(self, commit):
        v = commit(self.rlen)
        if v[0]:
            self.rlen = v[1]

    def offsets(self, n):
        o = self.out_base
        self.out_base += n
        return o, n


def parity_check(idx, reads, chunk_id, work, text: bytes, n_sample=384):
    """Untimed: the SAM text of a full-size chunk must start with exactly the lines the reference CLI (-t1) prints for the first
    `n_sample` reads of that chunk (the reference's results do not depend on what follows a read)."""
    from minialign_b200 import synth
    fa, sam = os.path.join(work, f"parity.{os.getpid()}.fa"), os.path.join(work, f"parity.{os.getpid()}.sam")
    sample = reads[:n_sample]
    with open(fa, "wb") as f:
        f.write(fasta_bytes(sample, chunk_id))
    ref_run(idx, fa, 1, out=sam)
    exp = b"".join(l for l in open(sam, "rb") if not l.startswith(b"@"))
    os.remove(fa); os.remove(sam)
    ok = text[:len(exp)] == exp and len(exp) > 0
    return {"ok": bool(ok), "reads": len(sample), "bytes": len(exp), "against": "oracle/_ref/minialign -x" + PRESET + " -t1, every SAM line of the first reads of a timed full-size chunk"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-reads", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=4, help="mapper contexts (in-flight chunks) per GPU")
    ap.add_argument("--workload", default=None, choices=[None, "ecoli", "saccer3", "dm6"], help="default: ecoli at one GPU (configs[1]), saccer3 across GPUs (configs[2])")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    which = args.workload or ("ecoli" if max(world, args.gpus) == 1 else "saccer3")
    global PRESET
    PRESET = GENOMES[which].get("preset", "pacbio")
    config = {"workload": GENOMES[which]["name"], "batch_reads": args.batch_reads, "read_model": "len N(20000,2000) acc N(0.88,0.07) sub:ins:del 10:60:30",
              "parallelism": f"read-shard x{world}: chunk c of the read set on rank c mod {world}", "contexts_per_gpu": args.contexts,
              "path": "FASTA text -> device reader -> seed/chain/extend -> device post-processing -> device SAM printer -> SAM text",
              "setup": "one untimed chunk per context before the warm-up steps (buffer allocation)", "l2": "every step maps a different chunk; text + reads + DP state per chunk >> 126 MB L2"}
    work = os.environ.get("MAB_BENCH_DIR", "/tmp/mab_bench")

    if args.impl == "reference":
        if rank != 0:
            return
        from minialign_b200 import synth
        g, idx, blob = build_genome(work, which)
        reads = chunk_reads(g, 0, min(args.batch_reads, 16384))      # bounded sample per step: one chunk of our arm, ~338 Mbases, ~1.2 s on 16 cores
        fa = os.path.join(work, "ref_step.fa")
        synth.write_fasta(fa, reads)
        bases = sum(r[1].size for r in reads)
        thr = host_threads()
        for _ in range(args.warmup):
            ref_run(idx, fa, thr)
        t = [ref_run(idx, fa, thr) for _ in range(args.steps)]
        secs = sum(t)
        v = bases * args.steps / 1e6 / secs
        cb = {"value": v, "unit": "Mbases/s", "cores": thr, "kind": "reference",
              "sample": f"{len(reads)} reads / {bases / 1e6:.1f} Mbases per step, oracle/_ref/minialign -x{PRESET} -t{thr} FASTA file -> SAM text, index load excluded"}
        try:
            hp = ref_hot_path(idx, reads[:max(256, len(reads) // 4)], thr)
            if hp:
                cb["hot_path"] = hp
        except Exception as e:
            cb["hot_path"] = {"value": None, "unavailable": str(e)}
        print(json.dumps({"impl": "reference", "metric": "Mbases aligned/sec", "value": v, "unit": "Mbases/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int8", "data": "synthetic", "config": dict(config, batch_reads=len(reads)), "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "Mbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # native libraries (NCCL's version banner, ...) write to fd 1: keep the real stdout for the one JSON line, send the rest to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    from minialign_b200 import api, pipeline, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_binding = shard.bind_near_gpu(local) if world > 1 else "unchanged (one rank)"
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # rlen chain (host data, on the critical path); its own group: the two exchanges run on different threads
    n_distinct = min(3, args.warmup + args.steps)
    t0 = time.time()
    g, idx, blob = build_genome(work + f"/r{rank}", which)
    # the (virtual) read file: chunk c = reads seeded by c; this rank maps chunks rank, rank + world, ... (a few distinct ones, cycled)
    my_reads = [chunk_reads(g, d * world + rank, args.batch_reads) for d in range(n_distinct)]
    texts = [fasta_bytes(r, d * world + rank) for d, r in enumerate(my_reads)]
    bases_of = [int(sum(s.size for _, s in r)) for r in my_reads]
    pinned = []
    for t in texts:
        p = torch.empty(len(t) + 64, dtype=torch.uint8).pin_memory()
        p[:len(t)] = torch.frombuffer(bytearray(t), dtype=torch.uint8)
        p[len(t):] = 10
        pinned.append(p)
    log(f"[rank {rank}] workload ready in {time.time() - t0:.1f}s: {n_distinct} chunks x {args.batch_reads} reads, {len(texts[0]) / 1e6:.0f} MB of FASTA each")
    # four mapper contexts per GPU, each driven by its own host thread (pipeline.py): while one chunk is being parsed / copied / printed,
    # another one's extension runs.  The pipelined contexts launch 3 of the 6 possible k_extend CTAs per SM (the 128-register build):
    # at 4 the persistent kernel holds every register of the SM and the other chunks' parser / seeding / sort kernels wait for it to
    # end; at 3 they run underneath (value 3945 -> 3918, e2e 3546 -> 3762 Mbases/s); alone, 6 is faster
    # (profiles/r02_sweep_contexts.txt)
    ext_pipe = os.environ.get("MAB_EXT_CTAS", "3" if args.contexts > 1 else "6")
    os.environ["MAB_EXT_CTAS"] = ext_pipe
    m0 = api.Mapper(blob, PRESET, device=local)
    ms = [m0] + [m0.clone() for _ in range(max(1, args.contexts) - 1)]
    config["extend_ctas_per_sm"] = int(ext_pipe)
    config["cpu_binding"] = cpu_binding

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last_out = {}

    def run(mode_device: bool, steps: int, warmup: int):
        for m in ms:
            m.lib.mab_set_device_input(m.h, 1 if mode_device else 0)
        dev_text = [p.cuda(non_blocking=False) for p in pinned] if mode_device else None
        flags = api.TEXT_DEVICE_OUT if mode_device else 0

        def get_chunk(first):
            def f(w):
                d = (first + w) % n_distinct
                return pipeline.Chunk(dev_text[d].data_ptr() if mode_device else pinned[d].data_ptr(), len(texts[d]))
            return f

        def sink(first):
            def f(w, ptr, n, ofs):
                if not mode_device:
                    last_out[(first + w) % n_distinct] = (ptr, n)       # stays valid until the context's next finish: only read after the run
            return f

        def drive(first, count):
            ex = shard.WaveExchange(device=dev, host_group=host_group)
            pipe = pipeline.WavePipeline(ms, ex, flags, device_index=local)
            pipe.out, pipe.out_cap = run.out, run.out_cap                # page-locked output buffers live across drives
            tot = pipe.run(count, get_chunk(first), sink(first))
            run.out, run.out_cap = pipe.out, pipe.out_cap
            tot["collectives"] = ex.n_collectives
            tot["collective_ms"] = 1e3 * ex.t_collectives
            return tot

        drive(0, len(ms))                                   # set-up: one untimed chunk per context sizes its device / pinned buffers
        drive(len(ms), max(warmup, len(ms)))                # the W warm-up steps
        sampler = ClockSampler(local); sampler.start()
        # device-side timing: the events sit on torch's (idle) stream; the first is recorded after a full device sync, the second
        # after the next one, so the interval covers everything the contexts ran on their own streams in between
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        tot = drive(warmup, steps)
        torch.cuda.synchronize()
        ev1.record(); ev1.synchronize()
        secs = ev0.elapsed_time(ev1) / 1e3
        barrier()
        sampler.stop_flag = True; sampler.join(timeout=2)
        if world > 1:
            tt = torch.tensor([secs], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); secs = float(tt.item())
            bb = torch.tensor([tot["bases"]], dtype=torch.float64, device="cuda"); dist.all_reduce(bb); total_bases = float(bb.item())
        else:
            total_bases = float(tot["bases"])
        return secs, total_bases, tot, sampler.summary()
    run.out, run.out_cap = [None] * len(ms), [0] * len(ms)

    secs_dev, bases_dev, tot_dev, clocks = run(True, args.steps, args.warmup)
    secs_e2e, bases_e2e, tot_e2e, _ = run(False, args.steps, args.warmup)
    # untimed parity: (1) a full-size chunk mapped from a fresh reference-thread state (rlen = 0, like the reference CLI's first read)
    # must start with exactly the lines the reference prints for a sample of its reads; (2) the text the TIMED run produced for the same
    # chunk must equal it from the second read on (in the timed run the chunk's first read inherits the previous chunk's rlen word)
    parity = None
    if rank == 0:
        try:
            d = sorted(last_out)[0]
            timed = C.string_at(*last_out[d])
            for m in ms:
                m.lib.mab_set_device_input(m.h, 0)
            fresh = {}
            pipe = pipeline.WavePipeline(ms, shard.WaveExchange(device=dev) if world == 1 else _Solo(), 0, device_index=local)
            pipe.out, pipe.out_cap = run.out, run.out_cap
            pipe.run(1, lambda w: pipeline.Chunk(pinned[d].data_ptr(), len(texts[d])), lambda w, ptr, n, ofs: fresh.update(t=C.string_at(ptr, n)))
            run.out, run.out_cap = pipe.out, pipe.out_cap
            second = b"\nc%d_%s\t" % (d * world + rank, my_reads[d][1][0].encode())
            a, b = timed.find(second), fresh["t"].find(second)
            parity = {"timed_equals_fresh_from_second_read": bool(a > 0 and b > 0 and timed[a:] == fresh["t"][b:]), "sam_bytes": len(timed)}
            if os.path.exists(REF_BIN):
                parity.update(parity_check(idx, my_reads[d], d * world + rank, work, fresh["t"]))
            else:
                parity.update({"ok": None, "against": "oracle/_ref/minialign not present"})
        except Exception as e:
            parity = {"ok": False, "error": repr(e)}
        if parity.get("ok") is False or parity.get("timed_equals_fresh_from_second_read") is False:
            raise SystemExit(f"bench.py: SAM text of the timed path differs from the reference: {parity}")

    # roofline pass: the dominant kernel timed ALONE (one context, CUDA events on its stream inside the library); in the pipelined
    # runs above the kernels of three contexts overlap on the GPU, which stretches every per-kernel event interval
    def solo(n):
        os.environ["MAB_EXT_CTAS"] = "6"
        m = api.Mapper(blob, PRESET, device=local)
        os.environ["MAB_EXT_CTAS"] = ext_pipe
        a = dict(bases=0, ms_ext_r0=0.0, ms_ext=0.0, vec=0, ms_post=0.0, ms_total=0.0)
        m.map_text(pinned[0][:len(texts[0])].numpy().tobytes()[:1 << 22].rsplit(b"\n>", 1)[0] + b"\n")      # warm-up: small allocations
        peak = m.fill_peak(True, 4000)
        info = api.MabTextInfo()
        for i in range(n + 1):
            d = i % n_distinct
            rc = m.lib.mab_map_text(m.h, pinned[d].data_ptr(), len(texts[d]), 0, None, 0, None, C.byref(info))
            if rc != 0:
                raise RuntimeError("mab_map_text failed: " + m.lib.mab_last_error().decode())
            st = m.stats()
            if i == 0:
                continue                                    # allocation pass
            a["bases"] += bases_of[d]; a["ms_ext_r0"] += st["ms_extend_r0"]; a["ms_ext"] += st["ms_extend"]; a["vec"] += st["n_vectors"]; a["ms_post"] += st["ms_post"]; a["ms_total"] += st["ms_total"]
        m.close()
        return a, n, peak
    for m in ms[1:]:
        m.close()                                                   # free their arenas before the solo context allocates its own
    ms = ms[:1]
    agg_solo, n_solo, peak_vps = solo(2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k_s = agg_solo["ms_ext_r0"] / 1e3 / n_solo                             # average k_extend (round 0) launch duration, kernel alone
    alg_bytes = BYTES_PER_BASE * agg_solo["bases"] / n_solo                # algorithmic bytes one launch processes
    hbm_achieved = alg_bytes / k_s / 1e9 if k_s > 0 else 0.0
    gcups = 64.0 * agg_solo["vec"] / max(1e-9, agg_solo["ms_ext"] / 1e3) / 1e9
    peak_gcups = 64.0 * peak_vps / 1e9
    line = {
        "metric": "Mbases aligned/sec", "value": bases_dev / 1e6 / secs_dev, "unit": "Mbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": bases_e2e / 1e6 / secs_e2e, "unit": "Mbases/s", "h2d_bytes_per_step": tot_e2e["h2d"] // args.steps, "d2h_bytes_per_step": tot_e2e["d2h"] // args.steps,
                "what": "FASTA bytes in page-locked host memory -> mab_text_begin/commit/finish -> SAM bytes in page-locked host memory", "sam_bytes_per_step": tot_e2e["sam_bytes"] // args.steps,
                "collectives_in_timed_region": tot_e2e["collectives"], "collective_ms_per_step": tot_e2e["collective_ms"] / args.steps, "reads_remapped_for_rlen_chain": tot_e2e["redo"]},
        "gpu_launches": tot_dev["launches"],
        # the bound that limits the dominant kernel: issue slots of the integer pipes.  peak = cell updates/s of the DP step
        # alone (k_fill_peak: the product's own traced step code, register-resident, k_extend's launch shape), achieved = what k_extend
        # sustains including block bookkeeping, search, trace and the state machine
        "roofline": {"bound": "int-alu", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": gcups / peak_gcups if peak_gcups else None,
                     "traffic": TRAFFIC_PER_BASE * agg_solo["bases"] / n_solo, "kernel": "k_extend (round 0)", "ms_per_launch": 1e3 * k_s,
                     "peak_source": "k_fill_peak microbenchmark, same run (profiles/: SASS opcode histogram of the step and the ceiling derived from it)",
                     "ncu_pipe_alu_pct": NCU_PIPE_ALU_PCT, "traffic_source": "ncu --set full capture under profiles/: dram bytes per read base x bases per launch",
                     "timing": "k_extend timed alone (one context, all 6 CTAs per SM resident) after the pipelined runs, CUDA events on its stream",
                     "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak if hbm_peak else None,
                             "algorithmic_bytes_per_base": BYTES_PER_BASE, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s"},
                     "text_stages_ms_per_chunk": {"post+sam kernels": agg_solo["ms_post"] / n_solo, "whole chunk, one context": agg_solo["ms_total"] / n_solo}},
        "parity": parity,
    }
    if rank == 0:
        if not args.no_cpu_baseline and os.path.exists(REF_BIN):
            try:
                line["cpu_baseline"] = cpu_baseline(idx, my_reads[0], work)
            except Exception as e:   # the reference binary is test infrastructure: report its absence, never fake it
                line["cpu_baseline"] = {"value": None, "unit": "Mbases/s", "cores": host_threads(), "kind": "reference", "sample": f"unavailable: {e}"}
        emit(line)
    [m.close() for m in ms]
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
