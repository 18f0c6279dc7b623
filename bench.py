#!/usr/bin/env python
"""bench.py -- Mbases aligned / s of the read->reference mapping hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation (oracle/_ref/minialign)

Workload (config.workload): BASELINE.json configs[1] -- E.coli-MG1655-sized reference (4.64 Mb, synthetic: no genomes or
PBSIM exist offline) x100 coverage of PBSIM-CLR-like reads (20k +- 2k, accuracy 0.88 +- 0.07), -xpacbio.
A step = one pass of the hot path (seed -> sort/chain -> extend -> results) over one batch of `--batch-reads` reads; batches
cycle through the read set, every batch's reads + DP state are far larger than L2 (config.l2 says so).
value   : kernel-side throughput, reads already resident in HBM when the timed region starts (device-input mode of the C ABI)
e2e     : the same through mab_map_batch with HOST (pinned) buffers, H2D of the reads and D2H of the results inside the timing
roofline: the dominant kernel (k_extend, round 0) timed with CUDA events on its own stream inside the library
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

GENOME_BP = 4_640_000
BYTES_PER_BASE = 160.0          # algorithmic HBM bytes per read base
TRAFFIC_PER_BASE = 292.0        # dram__bytes_read.sum + dram__bytes_write.sum of k_extend per read base, ncu --set full capture
                                # in profiles/r01_ncu_k_extend_full.json (49.3 GB for the 169 Mbase launch)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minialign")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_workload(work: str, n_batches: int, batch_reads: int, rank: int = 0, seed: int = 1):
    """Synthetic genome + index (built by the reference, like BASELINE.md section 3 step 3) + read batches."""
    from minialign_b200 import mai, synth
    os.makedirs(work, exist_ok=True)
    g = synth.make_genome(GENOME_BP, 1, seed=seed)
    fa, idx = os.path.join(work, "ecoli_like.fa"), os.path.join(work, "ecoli_like.mai")
    if not os.path.exists(idx):
        tmp = idx + f".tmp{os.getpid()}.mai"
        synth.write_fasta(fa + f".{os.getpid()}", g, 80)
        subprocess.check_call([REF_BIN, "-xpacbio", "-d", tmp, fa + f".{os.getpid()}"], stderr=subprocess.DEVNULL)
        os.replace(tmp, idx)
    blob = mai.load_mai(idx)
    batches = []
    for b in range(n_batches):
        reads = synth.make_reads(g, batch_reads * 20_600, seed=1000 + b)[:batch_reads]
        batches.append(reads)
    return g, idx, blob, batches


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


def ref_run(idx: str, fasta: str, threads: int):
    """One run of the reference CLI; returns mapping seconds = final Real time - index-load timestamp."""
    with open(os.devnull, "wb") as null:
        p = subprocess.run([REF_BIN, "-xpacbio", f"-t{threads}", idx, fasta], stdout=null, stderr=subprocess.PIPE, text=True)
    m1 = re.search(r"main_align::([0-9.]+)\*[0-9.]+\] loaded/built index", p.stderr)
    m2 = re.search(r"Real time: ([0-9.]+) sec", p.stderr)
    if p.returncode != 0 or not m1 or not m2:
        raise RuntimeError("reference run failed: " + p.stderr[-400:])
    return float(m2.group(1)) - float(m1.group(1))


def host_threads():
    return max(1, min(os.cpu_count() or 1, 127))          # MAX_THREADS requires -t < 128 (minialign.c:23, 5971)


def cpu_baseline(idx, batches, work, budget_core_s=20.0):
    from minialign_b200 import synth
    thr = host_threads()
    # ~10 Mbases/s/core: bound the sample to roughly `budget_core_s` core-seconds
    reads, bases = [], 0
    for b in batches:
        for r in b:
            reads.append(r); bases += r[1].size
            if bases >= budget_core_s * 10e6:
                break
        if bases >= budget_core_s * 10e6:
            break
    fa = os.path.join(work, f"cpu_sample.{os.getpid()}.fa")
    synth.write_fasta(fa, reads)
    secs = min(ref_run(idx, fa, thr) for _ in range(2))
    os.remove(fa)
    return {"value": bases / 1e6 / secs, "unit": "Mbases/s", "cores": thr, "kind": "reference",
            "sample": f"{len(reads)} reads / {bases / 1e6:.1f} Mbases of the same workload, oracle/_ref/minialign -xpacbio -t{thr}, best of 2, index load excluded"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-reads", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=3, help="mapper contexts (in-flight batches) per GPU")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": "ecoli-like 4.64 Mb synthetic reference x100 PBSIM-CLR-like reads (20k+-2k, acc 0.88+-0.07), -xpacbio (BASELINE configs[1])",
              "batch_reads": args.batch_reads, "read_model": "len N(20000,2000) acc N(0.88,0.07) sub:ins:del 10:60:30",
              "parallelism": f"read-shard x{world}", "contexts_per_gpu": args.contexts, "setup": "one untimed allocation batch per context before the warm-up steps", "l2": "every step maps a different batch; reads + DP state per batch >> 126 MB L2"}
    work = os.environ.get("MAB_BENCH_DIR", "/tmp/mab_bench")

    if args.impl == "reference":
        if rank != 0:
            return
        n_b = 1
        g, idx, blob, batches = build_workload(work, n_b, min(args.batch_reads, 16384))   # bounded sample per step: one batch of our arm, ~338 Mbases, ~1.2 s on 16 cores
        from minialign_b200 import synth
        fa = os.path.join(work, "ref_step.fa")
        synth.write_fasta(fa, batches[0])
        bases = sum(r[1].size for r in batches[0])
        thr = host_threads()
        for _ in range(args.warmup):
            ref_run(idx, fa, thr)
        t = [ref_run(idx, fa, thr) for _ in range(args.steps)]
        secs = sum(t)
        v = bases * args.steps / 1e6 / secs
        print(json.dumps({"impl": "reference", "metric": "Mbases aligned/sec", "value": v, "unit": "Mbases/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "int8", "data": "synthetic", "config": dict(config, batch_reads=len(batches[0])),
                          "cpu_baseline": {"value": v, "unit": "Mbases/s", "cores": thr, "kind": "reference",
                                           "sample": f"{len(batches[0])} reads / {bases / 1e6:.1f} Mbases per step, oracle/_ref/minialign -xpacbio -t{thr}, index load excluded"},
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
    from minialign_b200 import api, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    os.environ.setdefault("MAB_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // (world * max(1, args.contexts)))))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_b = min(3, args.warmup + args.steps)
    t0 = time.time()
    g, idx, blob, batches = build_workload(work + f"/r{rank}", n_b, args.batch_reads, rank, seed=1)
    # weak scaling: every rank maps its own batches (different read seeds per rank)
    if world > 1:
        from minialign_b200 import synth
        batches = [synth.make_reads(g, args.batch_reads * 20_600, seed=1000 + 100 * rank + b)[:args.batch_reads] for b in range(n_b)]
    from minialign_b200 import synth
    packed = []
    for b in batches:
        block, ofs, lens = api.pack_reads([synth.encode_2bit(r) for _, r in b])
        pinned = torch.from_numpy(block).pin_memory()
        packed.append((pinned, ofs, lens, int(lens.sum())))
    log(f"[rank {rank}] workload ready in {time.time() - t0:.1f}s: {n_b} batches x {args.batch_reads} reads")
    # two mapper contexts per GPU, each driven by its own host thread: while one batch sits in D2H / host post-processing
    # (MAPQ etc., minialign.c:4175-4396) the other one's kernels run -- the reference's source/worker/drain pipeline in two stages
    # the pipelined contexts launch 5 of the 6 possible k_extend CTAs per SM: the registers / shared memory left over let the next
    # batch's scan and sort/chain kernels run under the current batch's extend (+3 % end to end); alone, 6 is faster
    ext_pipe = os.environ.get("MAB_EXT_CTAS", "5" if args.contexts > 1 else "6")
    os.environ["MAB_EXT_CTAS"] = ext_pipe
    ms = [api.Mapper(blob, "pacbio", device=local) for _ in range(max(1, args.contexts))]
    config["extend_ctas_per_sm"] = int(ext_pipe)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(mode_device: bool, steps: int, warmup: int):
        for m in ms:
            m.lib.mab_set_device_input(m.h, 1 if mode_device else 0)
        dev_blocks = [p[0].cuda(non_blocking=False) for p in packed] if mode_device else None
        agg = dict(bases=0, launches=0, h2d=0, d2h=0, ms_ext=0.0, ms_ext_r0=0.0, ms_dev=0.0, vec=0, out_words=0)
        lock = threading.Lock()

        def one(m, i, timed):
            p = packed[i % n_b]
            tw = time.perf_counter()
            m.map_packed(dev_blocks[i % n_b].data_ptr() if mode_device else p[0].data_ptr(), p[0].numel(), p[1], p[2])
            st = m.stats()
            st["py_call_ms"] = 1e3 * (time.perf_counter() - tw)
            words = sum(int(m.lib.mab_result(m.h, j, None)) for j in range(0, len(p[2]), max(1, len(p[2]) // 64)))
            m.lib.mab_release_batch(m.h)
            if timed:
                if os.environ.get("MAB_BENCH_VERBOSE"):
                    log(f"[rank {rank}] step {i} device={mode_device} " + " ".join(f"{k}={v:.2f}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()))
                with lock:
                    agg["bases"] += p[3]; agg["launches"] += st["n_launches"]; agg["h2d"] += st["h2d_bytes"]; agg["d2h"] += st["d2h_bytes"]
                    agg["ms_ext"] += st["ms_extend"]; agg["ms_ext_r0"] += st["ms_extend_r0"]; agg["ms_dev"] += st["ms_total"]; agg["vec"] += st["n_vectors"]
                    agg["out_words"] += words

        def drive(first, count, timed, static=False):
            errs = []
            nxt = iter(range(first, first + count))     # steps are handed out as contexts become free

            def worker(t):
                try:
                    torch.cuda.set_device(local)
                    if static:
                        for i in range(first + t, first + count, len(ms)):
                            one(ms[t], i, timed)
                        return
                    while True:
                        with lock:
                            i = next(nxt, None)
                        if i is None:
                            break
                        one(ms[t], i, timed)
                except Exception as e:       # a failed batch must fail the run, not shorten it
                    errs.append(e)
            th = [threading.Thread(target=worker, args=(t,)) for t in range(len(ms))]
            [x.start() for x in th]
            [x.join() for x in th]
            if errs:
                raise errs[0]

        drive(0, len(ms), False, static=True)          # set-up: one untimed batch per context sizes its device / pinned buffers
        drive(len(ms), max(warmup, len(ms)), False)    # the W warm-up steps
        sampler = ClockSampler(local); sampler.start()
        # device-side timing: the events sit on torch's (idle) stream; the first is recorded after a full device sync, the second
        # after the next one, so the interval covers everything the contexts ran on their own streams in between
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        drive(warmup, steps, True)
        torch.cuda.synchronize()
        ev1.record(); ev1.synchronize()
        secs = ev0.elapsed_time(ev1) / 1e3
        # the one exchange step of the sharded path: output offsets of this wave (8 B per rank)
        shard.output_offsets(4 * agg["out_words"], device=torch.device("cuda", local))
        barrier()
        sampler.stop_flag = True; sampler.join(timeout=2)
        if world > 1:
            tt = torch.tensor([secs], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); secs = float(tt.item())
            bb = torch.tensor([agg["bases"]], dtype=torch.float64, device="cuda"); dist.all_reduce(bb); total_bases = float(bb.item())
        else:
            total_bases = float(agg["bases"])
        return secs, total_bases, agg, sampler.summary()

    # (the integer roofline of the DP step -- the product's own step code on register-resident state, k_fill_peak, traced variant --
    # is measured in the solo pass after the timed runs)
    secs_dev, bases_dev, agg_dev, clocks = run(True, args.steps, args.warmup)
    secs_e2e, bases_e2e, agg_e2e, _ = run(False, args.steps, args.warmup)

    # roofline pass: the dominant kernel timed ALONE (one context, CUDA events on its stream inside the library); in the pipelined
    # runs above the kernels of two contexts overlap on the GPU, which stretches every per-kernel event interval
    def solo(n):
        os.environ["MAB_EXT_CTAS"] = "6"
        m = api.Mapper(blob, "pacbio", device=local)
        os.environ["MAB_EXT_CTAS"] = ext_pipe
        m.lib.mab_set_device_input(m.h, 1)
        p = packed[0]; d = p[0].cuda()
        m.map_packed(d.data_ptr(), p[0].numel(), p[1], p[2]); m.lib.mab_release_batch(m.h)        # warm-up: allocations
        peak = m.fill_peak(True, 4000)
        a = dict(bases=0, ms_ext_r0=0.0, ms_ext=0.0, vec=0)
        for i in range(n):
            p = packed[i % n_b]
            d = p[0].cuda()
            m.map_packed(d.data_ptr(), p[0].numel(), p[1], p[2])
            st = m.stats(); m.lib.mab_release_batch(m.h)
            a["bases"] += p[3]; a["ms_ext_r0"] += st["ms_extend_r0"]; a["ms_ext"] += st["ms_extend"]; a["vec"] += st["n_vectors"]
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
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_s = agg_solo["ms_ext_r0"] / 1e3 / n_solo                             # average k_extend (round 0) launch duration, kernel alone
    alg_bytes = BYTES_PER_BASE * agg_solo["bases"] / n_solo                # algorithmic bytes one launch processes
    achieved = alg_bytes / k_s / 1e9 if k_s > 0 else 0.0
    line = {
        "metric": "Mbases aligned/sec", "value": bases_dev / 1e6 / secs_dev, "unit": "Mbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": bases_e2e / 1e6 / secs_e2e, "unit": "Mbases/s", "h2d_bytes_per_step": agg_e2e["h2d"] // args.steps, "d2h_bytes_per_step": agg_e2e["d2h"] // args.steps},
        "gpu_launches": agg_dev["launches"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": TRAFFIC_PER_BASE * agg_solo["bases"] / n_solo,
                     "traffic_source": "profiles/r01_ncu_k_extend_full.json: 292 B per read base (mask stream widened to 1 B per cell, DESIGN.md section 5)",
                     "kernel": "k_extend (round 0)", "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "note": "algorithmic bytes = 160 B/read base (SURVEY 8d); the kernel is integer-issue bound, not HBM bound: see DESIGN.md section 5",
                     "ms_per_launch": 1e3 * k_s, "gcups": 64.0 * agg_solo["vec"] / max(1e-9, agg_solo["ms_ext"] / 1e3) / 1e9,
                     "timing": "k_extend timed alone (one context, all 6 CTAs per SM resident) after the pipelined runs, CUDA events on its stream",
                     # the bound that actually limits k_extend: issue slots of the integer pipes.  peak = vectors/s of the DP step alone
                     # (k_fill_peak, same code, no memory), achieved = vectors/s k_extend sustains including search, trace and bookkeeping
                     "integer": {"achieved_gcups": 64.0 * agg_solo["vec"] / max(1e-9, agg_solo["ms_ext"] / 1e3) / 1e9, "peak_gcups": 64.0 * peak_vps / 1e9,
                                 "frac": (agg_solo["vec"] / max(1e-9, agg_solo["ms_ext"] / 1e3)) / peak_vps if peak_vps else None,
                                 "peak_source": "k_fill_peak microbenchmark (traced DP step, register-resident, k_extend launch shape), same run"}},
    }
    if rank == 0:
        if not args.no_cpu_baseline and os.path.exists(REF_BIN):
            try:
                line["cpu_baseline"] = cpu_baseline(idx, batches, work)
            except Exception as e:   # the reference binary is test infrastructure: report its absence, never fake it
                line["cpu_baseline"] = {"value": None, "unit": "Mbases/s", "cores": host_threads(), "kind": "reference", "sample": f"unavailable: {e}"}
        emit(line)
    [m.close() for m in ms]
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
