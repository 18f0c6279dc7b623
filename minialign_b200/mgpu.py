"""minialign-b200 across the GPUs of one box, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m minialign_b200.mgpu -xpacbio [-T tags] [-Q] [-c K] [-N MiB] -o out.sam ref.mai reads.fa [reads2.fa ...]

ONE read file (set) is cut into chunks at record boundaries; chunk c goes to rank c mod N (shard.py).  Every rank holds its own
copy of the index on its GPU, reads its chunks itself (pread into page-locked memory), maps them through the text path of the C
ABI on K contexts (pipeline.py) and pwrite()s its SAM text at the offset the per-wave exchange gives it.  The result is ONE SAM
file, byte-identical to `minialign -t1` (and to a one-GPU run): the only things that cross ranks are the reference thread's `rlen`
word and the byte counts, a few int64 per rank and wave over NCCL (gloo with --backend gloo, used by the CPU tests through the
emulation library).

This is the multi-GPU form of the reference's source / worker / drain pipeline (minialign.c:4565-4643): the strictly ordered drain
(4633-4643) becomes the exclusive prefix sum of byte counts.
"""
from __future__ import annotations

import ctypes as C
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA starts: see DESIGN.md section 6
import sys
import time

import numpy as np

from . import api, mai, pipeline, shard


def parse_args(argv):
    """The subset of minialign's options that select the mapping (presets of minialign.c:5853-5878) + the launcher's own."""
    o = dict(prm=dict(wlen=7000, glen=7000, min_score=50, min_ratio=0.3, gi=1, ge=1, gfa=0, gfb=0, xdrop=50, score_matrix=[1 if i % 5 == 0 else -1 for i in range(16)]),
             tags=0, keep_qual=False, contexts=3, chunk_mb=320.0, out=None, backend=None, lib=None, pos=[], verbose=False)
    p = o["prm"]

    def set_m(m):
        p["score_matrix"] = [m if i % 5 == 0 else p["score_matrix"][i] for i in range(16)]

    def set_x(x):
        p["score_matrix"] = [p["score_matrix"][i] if i % 5 == 0 else -x for i in range(16)]

    def line(s):
        for tok in s.split():
            opt(tok[1], tok[2:])

    def preset(a):
        tok = a.replace(":", ".").split(".")
        if tok[0] == "pacbio":
            line("-a2 -b4 -p4 -q2 -r3,3 -Y50 -s50 -m0.3")
            if len(tok) > 1 and tok[1] == "ccs":
                line("-b5 -p6 -p2")
        elif tok[0] == "ont":
            line("-a3 -b5 -p6 -q2 -r3,3 -Y50 -s50 -m0.3")
            for i, t in enumerate(tok[1:], 1):
                if t == "r7":
                    line("-b4")
                elif t in ("4", "5"):
                    line("-a2")
                elif t in ("1dsq", "2d"):
                    if i == 1:
                        line("-a2")
                    if not (i >= 2 and tok[1] == "r7"):
                        line("-b6 -r4,4")
                elif t == "1d" and i == 1:
                    line("-a2")
        else:
            raise SystemExit(f"[E::main] unknown preset `{a}'.")

    def opt(c, a):
        if c == "x":
            preset(a)
        elif c == "a":
            set_m(int(a))
        elif c == "b":
            set_x(int(a))
        elif c == "p":
            p["gi"] = int(a)
        elif c == "q":
            p["ge"] = int(a)
        elif c == "r":
            v = a.split(",")
            p["gfa"], p["gfb"] = int(v[0]), int(v[-1])
        elif c == "Y":
            p["xdrop"] = int(a)
        elif c == "s":
            p["min_score"] = int(a)
        elif c == "m":
            p["min_ratio"] = float(a)
        elif c == "W":
            p["wlen"] = int(a)
        elif c == "G":
            p["glen"] = int(a)
        elif c == "T":
            o["tags"] |= api.parse_tags(a)
        elif c == "c":
            o["contexts"] = max(1, int(a))
        elif c == "N":
            o["chunk_mb"] = max(0.001, float(a))
        elif c == "o":
            o["out"] = a
        elif c in "tkwBf":
            pass                                    # host threads / index-time parameters: the .mai carries its own
        else:
            raise SystemExit(f"[E::main] unknown or unsupported option `-{c}'.")

    i = 0
    while i < len(argv):
        a = argv[i]
        if a.startswith("--backend"):
            o["backend"] = a.split("=", 1)[1] if "=" in a else argv[(i := i + 1)]
        elif a.startswith("--lib"):
            o["lib"] = a.split("=", 1)[1] if "=" in a else argv[(i := i + 1)]
        elif a == "-Q":
            o["keep_qual"] = True
        elif a == "-v":
            o["verbose"] = True
        elif a.startswith("-") and len(a) > 1:
            arg = a[2:] if len(a) > 2 else argv[(i := i + 1)]
            opt(a[1], arg)
        else:
            o["pos"].append(a)
        i += 1
    if len(o["pos"]) < 2 or not o["out"]:
        raise SystemExit("usage: torchrun ... -m minialign_b200.mgpu [-x preset] [-T tags] [-Q] [-c contexts] [-N chunk MiB] -o out.sam ref.mai reads.fa [...]")
    return o


def record_start(fd: int, size: int, pos: int) -> int:
    """First record header at or after `pos` (0 stays 0, >= size gives size): every rank evaluates the same function, so all agree
    on the chunk boundaries without talking.  FASTA: '>' at a line start.  FASTQ: '@' at a line start whose line after next
    starts with '+' (a quality line may start with '@' too, but then the line after next is a sequence)."""
    if pos <= 0:
        return 0
    if pos >= size:
        return size
    delim = os.pread(fd, 1, 0)
    win = 1 << 20
    start = pos - 1                                                # the newline in front of a header sitting exactly at `pos`
    buf = b""
    while True:
        more = os.pread(fd, win, start + len(buf))
        buf += more
        p = 0
        while True:
            p = buf.find(b"\n" + delim, p)
            if p < 0:
                break
            if delim != b"@":
                return start + p + 1
            e1 = buf.find(b"\n", p + 1)
            e2 = buf.find(b"\n", e1 + 1) if e1 >= 0 else -1
            if e2 >= 0 and e2 + 1 < len(buf):
                if buf[e2 + 1:e2 + 2] == b"+":
                    return start + p + 1
                p += 1
                continue
            break                                                  # need more text to decide
        if not more:
            return size
        win *= 2


def main(argv=None):
    import torch
    import torch.distributed as dist
    o = parse_args(sys.argv[1:] if argv is None else argv)
    t0 = time.time()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    backend = o["backend"] or ("nccl" if torch.cuda.is_available() else "gloo")
    dev = None
    if backend == "nccl":
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if world > 1:
            shard.bind_near_gpu(local)
    if world > 1:
        dist.init_process_group(backend, **({"device_id": dev} if dev is not None else {}))
    if o["contexts"] > 1:
        os.environ.setdefault("MAB_EXT_CTAS", "3")     # contexts that run side by side leave registers for each other's small kernels
    m0 = api.Mapper.from_mai(o["pos"][0], o["prm"], device=local if backend == "nccl" else 0, lib_path=o["lib"])    # upload under the inflation
    ms = [m0] + [m0.clone() for _ in range(o["contexts"] - 1)]
    cmdline = "minialign-b200 " + " ".join(sys.argv[1:] if argv is None else argv)
    header = m0.sam_header(cmdline)
    flags = o["tags"] | (api.TEXT_KEEP_QUAL if o["keep_qual"] else 0)
    out_fd = None
    if rank == 0:
        out_fd = os.open(o["out"], os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        os.pwrite(out_fd, header, 0)
    if world > 1:
        dist.barrier()
    if out_fd is None:
        out_fd = os.open(o["out"], os.O_WRONLY)
    host_group = dist.new_group(backend="gloo") if world > 1 else None      # its own group: the two exchanges run on different threads
    ex = shard.WaveExchange(device=dev, host_group=host_group)
    ex.out_base = len(header)
    chunk_bytes = max(1024, int(o["chunk_mb"] * 1048576))
    m0.text_reserve(min(chunk_bytes + (16 << 20), 0xfffffff0 - 1))       # every context sizes its buffers for a full chunk at once
    tot = dict(reads=0, bases=0)
    for path in o["pos"][1:]:
        with open(path, "rb") as f:
            if f.read(2) == b"\x1f\x8b":
                raise SystemExit("[E::main_align] mgpu reads its chunks with pread: gzip input is not seekable, inflate it first (the single-process CLI takes .gz).")
        fd = os.open(path, os.O_RDONLY)
        size = os.fstat(fd).st_size
        n_chunks = (size + chunk_bytes - 1) // chunk_bytes
        n_waves = (n_chunks + world - 1) // world
        inbuf = [None] * len(ms)                                    # page-locked input buffer per context

        def get_chunk(w, fd=fd, size=size, n_chunks=n_chunks):
            c = w * world + rank
            if c >= n_chunks:
                return None
            a, b = record_start(fd, size, c * chunk_bytes), record_start(fd, size, (c + 1) * chunk_bytes)
            n = b - a
            if n <= 0:
                return None
            k = w % len(ms)
            if inbuf[k] is None or inbuf[k][1] < n:
                if inbuf[k] is not None:
                    m0.lib.mab_host_free(inbuf[k][0])
                cap = n + n // 8 + (1 << 20)
                inbuf[k] = (m0.lib.mab_host_alloc(cap), cap)
            view = memoryview((C.c_char * n).from_address(inbuf[k][0])).cast("B")
            got = 0
            while got < n:
                r = os.preadv(fd, [view[got:]], a + got)
                if r <= 0:
                    raise IOError("short read")
                got += r
            while n > 1 and view[n - 1] == 10 and view[n - 2] == 10:   # blank lines at the end of the file
                n -= 1
            return pipeline.Chunk(inbuf[k][0], n)

        def sink(w, ptr, n, ofs):
            view = memoryview((C.c_char * n).from_address(ptr)).cast("B") if n else b""
            done = 0
            while done < n:
                done += os.pwrite(out_fd, view[done:], ofs + done)

        pipe = pipeline.WavePipeline(ms, ex, flags, device_index=local if backend == "nccl" else None)
        t = pipe.run(n_waves, get_chunk, sink)
        tot["reads"] += t["reads"]; tot["bases"] += t["bases"]
        pipe.close()
        os.close(fd)
        for b in inbuf:
            if b is not None:
                m0.lib.mab_host_free(b[0])
    if world > 1:
        v = torch.tensor([tot["reads"], tot["bases"]], dtype=torch.int64, device=dev)
        dist.all_reduce(v)
        tot["reads"], tot["bases"] = int(v[0]), int(v[1])
        dist.barrier()
    os.close(out_fd)
    if rank == 0:
        dt = time.time() - t0
        print(f"[M::main] mapped {tot['reads']} reads / {tot['bases'] / 1e6:.1f} Mbases on {world} rank(s) in {dt:.3f} sec (index load and device set-up included)", file=sys.stderr)
    for m in ms[1:] + ms[:1]:
        m.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
