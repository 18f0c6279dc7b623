"""Read sharding across ranks (one process per GPU): the exchange step of the mapping path.

The chunks of ONE read file are dealt round-robin: wave w = chunks w*N .. w*N+N-1, rank r maps chunk w*N + r.  Reads are
independent except for the single word of state the reference's worker thread carries from read to read (`rlen`, minialign.c:3865
vs 3873; with -t1 that is file order), and the merged SAM needs every chunk's byte offset.  Both are a few bytes per rank per
wave and both are prefix problems in chunk order:

  rlen chain      each rank contributes (loaded-a-chain?, value it leaves behind); chunk c starts from the value left by the last
                  chain-loading chunk before it.  A rank whose chunk's first seed test flips under the true value re-maps that read
                  (mab_text_commit); in the rare case this changes what the chunk leaves behind, a second round propagates it.
  output offsets  all-gather of the SAM byte counts -> exclusive prefix sum -> every rank pwrite()s its text at its own offset.

The collectives are torch.distributed all_gathers of a handful of int64 (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Everything else -- the index, the reads, the DP -- stays on the rank's own GPU: no data-path collective.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def deal_chunks(n_chunks: int, world: int, rank: int):
    """Chunk ids of this rank, in wave order."""
    return list(range(rank, n_chunks, world))


def exclusive_offsets(sizes):
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += int(s)
    return out, acc


class WaveExchange:
    """The per-wave collectives.  Every rank must call the methods in the same order, once per wave (ranks without a chunk in the
    last, partial wave pass valid=False / 0 bytes)."""

    def __init__(self, device=None, host_group=None):
        """device: where the default group's tensors live (cuda device for NCCL, None for gloo).  host_group: an optional gloo group
        for the rlen chain -- that exchange is on the critical path of every wave (the chunks cannot be printed before it) and its
        payload is host data; an NCCL kernel for it would queue behind the mapping kernels that keep the GPU full.  The byte-count
        all-gather for the output offsets stays on the default (NCCL) group: nothing waits for it but the writer."""
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device
        self.host_group = host_group
        self.rlen = 0               # a fresh reference thread starts with rlen = 0
        self.out_base = 0           # bytes of SAM text before this wave
        self.n_collectives = 0
        self.t_collectives = 0.0    # seconds the calling thread spent inside the collectives

    def _gather(self, vals, host=False):
        if self.world == 1:
            return [list(vals)]
        import time
        t0 = time.perf_counter()
        if host and self.host_group is None and self.world > 1:
            raise RuntimeError("WaveExchange: the rlen chain needs its own (gloo) group when world > 1")
        if host:
            t = torch.tensor(list(vals), dtype=torch.int64)
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t, group=self.host_group)
            rows = [p.tolist() for p in parts]
        else:
            t = torch.tensor(list(vals), dtype=torch.int64, device=self.device)
            out = torch.empty(self.world * t.numel(), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(out, t)
            rows = out.view(self.world, -1).tolist()
        self.n_collectives += 1
        self.t_collectives += time.perf_counter() - t0
        return rows

    def rlen_inputs(self, valid: bool, value: int):
        """(valid, value) = what this rank's chunk leaves behind under its current assumption.  Returns the value this rank's chunk
        must be committed with; `self.rlen` is not advanced before settle() confirms the wave."""
        rows = self._gather((1 if valid else 0, int(value)), host=True)
        run, mine = self.rlen, None
        for q, (v, x) in enumerate(rows):
            if q == self.rank:
                mine = run
            if v:
                run = x
        self._pending = run
        return mine

    def settle(self, commit):
        """commit(rlen_in) -> (valid, value) re-maps whatever depends on rlen_in and returns what the chunk then leaves behind.
        Repeats the exchange until no rank's value moved (normally one round: a corrected first read almost never is the last
        chain-loading read of its chunk).  Advances self.rlen to the wave's final value."""
        valid, value = self._last
        while True:
            rlen_in = self.rlen_inputs(valid, value)
            nv, nx = commit(rlen_in)
            changed = (bool(nv), int(nx)) != (bool(valid), int(value))
            valid, value = nv, nx
            if self.world == 1:
                any_changed = changed
            else:
                any_changed = any(r[0] for r in self._gather((1 if changed else 0,), host=True))
            if not any_changed:
                break
        self.rlen = self._pending
        return rlen_in

    def begin_wave(self, valid: bool, value: int):
        self._last = (bool(valid), int(value))

    def offsets(self, n_bytes: int):
        """SAM bytes of this rank's chunk -> (absolute offset of this rank's text, total bytes of the wave)."""
        rows = self._gather((int(n_bytes),))
        ofs, total = exclusive_offsets([r[0] for r in rows])
        mine = self.out_base + ofs[self.rank]
        self.out_base += total
        return mine, total


def bind_near_gpu(device_index: int) -> str:
    """Put this process next to its GPU BEFORE it allocates page-locked buffers and starts threads: (1) its threads on the cores
    of the GPU's NUMA node (sysfs local_cpulist of the device's PCI function, cut down to the cores the process may use at all),
    (2) its memory preferably on that node (set_mempolicy(MPOL_PREFERRED): a soft preference, also when the cores cannot follow
    -- a container whose cores all sit on one socket still has the other socket's GPUs copy 0.8 GB per chunk across the
    inter-socket link otherwise).  Returns what was done, for the record; leaves things alone when the topology cannot be read."""
    done = []
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception as e:
        return f"unchanged ({type(e).__name__})"
    try:
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            txt = f.read().strip()
        near = set()
        for part in txt.split(","):
            if part:
                a, _, b = part.partition("-")
                near.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(near & allowed)
        if len(use) < 4 or len(use) == len(allowed):
            done.append(f"cores unchanged ({len(allowed)} allowed, {len(use)} of them near {bus})")
        else:
            os.sched_setaffinity(0, use)
            done.append(f"{len(use)} of {len(allowed)} cores, near {bus}")
    except Exception as e:          # no sysfs, ...: leave the scheduler alone
        done.append(f"cores unchanged ({type(e).__name__})")
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        import platform
        if platform.machine() != "x86_64":
            done.append("memory policy unchanged (syscall number known for x86-64 only)")
        elif 0 <= node < 64:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), 65)       # x86-64 set_mempolicy(MPOL_PREFERRED, &mask, maxnode)
            done.append(f"memory preferred on node {node}" if rc == 0 else f"memory policy unchanged (errno {ctypes.get_errno()})")
        else:
            done.append("memory policy unchanged (no NUMA node reported)")
    except Exception as e:
        done.append(f"memory policy unchanged ({type(e).__name__})")
    return "; ".join(done)
