"""One rank's mapping pipeline over the text path of the C ABI (mab_text_begin / commit / finish): several contexts on the rank's
GPU, each driven by its own host thread, chunks ("waves": this rank's share of the file, see shard.py) handed to them in order.

    worker (one per context)       begin(chunk)  ->  [main: exchange]  ->  commit(rlen)  ->  finish  ->  [main: offsets]  ->  sink
    main thread                    does every collective, in wave order, so that all ranks issue them in the same order

The host side mirrors the reference's source -> worker -> drain pipeline (minialign.c:4565-4643); the part the reference's drain
enforces with a heap on batch ids (4633-4643) is the in-order commit / offset exchange here.  bench.py (synthetic chunks in
page-locked memory) and mgpu.py (chunks of a read file) both run on it.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading

from . import api
from .shard import WaveExchange


class Chunk:
    """Text of one chunk: `ptr` / `n` (host pointer, or device pointer when the mappers are in device-input mode); `keep` holds
    whatever owns the memory."""

    def __init__(self, ptr: int, n: int, keep=None):
        self.ptr, self.n, self.keep = ptr, n, keep


class WavePipeline:
    def __init__(self, mappers, exchange: WaveExchange | None = None, flags: int = 0, out_cap: int = 0, device_index: int | None = None):
        self.ms = mappers
        self.ex = exchange or WaveExchange()
        self.flags = flags
        self.device_index = device_index
        self.out = [None] * len(mappers)            # page-locked output buffer per context (grown on demand)
        self.out_cap = [0] * len(mappers)
        self.out_cap0 = out_cap
        self.totals = dict(reads=0, bases=0, sam_bytes=0, waves=0, launches=0, h2d=0, d2h=0, redo=0, failed=0, ms_extend_r0=0.0, ms_extend=0.0, vectors=0,
                           ms_post=0.0)
        self.lock = threading.Lock()

    def _out_buffer(self, k: int, need: int):
        if self.out_cap[k] < need:
            m = self.ms[k]
            if self.out[k]:
                m.lib.mab_host_free(self.out[k])
            cap = need + need // 8 + (1 << 20)
            p = m.lib.mab_host_alloc(cap)
            if not p:
                raise RuntimeError("mab_host_alloc failed: " + m.lib.mab_last_error().decode())
            self.out[k], self.out_cap[k] = p, cap
        return self.out[k], self.out_cap[k]

    def close(self):
        for k, m in enumerate(self.ms):
            if self.out[k]:
                m.lib.mab_host_free(self.out[k])
                self.out[k] = None

    def run(self, n_waves: int, get_chunk, sink=None):
        """get_chunk(w) -> Chunk or None (no chunk for this rank in wave w); sink(w, ptr, n_bytes, offset) consumes the SAM text of
        wave w (ptr is valid until it returns; ptr = 0 when the text stays on the device).  Returns the totals dict."""
        K = len(self.ms)
        cmd = [queue.Queue() for _ in range(K)]
        begun, committed, finished = {}, {}, {}
        evs = [dict(b=threading.Event(), c=queue.Queue(), f=threading.Event()) for _ in range(n_waves)]
        errors = []
        device_out = bool(self.flags & api.TEXT_DEVICE_OUT)

        def worker(k):
            m = self.ms[k]
            try:
                if self.device_index is not None:
                    import torch
                    torch.cuda.set_device(self.device_index)
                for w in range(k, n_waves, K):
                    ch = get_chunk(w)
                    info = m.text_begin(ch.ptr, ch.n, self.flags, 0, False) if ch is not None and ch.n else None
                    begun[w] = info
                    evs[w]["b"].set()
                    while True:
                        op, val = cmd[k].get()
                        if op == "commit":
                            if info is not None:
                                info = m.text_commit(val)
                            evs[w]["c"].put(info)
                        elif op == "finish":
                            break
                        else:
                            return
                    ptr, st = 0, None
                    if info is not None:
                        if device_out:
                            info, ptr = m.text_finish()
                        else:
                            buf, cap = self._out_buffer(k, max(self.out_cap0, ch.n + ch.n // 2 + (1 << 20)))
                            info, ptr = m.text_finish(buf, cap)
                            if ptr is None:         # more text than estimated: a larger buffer, format again
                                buf, cap = self._out_buffer(k, int(info.sam_bytes))
                                info, ptr = m.text_finish(buf, cap)
                        st = m.stats()
                    finished[w] = (info, ptr, st)
                    evs[w]["f"].set()
                    op, val = cmd[k].get()          # ("sink", offset): consume the text, then the context is free again
                    if op != "sink":
                        return
                    if info is not None and sink is not None:
                        sink(w, 0 if device_out else ptr, int(info.sam_bytes), val)
                    if info is not None:
                        with self.lock:
                            t = self.totals
                            t["reads"] += int(info.n_reads); t["bases"] += int(info.n_bases); t["sam_bytes"] += int(info.sam_bytes); t["waves"] += 1
                            t["launches"] += st["n_launches"]; t["h2d"] += st["h2d_bytes"]; t["d2h"] += st["d2h_bytes"]; t["redo"] += st["n_retry"]
                            t["failed"] += st["n_failed"]; t["ms_extend_r0"] += st["ms_extend_r0"]; t["ms_extend"] += st["ms_extend"]; t["vectors"] += st["n_vectors"]
                            t["ms_post"] += st["ms_post"]
            except Exception as e:      # a failed chunk must fail the run, not shorten it
                errors.append(e)
                for w in range(n_waves):
                    evs[w]["b"].set(); evs[w]["f"].set(); evs[w]["c"].put(None)

        threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(K)]
        [t.start() for t in threads]

        def check():
            if errors:
                for q in cmd:
                    q.put(("stop", None))
                raise errors[0]

        # the offsets exchange (byte counts -> prefix sum) runs on its own thread, in wave order: nothing but the sink of a wave waits
        # for it, so the commit loop below never blocks behind a collective that has to find room on a busy GPU
        def offsets_loop():
            try:
                if self.device_index is not None:
                    import torch
                    torch.cuda.set_device(self.device_index)
                for w in range(n_waves):
                    evs[w]["f"].wait()
                    if errors:
                        return
                    info = finished[w][0]
                    ofs, _total = self.ex.offsets(int(info.sam_bytes) if info is not None else 0)
                    cmd[w % K].put(("sink", ofs))
            except Exception as e:
                errors.append(e)
                for q in cmd:
                    q.put(("stop", None))

        ot = threading.Thread(target=offsets_loop, daemon=True)
        ot.start()
        for w in range(n_waves):
            evs[w]["b"].wait(); check()
            info = begun[w]
            self.ex.begin_wave(bool(info.rlen_valid) if info is not None else False, int(info.rlen_next) if info is not None else 0)

            def commit(v, w=w):
                cmd[w % K].put(("commit", v))
                r = evs[w]["c"].get(); check()
                return (bool(r.rlen_valid), int(r.rlen_next)) if r is not None else (False, 0)
            self.ex.settle(commit)
            cmd[w % K].put(("finish", None))
        ot.join()
        [t.join() for t in threads]
        check()
        return self.totals
