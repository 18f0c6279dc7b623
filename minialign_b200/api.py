"""ctypes binding of the C ABI in include/minialign_b200.h (host-side mirror of the reference's mapper interface).

`Mapper(mai_blob, params)` corresponds to mm_align_init (minialign.c:4671), `Mapper.map_batch(reads)` to one
mm_align_worker call over a bseq_t batch (minialign.c:4589-4601).  The library is the in-tree CUDA build
(minialign_b200/libminialign_b200.so); there is no CPU fallback: loading fails loudly when it is missing and mab_init
fails when no sm_100 device is usable.  Tests may pass `lib_path` to load the CUDA-on-CPU emulation build of the same
sources (tests/emu), which is never used by the product.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libminialign_b200.so")


class MabParams(C.Structure):
    _fields_ = [("wlen", C.c_int32), ("glen", C.c_int32), ("min_score", C.c_uint32), ("min_ratio", C.c_float),
                ("score_matrix", C.c_int8 * 16), ("gi", C.c_int8), ("ge", C.c_int8), ("gfa", C.c_int8), ("gfb", C.c_int8),
                ("xdrop", C.c_int8), ("_pad", C.c_uint8 * 3), ("flags", C.c_uint32)]


class MabStats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_seed", C.c_float), ("ms_sortchain", C.c_float),
                ("ms_extend", C.c_float), ("ms_d2h", C.c_float), ("ms_post", C.c_float), ("ms_extend_r0", C.c_float),
                ("n_vectors", C.c_uint64), ("n_fill_calls", C.c_uint64), ("n_trace", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("n_launches", C.c_uint32), ("n_retry", C.c_uint32),
                ("ms_wall", C.c_float), ("ms_wall_sizing", C.c_float), ("ms_wall_wait", C.c_float), ("ms_wall_submit", C.c_float),
                ("n_failed", C.c_uint32), ("_pad", C.c_uint32)]


class MabTextInfo(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("sam_bytes", C.c_uint64), ("rlen_valid", C.c_uint32), ("rlen_next", C.c_uint32)]


# optional SAM tags (include/minialign_b200.h) and text-path flags
TAGS = {"RG": 1 << 0, "NH": 1 << 2, "IH": 1 << 3, "AS": 1 << 4, "XS": 1 << 5, "NM": 1 << 6, "SA": 1 << 7, "MD": 1 << 8}
OMIT_REP = 1 << 30
TEXT_KEEP_QUAL, TEXT_DEVICE_OUT = 0x01000000, 0x02000000


def parse_tags(s: str) -> int:
    t = 0
    for tok in s.replace(";", ",").replace(":", ",").replace("/", ",").split(","):
        t |= TAGS.get(tok, 0)
    return t


class MabPair(C.Structure):
    _fields_ = [("a_ofs", C.c_uint64), ("b_ofs", C.c_uint64), ("alen", C.c_uint32), ("blen", C.c_uint32), ("apos", C.c_uint32),
                ("bpos", C.c_uint32), ("brev", C.c_uint32), ("narrow", C.c_uint32), ("min_score", C.c_int64)]


# presets of the reference CLI (minialign.c:5853-5878, defaults 6141-6161)
PRESETS = {
    "pacbio": dict(wlen=7000, glen=7000, min_score=50, min_ratio=0.3, gi=4, ge=2, gfa=3, gfb=3, xdrop=50,
                   score_matrix=[2 if i % 5 == 0 else -4 for i in range(16)]),
    "ont.1dsq": dict(wlen=7000, glen=7000, min_score=50, min_ratio=0.3, gi=6, ge=2, gfa=4, gfb=4, xdrop=50,
                     score_matrix=[2 if i % 5 == 0 else -6 for i in range(16)]),
}


def make_params(d: dict) -> MabParams:
    p = MabParams()
    p.wlen, p.glen, p.min_score, p.min_ratio = d["wlen"], d["glen"], d["min_score"], d["min_ratio"]
    for i, s in enumerate(d["score_matrix"]):
        p.score_matrix[i] = s
    p.gi, p.ge, p.gfa, p.gfb, p.xdrop = d["gi"], d["ge"], d["gfa"], d["gfb"], d["xdrop"]
    p.flags = 0
    return p


def load_library(lib_path: str | None = None):
    path = lib_path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(path)
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.mab_init.restype = C.c_void_p
    L.mab_init.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(MabParams), C.c_int]
    L.mab_clone.restype = C.c_void_p
    L.mab_clone.argtypes = [C.c_void_p]
    L.mab_destroy.argtypes = [C.c_void_p]
    L.mab_device_memory.restype = C.c_int
    L.mab_device_memory.argtypes = [C.c_void_p, u64p, u64p]
    L.mab_set_arena_budget.argtypes = [C.c_void_p, C.c_uint64]
    L.mab_last_error.restype = C.c_char_p
    L.mab_n_ref.restype = C.c_uint32
    L.mab_n_ref.argtypes = [C.c_void_p]
    L.mab_map_batch.restype = C.c_int
    L.mab_map_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u32p, C.c_uint32]
    L.mab_result.restype = C.c_uint64
    L.mab_result.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(u32p)]
    L.mab_release_batch.argtypes = [C.c_void_p]
    L.mab_detach_batch.restype = C.c_void_p
    L.mab_detach_batch.argtypes = [C.c_void_p]
    L.mab_results_get.restype = C.c_uint64
    L.mab_results_get.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(u32p)]
    L.mab_results_free.argtypes = [C.c_void_p]
    L.mab_last_stats.argtypes = [C.c_void_p, C.POINTER(MabStats)]
    L.mab_set_device_input.argtypes = [C.c_void_p, C.c_int]
    L.mab_sketch.restype = C.c_uint64
    L.mab_sketch.argtypes = [C.c_void_p, u8p, C.c_uint32, u64p, C.c_uint64]
    L.mab_seed_chain.restype = C.c_uint64
    L.mab_seed_chain.argtypes = [C.c_void_p, u8p, C.c_uint32, C.c_uint32, u32p, C.c_uint64, u64p, u32p, C.c_uint64, u64p]
    L.mab_extend_pairs.restype = C.c_int
    L.mab_extend_pairs.argtypes = [C.c_void_p, u8p, C.c_uint64, C.POINTER(MabPair), C.c_uint32, u32p, u32p, C.c_uint64, u64p]
    L.mab_fill_peak.restype = C.c_int
    L.mab_fill_peak.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_double)]
    L.mab_selftest.restype = C.c_int
    L.mab_selftest.argtypes = [C.c_void_p, u32p]
    ti = C.POINTER(MabTextInfo)
    L.mab_load_begin.restype = C.c_void_p
    L.mab_load_begin.argtypes = [C.c_uint64, C.POINTER(MabParams), C.POINTER(C.c_int), C.c_int]
    L.mab_load_put.restype = C.c_int
    L.mab_load_put.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    L.mab_load_end.restype = C.c_int
    L.mab_load_end.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    L.mab_load_abort.restype = None
    L.mab_load_abort.argtypes = [C.c_void_p]
    L.mab_text_reserve.restype = C.c_int
    L.mab_text_reserve.argtypes = [C.c_void_p, C.c_uint64]
    L.mab_text_begin.restype = C.c_int
    L.mab_text_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, ti]
    L.mab_text_commit.restype = C.c_int
    L.mab_text_commit.argtypes = [C.c_void_p, C.c_uint32, ti]
    L.mab_text_finish.restype = C.c_int
    L.mab_text_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), ti]
    L.mab_map_text.restype = C.c_int
    L.mab_map_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), ti]
    L.mab_sam_header_text.restype = C.c_uint64
    L.mab_sam_header_text.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint64]
    L.mab_host_alloc.restype = C.c_void_p
    L.mab_host_alloc.argtypes = [C.c_uint64]
    L.mab_host_free.argtypes = [C.c_void_p]
    return L


def pack_reads(reads, margin: int = 64):
    """Concatenate encoded reads into one block with 64-byte zero margins (the layout of bseq_t, minialign.c:2109-2146)."""
    lens = np.array([r.size for r in reads], dtype=np.uint32)
    ofs = np.zeros(len(reads), dtype=np.uint64)
    total = margin
    for i, r in enumerate(reads):
        ofs[i] = total
        total += r.size + margin
    block = np.zeros(total + margin, dtype=np.uint8)
    for i, r in enumerate(reads):
        block[int(ofs[i]):int(ofs[i]) + r.size] = r
    return block, ofs, lens


class Mapper:
    def __init__(self, mai_blob: np.ndarray, params: dict | str = "pacbio", device: int = 0, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        self.params = make_params(PRESETS[params] if isinstance(params, str) else params)
        self.blob = np.ascontiguousarray(mai_blob, dtype=np.uint8)
        self.h = self.lib.mab_init(self.blob.ctypes.data, self.blob.size, C.byref(self.params), device)
        if not self.h:
            raise RuntimeError("mab_init failed: " + self.lib.mab_last_error().decode())

    @classmethod
    def staged(cls, mai_blob: np.ndarray, params: dict | str = "pacbio", devices=(0,), piece: int = 1 << 20, order=None, lib_path: str | None = None):
        """The staged set-up (mab_load_begin / put / end): the image goes to every listed device piece by piece (`order`: the
        sequence the pieces are handed over in, default ascending); returns one Mapper per device."""
        lib = load_library(lib_path)
        prm = make_params(PRESETS[params] if isinstance(params, str) else params)
        blob = np.ascontiguousarray(mai_blob, dtype=np.uint8)
        devs = (C.c_int * len(devices))(*devices)
        ld = lib.mab_load_begin(blob.size, C.byref(prm), devs, len(devices))
        if not ld:
            raise RuntimeError("mab_load_begin failed: " + lib.mab_last_error().decode())
        starts = list(range(0, blob.size, piece))
        for k in (order(starts) if order else starts):
            if lib.mab_load_put(ld, k, blob.ctypes.data + k, min(piece, blob.size - k)) != 0:
                lib.mab_load_abort(ld)
                raise RuntimeError("mab_load_put failed: " + lib.mab_last_error().decode())
        out = (C.c_void_p * len(devices))()
        if lib.mab_load_end(ld, blob.ctypes.data, blob.size, out) != 0:
            raise RuntimeError("mab_load_end failed: " + lib.mab_last_error().decode())
        ms = []
        for h in out:
            m = cls.__new__(cls)
            m.lib, m.params, m.blob, m.h = lib, prm, blob, h
            ms.append(m)
        return ms

    @classmethod
    def from_mai(cls, path: str, params: dict | str = "pacbio", device: int = 0, lib_path: str | None = None) -> "Mapper":
        """Context from a .mai file with the upload running under the inflation (mai.inflate_mai + the staged set-up): a
        human-sized index is on the GPU when its last frame is inflated."""
        from . import mai
        lib = load_library(lib_path)
        prm = make_params(PRESETS[params] if isinstance(params, str) else params)
        prm.flags |= 1                      # MAB_FLAG_BORROW_INDEX: the image stays with the Mapper (self.blob), no second host copy
        st = {"ld": None, "err": None}
        devs = (C.c_int * 1)(device)

        def on_size(size):
            st["ld"] = lib.mab_load_begin(size, C.byref(prm), devs, 1)
            if not st["ld"]:
                st["err"] = "mab_load_begin failed: " + lib.mab_last_error().decode()

        def on_piece(off, addr, n):
            if st["ld"] and st["err"] is None and lib.mab_load_put(st["ld"], off, addr, n) != 0:
                st["err"] = "mab_load_put failed: " + lib.mab_last_error().decode()

        try:
            blob = mai.inflate_mai(path, on_size, on_piece)
        except Exception:
            if st["ld"]:
                lib.mab_load_abort(st["ld"])
            raise
        if st["err"] is not None:
            if st["ld"]:
                lib.mab_load_abort(st["ld"])
            raise RuntimeError(st["err"])
        out = (C.c_void_p * 1)()
        if lib.mab_load_end(st["ld"], blob.ctypes.data, blob.size, out) != 0:
            raise RuntimeError("mab_load_end failed: " + lib.mab_last_error().decode())
        m = cls.__new__(cls)
        m.lib, m.params, m.blob, m.h = lib, prm, blob, out[0]
        return m

    def clone(self) -> "Mapper":
        """Another context on the same device sharing this one's index image (keep this one alive while the clone is used)."""
        m = Mapper.__new__(Mapper)
        m.lib, m.params, m.blob = self.lib, self.params, self.blob
        m.h = self.lib.mab_clone(self.h)
        if not m.h:
            raise RuntimeError("mab_clone failed: " + self.lib.mab_last_error().decode())
        return m

    def close(self):
        if self.h:
            self.lib.mab_destroy(self.h)
            self.h = None

    def map_packed(self, block_ptr: int, block_size: int, ofs: np.ndarray, lens: np.ndarray) -> int:
        rc = self.lib.mab_map_batch(self.h, block_ptr, block_size, ofs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    lens.ctypes.data_as(C.POINTER(C.c_uint32)), len(lens))
        if rc != 0:
            raise RuntimeError(f"mab_map_batch failed ({rc}): " + self.lib.mab_last_error().decode())
        return rc

    def result(self, i: int) -> np.ndarray:
        p = C.POINTER(C.c_uint32)()
        n = self.lib.mab_result(self.h, i, C.byref(p))
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def map_batch(self, reads):
        """reads: list of uint8 code arrays -> list of flat result arrays (empty = unmapped)."""
        block, ofs, lens = pack_reads(reads)
        self.map_packed(block.ctypes.data, block.size, ofs, lens)
        out = [self.result(i) for i in range(len(reads))]
        self.lib.mab_release_batch(self.h)
        return out

    # ---- text path: FASTA / FASTQ bytes in, SAM bytes out (mab_text_* in include/minialign_b200.h) ----
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): " + self.lib.mab_last_error().decode())

    def map_text(self, text: bytes, tags: int = 0, keep_qual: bool = False) -> bytes:
        """One chunk, sequential use: the reference thread's state is chained through the context."""
        info, ptr = MabTextInfo(), C.c_void_p()
        flags = tags | (TEXT_KEEP_QUAL if keep_qual else 0)
        buf = C.create_string_buffer(text, len(text)) if not isinstance(text, C.Array) else text
        self._check(self.lib.mab_map_text(self.h, C.addressof(buf), len(text), flags, None, 0, C.byref(ptr), C.byref(info)), "mab_map_text")
        self.last_info = info
        return C.string_at(ptr.value, info.sam_bytes) if info.sam_bytes else b""

    def text_reserve(self, max_chunk_bytes: int):
        """announce the largest chunk (bytes) this mapper and its clones will be given: buffers are sized for it at once"""
        self._check(self.lib.mab_text_reserve(self.h, max_chunk_bytes), "mab_text_reserve")

    def text_begin(self, ptr: int, n: int, flags: int = 0, rlen_prev: int = 0, rlen_known: bool = False) -> MabTextInfo:
        info = MabTextInfo()
        self._check(self.lib.mab_text_begin(self.h, ptr, n, flags, rlen_prev, 1 if rlen_known else 0, C.byref(info)), "mab_text_begin")
        return info

    def text_commit(self, rlen_prev: int) -> MabTextInfo:
        info = MabTextInfo()
        self._check(self.lib.mab_text_commit(self.h, rlen_prev, C.byref(info)), "mab_text_commit")
        return info

    def text_finish(self, out_ptr: int = 0, out_cap: int = 0):
        """-> (info, pointer to the SAM text).  out_ptr = 0: the text stays in the context's pinned buffer."""
        info, ptr = MabTextInfo(), C.c_void_p()
        rc = self.lib.mab_text_finish(self.h, out_ptr or None, out_cap, C.byref(ptr), C.byref(info))
        if rc == -3 and out_ptr and info.sam_bytes > out_cap:
            return info, None           # buffer too small: the chunk stays in flight, call again with info.sam_bytes of room
        self._check(rc, "mab_text_finish")
        return info, ptr.value

    def sam_header(self, cmdline: str = "", version: str = "0.6.0-devel") -> bytes:
        n = self.lib.mab_sam_header_text(self.h, version.encode(), cmdline.encode(), None, 0)
        buf = C.create_string_buffer(n)
        self.lib.mab_sam_header_text(self.h, version.encode(), cmdline.encode(), buf, n)
        return buf.raw[:n]

    def stats(self) -> dict:
        s = MabStats()
        self.lib.mab_last_stats(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in MabStats._fields_}

    def fill_peak(self, masks: bool = True, n_blocks: int = 2000) -> float:
        """Vectors/s ceiling of the DP step's instruction mix (k_fill_peak); see include/minialign_b200.h."""
        v = C.c_double(0.0)
        rc = self.lib.mab_fill_peak(self.h, 1 if masks else 0, n_blocks, C.byref(v))
        if rc != 0:
            raise RuntimeError("mab_fill_peak failed: " + self.lib.mab_last_error().decode())
        return v.value

    def selftest(self) -> np.ndarray:
        out = np.zeros(64 * 32, dtype=np.uint32)
        rc = self.lib.mab_selftest(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        if rc != 0:
            raise RuntimeError("mab_selftest failed: " + self.lib.mab_last_error().decode())
        return out.reshape(64, 32)

    def sort_check(self, elems: np.ndarray):
        """elems: (n, 4) uint32 -> (cycle-walking sort, parallel form), both (n, 4)"""
        e = np.ascontiguousarray(elems, dtype=np.uint32)
        a, b = np.zeros_like(e), np.zeros_like(e)
        u32p = C.POINTER(C.c_uint32)
        self.lib.mab_sort_check.restype = C.c_int
        self.lib.mab_sort_check.argtypes = [C.c_void_p, u32p, C.c_uint32, u32p, u32p]
        self._check(self.lib.mab_sort_check(self.h, e.ctypes.data_as(u32p), e.shape[0], a.ctypes.data_as(u32p), b.ctypes.data_as(u32p)), "mab_sort_check")
        return a, b

    # ---- stage-level entry points (parity tests) ----
    def sketch(self, seq: np.ndarray) -> np.ndarray:
        buf = np.zeros(seq.size + 128, dtype=np.uint8)
        buf[64:64 + seq.size] = seq
        cap = seq.size + 16
        out = np.zeros(cap, dtype=np.uint64)
        n = self.lib.mab_sketch(self.h, buf[64:].ctypes.data_as(C.POINTER(C.c_uint8)), seq.size, out.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
        return out[:n]

    def seed_chain(self, seq: np.ndarray, rnd: int = 0):
        buf = np.zeros(seq.size + 128, dtype=np.uint8)
        buf[64:64 + seq.size] = seq
        cap = 1 << 20
        seeds = np.zeros(cap * 4, dtype=np.uint32)
        roots = np.zeros(cap * 2, dtype=np.uint32)
        nt, nr = C.c_uint64(0), C.c_uint64(0)
        ns = self.lib.mab_seed_chain(self.h, buf[64:].ctypes.data_as(C.POINTER(C.c_uint8)), seq.size, rnd,
                                     seeds.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(nt),
                                     roots.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(nr))
        return ns, seeds[: nt.value * 4].reshape(-1, 4).copy(), roots[: nr.value * 2].reshape(-1, 2).copy()

    def extend_pairs(self, pairs):
        """pairs: list of (a, b, apos, bpos, brev, narrow, min_score) -> list of (res[16], aln words)."""
        seqs = []
        for a, b, *_ in pairs:
            seqs += [a, b]
        block, ofs, lens = pack_reads(seqs)
        arr = (MabPair * len(pairs))()
        for i, (a, b, apos, bpos, brev, narrow, ms) in enumerate(pairs):
            arr[i] = MabPair(int(ofs[2 * i]), int(ofs[2 * i + 1]), a.size, b.size, apos, bpos, brev, narrow, ms)
        res = np.zeros(16 * len(pairs), dtype=np.uint32)
        cap = int(sum(a.size + b.size for a, b, *_ in pairs)) // 8 + 64 * len(pairs) + 4096
        aln = np.zeros(cap, dtype=np.uint32)
        ao = np.zeros(len(pairs) + 1, dtype=np.uint64)
        rc = self.lib.mab_extend_pairs(self.h, block.ctypes.data_as(C.POINTER(C.c_uint8)), block.size, arr, len(pairs),
                                       res.ctypes.data_as(C.POINTER(C.c_uint32)), aln.ctypes.data_as(C.POINTER(C.c_uint32)), cap,
                                       ao.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc != 0:
            raise RuntimeError(f"mab_extend_pairs failed ({rc}): " + self.lib.mab_last_error().decode())
        return [(res[16 * i:16 * i + 16].copy(), aln[int(ao[i]):int(ao[i + 1])].copy()) for i in range(len(pairs))]
