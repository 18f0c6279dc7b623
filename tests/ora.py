"""ctypes wrapper around oracle/liboracle.so (the CPU restatement; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "liboracle.so")


class OraParams(C.Structure):
    _fields_ = [("score_matrix", C.c_int8 * 16), ("gi", C.c_int8), ("ge", C.c_int8), ("gfa", C.c_int8), ("gfb", C.c_int8), ("xdrop", C.c_int8)]


class MmoParams(C.Structure):
    _fields_ = [("k", C.c_uint32), ("w", C.c_uint32), ("b", C.c_uint32), ("n_occ", C.c_uint32), ("occ", C.c_uint32 * 8),
                ("wlen", C.c_int32), ("glen", C.c_int32), ("min_score", C.c_uint32), ("min_ratio", C.c_float), ("gp", OraParams)]


def build():
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("gaba_oracle.c", "mm_oracle.c", "gaba_oracle.h", "mm_oracle.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-f", "oracle/Makefile"], cwd=ROOT, stdout=subprocess.DEVNULL)


def make_params(d: dict) -> MmoParams:
    p = MmoParams()
    p.k, p.w, p.b, p.n_occ = d["k"], d["w"], d["b"], d["n_occ"]
    for i, o in enumerate(d["occ"]):
        p.occ[i] = o & 0xFFFFFFFF
    p.wlen, p.glen, p.min_score, p.min_ratio = d["wlen"], d["glen"], d["min_score"], d["min_ratio"]
    for i, s in enumerate(d["score_matrix"]):
        p.gp.score_matrix[i] = s
    p.gp.gi, p.gp.ge, p.gp.gfa, p.gp.gfb, p.gp.xdrop = d["gi"], d["ge"], d["gfa"], d["gfb"], d["xdrop"]
    return p


PACBIO = dict(k=15, w=10, b=14, n_occ=3, occ=[0, 0, 0], wlen=7000, glen=7000, min_score=50, min_ratio=0.3,
              gi=4, ge=2, gfa=3, gfb=3, xdrop=50, score_matrix=[2, -4, -4, -4, -4, 2, -4, -4, -4, -4, 2, -4, -4, -4, -4, 2])
ONT = dict(PACBIO, gi=6, ge=2, gfa=4, gfb=4, score_matrix=[2 if i % 5 == 0 else -6 for i in range(16)])


def custom(a, b, gi, ge, gfa, gfb, xdrop):
    """A non-preset scoring scheme (-a -b -p -q -r -Y, minialign.c:5950-6000) and the command line that selects it."""
    prm = dict(PACBIO, gi=gi, ge=ge, gfa=gfa, gfb=gfb, xdrop=xdrop, score_matrix=[a if i % 5 == 0 else -b for i in range(16)])
    return prm, ("-xpacbio", f"-a{a}", f"-b{b}", f"-p{gi}", f"-q{ge}", f"-r{gfa},{gfb}", f"-Y{xdrop}")


CUSTOM = [custom(1, 2, 2, 1, 2, 2, 30), custom(3, 5, 5, 3, 4, 5, 70), custom(4, 6, 6, 2, 4, 4, 90), custom(1, 1, 1, 1, 2, 2, 20)]
API_KEYS = ("wlen", "glen", "min_score", "min_ratio", "gi", "ge", "gfa", "gfb", "xdrop", "score_matrix")


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def pad(seq, margin=64):
    buf = np.zeros(seq.size + 2 * margin, dtype=np.uint8)
    buf[margin:margin + seq.size] = seq
    return buf


class Oracle:
    def __init__(self, params: dict, mai_blob: np.ndarray | None = None):
        build()
        self.lib = L = C.CDLL(SO)
        self.p = make_params(params)
        self.h = None
        L.mmo_extend.restype = C.c_uint64
        L.mmo_extend.argtypes = [C.POINTER(MmoParams), C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_uint32, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint64]
        if not hasattr(L, "mmo_init"):
            return
        L.mmo_init.restype = C.c_void_p
        L.mmo_init.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(MmoParams)]
        L.mmo_destroy.argtypes = [C.c_void_p]
        L.mmo_sketch.restype = C.c_uint64
        L.mmo_sketch.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint64), C.c_uint64]
        L.mmo_get.restype = C.c_uint32
        L.mmo_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
        L.mmo_seed_chain.restype = C.c_uint64
        L.mmo_seed_chain.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint64)]
        L.mmo_align.restype = C.c_uint64
        L.mmo_align.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64]
        L.mmo_vec_count.restype = C.c_uint64
        L.mmo_vec_count.argtypes = [C.c_void_p]
        self.h = None
        if mai_blob is not None:
            self.blob = np.ascontiguousarray(mai_blob)
            self.h = L.mmo_init(self.blob.ctypes.data, self.blob.size, C.byref(self.p))
            if not self.h:
                raise RuntimeError("mmo_init failed")

    def close(self):
        if self.h:
            self.lib.mmo_destroy(self.h)
            self.h = None

    def extend(self, a, b, apos, bpos, brev=0, narrow=0, min_score=0):
        pa, pb = pad(a), pad(b)
        cap = 1 << 20
        res = np.zeros(16, dtype=np.uint32)
        out = np.zeros(cap, dtype=np.uint32)
        n = self.lib.mmo_extend(C.byref(self.p), _u8(pa[64:]), a.size, _u8(pb[64:]), b.size, apos, bpos, brev, narrow, min_score, _u32(res), _u32(out), cap)
        return res, out[:n].copy()

    def sketch(self, seq):
        p = pad(seq)
        cap = 4 * seq.size // 5 + 512
        out = np.zeros(cap, dtype=np.uint64)
        n = self.lib.mmo_sketch(self.h, _u8(p[64:]), seq.size, _u64(out), cap)
        assert n <= cap
        return out[:n]

    def get(self, minier, cap=4096):
        out = np.zeros(cap, dtype=np.uint64)
        n = self.lib.mmo_get(self.h, minier, _u64(out), cap)
        return out[:min(n, cap)], n

    def seed_chain(self, seq, rnd=0):
        p = pad(seq)
        cap = 1 << 20
        seeds = np.zeros(cap * 4, dtype=np.uint32)
        roots = np.zeros(cap * 2, dtype=np.uint32)
        nt, nr = C.c_uint64(0), C.c_uint64(0)
        ns = self.lib.mmo_seed_chain(self.h, _u8(p[64:]), seq.size, rnd, _u32(seeds), cap, C.byref(nt), _u32(roots), cap, C.byref(nr))
        return ns, seeds[: nt.value * 4].reshape(-1, 4).copy(), roots[: nr.value * 2].reshape(-1, 2).copy()

    def align(self, seq, qid=0):
        p = pad(seq)
        cap = 1 << 22
        out = np.zeros(cap, dtype=np.uint32)
        n = self.lib.mmo_align(self.h, _u8(p[64:]), seq.size, qid, _u32(out), cap)
        assert n <= cap
        return out[:n].copy()
