"""ctypes wrapper around oracle/_ref/libref_harness.so (the UNMODIFIED reference compiled by oracle/Makefile.ref).

Test infrastructure only.  Present in this container and (as a prebuilt .so) on the GPU box; tests that need it skip
when it is absent.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
BIN = os.path.join(ROOT, "oracle", "_ref", "minialign")


def available() -> bool:
    return os.path.exists(SO)


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _u64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def pad(seq: np.ndarray, margin: int = 64) -> np.ndarray:
    """The reference's readers keep 64-byte zero margins around every sequence (minialign.c:2109-2146)."""
    buf = np.zeros(seq.size + 2 * margin, dtype=np.uint8)
    buf[margin:margin + seq.size] = seq
    return buf


class RefHarness:
    def __init__(self, mai_path: str, args=("-xpacbio",)):
        self.lib = C.CDLL(SO)
        L = self.lib
        L.refh_open.restype = C.c_void_p
        L.refh_open.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.refh_close.argtypes = [C.c_void_p]
        L.refh_params.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.refh_sketch.restype = C.c_uint64
        L.refh_sketch.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint64), C.c_uint64]
        L.refh_get.restype = C.c_uint32
        L.refh_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32]
        L.refh_seed_chain.restype = C.c_uint64
        L.refh_seed_chain.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_uint64)]
        L.refh_align.restype = C.c_uint64
        L.refh_align.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint64]
        L.refh_extend.restype = C.c_uint64
        L.refh_extend.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint64]
        argv = [b"minialign", *[a.encode() for a in args], mai_path.encode()]
        arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
        self.h = L.refh_open(len(argv), arr)
        if not self.h:
            raise RuntimeError("refh_open failed")

    def close(self):
        if self.h:
            self.lib.refh_close(self.h)
            self.h = None

    def params(self):
        out = (C.c_int32 * 48)()
        self.lib.refh_params(self.h, out)
        v = list(out)
        return dict(k=v[0], w=v[1], b=v[2], n_occ=v[3], occ=v[4:8], wlen=v[8], glen=v[9], min_score=v[10],
                    min_ratio=struct.unpack("f", struct.pack("i", v[11]))[0],
                    gi=v[12], ge=v[13], gfa=v[14], gfb=v[15], xdrop=v[16], score_matrix=v[17:33])

    def sketch(self, seq: np.ndarray) -> np.ndarray:
        p = pad(seq)
        cap = 4 * seq.size // 5 + 512
        out = np.zeros(cap, dtype=np.uint64)
        n = self.lib.refh_sketch(self.h, _u8(p[64:]), seq.size, _u64(out), cap)
        assert n <= cap
        return out[:n]

    def get(self, minier: int, cap: int = 4096) -> np.ndarray:
        out = np.zeros(cap, dtype=np.uint64)
        n = self.lib.refh_get(self.h, minier, _u64(out), cap)
        return out[:min(n, cap)], n

    def seed_chain(self, seq: np.ndarray, rnd: int = 0):
        p = pad(seq)
        cap = 1 << 20
        seeds = np.zeros(cap * 4, dtype=np.uint32)
        roots = np.zeros(cap * 2, dtype=np.uint32)
        nt = C.c_uint64(0)
        nr = C.c_uint64(0)
        ns = self.lib.refh_seed_chain(self.h, _u8(p[64:]), seq.size, rnd, _u32(seeds), cap, C.byref(nt), _u32(roots), cap, C.byref(nr))
        return ns, seeds[: nt.value * 4].reshape(-1, 4).copy(), roots[: nr.value * 2].reshape(-1, 2).copy()

    def align(self, seq: np.ndarray, qid: int = 0) -> np.ndarray:
        p = pad(seq)
        cap = 1 << 22
        out = np.zeros(cap, dtype=np.uint32)
        n = self.lib.refh_align(self.h, _u8(p[64:]), seq.size, qid, _u32(out), cap)
        assert n <= cap
        return out[:n].copy()

    def extend(self, a: np.ndarray, b: np.ndarray, apos: int, bpos: int, brev: int = 0, narrow: int = 0, min_score: int = 0):
        pa, pb = pad(a), pad(b)
        cap = 1 << 20
        res = np.zeros(16, dtype=np.uint32)
        out = np.zeros(cap, dtype=np.uint32)
        n = self.lib.refh_extend(self.h, _u8(pa[64:]), a.size, _u8(pb[64:]), b.size, apos, bpos, brev, narrow, min_score, _u32(res), _u32(out), cap)
        return res, out[:n].copy()
