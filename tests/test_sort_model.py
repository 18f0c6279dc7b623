"""The parallel form of the reference's unstable radix sort (mab_scalar.cuh, radix_sort_walk_warp).

1. The model it rests on -- one distribution pass of ksort.h:97-118 moves elements exactly as a walk over the *foreign*
   elements' destination digits says -- against the literal pass, in plain Python.
2. The device routine against the cycle-walking routine (itself pinned to the reference by the golden seed / SAM tests) on the
   CUDA-on-CPU shim: random keys with many ties, clustered keys (deep recursion), sorted and reversed input."""
import bisect
import random

import numpy as np
import pytest

from conftest import build_emu
from minialign_b200 import api


def exact_pass(a, dig):
    a = list(a)
    cnt = [0] * 256
    for x in a:
        cnt[dig(x)] += 1
    B = [0] * 257
    for k in range(256):
        B[k + 1] = B[k] + cnt[k]
    head, end = B[:256], B[1:]
    k = 0
    while k < 256:
        if head[k] != end[k]:
            l = dig(a[head[k]])
            if l != k:
                tmp = a[head[k]]
                while True:
                    swp = tmp
                    tmp = a[head[l]]
                    a[head[l]] = swp
                    head[l] += 1
                    l = dig(tmp)
                    if l == k:
                        break
                a[head[k]] = tmp
                head[k] += 1
            else:
                head[k] += 1
        else:
            k += 1
    return a


def model_pass(a, dig):
    n = len(a)
    d = [dig(x) for x in a]
    cnt = [0] * 256
    for x in d:
        cnt[x] += 1
    B = [0] * 257
    for k in range(256):
        B[k + 1] = B[k] + cnt[k]
    reg = [0] * n
    for k in range(256):
        for i in range(B[k], B[k + 1]):
            reg[i] = k
    fpos = [i for i in range(n) if d[i] != reg[i]]
    fdig = [d[i] for i in fpos]
    qs = [bisect.bisect_left(fpos, B[l]) for l in range(257)]
    qh, tl = [0] * 256, [0] * 256
    where = [None] * len(fpos)
    for k in range(256):
        tl[k] = qh[k]
        while qs[k] + qh[k] < qs[k + 1]:
            js = qs[k] + qh[k]
            qh[k] += 1
            cur = js
            while True:
                l = fdig[cur]
                if l == k:
                    where[cur] = ("close", js)
                    break
                j2 = qs[l] + qh[l]
                qh[l] += 1
                where[cur] = ("early", j2)
                cur = j2
    out = [None] * n
    fi = 0
    for i in range(n):
        if d[i] == reg[i]:
            r = reg[i]
            dst = i + (1 if tl[r] >= 1 and i < fpos[qs[r] + tl[r] - 1] else 0)
        else:
            kind, w = where[fi]
            fi += 1
            dst = fpos[w] if kind == "close" else (B[d[i]] if w == qs[d[i]] else fpos[w - 1] + 1)
        assert out[dst] is None
        out[dst] = a[i]
    return out


def test_walk_model_equals_literal_pass():
    rng = random.Random(1)
    for _ in range(1500):
        n = rng.choice([1, 2, 3, 5, 10, 50, 200, 700])
        nd = rng.choice([1, 2, 3, 8, 40, 256])
        a = [(rng.randrange(nd), i) for i in range(n)]
        if rng.random() < 0.3:
            a.sort(key=lambda x: x[0])
        if rng.random() < 0.3:
            for _ in range(n // 10 + 1):
                i, j = rng.randrange(n), rng.randrange(n)
                a[i], a[j] = a[j], a[i]
        assert exact_pass(a, lambda x: x[0]) == model_pass(a, lambda x: x[0])


def _cases():
    rng = np.random.default_rng(7)
    out = []
    for n, kind in [(1, "rand"), (2, "rand"), (64, "ties"), (65, "ties"), (300, "ties"), (1000, "rand"), (1000, "ties"), (2500, "cluster"),
                    (5000, "cluster"), (3000, "sorted"), (3000, "reversed"), (9000, "ties"), (20000, "cluster"), (32767, "rand")]:
        e = np.zeros((n, 4), dtype=np.uint32)
        if kind == "rand":
            e[:, 0] = rng.integers(0, 1 << 32, n, dtype=np.uint64)
            e[:, 1] = rng.integers(0, 3, n)
        elif kind == "ties":
            e[:, 0] = rng.integers(0, max(2, n // 4), n) * 977
            e[:, 1] = rng.integers(0, 2, n)
        elif kind == "cluster":                      # most keys share the upper digits: deep recursion, big sub-frames
            e[:, 0] = 0x40000000 + rng.integers(0, 40000, n) // rng.integers(1, 4, n)
            e[: n // 5, 0] = rng.integers(0, 1 << 31, n // 5)
            e[:, 1] = rng.integers(0, 2, n) * (rng.random(n) < 0.1)
        else:
            e[:, 0] = np.sort(rng.integers(0, 1 << 20, n))
            if kind == "reversed":
                e[:, 0] = e[::-1, 0]
        e[:, 2] = np.arange(n)                       # payload: tells equal keys apart
        e[:, 3] = 0x7fffffff
        out.append((n, kind, e))
    return out


def test_emu_parallel_sort_equals_cycle_walking_sort(gold):
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    for n, kind, e in _cases():
        a, b = m.sort_check(e)
        key = lambda x: (x[:, 1].astype(np.uint64) << 32) | x[:, 0]
        assert np.all(np.diff(key(a).astype(np.int64) >> 1) >= 0) or np.all(key(a)[1:] >= key(a)[:-1]), (n, kind)
        assert np.array_equal(a, b), (n, kind, int(np.argmax(np.any(a != b, axis=1))))
    m.close()


@pytest.mark.gpu
def test_gpu_parallel_sort_equals_cycle_walking_sort(gold):
    m = api.Mapper(gold["blob"], "pacbio")
    for n, kind, e in _cases():
        a, b = m.sort_check(e)
        assert np.array_equal(a, b), (n, kind, int(np.argmax(np.any(a != b, axis=1))))
    m.close()


def test_emu_wide_chain_scan_equals_single_lane(gold, monkeypatch):
    """the A/B switches -- window scan 32 candidates at a time, the fused k_sortchain, k_chain on a shared-memory copy -- give the default's chains"""
    so = build_emu()
    idx = [i for i, s in enumerate(gold["enc"]) if s.size <= 12000][:12]
    m = api.Mapper(gold["blob"], "pacbio", lib_path=so)
    exp = [m.seed_chain(gold["enc"][i], 2) for i in idx]
    m.close()
    for var, val in (("MAB_CHAIN_WARP", "1"), ("MAB_SORT_WALK", "0"), ("MAB_CHAIN_STAGED", "1"), ("MAB_CLASS_STREAMS", "0")):
        monkeypatch.setenv(var, val)
        m = api.Mapper(gold["blob"], "pacbio", lib_path=so)
        got = [m.seed_chain(gold["enc"][i], 2) for i in idx]
        m.close()
        monkeypatch.delenv(var)
        assert sum(e[0] for e in exp) > 1000
        for e, g in zip(exp, got):
            assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2])
