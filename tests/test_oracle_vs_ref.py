"""Live differential tests: oracle vs the unmodified reference compiled into oracle/_ref (skipped where it is absent)."""
import numpy as np
import pytest

import ora
import refh
from minialign_b200 import synth

pytestmark = pytest.mark.skipif(not refh.available(), reason="oracle/_ref/libref_harness.so not built (needs /root/reference)")


@pytest.mark.parametrize("args,prm,n", [(("-xpacbio",), ora.PACBIO, 300), (("-xont.1dsq",), ora.ONT, 300)] + [(a, p, 120) for p, a in ora.CUSTOM])
def test_extend_fuzz(gold, args, prm, n):
    h = refh.RefHarness(gold["mai"], args=args)
    o = ora.Oracle(prm)
    rng = np.random.default_rng(1234)
    for it in range(n):
        L = max(2, int(rng.choice([3, 9, 33, 64, 65, 100, 400, 1200])) + int(rng.integers(0, 20)))
        a = rng.integers(0, 4, size=L).astype(np.uint8)
        b = synth.encode_2bit(synth._mutate(np.frombuffer(b"ACGT", dtype=np.uint8)[a], float(rng.choice([1.0, 0.9, 0.8, 0.6])), rng))
        if b.size < 2:
            continue
        brev = int(rng.integers(0, 2))
        if brev:
            b = (3 - b[::-1]).astype(np.uint8)
        apos, bpos = int(rng.integers(0, min(a.size, 40))), int(rng.integers(0, min(b.size, 40)))
        narrow = int(rng.integers(0, 3))
        r1, a1 = h.extend(a, b, apos, bpos, brev, narrow, 0)
        r2, a2 = o.extend(a, b, apos, bpos, brev, narrow, 0)
        assert np.array_equal(r1, r2) and np.array_equal(a1, a2), (it, L, apos, bpos, brev, narrow)
    h.close()


def test_align_state_carry_over(gold, oracle_params):
    """Reads processed in a different order than the golden file: the stale-rlen dependency must follow the new order."""
    h = refh.RefHarness(gold["mai"])
    o = ora.Oracle(oracle_params, gold["blob"])
    order = np.random.default_rng(5).permutation(len(gold["enc"]))[:60]
    for i in order:
        assert np.array_equal(h.align(gold["enc"][i]), o.align(gold["enc"][i]))
    h.close(); o.close()
