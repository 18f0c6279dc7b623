"""The oracle (oracle/*.c, CPU restatement) against golden vectors produced by the UNMODIFIED reference."""
import numpy as np
import pytest

import ora
from conftest import gold_pairs, unpack


def test_oracle_extend_matches_reference_golden(gold):
    for key, prm in (("pacbio", ora.PACBIO), ("ont", ora.ONT)):
        o = ora.Oracle(prm)
        pairs, res, alns = gold_pairs(gold["extend"], key)
        for p, r, a in zip(pairs, res, alns):
            r2, a2 = o.extend(*p[:6], p[6])
            assert np.array_equal(r, r2) and np.array_equal(a, a2)


def test_oracle_sketch_seed_chain_match_reference_golden(gold, oracle_params):
    o = ora.Oracle(oracle_params, gold["blob"])
    st = gold["stage"]
    sk = unpack(st["sketch"], st["sketch_ofs"])
    for rnd in (0, 2):
        seeds, roots = unpack(st[f"seed{rnd}"], st[f"seedo{rnd}"]), unpack(st[f"root{rnd}"], st[f"rooto{rnd}"])
        for i, s in enumerate(gold["enc"]):
            if s.size < 15:
                continue
            if rnd == 0:
                assert np.array_equal(o.sketch(s), sk[i])
            ns, sd, rt = o.seed_chain(s, rnd)
            assert ns == st[f"ns{rnd}"][i] and np.array_equal(sd.reshape(-1), seeds[i]) and np.array_equal(rt.reshape(-1), roots[i])
    o.close()


def test_oracle_align_matches_reference_golden(gold, oracle_params):
    """mm_align_seq for every golden read, in file order through ONE context (the reference's -t1 state carry-over)."""
    o = ora.Oracle(oracle_params, gold["blob"])
    n_mapped = 0
    for s, exp in zip(gold["enc"], gold["align"]):
        got = o.align(s)
        assert np.array_equal(got, exp)
        n_mapped += len(exp) > 0
    assert n_mapped > 50
    o.close()


def test_cigar_known_answers():
    """gaba.c:4297-4522 pins the CIGAR printer on literal path words."""
    import ctypes as C
    ora.build()
    L = C.CDLL(ora.SO)
    for fn in (L.ora_dump_cigar_forward, L.ora_dump_cigar_reverse):
        fn.restype = C.c_uint64
        fn.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64]

    def run(fn, words, ofs, ln):
        arr = (C.c_uint32 * (len(words) + 4))(*words, 0, 0, 0, 0)
        buf = C.create_string_buffer(256)
        fn(buf, arr, ofs, ln)
        return buf.value.decode()

    assert run(L.ora_dump_cigar_forward, [0x55555555], 0, 32) == "16M"
    assert run(L.ora_dump_cigar_forward, [0x55550555], 0, 32) == "6M4D8M"
    assert run(L.ora_dump_cigar_reverse, [0x55550555], 0, 32) == "8M4D6M"
    assert run(L.ora_dump_cigar_forward, [0x5555f555], 0, 32) == "6M4I8M"
