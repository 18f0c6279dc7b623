"""The product's host-side SAM writer (minialign_b200/csrc/host/mab_sam.cpp) against the reference CLI's own output
(tests/golden/golden_pacbio.sam and golden_tags.sam, produced by `minialign -xpacbio -t1 [-TAS,XS,NM,MD,NH,IH]`).
Inputs are the reference's per-read results (golden_align.npz), so this isolates the text formatting: CIGAR from the path
bit string, clipping, flags, MAPQ, AS/XS/NM/NH/IH and the MD walk."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT
from minialign_b200 import mai

SO = os.path.join(ROOT, "minialign_b200", "libmab_sam.so")
SRC = os.path.join(ROOT, "minialign_b200", "csrc", "host", "mab_sam.cpp")


class Ref(C.Structure):
    _fields_ = [("name", C.c_char_p), ("l_name", C.c_uint32), ("l_seq", C.c_uint32), ("seq", C.c_void_p)]


class Read(C.Structure):
    _fields_ = [("name", C.c_char_p), ("l_name", C.c_uint32), ("seq", C.c_void_p), ("l_seq", C.c_uint32), ("qual", C.c_char_p)]


@pytest.fixture(scope="module")
def sam():
    if not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.mab_sam_format_c.restype = C.c_void_p
    L.mab_sam_format_c.argtypes = [C.POINTER(Ref), C.c_uint32, C.POINTER(Read), C.POINTER(C.c_uint32), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
    L.mab_sam_free.argtypes = [C.c_void_p]
    L.mab_sam_parse_tags.restype = C.c_uint32
    L.mab_sam_parse_tags.argtypes = [C.c_char_p]
    return L


def fmt(L, refs, name, seq, words, tags):
    rd = Read(name.encode(), len(name), seq.ctypes.data, seq.size, None)
    w = np.ascontiguousarray(words, dtype=np.uint32)
    n = C.c_uint64(0)
    p = L.mab_sam_format_c(refs, len(refs), C.byref(rd), w.ctypes.data_as(C.POINTER(C.c_uint32)), w.size, tags, C.byref(n))
    s = C.string_at(p, n.value).decode()
    L.mab_sam_free(p)
    return s


@pytest.mark.parametrize("golden,taglist", [("golden_pacbio.sam", ""), ("golden_tags.sam", "AS,XS,NM,MD,NH,IH")])
def test_sam_records_match_reference_cli(sam, gold, golden, taglist):
    blob = gold["blob"]
    rs = mai.ref_seqs(blob)
    refs = (Ref * len(rs))()
    keep = []
    for i, (name, l_seq, ofs) in enumerate(rs):
        keep.append(name.encode())
        refs[i] = Ref(keep[-1], len(name), l_seq, blob.ctypes.data + ofs)
    tags = sam.mab_sam_parse_tags(taglist.encode())
    got = "".join(fmt(sam, refs, nm, gold["enc"][i], gold["align"][i], tags) for i, (nm, _) in enumerate(gold["reads"]))
    exp = "".join(l for l in open(os.path.join(GOLD, golden)) if not l.startswith("@"))
    gl, el = got.split("\n"), exp.split("\n")
    assert len(gl) == len(el)
    for a, b in zip(gl, el):
        assert a == b
