"""The text path (FASTA / FASTQ bytes in, SAM bytes out: device-side reader, post-processing and SAM printer) on the
CUDA-on-CPU shim, against (a) the host formatter fed with the record-level results of the same reads -- itself pinned to the
reference CLI's golden SAM in test_sam.py -- and (b) the reference's golden SAM directly."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLD, build_emu
from minialign_b200 import api, mai
from test_sam import Read, Ref, fmt, sam  # noqa: F401  (fixture)


def _refs(blob):
    rs = mai.ref_seqs(blob)
    refs = (Ref * len(rs))()
    keep = []
    for i, (name, l_seq, ofs) in enumerate(rs):
        keep.append(name.encode())
        refs[i] = Ref(keep[-1], len(name), l_seq, blob.ctypes.data + ofs)
    return refs, keep


def _fasta(reads, width=0):
    out = []
    for name, seq in reads:
        b = seq.tobytes()
        out.append(b">" + name.encode() + b"\n" + (b"\n".join(b[k:k + width] for k in range(0, len(b), width)) if width else b) + b"\n")
    return b"".join(out)


def _expected(sam, gold, idx, taglist):
    """record-level results through the emulated kernels + the host formatter"""
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    res = m.map_batch([gold["enc"][i] for i in idx])
    m.close()
    refs, _keep = _refs(gold["blob"])
    tags = sam.mab_sam_parse_tags(taglist.encode())
    return "".join(fmt(sam, refs, gold["reads"][i][0], gold["enc"][i], w, tags) for i, w in zip(idx, res))


@pytest.mark.parametrize("taglist,width", [("", 0), ("AS,XS,NM,MD,NH,IH", 60), ("AS,NM,MD,SA", 0)])
def test_emu_text_path_matches_host_formatter(sam, gold, taglist, width):
    idx = [i for i, s in enumerate(gold["enc"]) if s.size <= 5000][:40]
    exp = _expected(sam, gold, idx, taglist)
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    got = m.map_text(_fasta([gold["reads"][i] for i in idx], width), api.parse_tags(taglist)).decode()
    m.close()
    assert got.count("\n") == exp.count("\n") and sum(1 for l in exp.split("\n") if l and l.split("\t")[1] != "4") > 10
    for a, b in zip(got.split("\n"), exp.split("\n")):
        assert a == b


def test_emu_text_path_prefix_of_golden_sam(gold):
    """the first reads of the golden file, in file order: the device text equals the reference CLI's lines"""
    n = 0
    while n < len(gold["enc"]) and sum(s.size for s in gold["enc"][:n + 1]) <= 60000:
        n += 1
    assert n >= 8
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    got = m.map_text(_fasta(gold["reads"][:n])).decode().split("\n")
    m.close()
    exp = [l for l in open(os.path.join(GOLD, "golden_pacbio.sam")).read().split("\n") if l and not l.startswith("@")]
    assert got[-1] == "" and len(got) - 1 >= n and got[:-1] == exp[:len(got) - 1]
    assert exp[len(got) - 1].split("\t")[0] == gold["reads"][n][0]            # ... and they are exactly the lines of the first n reads


def test_emu_text_fastq_and_edge_records(sam, gold):
    """FASTQ with and without -Q, an empty record (dropped by the reader), a read below k, names with comments / leading
    spaces / a tab, no newline at the end of the chunk."""
    idx = [i for i, s in enumerate(gold["enc"]) if 200 <= s.size <= 2500][:6]
    recs = []
    for j, i in enumerate(idx):
        name, seq = gold["reads"][i]
        q = bytes(33 + (k * 7 + j) % 40 for k in range(seq.size))
        recs.append((name, seq.tobytes(), q))
    text = b""
    for j, (name, s, q) in enumerate(recs):
        hdr = name.encode() + (b" some comment" if j == 1 else b"")
        text += b"@" + (b"  " if j == 2 else b"") + hdr + b"\n" + s + b"\n+\n" + q + b"\n"
    text += b"@empty\n\n+\n\n" + b"@tiny\nACGT\n+\nIIII"                       # dropped record; read below k; no final newline
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    res = m.map_batch([gold["enc"][i] for i in idx] + [np.array([0, 1, 2, 3], dtype=np.uint8)])
    refs, _keep = _refs(gold["blob"])
    for keep_qual in (False, True):
        exp = ""
        for (name, s, q), w, i in zip(recs + [("tiny", b"ACGT", b"IIII")], res, idx + [None]):
            enc = gold["enc"][i] if i is not None else np.array([0, 1, 2, 3], dtype=np.uint8)
            rd = Read(name.encode(), len(name), enc.ctypes.data, enc.size, q if keep_qual else None)
            w = np.ascontiguousarray(w, dtype=np.uint32)
            n = C.c_uint64(0)
            p = sam.mab_sam_format_c(refs, len(refs), C.byref(rd), w.ctypes.data_as(C.POINTER(C.c_uint32)), w.size, 0, C.byref(n))
            exp += C.string_at(p, n.value).decode(); sam.mab_sam_free(p)
        m2 = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
        got = m2.map_text(text, 0, keep_qual).decode()
        assert m2.last_info.n_reads == len(recs) + 1
        m2.close()
        assert got == exp
    m.close()


def test_emu_text_rejects_what_the_device_reader_does_not_take(gold):
    m = api.Mapper(gold["blob"], "pacbio", lib_path=build_emu())
    for bad in (b"garbage\n>r\nACGT\n", b"@r\nAC\nGT\n+\nII\nII\n", b"\n>r\nACGT\n"):
        with pytest.raises(RuntimeError):
            m.map_text(bad)
    assert m.map_text(b"") == b""
    m.close()


def test_emu_text_chunks_on_several_contexts_chain_rlen(tmp_path):
    """Chunks mapped by different contexts without knowing what the previous chunk left behind (begin, rlen unknown), then
    committed in file order: the concatenated SAM equals one context mapping the whole file sequentially, which equals the
    oracle's -t1 results printed by the host formatter.  Uneven contigs make the first-seed test flip often."""
    import ora
    from minialign_b200 import synth
    from test_emu_kernels import _uneven_genome_index
    from test_sam import SO, SRC
    import subprocess
    g, blob = _uneven_genome_index(tmp_path, seed=71)
    hdr = mai.parse_header(blob)
    reads = synth.make_reads(g, 60_000, seed=72, len_mean=800, len_sd=300, len_min=60, len_max=2000)
    enc = [synth.encode_2bit(r) for _, r in reads]
    if not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.mab_sam_format_c.restype = C.c_void_p
    L.mab_sam_format_c.argtypes = [C.POINTER(Ref), C.c_uint32, C.POINTER(Read), C.POINTER(C.c_uint32), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
    L.mab_sam_free.argtypes = [C.c_void_p]
    o = ora.Oracle(dict(ora.PACBIO, occ=hdr["occ"][:3]), blob)
    refs, _keep = _refs(blob)
    exp = "".join(fmt(L, refs, nm, e, o.align(e), 0) for (nm, _), e in zip(reads, enc))
    o.close()
    one = api.Mapper(blob, "pacbio", lib_path=build_emu())
    whole = one.map_text(_fasta(reads)).decode()
    one.close()
    assert whole == exp
    cuts = [0, len(reads) // 3, len(reads) // 3 + 7, len(reads)]
    ms, bufs = [], []
    for k in range(3):
        m = api.Mapper(blob, "pacbio", lib_path=build_emu())
        t = _fasta(reads[cuts[k]:cuts[k + 1]])
        buf = C.create_string_buffer(t, len(t))
        m.text_begin(C.addressof(buf), len(t), 0, 0, False)
        ms.append(m); bufs.append(buf)
    rlen, out, redone = 0, "", 0
    for m in ms:
        info = m.text_commit(rlen)
        redone += m.stats()["n_retry"]
        if info.rlen_valid:
            rlen = info.rlen_next
        info, ptr = m.text_finish()
        out += C.string_at(ptr, info.sam_bytes).decode()
        m.close()
    assert out == exp


def test_emu_staged_load_and_reserve(gold):
    """mab_load_begin / put / end with the pieces out of order gives the context mab_init gives; with a chunk size announced
    (mab_text_reserve) a short chunk followed by a long one maps like without it"""
    so = build_emu()
    n = 0
    while n < len(gold["enc"]) and sum(s.size for s in gold["enc"][:n + 1]) <= 40000:
        n += 1
    short, long_ = _fasta(gold["reads"][:3]), _fasta(gold["reads"][:n])
    m0 = api.Mapper(gold["blob"], "pacbio", lib_path=so)
    exp = [m0.map_text(short), m0.map_text(long_)]
    m0.close()
    (m1,) = api.Mapper.staged(gold["blob"], "pacbio", piece=70001, order=lambda st: st[::-1], lib_path=so)
    m1.text_reserve(len(long_) + 100)
    c = m1.clone()
    got = [m1.map_text(short), m1.map_text(long_)]
    assert got == exp
    assert c.map_text(long_) == api.Mapper(gold["blob"], "pacbio", lib_path=so).map_text(long_)
    c.close(); m1.close()


def test_emu_mapper_from_mai_file(tmp_path, gold):
    """mai.inflate_mai (frames on a thread pool, pieces handed to the staged set-up as they appear) gives load_mai's payload, and
    Mapper.from_mai a context that maps like one from mab_init"""
    import subprocess
    from conftest import ROOT
    from minialign_b200 import mai, synth
    cli = os.path.join(ROOT, "minialign_b200", "minialign-b200")
    g = synth.make_genome(2_300_000, 3, seed=5)           # > 2 frames of 1 MiB
    fa, idx = str(tmp_path / "g.fa"), str(tmp_path / "g.mai")
    synth.write_fasta(fa, g, 80)
    subprocess.check_call([cli, "-xpacbio", "-d", idx, fa], stderr=subprocess.DEVNULL)
    a, pieces = mai.load_mai(idx), []
    b = mai.inflate_mai(idx, None, lambda off, addr, n: pieces.append((off, n)))
    assert np.array_equal(a, b) and sum(n for _, n in pieces) == a.size and len(pieces) >= 3
    reads = synth.make_reads(g, 60_000, seed=6)
    text = _fasta(reads)
    so = build_emu()
    m0 = api.Mapper(a, "pacbio", lib_path=so)
    m1 = api.Mapper.from_mai(idx, "pacbio", lib_path=so)
    assert m0.map_text(text) == m1.map_text(text)
    m0.close(); m1.close()
