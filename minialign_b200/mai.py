""".mai index container reader (host side, Python): the "PG00" framed zlib stream of minialign.c:1135-1502 and the
payload header of minialign.c:3040-3167.  Returns the raw relocatable index blob that the C-ABI (`mab_init`) and the
oracle (`mmo_init`) take.  The native CLI has its own C loader (csrc/host_io.cpp); this one serves tests and bench."""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAI_MAGIC = 0x0849414D


def read_pg_stream(path: str) -> bytes:
    out = []
    with open(path, "rb") as f:
        while True:
            hdr = f.read(8)
            if len(hdr) < 8 or hdr[:4] != b"PG00":
                break
            (n,) = struct.unpack("<I", hdr[4:])
            if n == 0xFFFFFFFF:
                break
            out.append(zlib.decompress(f.read(n), 15))
    return b"".join(out)


def load_mai(path: str) -> np.ndarray:
    """First index block of a .mai file as a uint8 array (the bytes that follow the 12-byte magic+size header)."""
    raw = read_pg_stream(path)
    magic, size = struct.unpack_from("<IQ", raw, 0)
    if magic != MAI_MAGIC:
        raise ValueError("not a minialign v8 index")
    blob = np.frombuffer(raw, dtype=np.uint8, count=size, offset=12).copy()
    return blob


def parse_header(blob: np.ndarray) -> dict:
    b = blob[:64].tobytes()
    bkt, mask = struct.unpack_from("<QQ", b, 0)
    bb, w, k, n_occ = struct.unpack_from("<BBBB", b, 16)
    occ = list(struct.unpack_from("<7I", b, 20))
    n_seq, mono, s = struct.unpack_from("<IIQ", b, 48)
    return dict(bkt=bkt, mask=mask, b=bb, w=w, k=k, n_occ=n_occ, occ=occ, n_seq=n_seq, s=s)


def ref_seqs(blob: np.ndarray):
    """[(name, l_seq, seq_offset)] for every reference sequence in the blob."""
    h = parse_header(blob)
    out = []
    raw = blob.tobytes()
    for i in range(h["n_seq"]):
        seq, name, l_seq, l_name, circ = struct.unpack_from("<QQIHH", raw, h["s"] + 24 * i)
        out.append((raw[name:name + l_name].decode(), l_seq, seq))
    return out
