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


def inflate_mai(path: str, on_size=None, on_piece=None, threads: int | None = None) -> np.ndarray:
    """The index payload of a .mai file, its frames inflated by a pool of threads (zlib releases the GIL) straight into one array.
    on_size(payload_bytes) is called once the header frame is read, on_piece(payload_offset, address, n_bytes) for every frame as
    soon as it is in place (from the thread that inflated it): a caller forwards the pieces to the GPU while the other frames are
    still in the works (api.Mapper.from_mai).  The native CLI does the same in C++ (csrc/host/mab_cli.cpp, load_mai)."""
    import mmap
    import os
    from concurrent.futures import ThreadPoolExecutor
    BS = 1 << 20
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    frames, p = [], 0
    while p + 8 <= len(mm) and mm[p:p + 4] == b"PG00":
        (n,) = struct.unpack_from("<I", mm, p + 4)
        if n in (0, 0xFFFFFFFF) or p + 8 + n > len(mm):
            break
        frames.append((p + 8, n))
        p += 8 + n
    if not frames:
        raise ValueError("not a minialign index container")
    raw = np.empty(len(frames) * BS + 64, dtype=np.uint8)
    first = zlib.decompress(mm[frames[0][0]:frames[0][0] + frames[0][1]], 15)
    if len(first) < 12:
        raise ValueError("not a minialign v8 index")
    magic, size = struct.unpack_from("<IQ", first, 0)
    if magic != MAI_MAGIC or size + 12 > len(frames) * BS:
        raise ValueError("not a minialign v8 index")
    lens = [0] * len(frames)

    def place(i, data):
        if i + 1 < len(frames) and len(data) != BS:
            raise ValueError("not the frame layout the reference writes")
        raw[i * BS:i * BS + len(data)] = np.frombuffer(data, dtype=np.uint8)
        lens[i] = len(data)
        lo, hi = max(i * BS, 12), min(i * BS + len(data), 12 + size)
        if on_piece is not None and lo < hi:
            on_piece(lo - 12, raw.ctypes.data + lo, hi - lo)

    if on_size is not None:
        on_size(size)
    place(0, first)
    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda i: place(i, zlib.decompress(mm[frames[i][0]:frames[i][0] + frames[i][1]], 15)), range(1, len(frames))))
    if (len(frames) - 1) * BS + lens[-1] < 12 + size:
        raise ValueError("truncated index")
    return raw[12:12 + size]


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
