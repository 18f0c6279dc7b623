import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libmab_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def build_emu():
    """g++ build of the product's device + host sources against the CUDA-on-CPU shim (tests/emu)."""
    srcs = [os.path.join(ROOT, "tests/emu", f) for f in ("mab_emu.cpp", "cuda_emu.cpp", "cuda_emu.h")]
    srcs += [os.path.join(ROOT, "minialign_b200/csrc", f) for f in sorted(f for f in os.listdir(os.path.join(ROOT, "minialign_b200/csrc")) if f.endswith((".inl", ".cuh", ".h")))]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-pthread", "-shared", "-I" + os.path.join(ROOT, "tests/emu"),
                               os.path.join(ROOT, "tests/emu/mab_emu.cpp"), os.path.join(ROOT, "tests/emu/cuda_emu.cpp"), "-o", EMU_SO])
    return EMU_SO


EMU_CLI = os.path.join(ROOT, "tests", "emu", "minialign-emu")


def build_emu_cli():
    """the product's command line (host pipeline: reader, contexts, rlen chain, writers) linked against the emulation build"""
    so = build_emu()
    host = os.path.join(ROOT, "minialign_b200/csrc/host")
    srcs = [os.path.join(host, f) for f in os.listdir(host)] + [so, os.path.join(ROOT, "include/minialign_b200.h")]
    if not os.path.exists(EMU_CLI) or any(os.path.getmtime(s) > os.path.getmtime(EMU_CLI) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EMU_CLI, *[os.path.join(host, f) for f in ("mab_cli.cpp", "mab_sam.cpp", "mab_index.cpp")],
                               "-L" + os.path.join(ROOT, "tests/emu"), "-lmab_emu", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN"])
    return EMU_CLI


def unpack(words, ofs):
    return [words[ofs[i]:ofs[i + 1]] for i in range(len(ofs) - 1)]


@pytest.fixture(scope="session")
def gold():
    """Golden fixtures generated from the unmodified reference by tests/golden/make_golden.py."""
    from minialign_b200 import mai, synth
    blob = mai.load_mai(os.path.join(GOLD, "small.mai"))
    reads = []
    with open(os.path.join(GOLD, "reads.fa"), "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 1, 2):
        reads.append((lines[i][1:].decode(), np.frombuffer(lines[i + 1], dtype=np.uint8)))
    enc = [synth.encode_2bit(r) for _, r in reads]
    al = np.load(os.path.join(GOLD, "golden_align.npz"))
    st = np.load(os.path.join(GOLD, "golden_stage.npz"))
    ex = np.load(os.path.join(GOLD, "golden_extend.npz"))
    return dict(blob=blob, hdr=mai.parse_header(blob), reads=reads, enc=enc, align=unpack(al["words"], al["ofs"]), stage=st, extend=ex,
                mai=os.path.join(GOLD, "small.mai"))


def gold_pairs(ex, key):
    a, b = unpack(ex[f"{key}_a"], ex[f"{key}_ao"]), unpack(ex[f"{key}_b"], ex[f"{key}_bo"])
    alns = unpack(ex[f"{key}_aln"], ex[f"{key}_alno"])
    return [(a[i], b[i], *[int(x) for x in ex[f"{key}_args"][i]], 0) for i in range(len(a))], ex[f"{key}_res"], alns


@pytest.fixture(scope="session")
def oracle_params(gold):
    import ora
    return dict(ora.PACBIO, occ=gold["hdr"]["occ"][:3])
