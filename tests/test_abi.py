"""The C-ABI library: builds for sm_100a, loads without a GPU, exports every symbol include/minialign_b200.h declares, and
fails loudly (no CPU fallback) when no device is usable."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from minialign_b200 import api

HDR = os.path.join(ROOT, "include", "minialign_b200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mab_[a-z_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return api.load_library()


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s


def test_cuda_library_has_sm100a_code():
    if not os.path.exists(api.LIB_PATH):
        pytest.skip("library not built")
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(lib, gold):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = api.make_params(api.PRESETS["pacbio"])
    h = lib.mab_init(gold["blob"].ctypes.data, gold["blob"].size, C.byref(p), 0)
    assert not h
    assert b"CUDA" in lib.mab_last_error() or b"device" in lib.mab_last_error()


def test_rejects_unsupported_params(gold):
    bad = dict(api.PRESETS["pacbio"], gfa=0, gfb=0)          # affine model: outside the combined-gap path
    lib = api.load_library()
    p = api.make_params(bad)
    assert not lib.mab_init(gold["blob"].ctypes.data, gold["blob"].size, C.byref(p), 0)
    assert b"unsupported" in lib.mab_last_error()
