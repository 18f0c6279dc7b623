"""The emulation build once more with the vector types aligned like CUDA's (uint2: 8, uint4: 16 bytes) and the alignment sanitizer
on: a misaligned vector load / store traps on the GPU ("misaligned address") but goes unnoticed on x86, so the kernels are run
here through the text path (reader, mapper, post-processing, SAM printer) under -fsanitize=alignment."""
import os
import subprocess
import sys

from conftest import GOLD, ROOT


def test_text_path_under_alignment_sanitizer(tmp_path):
    so = str(tmp_path / "libmab_emu_san.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-pthread", "-shared", "-fsanitize=alignment", "-fno-sanitize-recover=alignment",
                           "-I" + os.path.join(ROOT, "tests/emu"), os.path.join(ROOT, "tests/emu/mab_emu.cpp"), os.path.join(ROOT, "tests/emu/cuda_emu.cpp"), "-o", so])
    ubsan = subprocess.check_output(["gcc", "-print-file-name=libubsan.so"], text=True).strip()
    code = f"""
import os, sys
sys.path.insert(0, {ROOT!r})
from minialign_b200 import api, mai
blob = mai.load_mai({os.path.join(GOLD, 'small.mai')!r})
lines = open({os.path.join(GOLD, 'reads.fa')!r}).read().split("\\n")
recs = [(lines[i], lines[i + 1]) for i in range(0, len(lines) - 1, 2) if len(lines[i + 1]) <= 4000][:50]
m = api.Mapper(blob, "pacbio", lib_path={so!r})
out = m.map_text("".join(h + "\\n" + s + "\\n" for h, s in recs).encode(), api.parse_tags("AS,XS,NM,MD,SA"))
assert out.count(b"\\n") >= len(recs)
m.close()
print("ok")
"""
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, LD_PRELOAD=ubsan), timeout=900)
    assert p.returncode == 0 and "runtime error" not in p.stderr and p.stdout.strip().endswith("ok"), p.stderr[-800:]
