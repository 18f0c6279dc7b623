"""Measure the DP-step ceiling (k_fill_peak) and one solo k_extend batch for each library variant given on the command line."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from minialign_b200 import api, mai, synth
import bench

work = "/tmp/mab_bench_v"
g, idx, blob, batches = bench.build_workload(work, 1, 8192)
block, ofs, lens = api.pack_reads([synth.encode_2bit(r) for _, r in batches[0]])
for lib in sys.argv[1:]:
    m = api.Mapper(blob, "pacbio", lib_path=lib)
    pk = [m.fill_peak(True, 3000) for _ in range(2)][-1]
    pu = m.fill_peak(False, 3000)
    ext = []
    for _ in range(3):
        m.map_packed(block.ctypes.data, block.size, ofs, lens)
        st = m.stats(); m.lib.mab_release_batch(m.h)
        ext.append(st["ms_extend_r0"])
    print(f"{os.path.basename(lib)}: peak masked {64*pk/1e9:.1f} GCUPS, unmasked {64*pu/1e9:.1f} GCUPS, k_extend {min(ext):.2f} ms, sortchain {st['ms_sortchain']:.2f} ms, vec {st['n_vectors']}", flush=True)
    m.close()
