"""One context, a few full-size chunks through the text path (for `ncu --metrics gpu__time_duration.sum` launch lists and
stage timings): python scripts/gpu_text_once.py [n_chunks] [batch_reads]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from minialign_b200 import api

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
which = sys.argv[3] if len(sys.argv) > 3 else "ecoli"
g, idx, blob = bench.build_genome("/tmp/mab_bench/once", which)
reads = bench.chunk_reads(g, 0, batch)
text = bench.fasta_bytes(reads, 0)
p = api.load_library().mab_host_alloc(len(text) + 64)
C.memmove(p, text, len(text))
m = api.Mapper(blob, "pacbio")
info = api.MabTextInfo()
for i in range(n_chunks):
    t0 = time.time()
    rc = m.lib.mab_map_text(m.h, p, len(text), api.parse_tags(os.environ.get("TAGS", "")), None, 0, None, C.byref(info))
    assert rc == 0, m.lib.mab_last_error()
    st = m.stats()
    print(f"chunk {i}: wall {1e3 * (time.time() - t0):.1f} ms  " + " ".join(f"{k}={v:.2f}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()), flush=True)
print("sam bytes", info.sam_bytes, "reads", info.n_reads, "bases", info.n_bases)
m.close()
