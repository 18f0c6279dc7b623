import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("gpu_dbg", os.path.join(ROOT, "scripts/gpu_dbg.py")); dbg = importlib.util.module_from_spec(spec); spec.loader.exec_module(dbg)
from minialign_b200 import mai, api
import ora
work = dbg.setup()
blob = mai.load_mai(f"{work}/g.mai")
g = api.Mapper(blob, "pacbio").selftest()
e = api.Mapper(blob, "pacbio", lib_path=os.path.join(ROOT, "tests/emu/libmab_emu.so")).selftest()
n = int(e[63, 0]); print("functions", n, int(g[63, 0]))
for f in range(n):
    if not np.array_equal(g[f], e[f]):
        print("SELFTEST MISMATCH fn", f, "gpu", [hex(v) for v in g[f][:4]], "emu", [hex(v) for v in e[f][:4]])
print("selftest done")
# detail of one pair
hd = mai.parse_header(blob); o = ora.Oracle(dict(ora.PACBIO, occ=hd["occ"][:3]), blob)
m = api.Mapper(blob, "pacbio")
for p in dbg.mk_pairs(3, 7, 150):
    (r2, o2), = m.extend_pairs([p]); r1, o1 = o.extend(*p[:6], p[6])
    print("pair", p[0].size, p[1].size, p[2:6]); print(" ora", r1, o1[:12]); print(" gpu", r2, o2[:12])
