import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location("gpu_dbg", os.path.join(ROOT, "scripts/gpu_dbg.py")); dbg = importlib.util.module_from_spec(spec); spec.loader.exec_module(dbg)
from minialign_b200 import mai, api
work = dbg.setup()
blob = mai.load_mai(f"{work}/g.mai")
lib = sys.argv[1]
m = api.Mapper(blob, "pacbio", lib_path=lib)
p = dbg.mk_pairs(3, 7, 150)[0]
(r2, o2), = m.extend_pairs([p])
print("res", r2)
