"""Small search covering all first-stage modes, for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oswald_b200 as ob
from oswald_b200 import capi

rng = np.random.default_rng(1)
aa = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)
lens = rng.integers(1, 400, size=300).astype(np.uint64)
lens[:3] = (5000, 2500, 1)
db = ob.Database.from_lengths(lens, aa[rng.integers(0, 20, size=int(lens.sum()))])
with ob.Searcher(1) as s:
    s.load_db(db, max_chunk_residues=512)
    for ql in ([60], [144, 189], [700, 1500], [1400, 1350, 90]):
        q = ob.Queries.from_list([aa[rng.integers(0, 20, size=m)] for m in ql])
        for mask in (capi.OSW_K_DEFAULT, capi.OSW_K_DEFAULT | capi.OSW_K_TWO_TRACK, capi.OSW_K_DEFAULT | capi.OSW_K_PAIR_DB, capi.OSW_K_DEFAULT | capi.OSW_K_TRANSPOSED, capi.OSW_K_I32):
            s.set_kernels(mask)
            hits, tm = s.search(q, ob.matrix("blosum62"), 10, 2, top=5)
print("sanitize case done")
