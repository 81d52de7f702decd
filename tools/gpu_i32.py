#!/usr/bin/env python3
"""Throughput of the 32-bit stage (sw_i32_kernel) on one B200.   usage (under gpurun): python tools/gpu_i32.py
 (a) as a complete scorer (kernel mask OSW_K_I32: every pair at 32 bit) on a 20 000-sequence slice of the
     config-2 database with the 20 standard queries;
 (b) as the re-score of a stress workload in which more than 1 % of all pairs overflow 16 bits: tryptophan-rich
     queries and database sequences under PAM30 9/1; every score checked against the other kernel / the oracle on
     a sample."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import oswald_b200 as ob
from oswald_b200 import capi


def main():
    out = []
    qs = bench.make_queries()
    with ob.Searcher(1) as s:
        # (a) everything at 32 bit
        wl = bench.make_workload(2, n_override=20000)
        db, q = wl["db"], wl["queries"]
        s.load_db(db)
        s.set_kernels(capi.OSW_K_I32)
        s.search(q, ob.matrix("blosum62"), 10, 2, top=10)
        h32, tm, s32 = s.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
        s.set_kernels(capi.OSW_K_DEFAULT)
        h16, tm16, s16 = s.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
        out.append({"case": "all pairs at 32 bit (kernel mask 2), 20 queries x 20 000 sequences", "pairs": int(tm["rescored_pairs"]),
                    "cells": int(tm["cells"]), "device_ms": tm["device_ms"], "gcups": tm["cells"] / tm["device_ms"] / 1e6,
                    "identical_to_the_16_bit_path": bool(np.array_equal(s32, s16) and h32 == h16)})
        # (b) the re-score under stress
        rng = np.random.default_rng(5)
        W = np.uint8(19)
        rich = lambda n, f: np.where(rng.random(n) < f, W, bench.AA20_CODE[rng.integers(0, 20, size=n)]).astype(np.uint8)
        queries = [qs[i] for i in range(16)] + [np.full(n, W, dtype=np.uint8) for n in (5100, 5200, 5300, 5478)]
        extra = [np.full(int(n), W, dtype=np.uint8) for n in rng.integers(5100, 9000, size=1000)] + [rich(int(n), 0.97) for n in rng.integers(5600, 9000, size=100)]
        slens, off, codes, pos = bench.build_db(20000, bench.CONFIGS[2]["mu"], 77, extra)
        db = ob.Database(codes, off)
        q = ob.Queries.from_list(queries)
        s.load_db(db)
        s.search(q, ob.matrix("pam30"), 9, 1, top=10)
        hits, tm, scores = s.search(q, ob.matrix("pam30"), 9, 1, top=10, all_scores=True)
        flagged = scores + 9 + 2 + 32 >= 65504
        lens = np.diff(db.offsets.astype(np.int64))
        cells32 = int(sum(int(q.lengths()[qi]) * int(lens[flagged[qi]].sum()) for qi in range(q.n)))
        # oracle on a sample of the re-scored pairs and of the others
        import oracle_lib as O
        idx = sorted(set(int(x) for x in rng.choice(pos, size=12, replace=False)) | set(int(x) for x in rng.choice(db.n_seqs, size=100, replace=False)))
        o2 = np.zeros(len(idx) + 1, dtype=np.uint64)
        o2[1:] = np.cumsum([len(db.sequence(i)) for i in idx])
        want = O.search(q.residues, q.offsets, np.concatenate([db.sequence(i) for i in idx]), o2, O.matrix("pam30"), 9, 1)
        out.append({"case": "re-score under stress: 20 queries (4 all-tryptophan) x 21 100 sequences (1 000 all-tryptophan, 100 tryptophan-rich), PAM30 9/1",
                    "pairs_total": int(scores.size), "pairs_rescored": int(tm["rescored_pairs"]), "fraction_rescored": tm["rescored_pairs"] / scores.size,
                    "flagged_by_score": int(flagged.sum()), "rescore_ms": tm["rescore_ms"], "score_ms": tm["score_ms"], "device_ms": tm["device_ms"],
                    "cells_rescored": cells32, "rescore_gcups": cells32 / tm["rescore_ms"] / 1e6 if tm["rescore_ms"] else None,
                    "max_score": int(scores.max()), "oracle_pairs": int(want.size), "oracle_mismatches": int((scores[:, idx] != want).sum()),
                    "launches": int(tm["launches"])})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
