#!/usr/bin/env python3
"""Long randomised parity sweep on a GPU (not collected by pytest):  python tests/fuzz_gpu.py [cases] [seed]

TEST INFRASTRUCTURE (imports the oracle).  Every case: random database (sometimes with empty,
single-residue, non-standard-residue or very long sequences), random queries (sometimes empty,
sometimes thousands of residues, sometimes homologous to database sequences), random matrix / gap
penalties / chunk size / top-r / first-stage mode / shard split; all scores and all hit lists are
compared with the oracle."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import oswald_b200 as ob
from oswald_b200 import capi
from oswald_b200.host import merge_hits

AA = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)
MODES = [capi.OSW_K_DEFAULT, capi.OSW_K_DEFAULT | capi.OSW_K_TWO_TRACK, capi.OSW_K_DEFAULT | capi.OSW_K_PAIR_DB, capi.OSW_K_I32,
         capi.OSW_K_DEFAULT | capi.OSW_K_TRANSPOSED]


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    names = ob.matrix_names()
    searchers = [ob.Searcher(1), ob.Searcher(1)]
    bad_cases = 0
    t0 = time.time()
    for case in range(cases):
        n = int(rng.integers(1, 600))
        style = int(rng.integers(0, 6))
        if style == 0:
            lens = rng.integers(0, 6, size=n)
        elif style == 1:
            lens = np.clip(np.round(np.exp(rng.normal(5.0, 0.8, size=n))), 1, 6000)
        elif style == 2:
            lens = np.concatenate([rng.integers(1, 200, size=n), rng.integers(5000, 20000, size=2)])
        else:
            lens = rng.integers(1, int(rng.integers(2, 900)), size=n)
        alphabet = 24 if rng.random() < 0.2 else 20
        seqs = [(rng.integers(0, 24, size=int(l)).astype(np.uint8) if alphabet == 24 else AA[rng.integers(0, 20, size=int(l))]) for l in lens]
        nq = int(rng.integers(1, 10))
        mode = MODES[int(rng.integers(0, len(MODES)))]
        qlens = []
        for _ in range(nq):
            r = rng.random()
            if mode & capi.OSW_K_TRANSPOSED:          # the transposed form takes queries of up to 1024 residues, 4096 in all
                qlens.append(0 if r < 0.05 else int(rng.integers(1, 40)) if r < 0.3 else int(rng.integers(40, 250)) if r < 0.85
                             else int(rng.integers(250, 1025)))
                continue
            qlens.append(0 if r < 0.05 else int(rng.integers(1, 40)) if r < 0.3 else int(rng.integers(40, 700)) if r < 0.8
                         else int(rng.integers(700, 3200)) if r < 0.97 else int(rng.integers(3200, 7000)))
        qs = [AA[rng.integers(0, 20, size=m)] for m in qlens]
        for k in range(min(4, len(seqs), nq)):                      # homologs with indels
            if qlens[k] > 8 and rng.random() < 0.6:
                cut = int(rng.integers(1, qlens[k]))
                seqs[k] = np.concatenate([qs[k][:cut], AA[rng.integers(0, 20, size=int(rng.integers(0, 5)))], qs[k][cut:]])[:65535]
        lens_arr = np.array([len(s) for s in seqs], dtype=np.uint64)
        db = ob.Database.from_lengths(lens_arr, np.concatenate(seqs) if lens_arr.sum() else np.zeros(0, np.uint8))
        q = ob.Queries.from_list(qs)
        name = names[int(rng.integers(0, len(names)))]
        go, ge = (int(rng.integers(0, 40)), int(rng.integers(0, 8))) if rng.random() < 0.9 else (255, 127)
        top = int(rng.choice([0, 1, 5, 10, 64, len(seqs) + 3]))
        chunk = int(rng.choice([0, 0, 0, 16, 100, 700, 3000]))
        shards = 2 if rng.random() < 0.25 else 1
        want = O.search(q.residues, q.offsets, db.residues, db.offsets, O.matrix(name), go, ge)
        got = np.zeros_like(want)
        parts = []
        for r in range(shards):
            s = searchers[r]
            s.set_kernels(mode)
            s.load_db(db, shard_rank=r, shard_count=shards, max_chunk_residues=chunk)
            hits, tm, sc = s.search(q, ob.matrix(name), go, ge, top=top, all_scores=True)
            got += sc
            parts.append(hits)
        ok = np.array_equal(got, want)
        for qi in range(q.n):
            idx, val = O.top_r(want[qi], top) if top else ([], [])
            merged = merge_hits([p[qi] for p in parts], top) if top else []
            ok &= merged == [(int(v), int(i)) for v, i in zip(val, idx)]
        if not ok:
            bad_cases += 1
            bad = np.argwhere(got != want)
            print("MISMATCH case", case, "n", len(seqs), "style", style, "qlens", sorted(qlens), name, go, ge, "top", top, "chunk", chunk,
                  "mode", mode, "shards", shards, "bad scores", len(bad), bad[:3].tolist())
    print("fuzz: %d cases, %d failing, %.0f s" % (cases, bad_cases, time.time() - t0))
    return 1 if bad_cases else 0


if __name__ == "__main__":
    sys.exit(main())
