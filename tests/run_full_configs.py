#!/usr/bin/env python3
"""Runs the five BASELINE.json configurations at full size on one B200 and checks them.

TEST INFRASTRUCTURE (imports the oracle).  Not collected by pytest (too long for the suite);
run under gpurun:   python tests/run_full_configs.py [--configs 1,2,3,4,5] > gpurun_out/configs.json

For every configuration the GPU result is checked (i) against the oracle on a random sample of
database sequences plus every planted / titin-length sequence, for all queries, bit for bit;
(ii) the returned top-r lists against a host ranking of the GPU's own full score matrix with the
reference comparator; (iii) configuration-specific properties (self scores of planted copies).
Config 1 is additionally compared with the reference binary's printed top-10 when it is present.
Prints one JSON object per configuration.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # workload generator shared with the bench
import oracle_lib as O
import oswald_b200 as ob

AA = bench.AA20_CODE


def mutate(rng, seq, rate):
    out = seq.copy()
    hit = rng.random(len(seq)) < rate
    out[hit] = AA[rng.integers(0, 20, size=int(hit.sum()))]
    return out


def build_db(n_seqs, mu, sigma, seed, extra=()):
    """Synthetic canonical database with extra sequences merged in (stable by length); the
    residues are generated straight into canonical order (no gather over a 1.3 G array)."""
    S = bench.synth_lib()
    lens = np.empty(n_seqs, dtype=np.uint16)
    S.osw_synth_lengths(n_seqs, mu, sigma, 10, 65535, seed, lens.ctypes.data)
    if extra:
        lens = np.concatenate([lens, np.array([len(e) for e in extra], dtype=np.uint16)])
    perm = np.argsort(lens, kind="stable").astype(np.uint64)
    slens = np.ascontiguousarray(lens[perm])
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(slens, dtype=np.uint64)
    codes = np.empty(int(off[-1]), dtype=np.uint8)
    S.osw_synth_codes(len(lens), slens.ctypes.data, off.ctypes.data, perm.ctypes.data, seed, 0, codes.ctypes.data)
    where = np.empty(len(lens), dtype=np.int64)
    where[perm.astype(np.int64)] = np.arange(len(lens))
    pos = [int(where[n_seqs + k]) for k in range(len(extra))]
    for k, e in enumerate(extra):
        codes[int(off[pos[k]]):int(off[pos[k] + 1])] = np.asarray(e, dtype=np.uint8)
    return ob.Database(codes, off), pos


def rank_rows(scores, top):
    """Reference order on the host: score descending, higher index first."""
    out = []
    n = scores.shape[1]
    idx = np.arange(n, dtype=np.int64)
    for row in scores:
        key = row.astype(np.int64) * (1 << 32) + idx
        part = np.argpartition(key, n - min(top, n))[n - min(top, n):]
        best = part[np.argsort(key[part])[::-1]]
        out.append([(int(row[i]), int(i)) for i in best])
    return out


def check(db, queries, name, go, ge, top, scores, hits, must_check, rng, n_sample):
    sample = set(int(x) for x in rng.choice(db.n_seqs, size=min(n_sample, db.n_seqs), replace=False))
    sample |= set(must_check)
    sample = sorted(sample)
    off = np.zeros(len(sample) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(db.sequence(i)) for i in sample])
    res = np.concatenate([db.sequence(i) for i in sample])
    t0 = time.time()
    want = O.search(queries.residues, queries.offsets, res, off, O.matrix(name), go, ge)
    got = scores[:, sample]
    bad = int((got != want).sum())
    ranked_ok = rank_rows(scores, top) == hits
    return {"oracle_checked_pairs": int(want.size), "oracle_mismatches": bad, "oracle_seconds": round(time.time() - t0, 1),
            "top_r_matches_reference_order": bool(ranked_ok), "max_score": int(scores.max())}


def run(s, db, queries, name, go, ge, top):
    mat = ob.matrix(name)
    s.search(queries, mat, go, ge, top=top)                      # warm-up
    hits, tm, scores = s.search(queries, mat, go, ge, top=top, all_scores=True)
    gcups = tm["cells"] / (tm["device_ms"] / 1e3) / 1e9
    return hits, tm, scores, gcups


def reference_top(db_seqs, queries_list, name, go, ge, top):
    ref = os.path.join(ROOT, "oracle", "_ref", "oswald_ref")
    if not os.path.exists(ref):
        return None, None
    with tempfile.TemporaryDirectory() as tmp:
        bench.write_fasta(os.path.join(tmp, "db.fasta"), db_seqs, "s")
        bench.write_fasta(os.path.join(tmp, "q.fasta"), queries_list, "q")
        subprocess.run([ref, "-O", "preprocess", "-i", "db.fasta", "-o", "db", "-c", "4"], cwd=tmp, check=True, capture_output=True)
        env = dict(os.environ, OSWALD_ORACLE_TIMING=os.path.join(tmp, "t.txt"))
        cores = os.cpu_count() or 1
        out = subprocess.run([ref, "-O", "search", "-q", "q.fasta", "-d", "db", "-m", "1", "-v", "32", "-c", str(cores), "-p", "0.2",
                              "-r", str(top), "-s", name, "-g", str(go), "-e", str(ge)], cwd=tmp, check=True,
                             capture_output=True, env=env).stdout.decode(errors="replace")
        t_cpu, t_work = [float(x) for x in open(os.path.join(tmp, "t.txt")).read().split()]
        titles = [l.rstrip("\n").rstrip("\x00")[1:] for l in open(os.path.join(tmp, "db.desc"), errors="replace")]
    blocks = out.split("Query no.")[1:]
    tops = []
    for b in blocks:
        lines = b.split("Score\tSequence description\n")[1].split("\n")[:top]
        tops.append([(int(l.split("\t")[0]), l.split("\t")[1].rstrip("\x00")) for l in lines if "\t" in l])
    return (tops, titles), t_cpu + t_work


def run_config(cfg, s, args, rng, queries_all):
    t_start = time.time()
    if cfg == 1:
        db, _ = build_db(10_000, 5.6, 0.6, 101)
        qs = [queries_all[0]]
        queries = ob.Queries.from_list(qs)
        s.load_db(db)
        hits, tm, scores, gcups = run(s, db, queries, "blosum62", 10, 2, 10)
        r = check(db, queries, "blosum62", 10, 2, 10, scores, hits, [], rng, db.n_seqs)     # every pair
        seqs = [db.sequence(i) for i in range(db.n_seqs)]
        ref, secs = reference_top(seqs, qs, "blosum62", 10, 2, 10)
        if ref is not None:
            tops, titles = ref
            mine = [(sc, titles[i]) for sc, i in hits[0]]
            r["reference_binary_top10_identical"] = mine == tops[0]
            r["reference_binary_gcups"] = queries.total_length * db.n_residues / secs / 1e9
            r["reference_binary_cores"] = os.cpu_count()
        what = "144-residue query vs 10k-sequence DB, BLOSUM62 10/2, top 10"
    elif cfg in (2, 3):
        n = 570_000 if cfg == 2 else 6_900_000
        mu = bench.MU if cfg == 2 else 5.056          # config 3: mean length 188 (1.3 G residues)
        db, _ = build_db(n, mu, 0.6, 200 + cfg)
        queries = ob.Queries.from_list(queries_all)
        s.load_db(db)
        hits, tm, scores, gcups = run(s, db, queries, "blosum62", 10, 2, 10)
        r = check(db, queries, "blosum62", 10, 2, 10, scores, hits, [], rng, args.sample)
        what = ("Swiss-Prot-sized" if cfg == 2 else "Environmental-NR-sized") + " synthetic DB, 20 queries, BLOSUM62 10/2, 1xB200"
    elif cfg == 4:
        long_q = [q for q in queries_all if len(q) >= 3000]
        extra = []
        for q in long_q:
            extra += [q.copy(), mutate(rng, q, 0.10), mutate(rng, q, 0.30)]
        db, where = build_db(570_000, bench.MU, 0.6, 204, extra)
        queries = ob.Queries.from_list(queries_all)
        s.load_db(db)
        r = {"runs": []}
        gcups_all = []
        for name, go, ge in (("pam30", 9, 1), ("blosum45", 14, 2)):
            hits, tm, scores, gcups = run(s, db, queries, name, go, ge, 10)
            rr = check(db, queries, name, go, ge, 10, scores, hits, where, rng, args.sample // 2)
            m = ob.matrix(name).reshape(24, 32)
            ok = True
            for k, q in enumerate(long_q):                       # exact copies score their self score, on top
                qi = [i for i in range(queries.n) if len(queries.query(i)) == len(q)][0]
                self_score = int(sum(int(m[c, c]) for c in q))
                ok &= hits[qi][0] == (self_score, where[3 * k])
            rr.update({"matrix": name, "gap": [go, ge], "gcups": gcups, "planted_copies_on_top_with_self_score": bool(ok),
                       "rescored_pairs": tm["rescored_pairs"], "rescore_ms": tm["rescore_ms"], "device_ms": tm["device_ms"]})
            r["runs"].append(rr)
            gcups_all.append(gcups)
        gcups = float(np.mean(gcups_all))
        r["oracle_mismatches"] = sum(x["oracle_mismatches"] for x in r["runs"])
        what = "Swiss-Prot-sized DB + planted homologs of the long queries, PAM30 9/1 and BLOSUM45 14/2"
    else:
        extra = [AA[rng.integers(0, 20, size=int(L))] for L in rng.integers(35_000, 65_536, size=14)]
        extra.append(AA[rng.integers(0, 20, size=65_535)])
        big = queries_all[-1]
        tandem = np.concatenate([np.concatenate([big, AA[rng.integers(0, 20, size=50)]]) for _ in range(6)])
        extra.append(tandem)
        db, where = build_db(570_000, bench.MU, 0.6, 205, extra)
        queries = ob.Queries.from_list(queries_all)
        s.load_db(db)
        hits, tm, scores, gcups = run(s, db, queries, "blosum62", 10, 2, 10)
        r = check(db, queries, "blosum62", 10, 2, 10, scores, hits, where, rng, 200)
        r["titin_length_sequences"] = len(extra)
        what = "Swiss-Prot-sized DB + 16 sequences of 35k-65k residues (one with 6 tandem copies of the 5478 query), BLOSUM62"
    r.update({"config": cfg, "what": what, "sequences": db.n_seqs, "residues": db.n_residues, "gcups_device": gcups,
              "device_ms": tm["device_ms"], "score_ms": tm["score_ms"], "topr_ms": tm["topr_ms"], "launches": tm["launches"], "rescored_pairs_last_run": tm["rescored_pairs"],
              "wall_seconds_total": round(time.time() - t_start, 1)})
    return r



def reference_binary_matrix(s, queries_all, n_sample=20000):
    """The reference binary itself (host AVX2 path) on a 20 000-sequence database with all 20
    queries: every raw score it computes (dumped by the shim's sort_scores hook) against the GPU."""
    ref = os.path.join(ROOT, "oracle", "_ref", "oswald_ref")
    if not os.path.exists(ref):
        return {"config": "2-sample-vs-reference-binary", "skipped": "oracle/_ref/oswald_ref not built"}
    db, _ = build_db(n_sample, bench.MU, 0.6, 777)
    queries = ob.Queries.from_list(queries_all)
    seqs = [db.sequence(i) for i in range(db.n_seqs)]
    with tempfile.TemporaryDirectory() as tmp:
        bench.write_fasta(os.path.join(tmp, "db.fasta"), seqs, "s")
        bench.write_fasta(os.path.join(tmp, "q.fasta"), queries_all, "q")
        subprocess.run([ref, "-O", "preprocess", "-i", "db.fasta", "-o", "db", "-c", "4"], cwd=tmp, check=True, capture_output=True)
        env = dict(os.environ, OSWALD_ORACLE_DUMP=os.path.join(tmp, "dump.bin"))
        t0 = time.time()
        subprocess.run([ref, "-O", "search", "-q", "q.fasta", "-d", "db", "-m", "1", "-v", "32", "-c", str(os.cpu_count() or 1),
                        "-p", "0.1", "-r", "10"], cwd=tmp, check=True, capture_output=True, env=env)
        secs = time.time() - t0
        want = np.fromfile(os.path.join(tmp, "dump.bin"), dtype=np.int32).reshape(queries.n, db.n_seqs)
    s.load_db(db)
    hits, tm, scores = s.search(queries, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
    return {"config": "2-sample-vs-reference-binary", "what": "every raw score of the reference binary (20 queries x %d sequences)" % db.n_seqs,
            "pairs": int(want.size), "oracle_mismatches": int((scores != want).sum()),
            "top_r_matches_reference_order": rank_rows(want, 10) == hits, "reference_wall_seconds": round(secs, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--sample", type=int, default=1500)
    args = ap.parse_args()
    todo = [int(x) for x in args.configs.split(",")]
    rng = np.random.default_rng(99)
    queries_all = bench.make_queries()
    s = ob.Searcher(1)
    results = []
    for cfg in todo:
        try:
            r = run_config(cfg, s, args, rng, queries_all)
        except Exception as e:          # keep going: one configuration must not hide the others
            import traceback
            traceback.print_exc()
            r = {"config": cfg, "error": repr(e), "oracle_mismatches": 1}
        results.append(r)
        print(json.dumps(r), flush=True)
    if 2 in todo:
        r = reference_binary_matrix(s, queries_all)
        results.append(r)
        print(json.dumps(r), flush=True)
    s.close()
    bad = sum(x.get("oracle_mismatches", 0) for x in results)
    print(json.dumps({"summary": "all configurations bit-exact on the checked pairs" if bad == 0 else "MISMATCHES", "mismatches": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
