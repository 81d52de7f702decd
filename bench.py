#!/usr/bin/env python3
"""bench.py - GCUPS of the Smith-Waterman database search hot path on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     the reference's own host AVX2 path

Workload (default): BASELINE.json config 3 - ONE fixed Environmental-NR-sized synthetic database
(6.9 M sequences, ~1.3 G residues, log-normal lengths) searched with 20 protein queries (lengths
144..5478, sum 41 750), BLOSUM62, gap 10/2, top 10.  With N GPUs that same database is SPLIT:
its chunks are dealt round-robin to the N ranks (strong scaling; no data-path collective, only
the r hits per query are gathered and merged with the reference comparator).  A step = one
complete search of the whole database:  host buffers -> HBM (database streams, queries), scoring,
32-bit re-score, top-r, hits back to the host, merge.  `value` is device-timed (CUDA events
around everything the GPUs do for the search, database already resident when the events start);
`e2e` is the wall clock of the same K steps including the host->device copy of the database.
GCUPS = sum(query lengths) * sum(database residues) / seconds / 1e9 - the reference's own
definition (HybridSearch.c:1227), unpadded.

After the timed region (never inside it) rank 0 verifies the merged result against the oracle and
a host ranking, checks the one-process multi-GPU form of the library, and runs the other four
BASELINE.json configurations for three steps each (`extra.configs`).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUERY_LENGTHS = [144, 189, 222, 375, 464, 567, 657, 727, 850, 1000, 1500, 2005, 2504, 3005, 3564,
                 4061, 4548, 4743, 5147, 5478]
SIGMA = 0.6
SEED = 20261018
TOP = 10
AA20_CODE = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)

# BASELINE.json configs: sequences, mu of ln(length), queries, (matrix, go, ge) runs
CONFIGS = {
    1: {"name": "config 1", "what": "144-residue query vs 10k-sequence synthetic DB, BLOSUM62 10/2, top 10",
        "n": 10_000, "mu": 5.6, "queries": [144], "runs": [("blosum62", 10, 2)]},
    2: {"name": "config 2", "what": "Swiss-Prot-sized synthetic DB (570k sequences, ~205 M residues), 20 queries 144-5478, BLOSUM62 10/2, top 10",
        "n": 570_000, "mu": 5.706, "queries": QUERY_LENGTHS, "runs": [("blosum62", 10, 2)]},
    3: {"name": "config 3", "what": "Environmental-NR-sized synthetic DB (6.9 M sequences, ~1.3 G residues), 20 queries 144-5478, BLOSUM62 10/2, top 10",
        "n": 6_900_000, "mu": 5.056, "queries": QUERY_LENGTHS, "runs": [("blosum62", 10, 2)]},
    4: {"name": "config 4", "what": "Swiss-Prot-sized DB + planted homologs (exact, 10 % and 30 % mutated) of every query >= 3000 residues, PAM30 9/1 and BLOSUM45 14/2",
        "n": 570_000, "mu": 5.706, "queries": QUERY_LENGTHS, "runs": [("pam30", 9, 1), ("blosum45", 14, 2)], "planted": True},
    5: {"name": "config 5", "what": "Swiss-Prot-sized DB + 16 sequences of 35 000-65 535 residues (one with 6 tandem copies of the 5478 query), BLOSUM62 10/2",
        "n": 570_000, "mu": 5.706, "queries": QUERY_LENGTHS, "runs": [("blosum62", 10, 2)], "titin": True},
}
MAIN_CONFIG = 3


def synth_lib():
    path = os.path.join(ROOT, "tools", "libosw_synth.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ROOT, "tools"])
    L = C.CDLL(path)
    L.osw_synth_lengths.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    L.osw_synth_codes.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    return L


def build_db(n_seqs, mu, seed, extra=()):
    """Canonical (stable ascending length) synthetic database: (lengths u16, offsets u64, codes u8,
    positions of the `extra` sequences).  Residues are generated straight into canonical order."""
    S = synth_lib()
    lens = np.empty(n_seqs, dtype=np.uint16)
    S.osw_synth_lengths(n_seqs, mu, SIGMA, 10, 65535, seed, lens.ctypes.data)
    if len(extra):
        lens = np.concatenate([lens, np.array([len(e) for e in extra], dtype=np.uint16)])
    perm = np.argsort(lens, kind="stable").astype(np.uint64)      # canonical order
    slens = np.ascontiguousarray(lens[perm])
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(slens, dtype=np.uint64)
    codes = np.empty(int(off[-1]), dtype=np.uint8)
    S.osw_synth_codes(len(lens), slens.ctypes.data, off.ctypes.data, perm.ctypes.data, seed, 0, codes.ctypes.data)
    pos = []
    if len(extra):
        where = np.empty(len(lens), dtype=np.int64)
        where[perm.astype(np.int64)] = np.arange(len(lens))
        pos = [int(where[n_seqs + k]) for k in range(len(extra))]
        for k, e in enumerate(extra):
            codes[int(off[pos[k]]):int(off[pos[k] + 1])] = np.asarray(e, dtype=np.uint8)
    return slens, off, codes, pos


def make_database(n_seqs, seed=SEED, mu=CONFIGS[2]["mu"]):
    slens, off, codes, _ = build_db(n_seqs, mu, seed)
    return slens, off, codes, None


def make_queries(lengths=None, seed=SEED):
    """Query i of the standard set is the same sequence whichever subset of lengths is asked for."""
    S = synth_lib()
    lengths = list(QUERY_LENGTHS if lengths is None else lengths)
    out = []
    for m in lengths:
        stream = np.array([QUERY_LENGTHS.index(m) if m in QUERY_LENGTHS else 1000 + m], dtype=np.uint64)
        ln = np.array([m], dtype=np.uint16)
        off = np.zeros(1, dtype=np.uint64)
        codes = np.empty(m, dtype=np.uint8)
        S.osw_synth_codes(1, ln.ctypes.data, off.ctypes.data, stream.ctypes.data, seed ^ 0x51, 0, codes.ctypes.data)
        out.append(codes)
    return out


def mutate(rng, seq, rate):
    out = seq.copy()
    hit = rng.random(len(seq)) < rate
    out[hit] = AA20_CODE[rng.integers(0, 20, size=int(hit.sum()))]
    return out


def make_workload(cfg_id, n_override=0, query_lengths=None):
    """Database + queries of a BASELINE.json configuration (deterministic)."""
    import oswald_b200 as ob
    cfg = CONFIGS[cfg_id]
    rng = np.random.default_rng(SEED + cfg_id)
    qs = make_queries(query_lengths or cfg["queries"])
    extra, kinds = [], []
    if cfg.get("planted"):
        for q in qs:
            if len(q) >= 3000:
                extra += [q.copy(), mutate(rng, q, 0.10), mutate(rng, q, 0.30)]
                kinds += [("copy", len(q)), ("mut10", len(q)), ("mut30", len(q))]
    if cfg.get("titin"):
        extra += [AA20_CODE[rng.integers(0, 20, size=int(L))] for L in rng.integers(35_000, 65_536, size=14)]
        extra.append(AA20_CODE[rng.integers(0, 20, size=65_535)])
        big = qs[-1]
        extra.append(np.concatenate([np.concatenate([big, AA20_CODE[rng.integers(0, 20, size=50)]]) for _ in range(6)]))
        kinds += [("titin", 0)] * 15 + [("tandem", len(big))]
    n = n_override or cfg["n"]
    slens, off, codes, pos = build_db(n, cfg["mu"], SEED + 100 * cfg_id, extra)
    return {"cfg": cfg_id, "db": ob.Database(codes, off), "queries": ob.Queries.from_list(qs), "runs": cfg["runs"],
            "planted": list(zip(pos, kinds)), "n_seqs": n + len(extra)}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 7 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 7 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own host AVX2 path (oracle/_ref/oswald_ref*), or
# the scalar C port of it (oracle/liboswald_oracle.so) when the reference was not compiled.
# ------------------------------------------------------------------------------------------
def write_fasta(path, seqs_codes, prefix):
    lut = np.frombuffer(b"ABCDEFGHIKLMNPQRSTVWXYZJ", dtype=np.uint8)
    with open(path, "wb") as f:
        for i, c in enumerate(seqs_codes):
            f.write((">%s%d len%d\n" % (prefix, i, len(c))).encode())
            s = lut[c].tobytes()
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + b"\n")


_REF_BINARY = None


def reference_binary():
    """The reference binary to time: the -O3 -march=native build when it runs
    on this host's CPU, else the portable -O3 -mavx2 -mfma build.  Both are the reference's
    unmodified sources (oracle/Makefile).  Returns (path, build flags) or (None, None)."""
    global _REF_BINARY
    if _REF_BINARY is not None:
        return _REF_BINARY
    base = os.path.join(ROOT, "oracle", "_ref")
    _REF_BINARY = (None, None)
    for exe, flags in (("oswald_ref_native", "-O3 -march=native (built in the build container)"), ("oswald_ref", "-O3 -mavx2 -mfma")):
        path = os.path.join(base, exe)
        if not os.path.exists(path):
            continue
        try:        # a -march=native binary dies with SIGILL on a CPU that lacks the build host's extensions
            with tempfile.TemporaryDirectory() as tmp:
                rng = np.random.default_rng(1)
                write_fasta(os.path.join(tmp, "db.fasta"), [AA20_CODE[rng.integers(0, 20, size=int(l))] for l in range(20, 2020)], "s")
                write_fasta(os.path.join(tmp, "q.fasta"), [AA20_CODE[rng.integers(0, 20, size=100)]], "q")
                subprocess.run([path, "-O", "preprocess", "-i", "db.fasta", "-o", "db", "-c", "2"], cwd=tmp, check=True, capture_output=True, timeout=120)
                subprocess.run([path, "-O", "search", "-q", "q.fasta", "-d", "db", "-m", "1", "-v", "32", "-c", "2", "-p", "0.3", "-r", "3"],
                               cwd=tmp, check=True, capture_output=True, timeout=120)
            _REF_BINARY = (path, flags)
            break
        except Exception:
            continue
    return _REF_BINARY


def reference_cpu_gcups(sample_seqs, queries, cores, run=("blosum62", 10, 2)):
    """Times the reference's host path on (queries x sample).  Returns (gcups, kind, seconds, build)."""
    ref, flags = reference_binary()
    q_total = sum(len(q) for q in queries)
    d_total = sum(len(s) for s in sample_seqs)
    if ref:
        with tempfile.TemporaryDirectory() as tmp:
            write_fasta(os.path.join(tmp, "db.fasta"), sample_seqs, "s")
            write_fasta(os.path.join(tmp, "q.fasta"), queries, "q")
            subprocess.run([ref, "-O", "preprocess", "-i", "db.fasta", "-o", "db", "-c", str(min(cores, 8))],
                           cwd=tmp, check=True, capture_output=True)
            env = dict(os.environ, OSWALD_ORACLE_TIMING=os.path.join(tmp, "timing.txt"))
            env.pop("OSWALD_ORACLE_DUMP", None)
            # -p: the calibration sample must hold some 16-sequence groups per thread
            pct = max(0.05, min(0.5, 64.0 * cores * 400 / max(d_total, 1)))
            subprocess.run([ref, "-O", "search", "-q", "q.fasta", "-d", "db", "-m", "1", "-v", "32", "-c", str(cores),
                            "-p", "%.3f" % pct, "-r", str(TOP), "-s", run[0], "-g", str(run[1]), "-e", str(run[2])],
                           cwd=tmp, check=True, capture_output=True, env=env, timeout=3000)
            t_cpu, t_work = [float(x) for x in open(os.path.join(tmp, "timing.txt")).read().split()]
        secs = t_cpu + t_work
        return q_total * d_total / secs / 1e9, "reference", secs, flags
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    off = np.zeros(len(sample_seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in sample_seqs])
    q_off = np.zeros(len(queries) + 1, dtype=np.uint32)
    q_off[1:] = np.cumsum([len(q) for q in queries])
    t0 = time.time()
    O.search(np.concatenate(queries), q_off, np.concatenate(sample_seqs), off, O.matrix(run[0]), run[1], run[2], threads=cores)
    secs = time.time() - t0
    return q_total * d_total / secs / 1e9, "port", secs, "scalar C port (oracle/sw_oracle.c), -O2"


def cpu_sample(n_sample, cfg_id):
    """A sample of the configuration's own length distribution (another seed)."""
    slens, off, codes, _ = build_db(n_sample, CONFIGS[cfg_id]["mu"], SEED + 1)
    return [codes[int(off[i]):int(off[i + 1])] for i in range(n_sample)]


def sample_text(queries, seqs, cfg_id):
    return "%d queries (sum %d) x %d-sequence / %d-residue sample of %s's length distribution (the full database has %d sequences)" % (
        len(queries), sum(len(q) for q in queries), len(seqs), sum(len(s) for s in seqs), CONFIGS[cfg_id]["name"], CONFIGS[cfg_id]["n"])


def workload_config(n_gpus, cfg_id, n_seqs, query_lengths):
    cfg = CONFIGS[cfg_id]
    name = cfg["name"] if n_seqs == cfg["n"] and list(query_lengths) == list(cfg["queries"]) else \
        "%s's length distribution with %d sequences, %d queries (not a BASELINE.json size)" % (cfg["name"], n_seqs, len(query_lengths))
    return {"workload": "%s: %s" % (name, cfg["what"]) if name == cfg["name"] else name,
            "sequences": n_seqs, "length_distribution": "log-normal mu=%.3f sigma=%.1f clipped [10,65535]" % (cfg["mu"], SIGMA),
            "query_lengths": list(query_lengths), "matrix": cfg["runs"][0][0], "gap_open": cfg["runs"][0][1], "gap_extend": cfg["runs"][0][2],
            "top": TOP, "split": "one fixed database; its chunks (~8192 residues; 2048 and 512 for the shortest sequences, taken last) "
                                 "are dealt round-robin to the %d GPU(s) - strong scaling" % n_gpus,
            "l2": "inputs larger than L2 (database stream >= 160 MB per GPU, re-copied from the host every step)"}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    qlens = [int(x) for x in args.query_lengths.split(",")] if args.query_lengths else CONFIGS[args.config]["queries"]
    queries = make_queries(qlens)
    seqs = cpu_sample(args.ref_sample, args.config)
    run = CONFIGS[args.config]["runs"][0]
    vals, secs_all = [], []
    kind, build = "port", ""
    for it in range(args.warmup + args.steps):
        g, kind, secs, build = reference_cpu_gcups(seqs, queries, cores, run)
        if it >= args.warmup:
            vals.append(g)
            secs_all.append(secs)
    v = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": "GCUPS", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs_all)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int8/int16/int32 host SIMD cascade", "data": "synthetic",
        "config": workload_config(args.gpus, args.config, args.seqs or CONFIGS[args.config]["n"], qlens),
        "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind, "build": build, "sample": sample_text(queries, seqs, args.config)},
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------
# verification (after the timed region): oracle on sampled pairs, host ranking of full score rows
# ------------------------------------------------------------------------------------------
def host_ranking(scores, top):
    """Reference order (utils.c:3-69) on the host: score descending, higher canonical index first."""
    out = []
    n = scores.shape[1]
    idx = np.arange(n, dtype=np.int64)
    k = min(top, n)
    for row in scores:
        key = row.astype(np.int64) * (1 << 32) + idx
        part = np.argpartition(key, n - k)[n - k:]
        best = part[np.argsort(key[part])[::-1]]
        out.append([(int(row[i]), int(i)) for i in best])
    return out


class Job:
    """This rank's view of a search job: the shard of the database it holds, and the collectives."""

    def __init__(self, searcher, rank, world, dist, torch):
        self.s, self.rank, self.world, self.dist, self.torch = searcher, rank, world, dist, torch

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def merge(self, hits, nq):
        """Gather every rank's r hits per query; rank 0 merges them with the reference comparator."""
        from oswald_b200.host import merge_hits
        if self.dist is None:
            return hits
        parts = [None] * self.world
        self.dist.all_gather_object(parts, hits)
        return [merge_hits([p[q] for p in parts], TOP) for q in range(nq)] if self.rank == 0 else None

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def sum_to_all(self, arr):
        """Element-wise sum over ranks of an int32 array (score rows: every entry is owned by one rank, 0 elsewhere)."""
        if self.dist is None:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def verify(self, wl, run, merged_hits, n_sample, titin_checked=3):
        """One more search with the full score matrix; rank 0 compares (i) the scores of a random
        sample of sequences plus planted ones, all queries, with the oracle, bit for bit; (ii) the
        merged top-r lists of the timed steps with a host ranking of the gathered score rows."""
        import oswald_b200 as ob
        db, queries = wl["db"], wl["queries"]
        name, go, ge = run
        hits2, tm2, scores = self.s.search(queries, ob.matrix(name), go, ge, top=TOP, all_scores=True)
        merged2 = self.merge(hits2, queries.n)
        rng = np.random.default_rng(SEED + 7 * wl["cfg"])
        sample = set(int(x) for x in rng.choice(db.n_seqs, size=min(n_sample, db.n_seqs), replace=False))
        planted = [p for p, (kind, _) in wl["planted"] if kind != "titin"]
        planted += [p for p, (kind, _) in wl["planted"] if kind == "titin"][-titin_checked:]
        sample = sorted(sample | set(planted))
        rows = self.sum_to_all(scores)
        del scores
        if self.rank != 0:
            return None
        got = rows[:, sample]
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O        # the checker (test infrastructure; never in the timed region)
        off = np.zeros(len(sample) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(db.sequence(i)) for i in sample])
        res = np.concatenate([db.sequence(i) for i in sample])
        t0 = time.time()
        want = O.search(queries.residues, queries.offsets, res, off, O.matrix(name), go, ge)
        ranked = host_ranking(rows, TOP)
        out = {"pairs": int(want.size), "mismatches": int((got != want).sum()),
               "topr_ok": bool(ranked == merged_hits and merged2 == merged_hits),
               "topr_queries": queries.n, "max_score": int(rows.max()), "oracle_seconds": round(time.time() - t0, 1),
               "what": "oracle (scalar C restatement) on %d sampled + %d planted sequences x %d queries; merged top-%d of the timed "
                       "steps == host ranking (score desc, index desc) of the gathered full score rows, all queries" % (
                           len(sample) - len(planted), len(planted), queries.n, TOP)}
        # configuration-specific properties
        m = ob.matrix(name).reshape(24, 32)
        ok = True
        for pos, (kind, qlen) in wl["planted"]:
            if kind == "copy":            # an exact copy scores the query's self score and leads its list
                qi = [i for i in range(queries.n) if len(queries.query(i)) == qlen][0]
                self_score = int(sum(int(m[c, c]) for c in queries.query(qi)))
                ok &= merged_hits[qi][0] == (self_score, pos)
            if kind == "tandem":          # six tandem copies: at least the self score
                qi = queries.n - 1
                ok &= int(rows[qi, pos]) >= int(sum(int(m[c, c]) for c in queries.query(qi)))
        if wl["planted"]:
            out["planted_ok"] = bool(ok)
        return out


def single_process_multi_gpu_check(n_gpus):
    """The reference's own deployment form (-f N inside one process, arguments.c:108-112):
    osw_init(N) on all N devices, one small search, every score and hit list against the oracle."""
    import oswald_b200 as ob
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    rng = np.random.default_rng(31)
    lens = np.sort(rng.integers(10, 400, size=4000)).astype(np.uint64)
    seqs = [AA20_CODE[rng.integers(0, 20, size=int(l))] for l in lens]
    qs = [AA20_CODE[rng.integers(0, 20, size=m)] for m in (90, 250, 700)]
    seqs[1234] = np.concatenate([qs[1], seqs[1234]])[:int(lens[1234])]
    db = ob.Database.from_lengths(lens, np.concatenate(seqs), presorted=True)
    queries = ob.Queries.from_list(qs)
    with ob.Searcher(n_gpus) as s:
        s.load_db(db, max_chunk_residues=1024)
        st = s.stats()
        hits, tm, scores = s.search(queries, ob.matrix("blosum62"), 10, 2, top=TOP, all_scores=True)
    want = O.search(queries.residues, queries.offsets, db.residues, db.offsets, O.matrix("blosum62"), 10, 2)
    ok = st["n_seqs"] == db.n_seqs and bool(np.array_equal(scores, want))
    for q in range(queries.n):
        idx, sc = O.top_r(want[q], TOP)
        ok &= hits[q] == [(int(a), int(b)) for a, b in zip(sc, idx)]
    return bool(ok)


def kernel_tag():
    """Hash of the kernel and layout sources this library was built from (written by the Makefile)."""
    p = os.path.join(ROOT, "oswald_b200", "build_tag.txt")
    return open(p).read().strip() if os.path.exists(p) else None


def run_steps(job, wl, run, steps, warmup, sampler=None):
    """W warm-up steps, then K timed steps.  Returns per-rank sums and the last merged hit lists."""
    import oswald_b200 as ob
    s, queries = job.s, wl["queries"]
    name, go, ge = run
    mat = ob.matrix(name)
    merged = None
    for _ in range(warmup):
        s.upload_db()
        hits, tm = s.search(queries, mat, go, ge, top=TOP)
        merged = job.merge(hits, queries.n)
    if sampler is not None:
        sampler.start()
    job.barrier()
    t0 = time.time()
    dev_ms, tms, h2d = 0.0, [], 0
    for _ in range(steps):
        # host buffers -> HBM: the database streams and directories, then (inside osw_search) queries + matrix
        h2d = s.upload_db() + int(queries.residues.nbytes + queries.offsets.nbytes + mat.nbytes)
        hits, tm = s.search(queries, mat, go, ge, top=TOP)
        merged = job.merge(hits, queries.n)
        dev_ms += tm["device_ms"]
        tms.append(tm)
    job.barrier()
    wall_s = time.time() - t0
    if sampler is not None:
        sampler.stop_flag.set()
    return dev_ms, wall_s, tms, h2d, merged


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=MAIN_CONFIG, choices=sorted(CONFIGS), help="BASELINE.json configuration of the timed workload")
    ap.add_argument("--seqs", type=int, default=0, help="sequences in the whole database (experiments; 0 = the configuration's own size)")
    ap.add_argument("--ref-sample", type=int, default=200000, help="sequences in the CPU baseline's sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other four configurations (extra.configs)")
    ap.add_argument("--no-verify", action="store_true", help="skip the post-run verification (profiling runs)")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--kernels", type=int, default=3, help="kernel mask (1|2 = default, 2 = 32-bit only)")
    ap.add_argument("--chunk-cols", type=int, default=0, help="residues per chunk (0 = library default)")
    ap.add_argument("--window-mb", type=int, default=0, help="stream the database through two device windows of this size (experiments; 0 = resident)")
    ap.add_argument("--query-lengths", default="", help="comma-separated subset/alternative query lengths (experiments; default: the 20 standard lengths)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.gpus != world and world > 1:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("bench.py: for N > 1 launch with torch.distributed.run (one rank per GPU)")

    # stdout carries exactly one JSON line: native libraries that print there (NCCL announces its
    # version on stdout when the first communicator is made) are sent to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import oswald_b200 as ob
    from oswald_b200.host import calibrate
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: ONE database, split across the ranks ---------------------------------------
    qlens = [int(x) for x in args.query_lengths.split(",")] if args.query_lengths else None
    t_gen = time.time()
    wl = make_workload(args.config, args.seqs, qlens)
    t_gen = time.time() - t_gen
    db, queries = wl["db"], wl["queries"]
    run = wl["runs"][0]
    q_total, d_total = queries.total_length, db.n_residues

    s = ob.Searcher(devices=[local_rank])
    s.set_kernels(args.kernels)
    if args.window_mb:
        s.set_device_window(args.window_mb << 20)
    t_load = time.time()
    s.load_db(db, shard_rank=rank, shard_count=world, max_chunk_residues=args.chunk_cols)
    t_load = time.time() - t_load
    st = s.stats()
    job = Job(s, rank, world, dist, torch)

    # ---- W warm-up steps, K timed steps --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    dev_ms, e2e_s, tms, h2d, merged = run_steps(job, wl, run, args.steps, args.warmup, sampler)
    d2h = queries.n * TOP * 8
    dev_ms, e2e_s = job.max_over_ranks([dev_ms, e2e_s])          # max over ranks of the device time / the wall time

    # ---- after the timed region: verification, the one-process multi-GPU form, the other configs --
    verified = None if args.no_verify else job.verify(wl, run, merged, 600)
    sp_ok = None
    if not args.no_verify:
        n_vis = torch.cuda.device_count()
        job.barrier()
        if rank == 0 and n_vis >= 2:
            sp_ok = single_process_multi_gpu_check(min(n_vis, max(world, 2)))
        job.barrier()
    extra = []
    if not args.no_extra and not args.seqs and not qlens:
        for cid in sorted(CONFIGS):
            if cid == args.config:
                continue
            t0 = time.time()
            w2 = make_workload(cid)
            s.load_db(w2["db"], shard_rank=rank, shard_count=world)
            res = {"config": cid, "what": CONFIGS[cid]["what"], "sequences": w2["db"].n_seqs, "residues": w2["db"].n_residues, "runs": []}
            for r2 in w2["runs"]:
                dms, wall2, tms2, _, merged2 = run_steps(job, w2, r2, args.extra_steps, 1)
                dms, wall2 = job.max_over_ranks([dms, wall2])
                cells2 = w2["queries"].total_length * w2["db"].n_residues
                resc = job.max_over_ranks([float(tms2[-1]["rescored_pairs"]), tms2[-1]["rescore_ms"], tms2[-1]["score_ms"], tms2[-1]["topr_ms"]])
                v2 = job.verify(w2, r2, merged2, 300)
                res["runs"].append({"matrix": r2[0], "gap": [r2[1], r2[2]], "steps": args.extra_steps,
                                    "gcups_device": cells2 * args.extra_steps / (dms / 1e3) / 1e9,
                                    "gcups_e2e": cells2 * args.extra_steps / wall2 / 1e9, "ms_per_step": dms / args.extra_steps,
                                    "rescored_pairs_max_rank": int(resc[0]), "rescore_ms": resc[1], "score_ms": resc[2], "topr_ms": resc[3],
                                    "launches": int(tms2[-1]["launches"]), "verified": v2})
            res["seconds_total"] = round(time.time() - t0, 1)
            extra.append(res)

    if rank == 0:
        cells_per_step = q_total * d_total
        value = cells_per_step * args.steps / (dev_ms / 1e3) / 1e9
        e2e = cells_per_step * args.steps / e2e_s / 1e9
        tm = tms[-1]
        clocks = sampler.summary()
        cal = calibrate(local_rank)
        n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        # roofline of the dominant kernel (sw_u16_kernel): cell updates per SM-cycle against the
        # calibrated DPX issue rate; 3 packed instructions per cell.
        score_ms = float(np.mean([x["score_ms"] for x in tms]))
        sm_cycles = float(np.mean([x["sm_cycles"] for x in tms]))       # busy SM-cycles summed over SMs
        local_cells = tm["cells"]
        r_dpx = cal["viaddmnmx_u16x2_per_sm_clk"]
        peak_cells_clk = 2.0 * r_dpx / 6.0
        sm_mhz = clocks["sm_mhz"] or cal["sm_mhz"]
        peak_gcups = peak_cells_clk * n_sms * sm_mhz * 1e6 / 1e9
        achieved_gcups = (local_cells / (score_ms / 1e3) / 1e9) if score_ms else None
        busy_cells_clk = local_cells / sm_cycles if sm_cycles else None
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        roofline = {"bound": "alu", "kernel": "sw_u16_kernel (packed 16-bit DPX), rank 0's GPU",
                    "achieved": achieved_gcups, "peak": peak_gcups, "unit": "GCUPS",
                    "frac": (achieved_gcups / peak_gcups) if achieved_gcups else None,
                    "definition": "peak = 148 SMs x SM clock x (2 x R_dpx / 6) cell updates per SM-cycle: 3 packed DPX "
                                  "instructions per cell (SURVEY.md 8(d)), R_dpx = measured VIADDMNMX.U16x2 issue rate; "
                                  "achieved = useful (unpadded) cells of this GPU's shard / CUDA-event time of its first-stage launches",
                    "peak_cells_per_sm_clk": peak_cells_clk, "r_dpx_thread_instr_per_sm_clk": r_dpx,
                    "achieved_cells_per_busy_sm_clk": busy_cells_clk,
                    "sm_busy_fraction": sm_cycles / (n_sms * sm_mhz * 1e3 * score_ms) if score_ms else None,
                    "kernel_issue_bound_cells_per_sm_clk": 2.0 * r_dpx / 4.5,
                    "frac_of_kernel_issue_bound": (achieved_gcups / (2.0 * r_dpx / 4.5 * n_sms * sm_mhz * 1e-3)) if achieved_gcups else None,
                    "sm_mhz": sm_mhz, "padded_over_useful_cells": tm["padded_cells"] / max(local_cells, 1),
                    "calibration": cal, "peak_source": "calibrated on this GPU in this run (osw_calibrate)",
                    "hbm": {"achieved_GBps": (tm["db_stream_bytes"] + tm["bound_bytes"]) / (score_ms / 1e3) / 1e9 if score_ms else None,
                            "peak_GBps": json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0,
                            "algorithmic_bytes_per_cell": (tm["db_stream_bytes"] + tm["bound_bytes"]) / max(local_cells, 1),
                            "algorithmic_bytes_per_launch": (tm["db_stream_bytes"] + tm["bound_bytes"]) / max(tm["score_launches"], 1),
                            "what": "column stream (1 B per column per pass) + bottom rows handed between passes (8 B per column each way)"},
                    "traffic": None}
        # measured DRAM bytes per first-stage launch (ncu, tools/gpu_profile.sh): reported only when the
        # capture on file is of THIS build's kernels and of this workload
        tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tr_path):
            tr = json.load(open(tr_path))
            if tr.get("kernel_tag") == kernel_tag() and tr.get("config") == args.config and tr.get("n_gpus", 1) == world and not args.seqs and not qlens:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr["source"]
        out = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "u16x2 (packed 16-bit DPX) + int32 re-score", "data": "synthetic",
               "config": workload_config(world, args.config, args.seqs or CONFIGS[args.config]["n"], qlens or CONFIGS[args.config]["queries"]),
               "cells_per_step": cells_per_step, "wall_ms_per_step": e2e_s * 1e3 / args.steps,
               "timing": "one loop of K steps: each step copies the database from pinned host memory to HBM, then searches; value = "
                         "cells / CUDA-event time of the searches (max over ranks), e2e = cells / wall clock of the loop (max over ranks)",
               "clocks": clocks,
               "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d * world if world > 1 else h2d, "d2h_bytes_per_step": d2h * world},
               "gpu_launches": int(sum(x["launches"] for x in tms)) * world,
               "rescored_pairs_per_step": tm["rescored_pairs"],
               "breakdown_ms": {"score": score_ms, "rescore": tm["rescore_ms"], "topr": tm["topr_ms"]},
               "roofline": roofline, "shard": st, "build": kernel_tag(),
               "setup_seconds": {"generate_database": round(t_gen, 2), "osw_db_load_layout_and_first_upload": round(t_load, 2)},
               "verified": verified, "single_process_multi_gpu_ok": sp_ok,
               "extra": {"configs": extra}}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            seqs = cpu_sample(args.ref_sample, args.config)
            qs = [queries.query(i) for i in range(queries.n)]
            g, kind, secs, build = reference_cpu_gcups(seqs, qs, cores, run)
            out["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "build": build, "seconds": secs,
                                   "sample": sample_text(qs, seqs, args.config)}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
