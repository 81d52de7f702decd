#!/usr/bin/env python3
"""bench.py - GCUPS of the Smith-Waterman database search hot path on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     the reference's own host AVX2 path

A step = one complete search of the workload: 20 protein queries (lengths 144..5478, sum
41 750) against a Swiss-Prot-shaped synthetic database (config 2 of BASELINE.json: ~570k
sequences / ~205 M residues, log-normal lengths) PER GPU, BLOSUM62, gap 10/2, top 10.
With N GPUs the database is N times that size and is dealt chunk by chunk to the ranks
(weak scaling, no data-path collective; only the r hits per query are gathered).
GCUPS = sum(query lengths) * sum(database residues) / seconds / 1e9 - the reference's own
definition (HybridSearch.c:1227), unpadded.
"""
import argparse
import ctypes as C
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
SEQS_PER_GPU = 570_000
MU, SIGMA = 5.706, 0.6            # ln-length: mean 360 residues
SEED = 20261018
MATRIX, GAP_OPEN, GAP_EXTEND, TOP = "blosum62", 10, 2, 10
AA20 = "ACDEFGHIKLMNPQRSTVWY"
AA20_CODE = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)


def synth_lib():
    path = os.path.join(ROOT, "tools", "libosw_synth.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ROOT, "tools"])
    L = C.CDLL(path)
    L.osw_synth_lengths.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    L.osw_synth_codes.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    return L


def make_database(n_seqs, seed=SEED):
    """Canonical (length-sorted) synthetic database: (lengths u16 sorted, offsets u64, codes u8, perm)."""
    S = synth_lib()
    lens = np.empty(n_seqs, dtype=np.uint16)
    S.osw_synth_lengths(n_seqs, MU, SIGMA, 10, 65535, seed, lens.ctypes.data)
    perm = np.argsort(lens, kind="stable").astype(np.uint64)      # canonical order
    slens = np.ascontiguousarray(lens[perm])
    off = np.zeros(n_seqs + 1, dtype=np.uint64)
    off[1:] = np.cumsum(slens, dtype=np.uint64)
    codes = np.empty(int(off[-1]), dtype=np.uint8)
    S.osw_synth_codes(n_seqs, slens.ctypes.data, off.ctypes.data, perm.ctypes.data, seed, 0, codes.ctypes.data)
    return slens, off, codes, perm


def make_queries(seed=SEED):
    S = synth_lib()
    lens = np.array(QUERY_LENGTHS, dtype=np.uint16)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens, dtype=np.uint64)
    codes = np.empty(int(off[-1]), dtype=np.uint8)
    S.osw_synth_codes(len(lens), lens.ctypes.data, off.ctypes.data, None, seed ^ 0x51, 0, codes.ctypes.data)
    return [codes[int(off[i]):int(off[i + 1])] for i in range(len(lens))]


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
# reference arm / cpu baseline: the reference's own host AVX2 path (oracle/_ref/oswald_ref), or
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


def reference_cpu_gcups(sample_seqs, queries, cores):
    """Times the reference's host path on (queries x sample).  Returns (gcups, kind, seconds)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "oswald_ref")
    q_total = sum(len(q) for q in queries)
    d_total = sum(len(s) for s in sample_seqs)
    if os.path.exists(ref):
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
                            "-p", "%.3f" % pct, "-r", str(TOP), "-s", MATRIX, "-g", str(GAP_OPEN), "-e", str(GAP_EXTEND)],
                           cwd=tmp, check=True, capture_output=True, env=env, timeout=3000)
            t_cpu, t_work = [float(x) for x in open(os.path.join(tmp, "timing.txt")).read().split()]
        secs = t_cpu + t_work
        return q_total * d_total / secs / 1e9, "reference", secs
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    off = np.zeros(len(sample_seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in sample_seqs])
    q_off = np.zeros(len(queries) + 1, dtype=np.uint32)
    q_off[1:] = np.cumsum([len(q) for q in queries])
    t0 = time.time()
    O.search(np.concatenate(queries), q_off, np.concatenate(sample_seqs), off, O.matrix(MATRIX), GAP_OPEN, GAP_EXTEND, threads=cores)
    secs = time.time() - t0
    return q_total * d_total / secs / 1e9, "port", secs


def cpu_sample(n_sample, seed=SEED):
    slens, off, codes, _ = make_database(n_sample, seed + 1)
    return [codes[int(off[i]):int(off[i + 1])] for i in range(n_sample)]


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    queries = make_queries()
    n_sample = args.ref_sample
    seqs = cpu_sample(n_sample)
    vals, secs_all = [], []
    kind = "port"
    for it in range(args.warmup + args.steps):
        g, kind, secs = reference_cpu_gcups(seqs, queries, cores)
        if it >= args.warmup:
            vals.append(g)
            secs_all.append(secs)
    v = float(np.mean(vals))
    sample = "%d queries (sum %d) x %d-sequence / %d-residue sample of the same synthetic database" % (
        len(queries), sum(len(q) for q in queries), n_sample, sum(len(s) for s in seqs))
    print(json.dumps({
        "impl": "reference", "metric": "GCUPS", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(secs_all)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8/int16/int32 host SIMD cascade", "data": "synthetic",
        "config": workload_config(args.gpus, args.seqs_per_gpu),
        "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(n_gpus, seqs_per_gpu):
    return {"workload": "config 2 (Swiss-Prot-sized synthetic DB per GPU, 20 queries 144-5478, BLOSUM62 10/2, top 10)",
            "sequences": seqs_per_gpu * n_gpus, "length_distribution": "log-normal mu=%.3f sigma=%.1f clipped [10,65535]" % (MU, SIGMA),
            "query_lengths": QUERY_LENGTHS, "matrix": MATRIX, "gap_open": GAP_OPEN, "gap_extend": GAP_EXTEND, "top": TOP,
            "sharding": "chunks of ~8192 residues (2048 and 512 for the shortest sequences, taken last) dealt round-robin to %d GPU(s)" % n_gpus,
            "l2": "inputs larger than L2 (database stream >= 200 MB per GPU, score matrix 45 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs-per-gpu", type=int, default=SEQS_PER_GPU)
    ap.add_argument("--ref-sample", type=int, default=100000, help="sequences in the CPU baseline's sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernels", type=int, default=3, help="kernel mask (1|2 = default, 2 = 32-bit only)")
    ap.add_argument("--chunk-cols", type=int, default=0, help="residues per chunk (0 = library default)")
    ap.add_argument("--window-mb", type=int, default=0, help="stream the database through two device windows of this size (experiments; 0 = resident)")
    ap.add_argument("--query-lengths", default="", help="comma-separated subset/alternative query lengths (experiments; default: the 20 standard lengths)")
    args = ap.parse_args()
    if args.query_lengths:
        global QUERY_LENGTHS
        QUERY_LENGTHS = [int(x) for x in args.query_lengths.split(",")]
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
    from oswald_b200.host import calibrate, merge_hits
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload --------------------------------------------------------------------------
    n_total = args.seqs_per_gpu * world
    slens, off, codes, _ = make_database(n_total)
    db = ob.Database(codes, off)
    queries = ob.Queries.from_list(make_queries())
    mat = ob.matrix(MATRIX)
    q_total, d_total = queries.total_length, db.n_residues

    s = ob.Searcher(devices=[local_rank])
    s.set_kernels(args.kernels)
    if args.window_mb:
        s.set_device_window(args.window_mb << 20)
    s.load_db(db, shard_rank=rank, shard_count=world, max_chunk_residues=args.chunk_cols)
    st = s.stats()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def gather_and_merge(hits):
        if dist is None:
            return hits
        parts = [None] * world
        dist.all_gather_object(parts, hits)
        return [merge_hits([p[q] for p in parts], TOP) for q in range(queries.n)] if rank == 0 else None

    # ---- warm-up ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        hits, tm = s.search(queries, mat, GAP_OPEN, GAP_EXTEND, top=TOP)
    # ---- timed: K steps, database resident in HBM -----------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    dev_ms, tms = 0.0, []
    for _ in range(args.steps):
        hits, tm = s.search(queries, mat, GAP_OPEN, GAP_EXTEND, top=TOP)
        dev_ms += tm["device_ms"]
        tms.append(tm)
    barrier()
    wall_s = time.time() - t0
    sampler.stop_flag.set()
    # ---- timed: K steps end to end (host buffers: database H2D + search + hits D2H + merge) --
    barrier()
    t1 = time.time()
    h2d = 0
    for _ in range(args.steps):
        h2d = s.upload_db() + int(queries.residues.nbytes + queries.offsets.nbytes + mat.nbytes)
        hits, tm_e = s.search(queries, mat, GAP_OPEN, GAP_EXTEND, top=TOP)
        merged = gather_and_merge(hits)
    barrier()
    e2e_s = time.time() - t1
    d2h = queries.n * TOP * 8
    # max over ranks of the device time
    t = torch.tensor([dev_ms, wall_s, e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_s, e2e_s = [float(x) for x in t.cpu()]

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
        roofline = {"bound": "alu", "kernel": "sw_u16_kernel (packed 16-bit DPX)",
                    "achieved": achieved_gcups, "peak": peak_gcups, "unit": "GCUPS",
                    "frac": (achieved_gcups / peak_gcups) if achieved_gcups else None,
                    "definition": "peak = 148 SMs x SM clock x (2 x R_dpx / 6) cell updates per SM-cycle: 3 packed DPX "
                                  "instructions per cell (SURVEY.md 8(d)), R_dpx = measured VIADDMNMX.U16x2 issue rate; "
                                  "achieved = useful (unpadded) cells / CUDA-event time of the first-stage launches",
                    "peak_cells_per_sm_clk": peak_cells_clk, "r_dpx_thread_instr_per_sm_clk": r_dpx,
                    "achieved_cells_per_busy_sm_clk": busy_cells_clk,
                    "sm_busy_fraction": sm_cycles / (n_sms * sm_mhz * 1e3 * score_ms) if score_ms else None,
                    "kernel_issue_bound_cells_per_sm_clk": 2.0 * r_dpx / 4.5,
                    "frac_of_kernel_issue_bound": (achieved_gcups / (2.0 * r_dpx / 4.5 * n_sms * sm_mhz * 1e-3)) if achieved_gcups else None,
                    "sm_mhz": sm_mhz, "padded_over_useful_cells": tm["padded_cells"] / max(local_cells, 1),
                    "calibration": cal, "peak_source": "calibrated on this GPU in this run (osw_calibrate)",
                    "hbm": {"achieved_GBps": (tm["db_stream_bytes"] + tm["bound_bytes"]) / (score_ms / 1e3) / 1e9 if score_ms else None,
                            "peak_GBps": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0,
                            "algorithmic_bytes_per_cell": (tm["db_stream_bytes"] + tm["bound_bytes"]) / max(local_cells, 1),
                            "algorithmic_bytes_per_launch": (tm["db_stream_bytes"] + tm["bound_bytes"]) / max(tm["score_launches"], 1),
                            "what": "column stream (1 B per column per pass) + bottom rows handed between passes (8 B per column each way)"},
                    "traffic": None}
        # measured DRAM bytes per first-stage launch of this workload (ncu, tools/gpu_profile.sh), when on file
        tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tr_path) and args.seqs_per_gpu == SEQS_PER_GPU and not args.query_lengths:
            tr = json.load(open(tr_path))
            roofline["traffic"] = tr["dram_bytes_per_launch"]
            roofline["traffic_source"] = tr["source"]
        out = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u16x2 (packed 16-bit DPX) + int32 re-score", "data": "synthetic",
               "config": workload_config(world, args.seqs_per_gpu),
               "cells_per_step": cells_per_step, "wall_ms_per_step": wall_s * 1e3 / args.steps,
               "clocks": clocks,
               "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(sum(x["launches"] for x in tms)),
               "rescored_pairs_per_step": tm["rescored_pairs"],
               "breakdown_ms": {"score": score_ms, "rescore": tm["rescore_ms"], "topr": tm["topr_ms"]},
               "roofline": roofline, "shard": st}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            seqs = cpu_sample(args.ref_sample)
            g, kind, secs = reference_cpu_gcups(seqs, make_queries(), cores)
            out["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "seconds": secs,
                                   "sample": "%d queries (sum %d) x %d-sequence / %d-residue sample of the same synthetic database" % (
                                       queries.n, q_total, len(seqs), sum(len(x) for x in seqs))}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
