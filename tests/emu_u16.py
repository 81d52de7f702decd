"""Executable model of the packed 16-bit kernel's dataflow (oswald_b200/csrc/cuda/sw_u16.cu).

TEST INFRASTRUCTURE.  Pure Python, tiny inputs only.  It follows the kernel step by step -
column stream with FIRST/LAST flags, per-lane mailbox messages {H, F, column max, residue+flags},
station k working on column (step - k), biased unsigned 16-bit arithmetic with wrap-around,
per-pass bottom-row hand-over, lanes that start / end a query inside a pass, flagging at 65504 -
so that the scheme itself (not the CUDA code) can be checked against the oracle on the CPU.
One 16-bit half is modelled (the two halves of a word are independent); a lane with `chains`
row segments is `chains` consecutive stations.
"""
import ctypes as C

import numpy as np

FIRST, LAST, PAD = 0x20, 0x40, 23
THRESH = 65504
FLAGGED = 0x7FFFFFFF
M16 = 0xFFFF
LANE_START, LANE_EMIT = 1, 2


def build_stream(seqs):
    out = []
    for s in seqs:
        col = [int(c) & 31 for c in s]
        if col:
            col[0] |= FIRST
            col[-1] |= LAST
        out += col
    return out


def segment_rows(R, chains):
    """Row counts of the segments a lane's R rows are split into (seg_begin() in sw_u16.cu)."""
    q = R // 4
    begin = [(q * c + chains - 1) // chains for c in range(chains + 1)]
    return [4 * (begin[c + 1] - begin[c]) for c in range(chains)]


def run_pass(stream, n_seqs, stations, mat, go, ge, bound_in, want_out, scores):
    """One launch.  stations: list of dicts {rows: [codes], start: bool, emit: query index or None}.
    scores: dict query -> list of per-sequence scores (max-accumulated)."""
    goe = go + ge
    B = goe + ge + 32
    nge = (0x10000 - ge) & M16
    G = len(stations)
    Hl = [[B] * len(st["rows"]) for st in stations]
    E = [[B] * len(st["rows"]) for st in stations]
    diag_top = [B] * G
    run = [B] * G
    seq = [0] * G
    mail = [(B, B, B, PAD)] * G
    n = len(stream)
    bound_out = [None] * n if want_out else None
    for step in range(n + G - 1):
        new_mail = list(mail)
        for t in range(G):
            if t == 0:
                if step < n:
                    hb, fb = bound_in[step] if bound_in is not None else (B, B)
                    msg = (hb, fb, B, stream[step])
                else:
                    msg = (B, B, B, PAD)
            else:
                msg = mail[t - 1]
            hup, fup, cm, lf = msg
            if stations[t]["start"]:
                hup, fup, cm = B, B, B
            if lf & FIRST:
                Hl[t] = [B] * len(Hl[t])
                E[t] = [B] * len(E[t])
                diag_top[t] = B
            code = lf & 31
            F, diag = fup, diag_top[t]
            rows = stations[t]["rows"]
            for r in range(len(rows)):
                sc = int(mat[rows[r] * 32 + code]) & M16
                tt = max((diag + sc) & M16, E[t][r])
                H = max(tt, F, B)
                u = (H - goe) & M16
                E[t][r] = max((E[t][r] + nge) & M16, u)
                F = max((F + nge) & M16, u)
                diag = Hl[t][r]
                Hl[t][r] = H
                cm = max(cm, H)
            diag_top[t] = hup
            run[t] = max(run[t], cm)
            if t == G - 1 and want_out:
                col = step - (G - 1)
                if 0 <= col < n:
                    bound_out[col] = (Hl[t][-1], F)
            if (lf & LAST) and stations[t]["emit"] is not None:
                val = FLAGGED if run[t] >= THRESH else run[t] - B
                row = scores[stations[t]["emit"]]
                row[seq[t]] = max(row[seq[t]], val)
                seq[t] += 1
                run[t] = B
            new_mail[t] = (Hl[t][-1], F, cm, lf)
        mail = new_mail
    for t in range(G):
        assert stations[t]["emit"] is None or seq[t] == n_seqs
    return bound_out


def score_chunk(seqs, query, G, R, mat, go, ge, chains=1):
    """One query alone, rows dealt G*R per pass (the simplest plan)."""
    stream = build_stream(seqs)
    n_seqs = sum(1 for s in seqs if len(s))
    scores = {0: [0] * n_seqs}
    passes = max(1, -(-len(query) // (G * R)))
    seg = segment_rows(R, chains)
    bound = None
    for p in range(passes):
        stations, first = [], p * G * R
        for t in range(G):
            for c, n_rows in enumerate(seg):
                rows = [(int(query[first + r]) if first + r < len(query) else PAD) for r in range(n_rows)]
                stations.append({"rows": rows, "start": p == 0 and t == 0 and c == 0,
                                 "emit": 0 if (t == G - 1 and c == len(seg) - 1) else None})
                first += n_rows
        bound = run_pass(stream, n_seqs, stations, mat, go, ge, bound, p + 1 < passes, scores)
    return np.array(scores[0], dtype=np.int64)


# ---- the product's planner (plan.cu), through ctypes ----------------------------------------
class LaneDesc(C.Structure):
    _fields_ = [("query", C.c_uint32), ("q_len", C.c_uint32), ("row0", C.c_uint32), ("flags", C.c_uint32)]


class Pass(C.Structure):
    _fields_ = [("G", C.c_int), ("R", C.c_int), ("lane", (LaneDesc * 32) * 2), ("has_in", C.c_int), ("has_out", C.c_int),
                ("pair_db", C.c_int)]


PLAN_AUTO, PLAN_TWO_TRACK, PLAN_PAIR_DB = 0, 1, 2


def plan_passes(lib, q_lens, max_passes=256, mode=PLAN_TWO_TRACK, min_g=4):
    lib.osw_plan_passes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.osw_plan_passes.restype = C.c_int
    arr = (C.c_uint32 * len(q_lens))(*q_lens)
    out = (Pass * max_passes)()
    n = lib.osw_plan_passes(arr, len(q_lens), out, max_passes, mode, min_g)
    assert n >= 0
    return [out[i] for i in range(n)]


def build_pair_streams(seqs):
    """Pair-database mode (dbformat.c): sequences 2p and 2p+1 zipped, the shorter padded; returns
    the two per-half column streams (flags of the pair's columns on both) and the pair count."""
    a_cols, b_cols = [], []
    for p in range(0, len(seqs), 2):
        a = [int(c) & 31 for c in seqs[p]]
        b = [int(c) & 31 for c in seqs[p + 1]] if p + 1 < len(seqs) else []
        n = max(len(a), len(b))
        a += [PAD] * (n - len(a))
        b += [PAD] * (n - len(b))
        if n:
            a[0] |= FIRST; b[0] |= FIRST
            a[-1] |= LAST; b[-1] |= LAST
        a_cols += a
        b_cols += b
    return a_cols, b_cols, (len(seqs) + 1) // 2


def score_with_plan(passes, seqs, queries, mat, go, ge, chains=2):
    """All queries against one chunk, following the planner's passes for both halves."""
    n_seqs = sum(1 for s in seqs if len(s))
    if passes and passes[0].pair_db:
        assert n_seqs == len(seqs)
        sa, sb, n_pairs = build_pair_streams(seqs)
        out = np.zeros((len(queries), n_seqs), dtype=np.int64)
        for half, stream in ((0, sa), (1, sb)):
            scores = {q: [0] * n_pairs for q in range(len(queries))}
            _run_half(passes, half, stream, n_pairs, queries, mat, go, ge, chains, scores)
            for q in range(len(queries)):
                for p in range(n_pairs):
                    if 2 * p + half < n_seqs:
                        out[q, 2 * p + half] = scores[q][p]
        return out
    stream = build_stream(seqs)
    scores = {q: [0] * n_seqs for q in range(len(queries))}
    for half in (0, 1):
        _run_half(passes, half, stream, n_seqs, queries, mat, go, ge, chains, scores)
    return np.array([scores[q] for q in range(len(queries))], dtype=np.int64)


def _run_half(passes, half, stream, n_seqs, queries, mat, go, ge, chains, scores):
    if True:
        bound = None
        for p in passes:
            seg = segment_rows(p.R, chains)
            stations = []
            for t in range(p.G):
                d = p.lane[half][t]
                first = d.row0
                for c, n_rows in enumerate(seg):
                    q = queries[d.query] if d.q_len else []
                    rows = [(int(q[first + r]) if first + r < d.q_len else PAD) for r in range(n_rows)]
                    stations.append({"rows": rows, "start": bool(d.flags & LANE_START) and c == 0,
                                     "emit": d.query if (d.flags & LANE_EMIT) and d.q_len and c == len(seg) - 1 else None})
                    first += n_rows
            bound = run_pass(stream, n_seqs, stations, mat, go, ge, bound if p.has_in else None, bool(p.has_out), scores)
