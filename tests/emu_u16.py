"""Executable model of the packed 16-bit kernel's dataflow (oswald_b200/csrc/cuda/sw_u16.cu).

TEST INFRASTRUCTURE.  Pure Python, tiny inputs only.  It follows the kernel step by step -
column stream with FIRST/LAST flags, per-lane mailbox messages {H, F, column max, residue+flags},
lane t working on column (step - t), biased unsigned 16-bit arithmetic with wrap-around,
per-pass bottom-row hand-over, flagging at 65504 - so that the scheme itself (not the CUDA
code) can be checked against the oracle on the CPU.  One 16-bit half is modelled (the two
halves of a word are independent).
"""
import numpy as np

FIRST, LAST, PAD = 0x20, 0x40, 23
THRESH = 65504
FLAGGED = 0x7FFFFFFF
M16 = 0xFFFF


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


def run_pass(stream, n_seqs, query, row0, G, R, mat, go, ge, bound_in, want_out, scores, chains=1):
    """One launch: rows row0 .. row0+G*R-1 of `query` against the chunk `stream`.  With
    chains > 1 every lane is `chains` consecutive stations of the systolic array (its row
    segments), each one column behind the previous one."""
    goe = go + ge
    B = goe + ge + 32
    nge = (0x10000 - ge) & M16
    if chains > 1:
        seg = segment_rows(R, chains)
        stations, first = [], row0
        for t in range(G):
            for n_rows in seg:
                stations.append([(int(query[first + r]) if first + r < len(query) else PAD) for r in range(n_rows)])
                first += n_rows
        rows = stations
        G = len(rows)
    else:
        rows = [[(int(query[row0 + t * R + r]) if row0 + t * R + r < len(query) else PAD) for r in range(R)] for t in range(G)]
    Hl = [[B] * len(rows[t]) for t in range(G)]
    E = [[B] * len(rows[t]) for t in range(G)]
    diag_top = [B] * G
    run = B
    seq = 0
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
            if lf & FIRST:
                Hl[t] = [B] * len(rows[t])
                E[t] = [B] * len(rows[t])
                diag_top[t] = B
            code = lf & 31
            F, diag = fup, diag_top[t]
            for r in range(len(rows[t])):
                sc = int(mat[rows[t][r] * 32 + code]) & M16
                tt = max((diag + sc) & M16, E[t][r])
                H = max(tt, F, B)
                u = (H - goe) & M16
                E[t][r] = max((E[t][r] + nge) & M16, u)
                F = max((F + nge) & M16, u)
                diag = Hl[t][r]
                Hl[t][r] = H
                cm = max(cm, H)
            diag_top[t] = hup
            if t == G - 1:
                run = max(run, cm)
                col = step - (G - 1)
                if want_out and 0 <= col < n:
                    bound_out[col] = (Hl[t][-1], F)
                if lf & LAST:
                    val = FLAGGED if run >= THRESH else run - B
                    scores[seq] = max(scores[seq], val)
                    seq += 1
                    run = B
            new_mail[t] = (Hl[t][-1], F, cm, lf)
        mail = new_mail
    assert seq == n_seqs
    return bound_out


def score_chunk(seqs, query, G, R, mat, go, ge, chains=1):
    """Scores of `query` against every sequence of one chunk (FLAGGED where the kernel would flag)."""
    stream = build_stream(seqs)
    n_seqs = sum(1 for s in seqs if len(s))
    scores = [0] * n_seqs
    passes = max(1, -(-len(query) // (G * R)))
    bound = None
    for p in range(passes):
        bound = run_pass(stream, n_seqs, query, p * G * R, G, R, mat, go, ge, bound, p + 1 < passes, scores, chains)
    return np.array(scores, dtype=np.int64)
