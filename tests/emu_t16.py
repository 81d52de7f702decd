"""Executable model of the TRANSPOSED first stage's dataflow (oswald_b200/csrc/cuda/sw_t16.cu).

TEST INFRASTRUCTURE.  Pure Python, tiny inputs only.  It follows the kernel step by step - the rows of
a block are database residues (R per lane, 32 lanes), the columns the query's residues with 31 padding
columns before and after, lane t one column behind lane t - 1 (bottom row H, F handed down as the
shuffles do), blocks handing their bottom row on through a ring of m entries (in place), lane 0
re-reading the last column's entry once it is past the query's end, biased unsigned 16-bit arithmetic
with wrap-around, the flag at 65504 - so that the scheme itself (not the CUDA code) can be checked
against the oracle on the CPU.  The block geometry comes from the library's own planner hook
(osw_t16_block_geometry).  One 16-bit half is modelled: the two halves of a word are the two sequences
of a pair and do not interact, except that the block count follows the longer one.
"""
import ctypes as C

PAD = 23
THRESH = 65504
FLAGGED = 0x7FFFFFFF
M16 = 0xFFFF


def block_geometry(lib, length, gang, rmax):
    lib.osw_t16_block_geometry.argtypes = [C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    lib.osw_t16_block_geometry.restype = None
    nb, R = C.c_uint32(), C.c_int()
    lib.osw_t16_block_geometry(length, gang, rmax, C.byref(nb), C.byref(R))
    return nb.value, R.value


def score_half(lib, seq, pair_length, query, mat, go, ge, gang=1, rmax=4):
    """Score of `seq` (one half of a pair whose longer sequence has pair_length residues) against `query`."""
    m = len(query)
    if pair_length == 0 or m == 0:
        return 0
    goe = go + ge
    B = goe + ge + 32
    nge = (0x10000 - ge) & M16
    n_blocks, R = block_geometry(lib, pair_length, gang, rmax)
    assert 1 <= R <= rmax and n_blocks % gang == 0 and n_blocks * 32 * R >= pair_length
    cols = [PAD] * 31 + [int(c) for c in query] + [PAD] * 33
    ring = [None] * m                       # (H, F) of the block above, per column; rewritten in place
    best = B
    for blk in range(n_blocks):
        has_in, has_out = blk > 0, blk + 1 < n_blocks
        rows = [[int(seq[(blk * 32 + t) * R + r]) if (blk * 32 + t) * R + r < len(seq) else PAD for r in range(R)] for t in range(32)]
        Hl = [[B] * R for _ in range(32)]
        E = [[B] * R for _ in range(32)]
        diag = [B] * 32
        bot = [(B, B)] * 32                 # every lane's outputs of the previous step
        for s in range(m + 31):
            prev = list(bot)
            # lane 0 past the query's end keeps reading the last column's entry
            top = ring[min(s, m - 1)] if has_in else (B, B)
            for t in range(32):
                q = cols[31 + s - t]
                hup, fup = top if t == 0 else prev[t - 1]
                F, d = fup, diag[t]
                for r in range(R):
                    sc = int(mat[q * 32 + rows[t][r]]) & M16
                    tt = max((d + sc) & M16, E[t][r])
                    H = max(tt, F, B)
                    u = H - goe
                    E[t][r] = max((E[t][r] + nge) & M16, u)
                    F = max((F + nge) & M16, u)
                    d = Hl[t][r]
                    Hl[t][r] = H
                    best = max(best, H)
                diag[t] = hup
                bot[t] = (Hl[t][R - 1], F)
            if has_out and s >= 31:
                ring[s - 31] = bot[31]      # lane 31 is at column s - 31
    return FLAGGED if best >= THRESH else best - B
