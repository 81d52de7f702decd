// sw_t16.cu - TRANSPOSED form of the first stage: for short queries, above all against small databases.
//
// sw_u16_kernel walks a database sequence column by column, one dependent step per residue, with
// the query's rows spread over G lanes.  With a short query that walk is the whole cost of a small
// database (BASELINE.json config 1: 5 000 pairs of sequences for 9 472 lane groups, and a launch that
// lasts as long as the 2 647 steps along its longest pair), and with ONE query the two 16-bit halves
// of a word have to score two database sequences, which costs a second table read per row quad
// (DESIGN.md 10).  Here the problem is turned round:
//   * ROWS are database residues: a task is one PAIR of neighbouring sequences (low / high halves of
//     the packed words) against one query; lane t of a warp holds R <= 8 consecutive rows of the pair,
//     a warp 32 R rows (a "block" of the pair); longer pairs take several blocks, handing the bottom
//     row (H, F per column) from block to block through shared memory;
//   * COLUMNS are the query's residues: m + 31 steps per block whatever the sequence length, lane t one
//     column behind lane t - 1 (bottom row by two warp shuffles);
//   * the substitution scores of a block come from a table the warp builds for it in shared memory,
//     P[query residue][row] = M[q][a_row] | M[q][b_row] << 16 - ONE conflict-free 128-bit read per four
//     rows and step for both sequences of the pair (the table is indexed by the column's residue, and
//     here that is the query's, which all rows share);
//   * a CTA has 16 warps with tables for 4 rows per lane (queries up to about 250 residues: four warps
//     per scheduler hide the latency of the row chain), or 8 warps with 8 rows per lane (up to 1024);
//   * long pairs are shared by a GANG of 16, 8 or 4 warps of the CTA: block k goes to warp k mod g, the
//     blocks run at the same time, each about 32 columns behind the one above it, and read the bottom
//     row of the block above from a double-buffered ring whose entries carry (task, block) tags - a
//     reader polls the entry it needs, no barrier and no fence in the sweep.  The host sorts the pairs
//     into four classes by length (osw_t16_plan); a gang that finds its class's queue empty splits
//     and goes on with the next class, so one launch runs all of them, longest first.
// Same arithmetic as sw_u16_kernel (biased unsigned 16-bit halves, DPX add-max / max3, flag at a biased
// maximum of 65 504 and above); E and F swap roles under transposition, the recurrence is symmetric in
// them (reference HybridSearch.c:842-913), so every H is the same number.
#include "sw_t16.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

namespace osw_t16 {

constexpr uint32_t FLAG_THRESHOLD = 65504;
constexpr int RMAX = 8;                                  // rows per lane (8-warp CTAs; 4 in 16-warp CTAs)
constexpr int QUAD_BYTES = 32 * 16;                      // a row quad of one query residue: [lane 32] x 16 bytes
// per warp: [query residue 24][row quads rmax / 4][lane 32] x 16 bytes
constexpr __host__ __device__ int table_bytes(int rmax) { return 24 * (rmax / 4) * QUAD_BYTES; }
constexpr int Q_STRIDE_PAD = 64;                         // column-table entries per query beyond its length (31 before, 33 after)
constexpr int MT_BYTES = 24 * 48 + 128;                  // [database residue][query residue] 16-bit scores (+ slack)

struct TArgs {
    OswT16Params p;
    uint32_t class_begin[OSW_T16_CLASSES + 1];
    uint32_t n_pairs;
    uint32_t q_cols, m_pad;
    int      rmax;                 // rows per lane at most: 4 or 8
    uint32_t bias2, nge2, goe2, bias;
};

__device__ __forceinline__ uint4 lds_volatile(uint32_t addr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(x), "r"(y), "r"(z), "r"(w));
}
__device__ __forceinline__ void gang_barrier(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n_threads) : "memory");
}

// One block of a task: R rows per lane against the m columns of the query.  tbl = this lane's slot of the
// warp's table, qcol[s] = table offset of the column this lane works on at step s.
// IO = 0: the block is the whole pair.  IO = 1: a lone warp works through the pair's blocks; it leaves a block's
// bottom row (H, F per column, 8 bytes) in its ring and reads it back, in place, in the next block.  IO = 2: a
// gang's blocks run at the same time: 16-byte ring entries (H, F, task tag, block tag) that the reader polls.
template <int R, int IO>
__device__ __forceinline__ uint32_t sweep(const unsigned char *tbl, const uint16_t *qcol, const int m,
                                          const bool has_in, const bool has_out, unsigned char *ring_in, unsigned char *ring_out,
                                          const uint32_t tag_task, const uint32_t blk,
                                          const int lane, const uint32_t B2, const uint32_t NGE, const uint32_t GOE2, uint32_t C) {
    uint32_t E[R], Hl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { E[r] = B2; Hl[r] = B2; }
    uint32_t diag = B2, Hbot = B2, Fbot = B2;
    const bool out_lane = IO != 0 && has_out && lane == 31;
    const int n_steps = m + 31;
    uint32_t qo1 = qcol[1];
    uint4 v0 = *reinterpret_cast<const uint4 *>(tbl + qcol[0]), v1 = v0;
    if (R > 4) v1 = *reinterpret_cast<const uint4 *>(tbl + qcol[0] + QUAD_BYTES);
    // IO = 1
    const uint2 *rin = reinterpret_cast<const uint2 *>(ring_in);
    uint2 *rout = reinterpret_cast<uint2 *>(ring_out);
    uint2 top = make_uint2(B2, B2);
    if (IO == 1 && has_in) top = rin[0];
    // IO = 2
    const uint32_t gin = IO == 2 ? (uint32_t)__cvta_generic_to_shared(ring_in) : 0u, gout = IO == 2 ? (uint32_t)__cvta_generic_to_shared(ring_out) : 0u;
    uint4 ent = make_uint4(B2, B2, tag_task, blk);
    if (IO == 2 && has_in) ent = lds_volatile(gin);
    // One step; LOAD: lane 0's top-row entry of the next column is read (IO = 1: columns below m - 1 only - past
    // the query's end lane 0 keeps the last column's entry, whose H the block above has counted already; anything
    // else could raise the maximum), OUT: lane 31 is at a column >= 0 and leaves its bottom row.
    auto step = [&](const int s, auto load_c, auto out_c) {
        constexpr bool LOAD = decltype(load_c)::value, OUT = decltype(out_c)::value;
        // the next steps' reads first: the column entry two steps ahead, the table words one step ahead
        const uint32_t qo2 = qcol[s + 2];
        const uint4 n0 = *reinterpret_cast<const uint4 *>(tbl + qo1);
        uint4 n1 = n0;
        if (R > 4) n1 = *reinterpret_cast<const uint4 *>(tbl + qo1 + QUAD_BYTES);
        uint2 top_next = top;
        if (IO == 1 && LOAD && has_in) top_next = rin[s + 1];
        uint4 ent_next = ent;
        if (IO == 2 && has_in) ent_next = lds_volatile(gin + 16u * (uint32_t)min(s + 1, m - 1));
        // the row above: the lane above's outputs of the previous step, i.e. for this column
        uint32_t Hup = __shfl_up_sync(0xffffffffu, Hbot, 1);
        uint32_t Fup = __shfl_up_sync(0xffffffffu, Fbot, 1);
        if (IO == 2) {
            if (has_in)
                while (ent.z != tag_task || ent.w != blk) ent = lds_volatile(gin + 16u * (uint32_t)min(s, m - 1));     // (written by the block above, a few dozen steps ahead)
            top = make_uint2(ent.x, ent.y);
        }
        if (lane == 0) { Hup = top.x; Fup = top.y; }
        uint32_t F = Fup, d = diag;
        const uint32_t sc[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t Hprev = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t tt = __viaddmax_u16x2(d, sc[r], E[r]);
            const uint32_t H = __vimax3_u16x2(tt, F, B2);
            const uint32_t u = H - GOE2;                 // (no borrow between the halves: H >= bias > go + ge)
            E[r] = __viaddmax_u16x2(E[r], NGE, u);
            F = __viaddmax_u16x2(F, NGE, u);
            d = Hl[r];                                   // H of this row in the previous column = the next row's diagonal
            Hl[r] = H;
            if (r & 1) C = __vimax3_u16x2(C, Hprev, H); else Hprev = H;
        }
        if (R & 1) C = __vmaxu2(C, Hprev);
        diag = Hup;                                      // H of the row above in this column = the next column's diagonal
        Hbot = Hl[R - 1]; Fbot = F;
        if (IO == 1 && OUT && out_lane) rout[s - 31] = make_uint2(Hbot, Fbot);
        if (IO == 2 && out_lane && s >= 31) sts_volatile(gout + 16u * (uint32_t)(s - 31), Hbot, Fbot, tag_task, blk + 1);
        v0 = n0; v1 = n1; qo1 = qo2; top = top_next; ent = ent_next;
    };
    const std::true_type yes;
    const std::false_type no;
    if (IO == 1) {
        // three stretches of steps, so that no step has to test where it is: lane 31 before its first column,
        // (both ends inside the query), lane 0 past its last column
        const int b1 = min(31, m - 1), b2 = max(31, m - 1);
        int s = 0;
#pragma unroll 2
        for (; s < b1; ++s) step(s, yes, no);
        if (m > 32) {
#pragma unroll 2
            for (; s < b2; ++s) step(s, yes, yes);
        } else {
#pragma unroll 2
            for (; s < b2; ++s) step(s, no, no);
        }
#pragma unroll 2
        for (; s < n_steps; ++s) step(s, no, yes);
    } else {
#pragma unroll 2
        for (int s = 0; s < n_steps; ++s) step(s, yes, yes);
    }
    return C;
}

template <int R>
__device__ __forceinline__ uint32_t sweep_io(const int io, const unsigned char *tbl, const uint16_t *qcol, const int m,
                                             const bool has_in, const bool has_out, unsigned char *ring_in, unsigned char *ring_out,
                                             const uint32_t tag_task, const uint32_t blk,
                                             const int lane, const uint32_t B2, const uint32_t NGE, const uint32_t GOE2, const uint32_t C) {
    if (io == 0) return sweep<R, 0>(tbl, qcol, m, false, false, nullptr, nullptr, tag_task, blk, lane, B2, NGE, GOE2, C);
    if (io == 1) return sweep<R, 1>(tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C);
    return sweep<R, 2>(tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C);
}

// The warp's table for one block: P[q][row] = (M[q][a_row] & 0xffff) | M[q][b_row] << 16, from the 16-bit
// matrix rows mt[residue][q] (48 bytes each).  da / db = the lane's residues (padding code for rows past
// the end of a sequence or of the block).  Entry (q, quad) of lane l is 16 bytes at ((quads q + quad) 32 + l) 16.
__device__ __forceinline__ void build_table(unsigned char *tbl, const unsigned char *mt, const uint32_t (&da)[RMAX], const uint32_t (&db)[RMAX], const int R, const int quads) {
#pragma unroll
    for (int quad = 0; quad < 2; ++quad) {
        if (quad == 1 && R <= 4) break;
#pragma unroll
        for (int qc = 0; qc < 3; ++qc) {                  // eight query residues at a time
            uint4 wa[4], wb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                wa[r] = *reinterpret_cast<const uint4 *>(mt + da[4 * quad + r] * 48 + qc * 16);
                wb[r] = *reinterpret_cast<const uint4 *>(mt + db[4 * quad + r] * 48 + qc * 16);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t a4[4], b4[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    a4[r] = j == 0 ? wa[r].x : j == 1 ? wa[r].y : j == 2 ? wa[r].z : wa[r].w;
                    b4[r] = j == 0 ? wb[r].x : j == 1 ? wb[r].y : j == 2 ? wb[r].z : wb[r].w;
                }
                const int q = 8 * qc + 2 * j;
                *reinterpret_cast<uint4 *>(tbl + (quads * q + quad) * QUAD_BYTES) =
                    make_uint4(__byte_perm(a4[0], b4[0], 0x5410), __byte_perm(a4[1], b4[1], 0x5410), __byte_perm(a4[2], b4[2], 0x5410), __byte_perm(a4[3], b4[3], 0x5410));
                *reinterpret_cast<uint4 *>(tbl + (quads * (q + 1) + quad) * QUAD_BYTES) =
                    make_uint4(__byte_perm(a4[0], b4[0], 0x7632), __byte_perm(a4[1], b4[1], 0x7632), __byte_perm(a4[2], b4[2], 0x7632), __byte_perm(a4[3], b4[3], 0x7632));
            }
        }
    }
}

// Gang size of a class in a CTA of W warps: W, W / 2, W / 4 while that is at least 4 warps (the two rings of a
// gang take the ring space of four warps), then lone warps.
__host__ __device__ inline int gang_size(int W, int cls) { return cls < OSW_T16_CLASSES - 1 && (W >> cls) >= 4 ? W >> cls : 1; }

template <int RM>                 // rows per lane at most: 4 (16 warps) or 8 (8 warps)
__global__ void __launch_bounds__(RM == 4 ? 512 : 256, 1)
sw_t16_kernel(const TArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const OswT16Params &p = a.p;
    const int NT = blockDim.x, W = NT >> 5;
    constexpr int quads = RM / 4, tbl_bytes = table_bytes(RM);
    unsigned char *s_tables = smem;                                                        // [W][tbl_bytes]
    unsigned char *s_ring = s_tables + (size_t)W * tbl_bytes;                              // [W][m_pad] x 8 bytes
    unsigned char *s_mt = s_ring + (size_t)W * a.m_pad * 8;                                // [24][24] x 2 bytes
    uint16_t *s_q = reinterpret_cast<uint16_t *>(s_mt + MT_BYTES);                         // [q_cols]
    volatile uint32_t *s_task = reinterpret_cast<volatile uint32_t *>(s_q + ((a.q_cols + 7) & ~7u));   // [2][W]

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const long long t_start = clock64();
    // ---- per-launch tables: 16-bit matrix rows by database residue, column offsets of every query ----
    for (int i = tid; i < 24 * 24; i += NT) {
        const int d = i / 24, q = i % 24;
        reinterpret_cast<uint16_t *>(s_mt)[i] = (uint16_t)(int16_t)p.matrix[q * 32 + d];
    }
    const uint32_t q_pitch = (uint32_t)(quads * QUAD_BYTES);          // table bytes per query residue
    for (uint32_t i = tid; i < a.q_cols; i += NT) s_q[i] = (uint16_t)(OSW_PAD_CODE * q_pitch);
    for (uint32_t i = tid; i < (uint32_t)W * a.m_pad * 2; i += NT) reinterpret_cast<uint32_t *>(s_ring)[i] = 0u;
    __syncthreads();
    {
        const uint32_t q0 = p.q_off[0], total = p.q_off[p.nq] - q0;
        for (uint32_t j = tid; j < total; j += NT) {
            int lo = 0, hi = p.nq - 1;                      // the query residue j belongs to
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.q_off[mid] - q0 <= j) lo = mid; else hi = mid - 1; }
            s_q[j + (uint32_t)Q_STRIDE_PAD * lo + 31] = (uint16_t)(p.queries[q0 + j] * q_pitch);
        }
    }
    __syncthreads();

    const uint32_t B2 = a.bias2, NGE = a.nge2, GOE2 = a.goe2;
    unsigned char *tbl = s_tables + (size_t)w * tbl_bytes + lane * 16;
    const uint32_t ring_pitch = a.m_pad * 8;

    int cls = 0;
    uint32_t fetches = 0;
    for (;;) {
        // ---- next task of the gang's class; an empty class splits the gang ----
        const int g = gang_size(W, cls), w0 = w & ~(g - 1);
        uint32_t t;
        if (g == 1) {
            t = 0;
            if (lane == 0) t = atomicAdd(p.counters + cls, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
        } else {
            volatile uint32_t *slot = s_task + (fetches & 1u) * W + w0;
            if (w == w0 && lane == 0) *slot = atomicAdd(p.counters + cls, 1u);
            gang_barrier(1 + w0 / 4, 32 * g);                 // (gangs start on multiples of four warps: ids 1..4)
            t = *slot;
            ++fetches;
        }
        const uint32_t n_cls = (a.class_begin[cls + 1] - a.class_begin[cls]) * (uint32_t)p.nq;
        if (t >= n_cls) {
            if (cls == OSW_T16_CLASSES - 1) break;
            ++cls;
            continue;
        }
        const uint32_t rank = a.class_begin[cls] + t / (uint32_t)p.nq, q = t % (uint32_t)p.nq;
        const uint32_t pair = a.n_pairs - 1 - rank;
        const uint64_t sa = 2ull * pair, sb = sa + 1;
        const uint32_t la = p.seq_len[sa], lb = sb < p.n_seqs ? p.seq_len[sb] : 0u;
        const uint32_t L = max(la, lb);
        const int m = (int)(p.q_off[q + 1] - p.q_off[q]);
        if (L == 0 || m == 0) continue;                       // scores stay 0
        const uint8_t *ra = p.stream + p.seq_off[sa];
        const uint8_t *rb = p.stream + (lb ? p.seq_off[sb] : p.seq_off[sa]);
        const uint32_t rows32 = (L + 31) / 32;
        const uint32_t per_round = (uint32_t)(RM * g);
        const uint32_t n_blocks = (uint32_t)g * ((rows32 + per_round - 1) / per_round);
        const int R = (int)((rows32 + n_blocks - 1) / n_blocks);         // 1..rmax rows per lane, the same in every block
        const uint16_t *qcol = s_q + (p.q_off[q] - p.q_off[0]) + (uint32_t)Q_STRIDE_PAD * q + 31 - lane;
        const uint32_t tag_task = ((uint32_t)cls << 28) + t + 1u;
        const int io = n_blocks == 1 ? 0 : g == 1 ? 1 : 2;
        uint32_t C = B2;
        for (uint32_t blk = (uint32_t)(w - w0); blk < n_blocks; blk += (uint32_t)g) {
            uint32_t da[RMAX], db[RMAX];
            const uint32_t row0 = (blk * 32u + (uint32_t)lane) * (uint32_t)R;
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const uint32_t row = row0 + (uint32_t)r;
                da[r] = (r < R && row < la) ? (uint32_t)(__ldg(ra + row) & OSW_COL_CODE) : (uint32_t)OSW_PAD_CODE;
                db[r] = (r < R && row < lb) ? (uint32_t)(__ldg(rb + row) & OSW_COL_CODE) : (uint32_t)OSW_PAD_CODE;
            }
            __syncwarp();                                      // (the previous block's table reads are done)
            build_table(tbl, s_mt, da, db, R, quads);
            __syncwarp();
            const bool has_in = blk > 0, has_out = blk + 1 < n_blocks;
            // bottom-row rings: a lone warp hands over in place (8-byte entries); a gang alternates between two
            // rings of 16-byte entries laid over the ring space of its first four warps
            unsigned char *ring_in = s_ring + (size_t)w * ring_pitch, *ring_out = ring_in;
            if (g > 1) {
                ring_in = s_ring + (size_t)w0 * ring_pitch + (size_t)((blk - 1) & 1u) * 2 * ring_pitch;
                ring_out = s_ring + (size_t)w0 * ring_pitch + (size_t)(blk & 1u) * 2 * ring_pitch;
            }
            switch (R) {
            case 1: C = sweep_io<1>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            case 2: C = sweep_io<2>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            case 3: C = sweep_io<3>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            case 4: C = sweep_io<4>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            default:
                if (RM == 4) break;                              // (R <= 4 in this instance)
                switch (R) {
            case 5: C = sweep_io<5>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            case 6: C = sweep_io<6>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            case 7: C = sweep_io<7>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
            default: C = sweep_io<8>(io, tbl, qcol, m, has_in, has_out, ring_in, ring_out, tag_task, blk, lane, B2, NGE, GOE2, C); break;
                }
            }
        }
        // ---- the warp's maximum over its blocks; the scores of the pair are the maximum over the gang ----
#pragma unroll
        for (int o = 16; o; o >>= 1) C = __vmaxu2(C, __shfl_xor_sync(0xffffffffu, C, o));
        if (lane == 0) {
            const uint32_t lo = C & 0xffffu, hi = C >> 16;
            int32_t *row = p.scores + (size_t)q * p.n_seqs;
            atomicMax(row + sa, lo >= FLAG_THRESHOLD ? OSW_SCORE_FLAGGED : (int)(lo - a.bias));
            if (lb) atomicMax(row + sb, hi >= FLAG_THRESHOLD ? OSW_SCORE_FLAGGED : (int)(hi - a.bias));
        }
    }
    if (p.cycle_acc) {
        __syncthreads();
        if (tid == 0) atomicAdd(p.cycle_acc, (unsigned long long)(clock64() - t_start));
    }
}

// ---- host side: cost model, classes, launch ----
// Scheduler cycles one warp needs for a step of R rows (4.5 DPX instructions per row and about five more for the
// step, at one per two cycles, on a pipe that is 72 % busy with four warps per scheduler: measured, 64 cycles at
// R = 4 - profiles/r2_transposed_form.md), and the time a step takes from start to end (two shuffles, then the row
// chain; the scheduler's other warps take their turns).
inline double step_issue(int R) { return 2.8 * (4.5 * R + 5.0); }
inline double step_time(int R, int warps_per_scheduler) { return std::max(45.0 + 15.0 * R, warps_per_scheduler * step_issue(R)); }
inline void block_geometry(uint32_t rows32, int g, int rmax, uint32_t *n_blocks, int *R) {
    const uint32_t per_round = (uint32_t)(rmax * g);
    *n_blocks = (uint32_t)g * ((rows32 + per_round - 1) / per_round);
    *R = (int)((rows32 + *n_blocks - 1) / *n_blocks);
}

}  // namespace osw_t16

void osw_t16_histogram(const uint32_t *seq_len, uint64_t n_seqs, uint32_t *hist) {
    memset(hist, 0, (OSW_T16_MAX_ROWS32 + 1) * sizeof(uint32_t));
    for (uint64_t s = 0; s < n_seqs; s += 2) {
        const uint32_t L = std::max(seq_len[s], s + 1 < n_seqs ? seq_len[s + 1] : 0u);
        ++hist[std::min<uint32_t>((L + 31) / 32, OSW_T16_MAX_ROWS32)];
    }
}

void osw_t16_plan(const uint32_t *hist, uint64_t n_seqs, const uint32_t *q_off, int nq, int n_sms, double gang_fraction, OswT16Plan *plan) {
    using namespace osw_t16;
    memset(plan, 0, sizeof *plan);
    if (nq < 1 || n_seqs == 0) return;
    uint32_t m_max = 0;
    uint64_t m_sum = 0;
    for (int q = 0; q < nq; ++q) { m_max = std::max(m_max, q_off[q + 1] - q_off[q]); m_sum += q_off[q + 1] - q_off[q]; }
    if (m_sum == 0 || m_sum > 4096 || m_max > 1024) return;
    if ((n_seqs + 1) / 2 * (uint64_t)nq >= (1ull << 28)) return;          // (task numbers are 28-bit tags of the ring entries)
    const uint32_t q_cols = (uint32_t)m_sum + Q_STRIDE_PAD * (uint32_t)nq;
    const uint32_t m_pad = (m_max + 3) & ~3u;
    // 16 warps with 4 rows per lane when the rings of the query's length fit beside their tables, else 8 warps
    // with 8 rows per lane, else 8 with 4
    int W = 0, rmax = 0;
    size_t smem = 0;
    const int shapes[3][2] = {{16, 4}, {8, 8}, {8, 4}};
    for (const auto &sh : shapes) {
        smem = (size_t)sh[0] * table_bytes(sh[1]) + (size_t)sh[0] * m_pad * 8 + MT_BYTES + (size_t)((q_cols + 7) & ~7u) * 2 + 2 * sh[0] * sizeof(uint32_t);
        if (smem <= 227 * 1024) { W = sh[0]; rmax = sh[1]; break; }
    }
    if (!W) return;
    plan->warps = W; plan->rmax = rmax; plan->q_cols = q_cols; plan->m_max = m_max; plan->smem_bytes = smem;
    plan->n_pairs = (uint32_t)((n_seqs + 1) / 2);
    const int wps = W / 4;                // warps per scheduler
    // total work: every pair x query as a lone warp's task
    double work = 0;                      // ALU-pipe cycles
    uint64_t padded = 0;
    for (uint32_t r32 = 1; r32 <= OSW_T16_MAX_ROWS32; ++r32) {
        if (!hist[r32]) continue;
        uint32_t nb; int R;
        block_geometry(r32, 1, rmax, &nb, &R);
        for (int q = 0; q < nq; ++q) {
            const double steps = (double)(q_off[q + 1] - q_off[q]) + 31.0;
            work += (double)hist[r32] * nb * (steps * step_issue(R) + 300.0);          // + table build
            padded += (uint64_t)hist[r32] * nb * 32u * (uint32_t)R * 2u * (uint64_t)steps;
        }
    }
    const double t_thr = work / ((double)std::max(n_sms, 1) * 4.0) * (wps >= 4 ? 1.0 : 1.4) + 4000.0;      // (two warps per scheduler do not hide the row chain)
    // gangs: the smallest one that finishes a task in a fraction of the launch's time; classes are ranges
    // of the descending length order, so the gang size may only shrink along it
    const double steps_max = (double)m_max + 31.0;
    auto task_time = [&](uint32_t r32, int g) {
        uint32_t nb; int R;
        block_geometry(r32, g, rmax, &nb, &R);
        if (g == 1) return (double)nb * (steps_max * step_time(R, wps) + 1500.0);
        // (a gang's warps start 36 steps apart, and its steps poll the ring entries)
        return ((double)(nb / g) * steps_max + (g - 1) * 36.0 + 31.0) * step_time(R, wps) * 1.15 + (double)(nb / g) * 1500.0;
    };
    uint32_t count[OSW_T16_CLASSES] = {0, 0, 0, 0};          // pairs per class
    int cls_cur = 0;
    double t_longest = 0;
    for (uint32_t r32 = OSW_T16_MAX_ROWS32; r32 >= 1; --r32) {
        if (!hist[r32]) continue;
        int cls = OSW_T16_CLASSES - 1;                        // the smallest gang that is fast enough, not larger than the class before
        while (cls > cls_cur && task_time(r32, gang_size(W, cls)) > gang_fraction * t_thr) --cls;
        cls_cur = cls;
        if (t_longest == 0) t_longest = task_time(r32, gang_size(W, cls));
        count[cls] += hist[r32];
    }
    count[OSW_T16_CLASSES - 1] += hist[0];                    // pairs of empty sequences: skipped by the kernel
    plan->class_begin[0] = 0;
    for (int i = 0; i < OSW_T16_CLASSES; ++i) plan->class_begin[i + 1] = plan->class_begin[i] + count[i];
    plan->est_cycles = std::max(t_thr, t_longest + 4000.0);
    plan->padded_cells = padded;
}

int osw_launch_t16(const OswT16Params &p, const OswT16Plan &plan, int n_ctas, cudaStream_t st) {
    using namespace osw_t16;
    if (!plan.warps || n_ctas < 1 || p.nq < 1) return OSW_E_ARG;
    TArgs a;
    a.p = p;
    for (int i = 0; i <= OSW_T16_CLASSES; ++i) a.class_begin[i] = plan.class_begin[i];
    a.n_pairs = plan.n_pairs; a.q_cols = plan.q_cols; a.m_pad = (plan.m_max + 3) & ~3u; a.rmax = plan.rmax;
    const uint32_t goe = (uint32_t)p.gap_open_extend, ge = (uint32_t)p.gap_extend;
    const uint32_t B = goe + ge + 32u;
    a.bias = B; a.bias2 = B | (B << 16);
    const uint32_t nge = (0x10000u - ge) & 0xffffu;
    a.nge2 = nge | (nge << 16);
    a.goe2 = goe | (goe << 16);
    if (plan.rmax == 4) {
        if (cudaFuncSetAttribute(sw_t16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes) != cudaSuccess) return OSW_E_CUDA;
        sw_t16_kernel<4><<<n_ctas, 32 * plan.warps, plan.smem_bytes, st>>>(a);
    } else {
        if (cudaFuncSetAttribute(sw_t16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes) != cudaSuccess) return OSW_E_CUDA;
        sw_t16_kernel<8><<<n_ctas, 32 * plan.warps, plan.smem_bytes, st>>>(a);
    }
    return cudaGetLastError() == cudaSuccess ? OSW_OK : OSW_E_CUDA;
}

// ---- test hooks (CPU tests; not in include/oswald_cuda.h) ----
extern "C" int osw_t16_plan_probe(const uint32_t *seq_len, uint64_t n_seqs, const uint32_t *q_off, int nq, int n_sms, double gang_fraction,
                                  uint32_t *class_begin, int *gang_sizes, int *warps, int *rmax, double *est_cycles, uint64_t *padded_cells) {
    if (!seq_len || !q_off || !class_begin || !gang_sizes || !warps || !rmax) return OSW_E_ARG;
    uint32_t *hist = new uint32_t[OSW_T16_MAX_ROWS32 + 1];
    osw_t16_histogram(seq_len, n_seqs, hist);
    OswT16Plan plan;
    osw_t16_plan(hist, n_seqs, q_off, nq, n_sms, gang_fraction, &plan);
    delete[] hist;
    for (int i = 0; i <= OSW_T16_CLASSES; ++i) class_begin[i] = plan.class_begin[i];
    for (int i = 0; i < OSW_T16_CLASSES; ++i) gang_sizes[i] = plan.warps ? osw_t16::gang_size(plan.warps, i) : 0;
    *warps = plan.warps; *rmax = plan.rmax;
    if (est_cycles) *est_cycles = plan.est_cycles;
    if (padded_cells) *padded_cells = plan.padded_cells;
    return OSW_OK;
}
// Blocks and rows per lane of a pair whose longer sequence has `length` residues, in a gang of g warps.
extern "C" void osw_t16_block_geometry(uint32_t length, int g, int rmax, uint32_t *n_blocks, int *rows_per_lane) {
    osw_t16::block_geometry((length + 31) / 32, g, rmax, n_blocks, rows_per_lane);
}
