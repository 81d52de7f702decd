#ifndef OSW_SW_U16_KERNEL_CUH
#define OSW_SW_U16_KERNEL_CUH
// sw_u16_kernel.cuh - the first-stage kernel: packed 16-bit Gotoh scoring with DPX instructions.
//
// Takes over the role of the reference's narrow stages (8-bit HybridSearch.c:831-913, 16-bit
// :937-1030, and the FPGA kernel device/sw.cl) - not their structure.
//
// Work decomposition
//   * Every 32-bit word carries two independent DP problems.  Normally both are against the SAME
//     database residue: the low half works on rows of query track 0, the high half on rows of
//     track 1 (plan.cu lays the queries end to end on the two tracks, each query starting on a
//     lane boundary).  A launch ("pass") covers the next `G*R` rows of both tracks against every
//     database chunk, and one shared-memory read of the pair profile
//         prof[residue][row] = (M[track0[row]][residue], M[track1[row]][residue])
//     serves two cell updates, bank-conflict free by construction: the table is laid out
//     [residue][quad of rows] with a 128-byte multiple as row pitch and an odd number of quads per
//     lane, so the 8 lanes of a quarter warp hit 8 different 16-byte bank groups whatever their
//     residues are.  In pair-database mode (template flag PD; single or lopsided query sets) both
//     halves work on the same query rows against two database sequences zipped in the pair
//     stream, and a row's score word is the sum of a low-half and a high-half table entry.
//     The table image is built once per pass in global memory (profile_build_kernel) and every CTA
//     brings it into shared memory with TMA bulk copies completing on an mbarrier.
//   * G lanes (4, 8, 16 or 32) form a systolic array over the query rows: lane t owns rows
//     t*R .. t*R+R-1 (H of the previous column and E in registers), swept as two independent
//     segments one column apart (two dependency chains per thread keep the DPX pipe fed).  A
//     chunk's column stream flows through the array; sequences follow one another without
//     draining it (the FIRST flag of a column restarts a lane's state).  With G = 32 one warp
//     sweeps the anti-diagonals of one sequence at a time - the intra-task path for very long
//     sequences; with G < 32 several sequences share a warp (inter-task).
//   * Per step a lane reads a 16-byte message {H, F of the row above, running column maximum,
//     residue(s)+flags}, sweeps its R rows, and writes the same message for the lane below through
//     a per-warp shared-memory mailbox.  Lane 0 of a group reads its messages from a ring the
//     group fills 32 columns ahead from the chunk stream (coalesced loads); when the plan has
//     several passes the ring also carries the previous pass's bottom row (H, F) and the last
//     lane stages this pass's bottom row in shared memory, flushed every 32 steps (in place: a
//     warp writes a chunk's columns behind the ones it still has to read).
//   * A lane that holds a query's first row replaces the message from above by "no row"; the
//     last lane of a query in the pass sees, per column, the maximum over the query's rows, keeps
//     the running maximum of the sequence and publishes it (atomicMax into the score matrix) at
//     the column flagged LAST.
//
// Arithmetic: unsigned 16-bit halves with a bias B = go + 2*ge + 32 (value v is stored as v+B):
//   t  = VIADDMNMX.U16x2(Hdiag, score, E)      max(Hdiag + s, E)   (the add wraps: s is two's complement)
//   H  = VIMNMX3.U16x2(t, F, B)                max(t, F, 0)
//   u  = H - (go+ge | go+ge << 16)             one 32-bit subtract (VIADD, co-issues with the DPX
//                                              pipe): no borrow between halves since H >= B >= go+ge
//   E  = VIADDMNMX.U16x2(E, -ge, u)            max(E - ge, u)
//   F  = VIADDMNMX.U16x2(F, -ge, u)
//   cm = VIMNMX3.U16x2(cm, H_even, H_odd)      every second row
// i.e. 4.5 DPX-pipe instructions per word = per two cell updates.  A sequence whose biased
// maximum reaches 65504 may have wrapped and is flagged for the 32-bit kernel (scores grow by at
// most 17 per cell, so a wrap cannot be missed).
#include "osw_internal.h"
#include <stdlib.h>
#include <string.h>

namespace osw_u16 {

constexpr uint32_t FLAG_THRESHOLD = 65504;      // biased maximum at/above which a pair is re-scored
constexpr int RING = 64;                        // ring entries per group (two halves of 32)

__host__ __device__ constexpr int pitch_quads(int R) { return (R / 4) | 1; }          // odd, >= R/4
__host__ __device__ constexpr int prof_quads(int G, int R) { return (G * pitch_quads(R) + 7) / 8 * 8; }
__host__ __device__ constexpr int prof_copies(int G) { return G == 4 ? 2 : 1; }
__host__ __device__ constexpr size_t prof_copy_bytes(int G, int R) { return (size_t)24 * prof_quads(G, R) * 16; }
__host__ __device__ constexpr int block_threads(int R) { return R > 32 ? 384 : 512; }
// (OSW_EXP_* macros: timing experiments only - each one breaks the results; never set in the product build.)
// A lane's R rows are swept as NUM_CHAINS independent segments (segment c works one column
// behind segment c-1): two dependency chains per thread keep the DPX pipe fed.
#ifndef OSW_NUM_CHAINS
#define OSW_NUM_CHAINS 2
#endif
constexpr int NUM_CHAINS = OSW_NUM_CHAINS;
__host__ __device__ constexpr int seg_begin(int R, int NC, int c) { return ((R / 4) * c + NC - 1) / NC; }   // first quad of segment c

// mailbox / ring traffic: ordered against __syncwarp
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
// profile reads: the profile is constant once built, so the compiler may schedule these freely
__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 lds128_const(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct KArgs {
    U16Params p;
    OswLaneDesc lane[2][32]; // [half][lane of group]: which query rows the lane holds
    uint32_t has_in, has_out;
    uint32_t pair_db;        // PD mode (host-side dispatch only)
    uint32_t bias2;          // B | B<<16
    uint32_t nge2;           // (-ge) & 0xffff, both halves
    uint32_t ngoe_word;      // -(goe | goe<<16) as a 32-bit two's complement
    uint32_t bias;           // B
    uint32_t n_ctas, warps;  // launch geometry of the scoring kernel (for the first-chunk deal)
};

constexpr uint32_t NO_GROUP = 0x0fffffffu;       // deal-table entry of a warp that takes no chunk

// One 16-byte profile entry (4 rows x one residue) - and its high-half twin in pair-database mode -
// written at its place in a table image starting at `base` (shared or global memory).
template <int G, int R, bool PD>
__device__ __forceinline__ void profile_entry(int idx, const KArgs &a, const int *s_mat, unsigned char *base) {
    constexpr int P = pitch_quads(R);
    constexpr int PITCH_B = prof_quads(G, R) * 16;
    constexpr int COPY_B = (int)prof_copy_bytes(G, R);
    constexpr int TABLE_B = COPY_B * prof_copies(G) + (prof_copies(G) > 1 ? 128 : 0);
    const U16Params &p = a.p;
    const int copy = idx / (24 * prof_quads(G, R));
    const int rem = idx % (24 * prof_quads(G, R));
    const int b = rem / prof_quads(G, R), slot = rem % prof_quads(G, R);
    const int tt = slot / P, k = slot % P;
    uint32_t w[4] = {0, 0, 0, 0};
    if (tt < G && k < R / 4) {
        const OswLaneDesc da = a.lane[0][tt], db = a.lane[1][tt];
        const uint8_t *qa = da.q_len ? p.queries + p.q_off[da.query] : p.queries;
        const uint8_t *qb = db.q_len ? p.queries + p.q_off[db.query] : p.queries;
        for (int r = 0; r < 4; ++r) {
            const uint32_t ra = da.row0 + 4 * k + r, rb = db.row0 + 4 * k + r;
            const int ca = ra < da.q_len ? qa[ra] : OSW_PAD_CODE;
            const int cb = rb < db.q_len ? qb[rb] : OSW_PAD_CODE;
            if (PD) w[r] = (uint32_t)s_mat[ca * 32 + b];            // track 0 only; split into halves below
            else w[r] = ((uint32_t)s_mat[ca * 32 + b] & 0xffffu) | ((uint32_t)s_mat[cb * 32 + b] << 16);
        }
    }
    // the second copy (G == 4) sits 64 bytes further modulo 128, i.e. 4 bank groups away
    unsigned char *dst = base + copy * (COPY_B + 64) + b * PITCH_B + slot * 16;
    if (PD) {
        *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0] & 0xffffu, w[1] & 0xffffu, w[2] & 0xffffu, w[3] & 0xffffu);
        *reinterpret_cast<uint4 *>(dst + TABLE_B) = make_uint4(w[0] << 16, w[1] << 16, w[2] << 16, w[3] << 16);
    } else {
        *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Builds the pass's profile table image once, in global memory; every CTA of the scoring kernel
// then brings it into shared memory with TMA bulk copies (cp.async.bulk + mbarrier).
template <int G, int R, bool PD>
__global__ void __launch_bounds__(256) profile_build_kernel(const KArgs a) {
    __shared__ int s_mat[24 * 32];
    for (int i = threadIdx.x; i < 24 * 32; i += 256) s_mat[i] = a.p.matrix[i];
    __syncthreads();
    const int n = prof_copies(G) * 24 * prof_quads(G, R);
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < n; idx += gridDim.x * 256)
        profile_entry<G, R, PD>(idx, a, s_mat, a.p.profile);
    // The first chunk group of every warp of the scoring kernel, when the launcher asked for a
    // static deal (a.p.first_table != nullptr):
    //  * express CTAs (K = a.p.express_ctas > 0; chosen by the host when the launch would last as
    //    long as its longest chunk): the 4K longest chunk groups go to warps 0-3 of CTAs 0..K-1, one
    //    warp per scheduler, and the other warps of those CTAs take nothing - a warp alone on its
    //    scheduler walks its columns at the latency of one step instead of sharing the issue slots
    //    with three others;
    //  * on a database of only a few chunks per warp (a.p.static_first) warp w of CTA b takes group
    //    w * CTAs + b of the remaining list, so that every SM starts with the same mix of long and
    //    short chunks (measured + 2..6 % there, - 2.4 % on a large database, which uses the counter);
    //  * every other warp starts at the counter like later fetches do (entry = NO_GROUP + 1).
    if (a.p.first_table && blockIdx.x == 0) {
        const uint32_t K = a.p.all_express ? a.n_ctas : min(a.p.express_ctas, a.n_ctas - 1);
        for (uint32_t i = threadIdx.x; i < a.n_ctas * a.warps; i += 256) {
            const uint32_t b = i / a.warps, w = i % a.warps;
            uint32_t g;
            if (b < K) g = w < 4 ? w * K + b : NO_GROUP;
            else if (a.p.static_first) g = 4u * K + w * (a.n_ctas - K) + (b - K);
            else g = NO_GROUP + 1;
            a.p.first_table[i] = g;
        }
    }
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier --------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    }
}

// PD = "pair database" mode: both halves work on the SAME query rows (track 0) against two
// different database sequences zipped in the pair stream; the score word of a row is the sum of a
// low-half table entry (first sequence's residue) and a high-half table entry (second one's).
// DEAL = the first chunk of every warp comes from the deal table (express CTAs, small databases).
// A template flag rather than a run-time test: the R = 40 row sweep uses every register it can
// get, and even a loop-carried flag in the fetch path changed the schedule ptxas finds for the
// sweep (- 0.8..1.7 % at config 2, measured against the previous build on the same box).
// PIPE = pipelined passes: the launch runs beside the launches of the passes before and after it, each on
// SMs of its own, all walking the same (longest) chunks a few dozen columns apart.  Before a warp reads the
// bottom row of a 32-column block it waits until the pass before has written it (progress_in), and after it
// has flushed its own bottom row of a block it says so (progress_out).  Only the G = 32, DEAL variants exist.
__device__ __forceinline__ void wait_progress(const uint32_t *counter, uint32_t need) {
    if ((threadIdx.x & 31) == 0) {
        uint32_t have;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(counter) : "memory"); } while (have < need);
    }
    __syncwarp();
}

template <int G, int R, int THREADS, bool PD, bool DEAL, bool PIPE = false>
__global__ void __launch_bounds__(THREADS, 1)
sw_u16_kernel(const KArgs a) {
    constexpr int WARPS = THREADS / 32;
    constexpr int GROUPS = 32 / G;              // groups per warp
    constexpr int P = pitch_quads(R);
    constexpr int PITCH_B = prof_quads(G, R) * 16;
    constexpr int COPY_B = (int)prof_copy_bytes(G, R);
    constexpr int EPL = 32 / G;                 // ring entries each lane fills per 32-column block
    constexpr int NC = NUM_CHAINS;              // independent row segments per lane

    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [profile copies][mailbox: WARPS*32 uint4][rings: WARPS*GROUPS*RING uint4][out rings: WARPS*32 uint2]
    unsigned char *s_prof = smem;
    constexpr int TABLE_B = COPY_B * prof_copies(G) + (prof_copies(G) > 1 ? 128 : 0);
    constexpr int PROF_B = TABLE_B * (PD ? 2 : 1);           // PD: low-half table, then high-half table
    uint4 *s_mail = reinterpret_cast<uint4 *>(smem + PROF_B);
    uint4 *s_ring = s_mail + WARPS * 32;
    uint2 *s_oring = reinterpret_cast<uint2 *>(s_ring + WARPS * GROUPS * RING);     // [WARPS][32] bottom rows of the last 32 steps
    __shared__ uint32_t s_chunk[WARPS];

    const U16Params &p = a.p;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int t = lane % G, grp = lane / G;
    long long clk0 = clock64();

    // ---- the pass's profile table (built once by profile_build_kernel): TMA bulk copies into shared memory
    __shared__ __align__(8) unsigned long long s_mbar;
    {
        const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
        if (threadIdx.x == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(mbar, (uint32_t)PROF_B);
            constexpr uint32_t PIECE = 32768;
            for (uint32_t off = 0; off < (uint32_t)PROF_B; off += PIECE) {
                const uint32_t n = (uint32_t)PROF_B - off < PIECE ? (uint32_t)PROF_B - off : PIECE;
                bulk_g2s((uint32_t)__cvta_generic_to_shared(s_prof) + off, p.profile + off, n, mbar);
            }
        }
        mbar_wait(mbar, 0);
    }

    const uint32_t prof_lane = (uint32_t)__cvta_generic_to_shared(s_prof) +
                               (prof_copies(G) > 1 ? (grp & 1) * (COPY_B + 64) : 0) + t * P * 16;
    const uint32_t mail_self = (uint32_t)__cvta_generic_to_shared(s_mail + wib * 32 + lane);
    const uint32_t mail_up = mail_self - 16;     // lane-1's mailbox (unused when t == 0)
    const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(s_ring + (wib * GROUPS + grp) * RING);
    const uint32_t oring_base = (uint32_t)__cvta_generic_to_shared(s_oring + wib * 32);      // (G == 32 when used)
    const uint32_t B2 = a.bias2, NGE = a.nge2, GOE2 = 0u - a.ngoe_word;
    constexpr uint32_t PAD_MSG = OSW_COL_PADBYTE | (OSW_COL_PADBYTE << 8);    // "no column": padding residue(s), no flags
    const bool multi_in = a.has_in != 0, has_out = a.has_out != 0;
    // what this lane is, per half: first lane of a query (its input from above is replaced by
    // "no row"), last lane of a query in this pass (it publishes that query's maximum)
    const uint32_t fa = a.lane[0][t].flags, fb = a.lane[1][t].flags;
    const uint32_t keep = ((fa & OSW_LANE_START) ? 0u : 0x0000ffffu) | ((fb & OSW_LANE_START) ? 0u : 0xffff0000u);
    const uint32_t emit = ((fa & OSW_LANE_EMIT) && a.lane[0][t].q_len ? 1u : 0u) | ((fb & OSW_LANE_EMIT) && a.lane[1][t].q_len ? 2u : 0u);
    const uint32_t last_mask = emit ? OSW_COL_LAST : 0u;

    // Chunk hand-out: from a counter, longest first.  With DEAL a warp's FIRST chunk comes from the
    // table profile_build_kernel fills (express CTAs; small databases: see there).
    bool first_fetch = DEAL;
    for (;;) {
        // ---- fetch one chunk per group ---------------------------------------------------
        if (DEAL) {
            if (lane == 0) {
                uint32_t g = first_fetch ? p.first_table[blockIdx.x * WARPS + wib] : NO_GROUP + 1;
                if (g > NO_GROUP) g = p.dyn_base + atomicAdd(p.chunk_counter, 1u);
                s_chunk[wib] = g * GROUPS;
            }
            first_fetch = false;
        } else {
            if (lane == 0) s_chunk[wib] = atomicAdd(p.chunk_counter, (uint32_t)GROUPS);
        }
        __syncwarp();
        const uint32_t cbase = p.chunk_first + s_chunk[wib];
        __syncwarp();
        if (cbase >= p.chunk_end) break;
        const uint32_t ci = cbase + grp;
        const bool have = ci < p.chunk_end;
        osw_chunk ck;
        if (have) ck = p.chunks[ci];
        else { ck.stream_off = 0; ck.n_cols = 0; ck.n_seqs = 0; ck.seq0 = 0; ck.canon0 = 0; ck.pair_off = 0; ck.n_pair_cols = 0; }
        const uint32_t n_cols = PD ? ck.n_pair_cols : ck.n_cols;
        const uint64_t col0 = PD ? ck.pair_off : ck.stream_off;       // first column in the (pair) stream
        uint32_t steps = have ? n_cols + NC * G - 1 : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        const uint32_t n_blocks = (steps + 31) / 32;
        const uint8_t *col_src = PD ? p.pair_stream + 2 * (col0 - p.stream_col0 + t * EPL) : p.stream + (col0 - p.stream_col0) + t * EPL;
        const uint2 *bnd_src = multi_in ? p.bound + (col0 - p.bound_col0) + t * EPL : nullptr;
        constexpr uint32_t COL_ALIGN = PD ? 64 : OSW_CHUNK_ALIGN;
        const uint32_t cols_padded = have ? (n_cols + COL_ALIGN - 1) / COL_ALIGN * COL_ALIGN : 0;

        // ring fill helpers: block b covers columns [32b, 32b+32); this lane fills EPL of them
        uint32_t pre_cols[4];            // up to 8 stream bytes (16 in pair mode: two bytes per column)
        uint2 pre_bnd;                   // (G == 32 only: EPL == 1)
        auto prefetch = [&](uint32_t b) {
            pre_cols[0] = pre_cols[1] = pre_cols[2] = pre_cols[3] = 0x17171717u;     // pad residues
            pre_bnd = make_uint2(B2, B2);
            if (32 * b < cols_padded) {
                constexpr int BYTES = EPL * (PD ? 2 : 1);
                const uint8_t *src = col_src + 32 * b * (PD ? 2 : 1);
                if (BYTES == 16) { uint4 v = __ldg(reinterpret_cast<const uint4 *>(src)); pre_cols[0] = v.x; pre_cols[1] = v.y; pre_cols[2] = v.z; pre_cols[3] = v.w; }
                else if (BYTES == 8) { uint2 v = __ldg(reinterpret_cast<const uint2 *>(src)); pre_cols[0] = v.x; pre_cols[1] = v.y; }
                else if (BYTES == 4) pre_cols[0] = __ldg(reinterpret_cast<const uint32_t *>(src));
                else if (BYTES == 2) pre_cols[0] = __ldg(reinterpret_cast<const uint16_t *>(src));
                else pre_cols[0] = __ldg(src);
                if (EPL == 1 && multi_in) pre_bnd = __ldcg(bnd_src + 32 * b);
            }
        };
        auto commit = [&](uint32_t b) {
            const uint32_t dst = ring_base + ((b & 1) * 32 + t * EPL) * 16;
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                // message word 3: residue + flags in bits 0-7 (pair mode: second residue in bits 8-15)
                const uint32_t w = PD ? (pre_cols[e >> 1] >> (16 * (e & 1))) & 0xffffu
                                      : (pre_cols[e >> 2] >> (8 * (e & 3))) & 0xffu;
                sts128(dst + e * 16, make_uint4(pre_bnd.x, pre_bnd.y, B2, w));
            }
        };
        // (waiting and reporting do not depend on whether THIS pass reads or writes a bottom row: the buffer is
        // updated in place, so no pass may overtake the one before it)
        const uint32_t *prog_in = PIPE && p.progress_in && have ? p.progress_in + (ci - p.chunk_first) : nullptr;
        uint32_t *prog_out = PIPE && p.progress_out && have ? p.progress_out + (ci - p.chunk_first) : nullptr;
        if (PIPE && prog_in) wait_progress(prog_in, min(32u, n_cols));
        prefetch(0);
        commit(0);

        uint32_t Hl[R], E[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { Hl[r] = B2; E[r] = B2; }
        uint32_t diag[NC], run = B2;
        uint4 mid[NC];                   // mid[c]: message segment c-1 produced in the previous step
#pragma unroll
        for (int c = 0; c < NC; ++c) { diag[c] = B2; mid[c] = make_uint4(B2, B2, B2, PAD_MSG); }
        uint32_t seq = ck.seq0;
        uint2 *out_base = has_out ? p.bound + (col0 - p.bound_col0) : nullptr;
        sts128(mail_self, make_uint4(B2, B2, B2, PAD_MSG));
        __syncwarp();

        for (uint32_t blk = 0; blk < n_blocks; ++blk) {
            if (PIPE && prog_in) wait_progress(prog_in, min(32u * (blk + 2), n_cols));
            prefetch(blk + 1);
#pragma unroll 1
            for (uint32_t i = 0; i < 32; ++i) {
                const uint32_t step = blk * 32 + i;
                const uint32_t in_addr = t == 0 ? ring_base + (step & (RING - 1)) * 16 : mail_up;
                uint4 msg[NC];
#ifdef OSW_EXP_NO_MAIL
                msg[0] = make_uint4(B2 + step, B2, B2, (step * 7) % 20);
#else
                msg[0] = lds128(in_addr);
#endif
#ifndef OSW_EXPERIMENT_NO_KEEP
                msg[0].x = (msg[0].x & keep) | (B2 & ~keep);      // a query's first lane has no row above it
                msg[0].y = (msg[0].y & keep) | (B2 & ~keep);
                msg[0].z = (msg[0].z & keep) | (B2 & ~keep);
#endif
#pragma unroll
                for (int c = 1; c < NC; ++c) msg[c] = mid[c];
                uint32_t paddr[NC], paddr_hi[NC];
                // a column flagged FIRST restarts the DP state of the segment that works on it (one
                // test for all segments: the restart is rare)
                uint32_t any_first = 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) any_first |= msg[c].w;
#ifndef OSW_EXP_NO_FIRST
                if (any_first & OSW_COL_FIRST) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (msg[c].w & OSW_COL_FIRST) {
#pragma unroll
                            for (int r = 4 * seg_begin(R, NC, c); r < 4 * seg_begin(R, NC, c + 1); ++r) { Hl[r] = B2; E[r] = B2; }
                            diag[c] = B2;
                        }
                    }
                }
#endif
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    paddr[c] = prof_lane + (msg[c].w & OSW_COL_CODE) * PITCH_B + seg_begin(R, NC, c) * 16;
                    paddr_hi[c] = PD ? prof_lane + TABLE_B + ((msg[c].w >> 8) & OSW_COL_CODE) * PITCH_B + seg_begin(R, NC, c) * 16 : 0u;
                }
                // Row sweeps of the NC segments, interleaved: they are independent dependency
                // chains.  t (the diagonal term) of row r+1 is issued before H of row r is written,
                // so that H can overwrite Hl[r] in place.
                uint32_t F[NC], cm[NC], Heven[NC], t_next[NC];
                uint4 sv[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    F[c] = msg[c].y; cm[c] = msg[c].z; Heven[c] = B2;
                    sv[c] = lds128_const(paddr[c]);
                    if (PD) sv[c] = add4(sv[c], lds128_const(paddr_hi[c]));
                    t_next[c] = __viaddmax_u16x2(diag[c], sv[c].x, E[4 * seg_begin(R, NC, c)]);
                }
#pragma unroll
                for (int kk = 0; kk < seg_begin(R, NC, 1); ++kk) {
                    uint4 sn[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        sn[c] = sv[c];
                        if (kk + 1 < seg_begin(R, NC, c + 1) - seg_begin(R, NC, c)) {
                            sn[c] = lds128_const(paddr[c] + (kk + 1) * 16);
                            if (PD) sn[c] = add4(sn[c], lds128_const(paddr_hi[c] + (kk + 1) * 16));
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            if (kk < seg_begin(R, NC, c + 1) - seg_begin(R, NC, c)) {
                                const int r = 4 * (seg_begin(R, NC, c) + kk) + rr;
                                const uint32_t s_after = rr == 0 ? sv[c].y : rr == 1 ? sv[c].z : rr == 2 ? sv[c].w : sn[c].x;
                                const uint32_t tt = t_next[c];
                                if (r + 1 < 4 * seg_begin(R, NC, c + 1)) t_next[c] = __viaddmax_u16x2(Hl[r], s_after, E[r + 1]);
                                const uint32_t H = __vimax3_u16x2(tt, F[c], B2);
                                const uint32_t u = H - GOE2;
                                E[r] = __viaddmax_u16x2(E[r], NGE, u);
                                F[c] = __viaddmax_u16x2(F[c], NGE, u);
                                Hl[r] = H;
#ifdef OSW_EXP_NO_CM
                                if (r == 0) cm[c] = H;
#else
                                if (rr & 1) cm[c] = __vimax3_u16x2(cm[c], Heven[c], H); else Heven[c] = H;
#endif
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < NC; ++c) sv[c] = sn[c];
                }
                // hand the segments' bottom rows on: segment c -> segment c+1 (next step), last -> next lane
#pragma unroll
                for (int c = 0; c < NC; ++c) diag[c] = msg[c].x;
#pragma unroll
                for (int c = NC - 1; c >= 1; --c)
                    mid[c] = make_uint4(Hl[4 * seg_begin(R, NC, c) - 1], F[c - 1], cm[c - 1], msg[c - 1].w);
                const uint32_t lf = msg[NC - 1].w;
                const uint32_t Hbot = Hl[R - 1], Fbot = F[NC - 1], cmbot = cm[NC - 1];
                run = __vmaxu2(run, cmbot);
                if (has_out && t == G - 1)           // several passes: the group's last lane stages this pass's bottom row
                    sts64(oring_base + i * 8, make_uint2(Hbot, Fbot));
                if (lf & last_mask) {                // last lane of the group, last column of a sequence
                    const uint32_t lo = run & 0xffffu, hi = run >> 16;
                    const int sa = lo >= FLAG_THRESHOLD ? OSW_SCORE_FLAGGED : (int)(lo - a.bias);
                    const int sb = hi >= FLAG_THRESHOLD ? OSW_SCORE_FLAGGED : (int)(hi - a.bias);
                    if (PD) {       // low half: sequence 2p of the chunk, high half: sequence 2p+1 (if there is one)
                        int32_t *row = p.scores + (size_t)a.lane[0][t].query * p.n_seqs;
                        atomicMax(row + seq, sa);
                        if (seq + 1 < ck.seq0 + ck.n_seqs) atomicMax(row + seq + 1, sb);
                        ++seq;               // a pair is two sequences (second increment below)
                    } else {
                        if (emit & 1u) atomicMax(p.scores + (size_t)a.lane[0][t].query * p.n_seqs + seq, sa);
                        if (emit & 2u) atomicMax(p.scores + (size_t)a.lane[1][t].query * p.n_seqs + seq, sb);
                    }
                    ++seq;
                    run = B2;
                }
#ifndef OSW_EXP_NO_SYNC
                __syncwarp();
#endif
                sts128(mail_self, make_uint4(Hbot, Fbot, cmbot, lf));
#ifndef OSW_EXP_NO_SYNC
                __syncwarp();
#endif
            }
            if (has_out) {
                // flush the 32 bottom-row entries of this block: step s finished column s - (NC*G - 1)
                const uint32_t col = blk * 32 + lane - (NC * G - 1);
                if (col < cols_padded) __stcg(out_base + col, lds64(oring_base + lane * 8));
            }
            if (PIPE && prog_out) {              // the columns this block finished (and their bottom row) are done: tell the next pass
                __threadfence();
                __syncwarp();
                const uint32_t done = blk * 32 + 32;
                if (lane == 0 && done > (uint32_t)(NC * G - 1))
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(prog_out), "r"(blk + 1 == n_blocks ? 0xffffffffu : done - (NC * G - 1)) : "memory");
            }
            commit(blk + 1);
            __syncwarp();
        }
    }
    if (p.cycle_acc && threadIdx.x == 0) atomicAdd(p.cycle_acc, (unsigned long long)(clock64() - clk0));
}

template <int G, int R, int THREADS, bool PD, bool DEAL, bool PIPE = false>
int launch_kernel(const KArgs &k, int n_sms, size_t smem, cudaStream_t st) {
    // (set on every launch: the attribute is per device and per function, the call costs about a
    // microsecond, and a "configured" flag here would be shared state between host threads)
    if (smem > 227 * 1024) return OSW_E_ARG;
    if (cudaFuncSetAttribute(sw_u16_kernel<G, R, THREADS, PD, DEAL, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return OSW_E_CUDA;
    profile_build_kernel<G, R, PD><<<32, 256, 0, st>>>(k);
    sw_u16_kernel<G, R, THREADS, PD, DEAL, PIPE><<<n_sms, THREADS, smem, st>>>(k);
    return cudaGetLastError() == cudaSuccess ? OSW_OK : OSW_E_CUDA;
}

template <int G, int R, int THREADS, bool PD>
int launch_threads(const KArgs &a, int n_sms, cudaStream_t st) {
    const size_t prof = (prof_copy_bytes(G, R) * prof_copies(G) + (prof_copies(G) > 1 ? 128 : 0)) * (PD ? 2 : 1);
    const size_t smem = prof + (size_t)(THREADS / 32) * 32 * 16 + (size_t)(THREADS / 32) * (32 / G) * RING * 16 + (size_t)(THREADS / 32) * 32 * 8;
    KArgs k = a;
    constexpr uint32_t WARPS = THREADS / 32;
    const uint32_t slots = (uint32_t)n_sms * WARPS * (32 / G);          // chunks in flight
    const bool few_chunks = a.p.chunk_end - a.p.chunk_first < 12 * slots;
    k.n_ctas = (uint32_t)n_sms; k.warps = WARPS;
    if (a.p.all_express) {          // a long-chunk launch: every CTA gives four chunks a scheduler each, the rest from the counter
        k.p.express_ctas = (uint32_t)n_sms;
        k.p.static_first = 0u;
        k.p.dyn_base = 4u * (uint32_t)n_sms;
        if (a.p.progress_in || a.p.progress_out) {
            if constexpr (G == 32) return launch_kernel<G, R, THREADS, PD, true, true>(k, n_sms, smem, st);
            else return OSW_E_ARG;
        }
        return launch_kernel<G, R, THREADS, PD, true>(k, n_sms, smem, st);
    }
    const uint32_t K = std::min<uint32_t>(a.p.express_ctas, (uint32_t)n_sms - 1);
    if (K || few_chunks) {          // express CTAs and / or a static first deal: through the table
        k.p.express_ctas = K;
        k.p.static_first = few_chunks ? 1u : 0u;
        k.p.dyn_base = 4u * K + (few_chunks ? ((uint32_t)n_sms - K) * WARPS : 0u);       // groups dealt statically
        return launch_kernel<G, R, THREADS, PD, true>(k, n_sms, smem, st);
    }
    k.p.first_table = nullptr;          // large database: the counter alone
    return launch_kernel<G, R, THREADS, PD, false>(k, n_sms, smem, st);
}

// CTA size: 512 threads (128 registers each) up to R = 32, 384 (168 registers) above.
#ifndef OSW_PD_THREADS
#define OSW_PD_THREADS 0        // experiment: CTA size of the pair-database kernels (0 = same rule as two-track)
#endif
template <int G, int R>
int launch_one(const KArgs &a, int n_sms, cudaStream_t st) {
    constexpr int THREADS = block_threads(R);
    if (a.pair_db) return launch_threads<G, R, (OSW_PD_THREADS ? OSW_PD_THREADS : THREADS), true>(a, n_sms, st);
    return launch_threads<G, R, THREADS, false>(a, n_sms, st);
}

template <int G>
int launch_g(int R, const KArgs &a, int n_sms, cudaStream_t st) {
    switch (R) {
        case 8: return launch_one<G, 8>(a, n_sms, st);
        case 12: return launch_one<G, 12>(a, n_sms, st);
        case 16: return launch_one<G, 16>(a, n_sms, st);
        case 20: return launch_one<G, 20>(a, n_sms, st);
        case 24: return launch_one<G, 24>(a, n_sms, st);
        case 28: return launch_one<G, 28>(a, n_sms, st);
        case 32: return launch_one<G, 32>(a, n_sms, st);
        case 36: return launch_one<G, 36>(a, n_sms, st);
        case 40: return launch_one<G, 40>(a, n_sms, st);
        case 44: return launch_one<G, 44>(a, n_sms, st);
    }
    return OSW_E_ARG;
}

// one translation unit per group width (sw_u16_g*.cu) so that the 64 kernel instances compile in parallel
int launch_g4(int R, const KArgs &a, int n_sms, cudaStream_t st);
int launch_g8(int R, const KArgs &a, int n_sms, cudaStream_t st);
int launch_g16(int R, const KArgs &a, int n_sms, cudaStream_t st);
int launch_g32(int R, const KArgs &a, int n_sms, cudaStream_t st);

}  // namespace osw_u16

// bias, gap words and lane descriptors of a launch (sw_u16.cu)
void osw_fill_kargs(osw_u16::KArgs &a, const U16Params &p, const OswPass &pass);
#endif
