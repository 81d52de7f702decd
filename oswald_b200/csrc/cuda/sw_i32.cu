// sw_i32.cu - exact 32-bit Gotoh local score, one warp per (query, sequence) pair.
//
// Role: the overflow re-score pass (the reference's 32-bit stage, HybridSearch.c:1046-1134,
// and sw_host, FPGAsearch.c:377-506) for pairs the packed 16-bit kernel flags, and - with
// kernel mask OSW_K_I32 alone - a complete scorer used by the parity tests.
//
// Recurrence (reference HybridSearch.c:842-913, per cell of query row i / database column j):
//     H = max(0, Hdiag + M[a_i][b_j], E_i, F_j);  E_i = max(E_i - ge, H - (go+ge));  F_j likewise.
// Mapping: the 32 lanes of a warp own R consecutive query rows each (32*R rows per pass) and
// sweep the database columns as a systolic array: lane t works on column (step - t).  A lane's
// bottom row (H, F) and the column's residue go to the next lane by warp shuffle; lane 0 takes
// its residues from a 32-column window the warp loads with one coalesced read per 32 steps.
// Lane 31's bottom row is parked in global scratch for the next pass over the following 32*R
// query rows.  Substitution scores come from a per-warp, per-pass profile in shared memory,
// prof[residue][row] (int8): a lane's R = 8 scores for its column are ONE 8-byte read, and with a
// 256-byte row pitch the 32 lanes never conflict whatever their residues are (the first version
// indexed the 24x32 matrix by (query residue, database residue): an 8-way gather with random bank
// conflicts per step).  H/E/F state lives in registers; DPX VIADDMNMX / VIMNMX3 do the max/add-max.
#include "osw_internal.h"

namespace {

constexpr int R = 8;                   // query rows per lane
constexpr int ROWS_PER_PASS = 32 * R;
constexpr int BLOCK_THREADS = 128;
constexpr int WARPS = BLOCK_THREADS / 32;
constexpr int PROF_PITCH = ROWS_PER_PASS;          // bytes per residue row of the profile

__device__ __forceinline__ int sext8(uint32_t w, int k) { return (int)(int8_t)(w >> (8 * k)); }

__global__ void __launch_bounds__(BLOCK_THREADS)
sw_i32_kernel(I32Params p) {
    __shared__ int8_t sM[24 * 32];
    __shared__ __align__(16) uint8_t s_prof[WARPS][24 * PROF_PITCH];
    __shared__ unsigned long long s_task[WARPS];
    for (int i = threadIdx.x; i < 24 * 32; i += blockDim.x) sM[i] = p.matrix[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned warp_global = blockIdx.x * WARPS + wib;
    int2 *scr = p.scratch + (size_t)warp_global * p.max_len;
    uint8_t *prof = s_prof[wib];
    const int goe = p.gap_open_extend, ge = p.gap_extend;
    unsigned long long n_tasks = p.n_tasks;
    if (p.n_tasks_dev) { const unsigned long long n = *p.n_tasks_dev; if (n < n_tasks) n_tasks = n; }

    for (;;) {
        if (lane == 0) s_task[wib] = atomicAdd(p.task_counter, 1ull);
        __syncwarp();
        const unsigned long long task = s_task[wib];
        __syncwarp();
        if (task >= n_tasks) break;
        uint32_t q, s;
        if (p.pairs) { uint2 pr = p.pairs[task]; q = pr.x; s = pr.y; }
        else { q = (uint32_t)(task / p.n_seqs); s = (uint32_t)(task % p.n_seqs); }
        const uint8_t *a = p.queries + p.q_off[q];
        const int m = (int)(p.q_off[q + 1] - p.q_off[q]);
        const uint8_t *b = p.stream + (p.task_off ? p.task_off[task] : p.seq_off[s]);
        const int n = (int)p.seq_len[s];
        int best = 0;
        const int n_pass = (m + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
        for (int pass = 0; pass < n_pass; ++pass) {
            // the pass's profile: prof[residue][lane*R + r] = M[a_row][residue] (pad row beyond m)
            {
                int arow[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = pass * ROWS_PER_PASS + lane * R + r;
                    arow[r] = (i < m ? (int)a[i] : OSW_PAD_CODE) * 32;
                }
                for (int res = 0; res < 24; ++res) {
                    uint32_t w0 = 0, w1 = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        w0 |= (uint32_t)(uint8_t)sM[arow[r] + res] << (8 * r);
                        w1 |= (uint32_t)(uint8_t)sM[arow[4 + r] + res] << (8 * r);
                    }
                    *reinterpret_cast<uint2 *>(prof + res * PROF_PITCH + lane * R) = make_uint2(w0, w1);
                }
            }
            __syncwarp();
            int Hl[R], E[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { Hl[r] = 0; E[r] = 0; }
            int diag_top = 0;            // H[row above][column - 1]
            int Hbot = 0, Fbot = 0;      // my bottom row at the column I finished last
            uint32_t res_mine = OSW_PAD_CODE;      // residue of the column I finished last
            uint32_t window = 0;         // lane l: residue of column 32*block + l
            const bool last_pass = pass == n_pass - 1;
            for (int step = 0; step < n + 31; ++step) {
                if ((step & 31) == 0) { const int j = step + lane; window = j < n ? (uint32_t)(b[j] & OSW_COL_CODE) : (uint32_t)OSW_PAD_CODE; }
                int Hup = __shfl_up_sync(0xffffffffu, Hbot, 1);
                int Fup = __shfl_up_sync(0xffffffffu, Fbot, 1);
                uint32_t res = __shfl_up_sync(0xffffffffu, res_mine, 1);
                const uint32_t fresh = __shfl_sync(0xffffffffu, window, step & 31);
                if (lane == 0) res = fresh;
                res_mine = res;
                const int j = step - lane;
                if (j >= 0 && j < n) {
                    if (lane == 0) {
                        if (pass) { int2 v = __ldcg(scr + j); Hup = v.x; Fup = v.y; } else { Hup = 0; Fup = 0; }
                    }
                    const uint2 sc8 = *reinterpret_cast<const uint2 *>(prof + res * PROF_PITCH + lane * R);
                    int F = Fup, diag = diag_top;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int sc = sext8(r < 4 ? sc8.x : sc8.y, r & 3);
                        int t = __viaddmax_s32(diag, sc, E[r]);
                        int H = __vimax3_s32_relu(t, F, 0);
                        int u = H - goe;
                        E[r] = __viaddmax_s32(E[r], -ge, u);
                        F = __viaddmax_s32(F, -ge, u);
                        diag = Hl[r];
                        Hl[r] = H;
                        best = max(best, H);
                    }
                    diag_top = Hup;
                    Hbot = Hl[R - 1]; Fbot = F;
                    if (lane == 31 && !last_pass) __stcg(scr + j, make_int2(Hbot, Fbot));
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) p.scores[(size_t)q * p.n_seqs + s] = best;
    }
}

}  // namespace

int osw_i32_block_threads() { return BLOCK_THREADS; }

void osw_launch_i32(const I32Params &p, int n_blocks, cudaStream_t st) {
    sw_i32_kernel<<<n_blocks, BLOCK_THREADS, 0, st>>>(p);
}
