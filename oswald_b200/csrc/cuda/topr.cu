// topr.cu - device top-r selection with the reference's tie rule, and the flagged-pair scan.
//
// The reference ranks every query's N scores with a full merge sort whose comparisons put,
// among equal scores, the HIGHER canonical index first (utils.c:3-69; sort_scores :71-86).
// That total order is the descending order of the 64-bit key (score << 32 | index), so the
// top r hits are the r largest keys.  They are found by an MSB-first radix select over the key's
// bytes: round j histograms byte j of the keys that still match the prefix found so far, and the
// NEXT kernel (the following round, or the final gather) starts by picking, from that histogram,
// the digit that holds the r-th largest key.  Bytes that cannot be set are skipped: a score is
// below 2^24 (at most 65535 columns x 31 per column, oswald_cuda.h), an index below the size of
// the canonical database.  One kernel per round and one for the gather; the host orders the r
// keys it gets.
#include "osw_internal.h"

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ unsigned long long make_key(int score, uint32_t canon) {
    return ((unsigned long long)(uint32_t)score << 32) | canon;    // scores are >= 0
}

__device__ __forceinline__ void list_flagged(const FlagList &fl, uint32_t q, uint32_t i) {
    const uint32_t slot = atomicAdd(fl.count, 1u);
    if (slot < fl.capacity) fl.pairs[slot] = make_uint2(q, i);
}

// Round j's pick, by a whole block of 256 threads: the digit whose bin holds the need-th largest
// key among those matching the prefix.  Returns prefix | digit << shift and the rank left inside
// that bin.  sh: 256 words of shared memory.
__device__ __forceinline__ void pick_digit(const uint32_t *hist, int shift, unsigned long long pre, uint32_t need,
                                           uint32_t *sh, unsigned long long *pre_out, uint32_t *need_out) {
    __shared__ uint32_t s_digit, s_above;
    const int v = threadIdx.x;
    sh[v] = hist[v];
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {           // suffix sums: sh[v] = keys in bins >= v
        const uint32_t add = v + o < 256 ? sh[v + o] : 0u;
        __syncthreads();
        sh[v] += add;
        __syncthreads();
    }
    const uint32_t above = v < 255 ? sh[v + 1] : 0u;
    if (sh[v] >= need && above < need) { s_digit = (uint32_t)v; s_above = above; }
    __syncthreads();
    *pre_out = pre | ((unsigned long long)s_digit << shift);
    *need_out = need - s_above;
    __syncthreads();
}

struct Rounds { int n; int shift[8]; };

// State after round j is kept at slot j of prefix / remaining ([nq][8]); block x == 0 of a query
// records it for the kernels that follow.
__device__ __forceinline__ void state_before_round(const Rounds &rd, int j, int q, uint32_t r, const TopRWork &w, uint32_t *sh,
                                                   unsigned long long *pre, uint32_t *need) {
    if (j == 0) { *pre = 0; *need = r; return; }
    unsigned long long p0 = 0; uint32_t n0 = r;
    if (j >= 2) { p0 = w.prefix[q * 8 + j - 2]; n0 = w.remaining[q * 8 + j - 2]; }
    pick_digit(w.hist + ((size_t)q * 8 + j - 1) * 256, rd.shift[j - 1], p0, n0, sh, pre, need);
    if (blockIdx.x == 0 && threadIdx.x == 0) { w.prefix[q * 8 + j - 1] = *pre; w.remaining[q * 8 + j - 1] = *need; }
}

// round j: count byte values among the keys whose higher (examined) bytes equal the prefix's
__global__ void __launch_bounds__(THREADS)
topr_round_kernel(const int32_t *scores, const uint32_t *canon, uint64_t n, uint32_t r, int j, Rounds rd, TopRWork w, FlagList fl) {
    __shared__ uint32_t sh[256];
    const int q = blockIdx.y;
    unsigned long long pre; uint32_t need;
    state_before_round(rd, j, q, r, w, sh, &pre, &need);
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int shift = rd.shift[j];
    const unsigned long long himask = j ? ~0ull << (rd.shift[j - 1]) : 0ull;      // the bytes examined so far (skipped ones are 0 in every key)
    const int32_t *row = scores + (size_t)q * n;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        const int sc = row[i];
        if (sc == OSW_SCORE_FLAGGED && fl.count) list_flagged(fl, (uint32_t)q, (uint32_t)i);      // (first round only)
        unsigned long long k = make_key(sc, canon[i]);
        if ((k & himask) == pre) atomicAdd(&sh[(k >> shift) & 255], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&w.hist[((size_t)q * 8 + j) * 256 + threadIdx.x], sh[threadIdx.x]);
}

// after the last round the prefix is the r-th largest key itself: gather the keys >= it
__global__ void __launch_bounds__(THREADS)
topr_gather_kernel(const int32_t *scores, const uint32_t *canon, uint64_t n, uint32_t r, Rounds rd, TopRWork w) {
    __shared__ uint32_t sh[256];
    const int q = blockIdx.y;
    unsigned long long thr; uint32_t need;
    state_before_round(rd, rd.n, q, r, w, sh, &thr, &need);
    const int32_t *row = scores + (size_t)q * n;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        unsigned long long k = make_key(row[i], canon[i]);
        if (k >= thr) {
            uint32_t slot = atomicAdd(&w.out_count[q], 1u);
            if (slot < r) w.out_keys[(size_t)q * r + slot] = k;
        }
    }
}

// Small databases (n <= TOPR_SMALL_MAX): the whole selection in ONE launch, one CTA per query.  The
// keys are read from global memory once, into shared memory; the radix rounds and the final gather
// work there (a tiny search is bound by launch count, not by the scan: 7 launches -> 1).
constexpr int SMALL_THREADS = 1024, SMALL_WARPS = SMALL_THREADS / 32;

__global__ void __launch_bounds__(SMALL_THREADS)
topr_small_kernel(const int32_t *scores, const uint32_t *canon, uint32_t n, uint32_t r, Rounds rd, TopRWork w, FlagList fl) {
    extern __shared__ __align__(16) unsigned long long s_keys[];     // [n]
    // (a histogram per warp: the scores of a search cluster in a few dozen values, and thousands of atomics on
    // one shared-memory word serialise)
    __shared__ uint32_t hist_w[SMALL_WARPS][256], hist[256], suffix[256];
    __shared__ unsigned long long s_pre, s_or, s_and;
    __shared__ uint32_t s_need, s_count;
    const int q = blockIdx.x, wid = threadIdx.x >> 5;
    const int32_t *row = scores + (size_t)q * n;
    if (threadIdx.x == 0) { s_pre = 0; s_need = r; s_count = 0; s_or = 0; s_and = ~0ull; }
    __syncthreads();
    unsigned long long k_or = 0, k_and = ~0ull;
    for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * SMALL_THREADS) {          // (four independent loads in flight per thread)
        int sc[4]; uint32_t cn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * SMALL_THREADS;
            sc[u] = i < n ? row[i] : 0; cn[u] = i < n ? canon[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * SMALL_THREADS;
            if (i >= n) break;
            if (sc[u] == OSW_SCORE_FLAGGED && fl.count) list_flagged(fl, (uint32_t)q, i);
            const unsigned long long k = make_key(sc[u], cn[u]);
            s_keys[i] = k;
            k_or |= k; k_and &= k;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { k_or |= __shfl_xor_sync(0xffffffffu, k_or, o); k_and &= __shfl_xor_sync(0xffffffffu, k_and, o); }
    if ((threadIdx.x & 31) == 0) { atomicOr(&s_or, k_or); atomicAnd(&s_and, k_and); }
    __syncthreads();
    const unsigned long long differ = s_or ^ s_and;           // bits in which the keys are not all alike
    for (int j = 0; j < rd.n; ++j) {
        const int shift = rd.shift[j];
        const unsigned long long himask = j ? ~0ull << (rd.shift[j - 1]) : 0ull, pre = s_pre;
        const uint32_t need = s_need;
        if (((differ >> shift) & 255ull) == 0) {              // every key has the same digit here: no counting needed
            __syncthreads();
            if (threadIdx.x == 0) s_pre = pre | (s_or & (255ull << shift));
            __syncthreads();
            continue;
        }
        for (int v = threadIdx.x; v < SMALL_WARPS * 256; v += SMALL_THREADS) (&hist_w[0][0])[v] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += SMALL_THREADS) {
            const unsigned long long k = s_keys[i];
            if ((k & himask) == pre) atomicAdd(&hist_w[wid][(k >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t sum = 0;
#pragma unroll
            for (int v = 0; v < SMALL_WARPS; ++v) sum += hist_w[v][threadIdx.x];
            hist[threadIdx.x] = sum;
        }
        __syncthreads();
        if (threadIdx.x < 32) {          // one warp: suffix sums over the 256 bins (8 per lane), then the pick
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int k = 7; k >= 0; --k) { sum += hist[threadIdx.x * 8 + k]; loc[k] = sum; }
            uint32_t above = 0;          // keys in the bins of higher lanes
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_down_sync(0xffffffffu, sum, o);
                if ((int)threadIdx.x + o < 32) sum += v;
            }
            above = sum - loc[0];
#pragma unroll
            for (int k = 0; k < 8; ++k) suffix[threadIdx.x * 8 + k] = loc[k] + above;
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            const uint32_t at = suffix[threadIdx.x], above = threadIdx.x < 255 ? suffix[threadIdx.x + 1] : 0u;
            if (at >= need && above < need) { s_pre = pre | ((unsigned long long)threadIdx.x << shift); s_need = need - above; }
        }
        __syncthreads();
    }
    const unsigned long long thr = s_pre;       // the r-th largest key itself
    for (uint32_t i = threadIdx.x; i < n; i += SMALL_THREADS) {
        const unsigned long long k = s_keys[i];
        if (k >= thr) {
            const uint32_t slot = atomicAdd(&s_count, 1u);
            if (slot < r) w.out_keys[(size_t)q * r + slot] = k;
        }
    }
}

__global__ void __launch_bounds__(THREADS)
collect_flagged_kernel(const int32_t *scores, uint64_t n, uint2 *pairs, uint32_t *count, uint32_t capacity) {
    const int q = blockIdx.y;
    const int32_t *row = scores + (size_t)q * n;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        if (row[i] == OSW_SCORE_FLAGGED) {
            uint32_t slot = atomicAdd(count, 1u);
            if (slot < capacity) pairs[slot] = make_uint2((uint32_t)q, (uint32_t)i);
        }
    }
}

int grid_x(uint64_t n) {
    uint64_t b = (n + THREADS * 8 - 1) / (THREADS * 8);
    if (b < 1) b = 1;
    if (b > 1184) b = 1184;            // 148 SMs * 8
    return (int)b;
}

}  // namespace

int osw_topr_select(const int32_t *scores, const uint32_t *canon, uint64_t n_seqs, uint64_t n_canon, int nq,
                    uint32_t top_r, const TopRWork &w, const FlagList *flags, cudaStream_t st) {
    uint32_t r = top_r < n_seqs ? top_r : (uint32_t)n_seqs;
    const FlagList none = {nullptr, nullptr, 0};
    if (r == 0) {          // nothing to select: the flagged pairs still have to be found
        if (flags && n_seqs) return osw_collect_flagged(scores, n_seqs, nq, flags->pairs, flags->count, flags->capacity, st);
        return 0;
    }
    Rounds rd;
    rd.n = 0;
    for (int byte = 6; byte >= 0; --byte) {          // byte 7 = score bits 24-31: never set
        if (byte < 4 && byte > 0 && ((n_canon - 1) >> (8 * byte)) == 0) continue;     // index bytes above the database size
        rd.shift[rd.n++] = 8 * byte;
    }
    if (n_seqs <= OSW_TOPR_SMALL_MAX) {
        cudaFuncSetAttribute(topr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OSW_TOPR_SMALL_MAX * 8);
        topr_small_kernel<<<nq, SMALL_THREADS, (size_t)n_seqs * 8, st>>>(scores, canon, (uint32_t)n_seqs, r, rd, w, flags ? *flags : none);
        return 1;
    }
    cudaMemsetAsync(w.hist, 0, (size_t)nq * 8 * 256 * sizeof(uint32_t), st);
    cudaMemsetAsync(w.out_count, 0, (size_t)nq * sizeof(uint32_t), st);
    dim3 grid(grid_x(n_seqs), nq);
    for (int j = 0; j < rd.n; ++j)
        topr_round_kernel<<<grid, THREADS, 0, st>>>(scores, canon, n_seqs, r, j, rd, w, j == 0 && flags ? *flags : none);
    topr_gather_kernel<<<grid, THREADS, 0, st>>>(scores, canon, n_seqs, r, rd, w);
    return rd.n + 1;
}

int osw_collect_flagged(const int32_t *scores, uint64_t n_seqs, int nq, uint2 *pairs,
                        uint32_t *count, uint32_t capacity, cudaStream_t st) {
    dim3 grid(grid_x(n_seqs), nq);
    collect_flagged_kernel<<<grid, THREADS, 0, st>>>(scores, n_seqs, pairs, count, capacity);
    return 1;
}
