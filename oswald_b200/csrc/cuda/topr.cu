// topr.cu - device top-r selection with the reference's tie rule, and the flagged-pair scan.
//
// The reference ranks every query's N scores with a full merge sort whose comparisons put,
// among equal scores, the HIGHER canonical index first (utils.c:3-69; sort_scores :71-86).
// That total order is the descending order of the 64-bit key (score << 32 | index), so the
// top r hits are the r largest keys.  They are found by an MSB-first radix select (8 digits
// of 8 bits): each round histograms the digit of the keys that still match the prefix, a
// one-block kernel picks the digit that contains the r-th largest key, and a final pass
// gathers the keys >= the r-th largest.  The host orders those r keys.
#include "osw_internal.h"

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ unsigned long long make_key(int score, uint32_t canon) {
    return ((unsigned long long)(uint32_t)score << 32) | canon;    // scores are >= 0
}

// round `d` (0 = most significant byte): count digit values among keys whose higher bytes
// equal prefix[q]'s.
__global__ void __launch_bounds__(THREADS)
topr_hist_kernel(const int32_t *scores, const uint32_t *canon, uint64_t n, int d,
                 const unsigned long long *prefix, uint32_t *hist) {
    __shared__ uint32_t sh[256];
    const int q = blockIdx.y;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 56 - 8 * d;
    const unsigned long long pre = prefix[q];
    const unsigned long long himask = d ? ~0ull << (shift + 8) : 0ull;
    const int32_t *row = scores + (size_t)q * n;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        unsigned long long k = make_key(row[i], canon[i]);
        if ((k & himask) == (pre & himask)) atomicAdd(&sh[(k >> shift) & 255], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[q * 256 + threadIdx.x], sh[threadIdx.x]);
}

// one block per query: choose the digit holding the `remaining`-th largest key.
__global__ void topr_pick_kernel(int d, unsigned long long *prefix, uint32_t *remaining, uint32_t *hist) {
    const int q = blockIdx.x;
    if (threadIdx.x == 0) {
        uint32_t need = remaining[q];
        const int shift = 56 - 8 * d;
        int digit = 0;
        for (int v = 255; v >= 0; --v) {
            uint32_t c = hist[q * 256 + v];
            if (c >= need) { digit = v; break; }
            need -= c;
        }
        prefix[q] |= (unsigned long long)digit << shift;
        remaining[q] = need;
    }
    __syncthreads();
    hist[q * 256 + threadIdx.x] = 0;          // ready for the next round (blockDim.x == 256)
}

__global__ void __launch_bounds__(THREADS)
topr_gather_kernel(const int32_t *scores, const uint32_t *canon, uint64_t n, uint32_t top_r,
                   const unsigned long long *threshold, uint32_t *out_count, unsigned long long *out_keys) {
    const int q = blockIdx.y;
    const unsigned long long thr = threshold[q];
    const int32_t *row = scores + (size_t)q * n;
    for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) {
        unsigned long long k = make_key(row[i], canon[i]);
        if (k >= thr) {
            uint32_t slot = atomicAdd(&out_count[q], 1u);
            if (slot < top_r) out_keys[(size_t)q * top_r + slot] = k;
        }
    }
}

__global__ void topr_init_kernel(int nq, uint32_t r, unsigned long long *prefix, uint32_t *remaining,
                                 uint32_t *out_count, uint32_t *hist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { prefix[i] = 0; remaining[i] = r; out_count[i] = 0; }
    if (i < nq * 256) hist[i] = 0;
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

int osw_topr_select(const int32_t *scores, const uint32_t *canon, uint64_t n_seqs, int nq,
                    uint32_t top_r, const TopRWork &w, cudaStream_t st) {
    int launches = 0;
    uint32_t r = top_r < n_seqs ? top_r : (uint32_t)n_seqs;
    topr_init_kernel<<<(nq * 256 + 255) / 256, 256, 0, st>>>(nq, r, w.prefix, w.remaining, w.out_count, w.hist);
    ++launches;
    if (r == 0) return launches;
    dim3 grid(grid_x(n_seqs), nq);
    for (int d = 0; d < 8; ++d) {
        topr_hist_kernel<<<grid, THREADS, 0, st>>>(scores, canon, n_seqs, d, w.prefix, w.hist);
        topr_pick_kernel<<<nq, 256, 0, st>>>(d, w.prefix, w.remaining, w.hist);
        launches += 2;
    }
    topr_gather_kernel<<<grid, THREADS, 0, st>>>(scores, canon, n_seqs, r, w.prefix, w.out_count, w.out_keys);
    return launches + 1;
}

int osw_collect_flagged(const int32_t *scores, uint64_t n_seqs, int nq, uint2 *pairs,
                        uint32_t *count, uint32_t capacity, cudaStream_t st) {
    dim3 grid(grid_x(n_seqs), nq);
    collect_flagged_kernel<<<grid, THREADS, 0, st>>>(scores, n_seqs, pairs, count, capacity);
    return 1;
}
