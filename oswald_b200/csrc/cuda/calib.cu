// calib.cu - roofline calibration micro-kernels.
//
// The cell-update roofline of the packed 16-bit kernel is an instruction-issue bound: how
// many VIADDMNMX.U16x2 / VIMNMX3.U16x2 (the DPX max / add-max instructions) an SM retires per
// clock.  That rate is not documented for sm_100, so it is measured here with
// dependent-chain-free loops, one block per SM, timed with clock64 inside the kernel.
#include "osw_internal.h"
#include <stdio.h>

namespace {

constexpr int CAL_THREADS = 512;
constexpr int CAL_ITERS = 2048;

struct CalArgs { uint32_t a, b, one, ngoe, nge, B; unsigned long long *cycles; uint32_t *sink; };

__device__ __forceinline__ uint32_t imad_sub(uint32_t h, uint32_t one, uint32_t c) {
    uint32_t u;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(h), "r"(one), "r"(c));   // IMAD: fma pipe
    return u;
}

template <int MODE>
__global__ void __launch_bounds__(CAL_THREADS) calib_kernel(CalArgs g) {
    __shared__ uint4 prof[24 * 16];
    for (int i = threadIdx.x; i < 24 * 16; i += blockDim.x)
        prof[i] = make_uint4(g.a + i, g.a ^ i, g.b + i, g.b ^ i);
    __syncthreads();
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = g.a * (threadIdx.x + k + 1);
    uint32_t Hl[16], E[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { Hl[r] = g.B + r + threadIdx.x; E[r] = g.B + r; }
    uint32_t F = g.B, best = g.B, diag = g.B;
    uint32_t letter = threadIdx.x % 24;
    long long t0 = clock64();
    for (int it = 0; it < CAL_ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __viaddmax_u16x2(x[k], g.a, g.b);
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = __vimax3_u16x2(x[k], g.a + k, g.b);
        } else if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = imad_sub(x[k], g.one, g.ngoe);
        } else {
            // one database column against 16 query rows, as the u16 kernel issues it
            uint32_t sc[16];
            if (MODE == 5) {
                letter = (letter + 7) % 24;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint4 v = prof[letter * 16 + ((k * 4 + (threadIdx.x & 7)) & 15)];
                    sc[4 * k] = v.x; sc[4 * k + 1] = v.y; sc[4 * k + 2] = v.z; sc[4 * k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) sc[r] = g.a + r;
            }
            uint32_t Hprev = 0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                uint32_t t = __viaddmax_u16x2(diag, sc[r], E[r]);
                uint32_t H = __vimax3_u16x2(t, F, g.B);
                uint32_t u = (MODE == 4) ? H - g.ngoe : imad_sub(H, g.one, g.ngoe);
                E[r] = __viaddmax_u16x2(E[r], g.nge, u);
                F = __viaddmax_u16x2(F, g.nge, u);
                diag = Hl[r];
                Hl[r] = H;
                if (r & 1) best = __vimax3_u16x2(best, Hprev, H); else Hprev = H;
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = F ^ best ^ diag;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= x[k];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc ^= Hl[r] ^ E[r];
    g.sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) atomicMax(g.cycles, (unsigned long long)(t1 - t0));
}

__global__ void clock_probe(unsigned long long *out) {
    unsigned long long g0, g1;
    long long c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1)); } while (g1 - g0 < 2000000ull);   // 2 ms
    long long c1 = clock64();
    out[0] = (unsigned long long)(c1 - c0); out[1] = g1 - g0;
}

template <int MODE>
int run_mode(int n_sms, CalArgs g, double *cycles_out) {
    cudaMemset(g.cycles, 0, sizeof(unsigned long long));
    calib_kernel<MODE><<<n_sms, CAL_THREADS>>>(g);       // warm-up
    cudaMemset(g.cycles, 0, sizeof(unsigned long long));
    calib_kernel<MODE><<<n_sms, CAL_THREADS>>>(g);
    unsigned long long c = 0;
    if (cudaMemcpy(&c, g.cycles, sizeof c, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    *cycles_out = (double)c;
    return 0;
}

}  // namespace

extern "C" int osw_calibrate(int device, double out[8]) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OSW_E_NODEV;
    if (cudaSetDevice(device) != cudaSuccess) return OSW_E_CUDA;
    const int n_sms = prop.multiProcessorCount;
    CalArgs g;
    g.a = 0x00030005u; g.b = 0x00110013u; g.one = 1u; g.ngoe = 0x000c000cu; g.nge = 0xfffefffeu; g.B = 0x00400040u;
    if (cudaMalloc(&g.cycles, 2 * sizeof(unsigned long long)) != cudaSuccess) return OSW_E_NOMEM;
    if (cudaMalloc(&g.sink, (size_t)n_sms * CAL_THREADS * sizeof(uint32_t)) != cudaSuccess) return OSW_E_NOMEM;
    const double warp_instr = (double)(CAL_THREADS / 32) * CAL_ITERS;
    double cyc;
    int rc = 0;
    for (int k = 0; k < 8; ++k) out[k] = 0;
    rc |= run_mode<0>(n_sms, g, &cyc); out[0] = warp_instr * 8 * 32 / cyc;            // thread-instr / SM-clk
    rc |= run_mode<1>(n_sms, g, &cyc); out[1] = warp_instr * 8 * 32 / cyc;
    rc |= run_mode<2>(n_sms, g, &cyc); out[2] = warp_instr * 16 * 2 * 32 / cyc;       // cells / SM-clk (IMAD form)
    rc |= run_mode<3>(n_sms, g, &cyc); out[3] = warp_instr * 8 * 32 / cyc;
    rc |= run_mode<4>(n_sms, g, &cyc); out[5] = warp_instr * 16 * 2 * 32 / cyc;       // compiler-chosen subtract
    rc |= run_mode<5>(n_sms, g, &cyc); out[6] = warp_instr * 16 * 2 * 32 / cyc;       // + LDS.128 profile reads
    clock_probe<<<1, 1>>>(g.cycles);
    unsigned long long cp[2] = {0, 1};
    if (cudaMemcpy(cp, g.cycles, sizeof cp, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
    out[4] = (double)cp[0] / (double)cp[1] * 1000.0;     // MHz
    cudaFree(g.cycles); cudaFree(g.sink);
    return rc ? OSW_E_CUDA : OSW_OK;
}
