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
                for (int r = 0; r < 16; ++r) sc[r] = Hl[(r + 5) & 15] ^ g.a;     // any per-row value already in a register
            }
            uint32_t Hprev = 0;
            if (MODE == 9) {
                // E kept clamped at >= B: then t >= B and H needs only a 2-input max
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    uint32_t t = __viaddmax_u16x2(diag, sc[r], E[r]);
                    uint32_t H = __vmaxu2(t, F);
                    uint32_t u = H - g.ngoe;
                    uint32_t e1 = E[r] - (g.ngoe >> 2);
                    E[r] = __vimax3_u16x2(e1, u, g.B);
                    F = __viaddmax_u16x2(F, g.nge, u);
                    diag = Hl[r];
                    Hl[r] = H;
                    if (r & 1) best = __vimax3_u16x2(best, Hprev, H); else Hprev = H;
                }
                continue;
            }
            if (MODE == 6 || MODE == 7) {
                // signed variants: relu folded into the add-max, H by a 2-input max, packed 16-bit add
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    uint32_t t = __viaddmax_s16x2_relu(diag, sc[r], E[r]);
                    uint32_t H = __vmaxs2(t, F);
                    uint32_t u = __vadd2(H, g.ngoe);
                    E[r] = __viaddmax_s16x2(E[r], g.nge, u);
                    F = __viaddmax_s16x2(F, g.nge, u);
                    diag = Hl[r];
                    Hl[r] = H;
                    if (MODE == 6) { if (r & 1) best = __vimax3_s16x2(best, Hprev, H); else Hprev = H; }
                }
                continue;
            }
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (MODE == 8) {       // current arithmetic without the column-maximum tracking
                    uint32_t t = __viaddmax_u16x2(diag, sc[r], E[r]);
                    uint32_t H = __vimax3_u16x2(t, F, g.B);
                    uint32_t u = H - g.ngoe;
                    E[r] = __viaddmax_u16x2(E[r], g.nge, u);
                    F = __viaddmax_u16x2(F, g.nge, u);
                    diag = Hl[r];
                    Hl[r] = H;
                    continue;
                }
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

extern "C" int osw_calibrate(int device, double out[12]) {
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
    for (int k = 0; k < 12; ++k) out[k] = 0;
    rc |= run_mode<0>(n_sms, g, &cyc); out[0] = warp_instr * 8 * 32 / cyc;            // thread-instr / SM-clk
    rc |= run_mode<1>(n_sms, g, &cyc); out[1] = warp_instr * 8 * 32 / cyc;
    rc |= run_mode<2>(n_sms, g, &cyc); out[2] = warp_instr * 16 * 2 * 32 / cyc;       // cells / SM-clk (IMAD form)
    rc |= run_mode<3>(n_sms, g, &cyc); out[3] = warp_instr * 8 * 32 / cyc;
    rc |= run_mode<4>(n_sms, g, &cyc); out[5] = warp_instr * 16 * 2 * 32 / cyc;       // compiler-chosen subtract
    rc |= run_mode<5>(n_sms, g, &cyc); out[6] = warp_instr * 16 * 2 * 32 / cyc;       // + LDS.128 profile reads
    rc |= run_mode<6>(n_sms, g, &cyc); out[7] = warp_instr * 16 * 2 * 32 / cyc;       // signed relu variant
    rc |= run_mode<7>(n_sms, g, &cyc); out[8] = warp_instr * 16 * 2 * 32 / cyc;       // ... without max tracking
    rc |= run_mode<8>(n_sms, g, &cyc); out[9] = warp_instr * 16 * 2 * 32 / cyc;       // unsigned without max tracking
    rc |= run_mode<9>(n_sms, g, &cyc); out[10] = warp_instr * 16 * 2 * 32 / cyc;      // clamped E, 2-input max for H
    clock_probe<<<1, 1>>>(g.cycles);
    unsigned long long cp[2] = {0, 1};
    if (cudaMemcpy(cp, g.cycles, sizeof cp, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
    out[4] = (double)cp[0] / (double)cp[1] * 1000.0;     // MHz
    cudaFree(g.cycles); cudaFree(g.sink);
    return rc ? OSW_E_CUDA : OSW_OK;
}

// ---- co-issue probe: which instruction classes issue alongside the DPX add-max ----------------
// mix_kernel<A, B> runs 8 independent chains of op A and 8 of op B, interleaved.  If the total
// rate of <A,B> is about twice A's own rate the two classes use different issue pipes.
namespace {

enum { OP_NONE = -1, OP_VIADDMNMX = 0, OP_VIMNMX3, OP_VIADD, OP_IMAD, OP_HMNMX2, OP_VIMNMX2, OP_IMNMX, OP_LOP3,
       OP_FMNMX, OP_PRMT, OP_SHF, OP_HADD2, OP_COUNT, OP_IMADHI = OP_COUNT, OP_LEAHI, OP_IMADSHL, OP_COUNT_EXT };

template <int OP>
__device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t a, uint32_t b) {
    uint32_t r = x;
    if (OP == OP_VIADDMNMX) r = __viaddmax_u16x2(x, a, b);
    else if (OP == OP_VIMNMX3) r = __vimax3_u16x2(x, a, b);
    else if (OP == OP_VIADD) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
    else if (OP == OP_HMNMX2) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_VIMNMX2) r = __vmaxu2(x, a);
    else if (OP == OP_IMNMX) asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_LOP3) asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_FMNMX) asm volatile("max.f32 %0, %1, %2;" : "=f"(*(float *)&r) : "f"(__uint_as_float(x)), "f"(__uint_as_float(a)));
    else if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_HADD2) asm volatile("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
    else if (OP == OP_IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));      // (x * a >> 32) + b
    else if (OP == OP_LEAHI) r = (x >> 16) + a;                      // LEA.HI: the high half moved down, plus a word
    else if (OP == OP_IMADSHL) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(b), "r"(a));     // x * b + a, b = 65536 at run time
    return r;
}

template <int A, int B, int NB>
__global__ void __launch_bounds__(CAL_THREADS) mix_kernel(CalArgs g) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = g.a * (threadIdx.x + k + 1); y[k] = g.b + threadIdx.x * (k + 3); }
    long long t0 = clock64();
    for (int it = 0; it < CAL_ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = apply<A>(x[k], g.a, g.b);
            if (B != OP_NONE && k < NB) y[k] = apply<B>(y[k], g.b, g.a);
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= x[k] ^ y[k];
    g.sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) atomicMax(g.cycles, (unsigned long long)(t1 - t0));
}

template <int A, int B, int NB>
double run_mix(int n_sms, CalArgs g) {
    cudaMemset(g.cycles, 0, sizeof(unsigned long long));
    mix_kernel<A, B, NB><<<n_sms, CAL_THREADS>>>(g);
    cudaMemset(g.cycles, 0, sizeof(unsigned long long));
    mix_kernel<A, B, NB><<<n_sms, CAL_THREADS>>>(g);
    unsigned long long c = 1;
    cudaMemcpy(&c, g.cycles, sizeof c, cudaMemcpyDeviceToHost);
    const double n_instr = (double)(CAL_THREADS / 32) * CAL_ITERS * (8 + (B == OP_NONE ? 0 : NB)) * 32;
    return n_instr / (double)c;          // thread instructions per SM-cycle, both classes together
}

template <int B>
void probe_pair(int n_sms, CalArgs g, double *alone, double *with_dpx, double *with_dpx_half) {
    *alone = run_mix<B, OP_NONE, 0>(n_sms, g);
    *with_dpx = run_mix<OP_VIADDMNMX, B, 8>(n_sms, g);
    *with_dpx_half = run_mix<OP_VIADDMNMX, B, 4>(n_sms, g);
}

}  // namespace

// out[3*k + {0,1,2}] for k = op class: rate alone, total rate of 8 DPX + 8 of it, 8 DPX + 4 of it.
extern "C" int osw_calibrate_mix(int device, double *out, int n_out) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OSW_E_NODEV;
    if (cudaSetDevice(device) != cudaSuccess) return OSW_E_CUDA;
    if (n_out < 3 * OP_COUNT) return OSW_E_ARG;
    const int n_sms = prop.multiProcessorCount;
    CalArgs g;
    g.a = 0x00030005u; g.b = 0x00110013u; g.one = 1u; g.ngoe = 0x000c000cu; g.nge = 0xfffefffeu; g.B = 0x00400040u;
    if (cudaMalloc(&g.cycles, 2 * sizeof(unsigned long long)) != cudaSuccess) return OSW_E_NOMEM;
    if (cudaMalloc(&g.sink, (size_t)n_sms * CAL_THREADS * sizeof(uint32_t)) != cudaSuccess) return OSW_E_NOMEM;
    probe_pair<OP_VIADDMNMX>(n_sms, g, out + 0, out + 1, out + 2);
    probe_pair<OP_VIMNMX3>(n_sms, g, out + 3, out + 4, out + 5);
    probe_pair<OP_VIADD>(n_sms, g, out + 6, out + 7, out + 8);
    probe_pair<OP_IMAD>(n_sms, g, out + 9, out + 10, out + 11);
    probe_pair<OP_HMNMX2>(n_sms, g, out + 12, out + 13, out + 14);
    probe_pair<OP_VIMNMX2>(n_sms, g, out + 15, out + 16, out + 17);
    probe_pair<OP_IMNMX>(n_sms, g, out + 18, out + 19, out + 20);
    probe_pair<OP_LOP3>(n_sms, g, out + 21, out + 22, out + 23);
    probe_pair<OP_FMNMX>(n_sms, g, out + 24, out + 25, out + 26);
    probe_pair<OP_PRMT>(n_sms, g, out + 27, out + 28, out + 29);
    probe_pair<OP_SHF>(n_sms, g, out + 30, out + 31, out + 32);
    probe_pair<OP_HADD2>(n_sms, g, out + 33, out + 34, out + 35);
    if (n_out >= 3 * OP_COUNT_EXT) {          // the unpack candidates of the 16-bit pair-database tables
        g.b = 0x00010000u;
        probe_pair<OP_IMADHI>(n_sms, g, out + 36, out + 37, out + 38);
        probe_pair<OP_LEAHI>(n_sms, g, out + 39, out + 40, out + 41);
        probe_pair<OP_IMADSHL>(n_sms, g, out + 42, out + 43, out + 44);
    }
    cudaFree(g.cycles); cudaFree(g.sink);
    return cudaDeviceSynchronize() == cudaSuccess ? OSW_OK : OSW_E_CUDA;
}
