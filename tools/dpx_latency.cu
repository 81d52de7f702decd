// dpx_latency.cu - dependent-issue latency of the instructions on the 16-bit kernel's row chain
// (one warp, one chain; clock64 around 4096 dependent instructions).  Experiments only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dpx_latency tools/dpx_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(unsigned *out, unsigned a, unsigned b, unsigned c, long long *cycles) {
    unsigned x = a + threadIdx.x, f = b;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) {
        if (OP == 0) x = __viaddmax_u16x2(x, b, c);                 // VIADDMNMX.U16x2
        if (OP == 1) x = __vimax3_u16x2(x, b, c);                   // VIMNMX3.U16x2
        if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));   // IADD / VIADD
        if (OP == 3) {                                              // one row of the F chain: H -> u -> F
            unsigned h = __vimax3_u16x2(x, f, c);
            unsigned u = h - b;
            f = __viaddmax_u16x2(f, a, u);
            x = h ^ i;
        }
        if (OP == 4) x = __vmaxu2(x, b) + i;                          // 2-input packed max (+ add to keep it live)
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + f;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const char *names[5] = {"VIADDMNMX.U16x2", "VIMNMX3.U16x2", "IADD", "row of the F chain (VIMNMX3 -> sub -> VIADDMNMX)", "VIMNMX.U16x2 + IADD"};
    for (int op = 0; op < 5; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
                case 0: chain<0><<<1, 32>>>(out, 3, 5, 7, cyc); break;
                case 1: chain<1><<<1, 32>>>(out, 3, 5, 7, cyc); break;
                case 2: chain<2><<<1, 32>>>(out, 3, 5, 7, cyc); break;
                case 3: chain<3><<<1, 32>>>(out, 3, 5, 7, cyc); break;
                case 4: chain<4><<<1, 32>>>(out, 3, 5, 7, cyc); break;
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("%-55s %.2f cycles per dependent step\n", names[op], h / 4096.0);
    }
    return 0;
}
