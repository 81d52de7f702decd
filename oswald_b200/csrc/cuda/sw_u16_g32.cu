// sw_u16_g32.cu - instances of the first-stage kernel with G = 32 lanes per database sequence.
#include "sw_u16_kernel.cuh"
namespace osw_u16 {
int launch_g32(int R, const KArgs &a, int n_sms, cudaStream_t st) { return launch_g<32>(R, a, n_sms, st); }
}
