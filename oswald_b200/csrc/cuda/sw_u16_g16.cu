// sw_u16_g16.cu - instances of the first-stage kernel with G = 16 lanes per database sequence.
#include "sw_u16_kernel.cuh"
namespace osw_u16 {
int launch_g16(int R, const KArgs &a, int n_sms, cudaStream_t st) { return launch_g<16>(R, a, n_sms, st); }
}
