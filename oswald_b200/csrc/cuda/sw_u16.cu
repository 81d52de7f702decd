// sw_u16.cu - launch of the first-stage kernel (the kernel itself: sw_u16_kernel.cuh; its
// instances are compiled in sw_u16_g4.cu ... sw_u16_g32.cu, one translation unit per group width).
#include "sw_u16_kernel.cuh"

void osw_fill_kargs(osw_u16::KArgs &a, const U16Params &p, const OswPass &pass) {
    a.p = p;
    memcpy(a.lane, pass.lane, sizeof a.lane);
    a.has_in = pass.has_in && p.bound ? 1u : 0u;
    a.has_out = pass.has_out && p.bound ? 1u : 0u;
    a.pair_db = pass.pair_db ? 1u : 0u;
    const uint32_t goe = (uint32_t)p.gap_open_extend, ge = (uint32_t)p.gap_extend;
    const uint32_t B = goe + ge + 32u;
    a.bias = B; a.bias2 = B | (B << 16);
    const uint32_t nge = (0x10000u - ge) & 0xffffu;
    a.nge2 = nge | (nge << 16);
    a.ngoe_word = 0u - (goe | (goe << 16));
    a.n_ctas = 0; a.warps = 0;
}

int osw_launch_u16(const U16Params &p, const OswPass &pass, int n_sms, cudaStream_t st) {
    if ((pass.has_in || pass.has_out) && !p.bound) return OSW_E_ARG;
    osw_u16::KArgs a;
    osw_fill_kargs(a, p, pass);
    switch (pass.G) {
        case 4:  return osw_u16::launch_g4(pass.R, a, n_sms, st);
        case 8:  return osw_u16::launch_g8(pass.R, a, n_sms, st);
        case 16: return osw_u16::launch_g16(pass.R, a, n_sms, st);
        case 32: return osw_u16::launch_g32(pass.R, a, n_sms, st);
    }
    return OSW_E_ARG;
}
