/* vvb_tu_inv_march_small.cu -- istft_march_kernel for fft_size 512 / 1024 (whole-warp three-pass configurations;
 * used when VVB_NO_PAIR disables istft_pair_kernel). */
#include "vvb_launch_march.cuh"
namespace vvb {
int tu_inv_march_small(int nfft, size_t hop, const InvArgs& a, long long batch, int sms, void* stream)
{
    if (nfft == 512) return launch_inv_march<Cfg256m>(hop, a, batch, sms, stream);
    if (nfft == 1024) return launch_inv_march<Cfg512m>(hop, a, batch, sms, stream);
    return -1;
}
}
