/* vvb_tu_inv_march_2048.cu -- istft_march_kernel instantiations for fft_size 2048 (hop = N/8, N/4, N/2). */
#include "vvb_launch_march.cuh"
namespace vvb {
int tu_inv_march_2048(size_t hop, const InvArgs& a, long long batch, int sms, void* stream) { return launch_inv_march<Cfg1024>(hop, a, batch, sms, stream); }
}
