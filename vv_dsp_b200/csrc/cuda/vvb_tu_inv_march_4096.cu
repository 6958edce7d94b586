/* vvb_tu_inv_march_4096.cu -- istft_march_kernel instantiations for fft_size 4096 (hop = N/8, N/4, N/2). */
#include "vvb_launch_march.cuh"
namespace vvb {
int tu_inv_march_4096(size_t hop, const InvArgs& a, long long batch, int sms, void* stream) { return launch_inv_march<Cfg2048>(hop, a, batch, sms, stream); }
}
