/* vvb_tu_fwd_march_2048.cu -- stft_march_kernel instantiations for fft_size 2048 (hop = N/8, N/4, N/2; three output kinds). */
#include "vvb_launch_march.cuh"
namespace vvb {
int tu_fwd_march_2048(size_t hop, const FwdArgs& a, int kind, int sms, void* stream) { return launch_fwd_march<Cfg1024>(hop, a, kind, sms, stream); }
}
