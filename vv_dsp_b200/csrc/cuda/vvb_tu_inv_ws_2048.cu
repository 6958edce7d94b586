/* vvb_tu_inv_ws_2048.cu -- istft_ws_kernel instantiations: warp-specialised marching ISTFT, fft_size 2048 (hop = N/8, N/4, N/2). */
#include "vvb_rt.cuh"
#include "vvb_istft_ws.cuh"

namespace vvb {

template <int S> static int launch_inv_ws_s(InvArgs a, long long batch, int sms, void* stream)
{
    using C = Cfg1024;
    constexpr int NPAIR = 8;
    static OccCache occ;
    auto kern = istft_ws_kernel<C, S, NPAIR>;
    const size_t smem = WsLayout<C, NPAIR>::TOTAL;
    const int per_sm = occ.get(kern, 64 * NPAIR, smem);
    if (per_sm == 0) return rt_fail(4, "istft_ws_kernel", "does not fit on this device");
    if (batch > 0x7fffffffLL) return rt_fail(2, "vvb_stft_inverse", "batch");
    a.num_items = (int)batch;                                   /* the kernel partitions batch*frames itself */
    const long long total = batch * a.frames;
    const long long want = (total + 16 * NPAIR - 1) / (16 * NPAIR);      /* at least ~16 frames per warp pair */
    VVB_LAUNCH(kern, persistent_grid(want, per_sm, sms), 64 * NPAIR, smem, stream, a);
    return 0;
}

int tu_inv_ws_2048(size_t hop, const InvArgs& a, long long batch, int sms, void* stream)
{
    switch (hop) {
    case 256: return launch_inv_ws_s<4>(a, batch, sms, stream);
    case 512: return launch_inv_ws_s<8>(a, batch, sms, stream);
    case 1024: return launch_inv_ws_s<16>(a, batch, sms, stream);
    default: return -1;
    }
}

}  // namespace vvb
