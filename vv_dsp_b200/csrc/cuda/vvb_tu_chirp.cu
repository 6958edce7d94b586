/* vvb_tu_chirp.cu -- chirp_fused_kernel instantiations (Bluestein on the M-point register FFT, M = 128 ... 8192). */
#include "vvb_rt.cuh"

namespace vvb {

template <class C, int MODE> static int launch_chirp_fused_m(const ChirpFusedArgs& a, int sms, void* stream)
{
    constexpr int G = Teams<C>::G;
    static OccCache occ;
    auto kern = chirp_fused_kernel<C, G, MODE>;
    const size_t smem = smem_c2c<C>();
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "chirp_fused_kernel", "does not fit on this device");
    VVB_LAUNCH(kern, persistent_grid((a.count + G - 1) / G, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}
template <class C> static int launch_chirp_fused(const ChirpFusedArgs& a, int sms, void* stream)
{
    if (a.mode == CHIRP_STFT_FWD) return launch_chirp_fused_m<C, CHIRP_STFT_FWD>(a, sms, stream);
    if (a.mode == CHIRP_STFT_INV) return launch_chirp_fused_m<C, CHIRP_STFT_INV>(a, sms, stream);
    return launch_chirp_fused_m<C, CHIRP_C2C>(a, sms, stream);
}

int tu_chirp_fused(size_t M, const ChirpFusedArgs& a, int sms, void* stream)
{
    switch (M) {
    case 128: return launch_chirp_fused<Cfg128>(a, sms, stream);
    case 256: return launch_chirp_fused<Cfg256>(a, sms, stream);
    case 512: return launch_chirp_fused<Cfg512>(a, sms, stream);
    case 1024: return launch_chirp_fused<Cfg1024>(a, sms, stream);
    case 2048: return launch_chirp_fused<Cfg2048>(a, sms, stream);
    case 4096: return launch_chirp_fused<Cfg4096>(a, sms, stream);
    case 8192: return launch_chirp_fused<Cfg8192>(a, sms, stream);
    default: return rt_fail(6, "chirp_fused", "no Stockham kernel for this size");
    }
}

}  // namespace vvb
