/* vvb_tu_c2c.cu -- fft_c2c_kernel instantiations (plan API, power-of-two n = 128 ... 8192). */
#include "vvb_rt.cuh"

namespace vvb {

template <class C> static int launch_c2c(C2CArgs a, int sms, void* stream)
{
    constexpr int G = Teams<C>::G;
    static OccCache occ;
    auto kern = fft_c2c_kernel<C, G>;
    const size_t smem = smem_c2c<C>();
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "fft_c2c_kernel", "does not fit on this device");
    const long long groups = (a.batch + G - 1) / G;
    VVB_LAUNCH(kern, persistent_grid(groups, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}

int tu_c2c(int n, const C2CArgs& a, int sms, void* stream)
{
    switch (n) {
    case 128: return launch_c2c<Cfg128>(a, sms, stream);
    case 256: return launch_c2c<Cfg256>(a, sms, stream);
    case 512: return launch_c2c<Cfg512>(a, sms, stream);
    case 1024: return launch_c2c<Cfg1024>(a, sms, stream);
    case 2048: return launch_c2c<Cfg2048>(a, sms, stream);
    case 4096: return launch_c2c<Cfg4096>(a, sms, stream);
    case 8192: return launch_c2c<Cfg8192>(a, sms, stream);
    default: return rt_fail(6, "vvb_fft_exec", "no Stockham kernel for this size");
    }
}

}  // namespace vvb
