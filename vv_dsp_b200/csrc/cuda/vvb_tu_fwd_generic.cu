/* vvb_tu_fwd_generic.cu -- stft_forward_kernel instantiations (fft_size 256 ... 8192 and the mixed-radix sizes 320 / 400 / 480 / 640, any hop). */
#include "vvb_rt.cuh"

namespace vvb {

template <class C, int OUT> static int launch_forward_t(FwdArgs a, int sms, void* stream)
{
    constexpr int G = Teams<C>::G;
    a.groups_per_signal = (a.frames + G - 1) / G;
    static OccCache occ;
    auto kern = stft_forward_kernel<C, G, OUT>;
    const size_t smem = smem_fwd<C>();
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "stft_forward_kernel", "does not fit on this device");
    const long long groups = (long long)a.groups_per_signal * (long long)(a.num_groups);   /* num_groups carries batch here */
    if (groups > 0x7fffffffLL) return rt_fail(2, "vvb_stft_forward", "batch*frames too large for one launch");
    a.num_groups = (int)groups;
    if (groups == 0) return 0;
    VVB_LAUNCH(kern, persistent_grid(groups, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}
template <class C> static int launch_forward(const FwdArgs& a, int kind, int sms, void* stream)
{
    switch (kind) {
    case OUT_COMPLEX: return launch_forward_t<C, OUT_COMPLEX>(a, sms, stream);
    case OUT_POWER: return launch_forward_t<C, OUT_POWER>(a, sms, stream);
    case OUT_MAGNITUDE: return launch_forward_t<C, OUT_MAGNITUDE>(a, sms, stream);
    default: return rt_fail(3, "vvb_stft_forward", "bad out_kind");
    }
}

int tu_fwd_generic(int m, const FwdArgs& a, int kind, int sms, void* stream)
{
    switch (m) {
    case 128: return launch_forward<Cfg128>(a, kind, sms, stream);
    case 160: return launch_forward<Cfg160>(a, kind, sms, stream);
    case 200: return launch_forward<Cfg200>(a, kind, sms, stream);
    case 240: return launch_forward<Cfg240>(a, kind, sms, stream);
    case 320: return launch_forward<Cfg320>(a, kind, sms, stream);
    case 256: return launch_forward<Cfg256>(a, kind, sms, stream);
    case 512: return launch_forward<Cfg512>(a, kind, sms, stream);
    case 1024: return launch_forward<Cfg1024>(a, kind, sms, stream);
    case 2048: return launch_forward<Cfg2048>(a, kind, sms, stream);
    case 4096: return launch_forward<Cfg4096>(a, kind, sms, stream);
    default: return rt_fail(6, "vvb_stft_forward", "no Stockham kernel for this size");
    }
}

}  // namespace vvb
