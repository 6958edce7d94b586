/* vvb_tu_inv_pair.cu -- istft_pair_kernel instantiations: two frames per complex transform (fft_size 256 / 512 / 1024). */
#include "vvb_rt.cuh"

namespace vvb {

/* C = one-warp plan with M = fft_size, CO = the real plan whose window tables are reused */
template <class C> struct PairCfg;
template <> struct PairCfg<Cfg256m> { static constexpr int G = 8, MINB = 3; };
template <> struct PairCfg<Cfg512m> { static constexpr int G = 8, MINB = 2; };
template <> struct PairCfg<Cfg1024> { static constexpr int G = 8, MINB = 1; };

template <class C, class CO, int HS>
static int launch_inv_pair_s(const InvArgs& ia, long long batch, int sms, const float* tables_p, const float* tables_real, void* stream)
{
    constexpr int G = PairCfg<C>::G, MINB = PairCfg<C>::MINB;
    static OccCache occ;
    auto kern = istft_pair_kernel<C, HS, G, MINB>;
    const size_t smem = sizeof(float) * 2 * (C::TW2 + C::TW3 + G * C::XBUF);
    const int per_sm = occ.get(kern, 32 * G, smem);
    if (per_sm == 0) return rt_fail(4, "istft_pair_kernel", "does not fit on this device");
    if (batch > 0x7fffffffLL) return rt_fail(2, "vvb_stft_inverse", "batch");
    PairArgs a;
    a.spec = ia.spec; a.spec_pitch = ia.spec_pitch; a.frames = ia.frames; a.num_items = (int)batch;
    a.y = ia.y; a.y_pitch = ia.y_pitch; a.n_out = ia.n_out; a.inv_norm = ia.inv_norm;
    a.tables = tables_p;
    a.wsyn = tables_real + (ia.inv_norm ? Tables<CO>::WSYN_NORM : Tables<CO>::WSYN);
    a.midnorm = tables_real + Tables<CO>::MIDNORM;
    const long long total = batch * ((ia.frames + 1) / 2);
    const long long want = (total + 8 * G - 1) / (8 * G);
    VVB_LAUNCH(kern, persistent_grid(want, per_sm, sms), 32 * G, smem, stream, a);
    return 0;
}
template <class C, class CO>
static int launch_inv_pair(size_t hop, const InvArgs& a, long long batch, int sms, const float* tp, const float* tr, void* stream)
{
    if (hop % 32) return -1;
    switch (hop / 32) {
    case C::E / 8: return launch_inv_pair_s<C, CO, C::E / 8>(a, batch, sms, tp, tr, stream);
    case C::E / 4: return launch_inv_pair_s<C, CO, C::E / 4>(a, batch, sms, tp, tr, stream);
    case C::E / 2: return launch_inv_pair_s<C, CO, C::E / 2>(a, batch, sms, tp, tr, stream);
    default: return -1;
    }
}

int tu_inv_pair(int nfft, size_t hop, const InvArgs& a, long long batch, int sms, const float* tables_p, const float* tables_real, void* stream)
{
    if (nfft == 256) return launch_inv_pair<Cfg256m, Cfg128>(hop, a, batch, sms, tables_p, tables_real, stream);
    if (nfft == 512) return launch_inv_pair<Cfg512m, Cfg256>(hop, a, batch, sms, tables_p, tables_real, stream);
    if (nfft == 1024) return launch_inv_pair<Cfg1024, Cfg512>(hop, a, batch, sms, tables_p, tables_real, stream);
    return -1;
}

}  // namespace vvb
