/* vvb_tu_fwd_logmel.cu -- stft_forward_kernel<..., OUT_LOGMEL>: samples -> log-mel rows in one kernel for the sizes of the generic
 * forward kernel with sub-warp teams (fft_size 256, 512, 1024 and the speech framings 320 / 400 / 480 / 640), any hop: every warp
 * runs mel_phase on the power rows its own teams have just left in shared memory. */
#include <cstdlib>
#include "vvb_rt.cuh"

namespace vvb {

template <class C> static size_t logmel_smem(const FwdArgs& a)
{
    constexpr int G = Teams<C>::G, TPW = 32 / C::T, NF = TPW < VVB_MEL_NF_MAX ? TPW : VVB_MEL_NF_MAX;
    const int nmp = (a.n_mels + 31) & ~31;
    return ((smem_fwd<C>() + 15) & ~(size_t)15) + (size_t)a.mel_S * 32 * (16 * (size_t)a.mel_unit + 8) +
           sizeof(float) * (G / NF) * (size_t)mel_group_stride(a.mel_prow, nmp, C::T, NF, true);
}
template <class C> static int launch_forward_logmel(FwdArgs a, int sms, void* stream, bool probe)
{
    constexpr int G = Teams<C>::G;
    if (32 / C::T == 2 && a.mel_unit != MEL_U) return 6;       /* two frames per warp: mel_phase<2>, written for MEL_U quads per segment */
    const size_t smem = logmel_smem<C>(a);
    if (smem > 227 * 1024) return 6;                             /* schedule too long for the shared memory left: chained kernels */
    if (probe) return 0;
    a.mel_pair = getenv("VVB_MEL_ROW_SPREAD_OFF") ? 0 : 1;       /* A/B switch: power rows T banks apart (default) or back to back */
    a.groups_per_signal = (a.frames + G - 1) / G;
    static OccCache occ;
    auto kern = stft_forward_kernel<C, G, OUT_LOGMEL>;
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "stft_forward_kernel (log-mel)", "does not fit on this device");
    const long long groups = (long long)a.groups_per_signal * (long long)(a.num_groups);   /* num_groups carries the batch */
    if (groups > 0x7fffffffLL) return rt_fail(2, "vvb_stft_forward_logmel", "batch*frames too large for one launch");
    a.num_groups = (int)groups;
    if (groups == 0) return 0;
    VVB_LAUNCH(kern, persistent_grid(groups, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}

/* m = fft_size / 2; probe: only answer whether the kernel exists and its tables fit (0) or not (6) */
int tu_fwd_logmel(int m, const FwdArgs& a, int sms, void* stream, bool probe)
{
    switch (m) {
    case 128: return launch_forward_logmel<Cfg128>(a, sms, stream, probe);
    case 160: return launch_forward_logmel<Cfg160>(a, sms, stream, probe);
    case 200: return launch_forward_logmel<Cfg200>(a, sms, stream, probe);
    case 240: return launch_forward_logmel<Cfg240>(a, sms, stream, probe);
    case 320: return launch_forward_logmel<Cfg320>(a, sms, stream, probe);
    case 256: return launch_forward_logmel<Cfg256>(a, sms, stream, probe);
    case 512: return launch_forward_logmel<Cfg512>(a, sms, stream, probe);
    default: return 6;
    }
}

}  // namespace vvb
