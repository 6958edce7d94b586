/* vvb_launch_march.cuh -- launchers of the team-marching kernels, shared by the per-size translation units. */
#pragma once
#include "vvb_rt.cuh"

namespace vvb {

template <class C, int S, int OUT> static int launch_fwd_march_t(FwdArgs a, int sms, void* stream)
{
    constexpr int G = March<C>::G, MINB = March<C>::MINB;
    static OccCache occ;
    auto kern = stft_march_kernel<C, S, G, MINB, OUT>;
    size_t smem = sizeof(float) * 2 * (C::TW2 + C::TW3 + C::POST + 1 + G * C::XBUF + G * (C::E / S + 1) * C::T * S) + 8 * G;
    if constexpr (OUT == OUT_LOGMEL)      /* + schedule tables and one row of band sums per team (see mel_phase) */
        smem += mel_smem_bytes(G, a.mel_S, a.n_mels, a.mel_prow, a.mel_pair);
    if (smem > 227 * 1024) return 6;                             /* schedule too long for the shared memory left: unfused chain */
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "stft_march_kernel", "does not fit on this device");
    const long long total = (long long)a.num_groups * a.frames;  /* num_groups carries the batch */
    const long long want = (total + 16 * G - 1) / (16 * G);      /* at least ~16 frames per team */
    VVB_LAUNCH(kern, persistent_grid(want, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}
template <class C, int S> static int launch_fwd_march_s(const FwdArgs& a, int kind, int sms, void* stream)
{
    switch (kind) {
    case OUT_COMPLEX: return launch_fwd_march_t<C, S, OUT_COMPLEX>(a, sms, stream);
    case OUT_POWER: return launch_fwd_march_t<C, S, OUT_POWER>(a, sms, stream);
    case OUT_MAGNITUDE: return launch_fwd_march_t<C, S, OUT_MAGNITUDE>(a, sms, stream);
    case OUT_LOGMEL:
        if constexpr (C::T == 32) return launch_fwd_march_t<C, S, OUT_LOGMEL>(a, sms, stream);
        else return -1;
    default: return rt_fail(3, "vvb_stft_forward", "bad out_kind");
    }
}
/* returns -1 when this (fft_size, hop) has no marching kernel */
template <class C> static int launch_fwd_march(size_t hop, const FwdArgs& a, int kind, int sms, void* stream)
{
    const size_t unit = 2 * (size_t)C::T;                        /* samples per register slot of a team */
    if (hop % unit) return -1;
    switch (hop / unit) {
    case C::E / 8: return launch_fwd_march_s<C, C::E / 8>(a, kind, sms, stream);
    case C::E / 4: return launch_fwd_march_s<C, C::E / 4>(a, kind, sms, stream);
    case C::E / 2: return launch_fwd_march_s<C, C::E / 2>(a, kind, sms, stream);
    default: return -1;
    }
}

/* team-marching ISTFT: register-resident overlap-add (see istft_march_kernel) */
template <class C, int S> static int launch_inv_march_s(InvArgs a, long long batch, int sms, void* stream)
{
    constexpr int G = March<C>::G, MINB = March<C>::MINB;
    static OccCache occ;
    auto kern = istft_march_kernel<C, S, G, MINB>;
    const size_t smem = sizeof(float) * (2 * C::M + 2 * (C::TW2 + C::TW3) + 2 * G * (C::XBUF + C::M + 2)) + 8 * G;
    const int per_sm = occ.get(kern, C::T * G, smem);
    if (per_sm == 0) return rt_fail(4, "istft_march_kernel", "does not fit on this device");
    if (batch > 0x7fffffffLL) return rt_fail(2, "vvb_stft_inverse", "batch");
    a.num_items = (int)batch;                                   /* the kernel partitions batch*frames itself */
    const long long total = batch * a.frames;
    const long long want = (total + 16 * G - 1) / (16 * G);
    VVB_LAUNCH(kern, persistent_grid(want, per_sm, sms), C::T * G, smem, stream, a);
    return 0;
}
template <class C> static int launch_inv_march(size_t hop, const InvArgs& a, long long batch, int sms, void* stream)
{
    const size_t unit = 2 * (size_t)C::T;
    if (hop % unit) return -1;
    switch (hop / unit) {
    case C::E / 8: return launch_inv_march_s<C, C::E / 8>(a, batch, sms, stream);
    case C::E / 4: return launch_inv_march_s<C, C::E / 4>(a, batch, sms, stream);
    case C::E / 2: return launch_inv_march_s<C, C::E / 2>(a, batch, sms, stream);
    default: return -1;
    }
}

}  // namespace vvb
