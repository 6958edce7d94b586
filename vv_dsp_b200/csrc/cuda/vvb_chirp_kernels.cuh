/*
 * vvb_chirp_kernels.cuh -- the fused chirp-z (Bluestein) kernel for sizes without a Stockham kernel:
 * STFT analysis / synthesis frames and plan-API C2C on top of the power-of-two register FFT
 * (the reference serves these sizes with its O(n^2) DFT, src/spectral/fft_kiss.c:76-92,115; its own
 * Bluestein lives in src/spectral/czt.c:126-160).  Template only, so several translation units may include it.
 */
#pragma once
#include "vvb_stft_kernels.cuh"

namespace vvb {

/* The whole chirp-z transform in ONE kernel: a team (the same T = M/E threads and register layout as
 * fft_c2c_kernel) loads its transform with the pre-multiplication fused in, runs the forward FFT_M in
 * registers, multiplies by the chirp spectrum (1/M folded in), turns the result back into pass-1 order through
 * its exchange buffer, runs the second FFT_M on re/im-swapped data (= the inverse) and applies the
 * post-multiplication on the way out.  HBM sees the input samples and the n-point result only; the
 * multi-kernel version above moved 8 x M complex values per transform through HBM (kept for sizes whose
 * team does not fit, and as a cross-check: VVB_BLUESTEIN_UNFUSED=1). */
enum { CHIRP_STFT_FWD = 0, CHIRP_STFT_INV = 1, CHIRP_C2C = 2 };
struct ChirpFusedArgs {
    long long count;                 /* transforms */
    int n, mode;
    const float2* chirp; const float2* bspec_over_m; const float* win;
    const float* tables;             /* Tables<C> blob of the M-point C2C plan */
    /* STFT_FWD */
    const float* x; long long x_pitch, n_sig; int frames, hop, pad_mode, out_kind; void* out; long long out_pitch;
    /* STFT_INV */
    const float2* spec; long long spec_pitch; float* frames_out;
    /* C2C */
    const float2* cin; float2* cout; int inverse;
};

template <class C, int G, int MODE>
__global__ void __launch_bounds__(C::T* G) chirp_fused_kernel(const ChirpFusedArgs a)
{
    using TB = Tables<C>;
    using L = LastPass<C>;
    constexpr int M = C::M, E = C::E, T = C::T;
#ifdef VVB_EMU
    float* smem = reinterpret_cast<float*>(vvb_emu::g_dyn_smem);
#else
    extern __shared__ __align__(16) float smem[];
#endif
    float2* s_tw2 = reinterpret_cast<float2*>(smem);
    float2* s_tw3 = s_tw2 + C::TW2;
    float2* s_xb = s_tw3 + C::TW3;
    copy_table(reinterpret_cast<float*>(s_tw2), a.tables + TB::TW2, 2 * (C::TW2 + C::TW3));
    __syncthreads();
    const int team = threadIdx.x / T, t = threadIdx.x % T;
    float2* xb = s_xb + team * C::XBUF;
    const int n = a.n;
    const float invn = 1.0f / (float)n;
    const long long groups = (a.count + G - 1) / G;
    for (long long group = blockIdx.x; group < groups; group += gridDim.x) {
        const long long id = group * G + team;
        const bool active = id < a.count;
        float2 v[E];
        {
            /* phase 1: every global load of this transform is issued before any of them is used */
            constexpr int R = C::R1, NQ = E / R, STRIDE = M / R;
            if constexpr (MODE == CHIRP_STFT_FWD) {
                const long long b = active ? id / a.frames : 0;
                const int f = (int)(id - b * a.frames);
                const float* xs = a.x + b * a.x_pitch;
                const long long start = (long long)f * a.hop - (a.pad_mode == PAD_REFLECT ? n / 2 : 0);
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = t + T * q + r * STRIDE;
                        v[q * R + r].x = (active && i < n) ? fetch_sample(xs, a.n_sig, start + i, a.pad_mode) : 0.0f;
                    }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = t + T * q + r * STRIDE;
                        float2 s = make_float2(0.f, 0.f);
                        if (i < n) {
                            const float2 c = __ldg(a.chirp + i);
                            const float val = v[q * R + r].x * __ldg(a.win + i);
                            s = make_float2(val * c.x, val * c.y);
                        }
                        v[q * R + r] = s;
                    }
            } else {
                const float2* src = (MODE == CHIRP_STFT_INV) ? a.spec + (active ? id : 0) * a.spec_pitch : a.cin + (active ? id : 0) * n;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = t + T * q + r * STRIDE;
                        const int from = (MODE == CHIRP_STFT_INV && 2 * i > n) ? n - i : i;      /* Hermitian extension */
                        v[q * R + r] = (active && i < n) ? __ldg(src + from) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = t + T * q + r * STRIDE;
                        float2 h = v[q * R + r];
                        if constexpr (MODE == CHIRP_STFT_INV) {
                            /* feed conj(Xfull): the mirrored half is already a conjugate, the lower half gets one here */
                            if (2 * i <= n) h.y = -h.y;
                            if (i == 0 || 2 * i == n) h.y = 0.0f;
                        } else if (a.inverse) h.y = -h.y;
                        v[q * R + r] = (i < n) ? cmul(h, __ldg(a.chirp + i)) : make_float2(0.f, 0.f);
                    }
            }
        }
        team_fft<C>(v, xb, s_tw2, s_tw3, t, team);
        /* times the chirp spectrum, parked in natural order, re-read in pass-1 order with re/im swapped */
#pragma unroll
        for (int q = 0; q < L::NQ; ++q)
#pragma unroll
            for (int r = 0; r < L::R; ++r) {
                const int k = t + T * q + r * L::NS;
                xb[C::pad(k)] = cmul(v[q * L::R + ct_bitrev(r, L::R)], __ldg(a.bspec_over_m + k));
            }
        team_sync<T>(team);
        {
            constexpr int R = C::R1, NQ = E / R, STRIDE = M / R;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 z = xb[C::pad(t + T * q + r * STRIDE)];
                    v[q * R + r] = make_float2(z.y, z.x);
                }
        }
        team_sync<T>(team);
        team_fft<C>(v, xb, s_tw2, s_tw3, t, team);
        if (active) {
#pragma unroll
            for (int q = 0; q < L::NQ; ++q)
#pragma unroll
                for (int r = 0; r < L::R; ++r) {
                    const int k = t + T * q + r * L::NS;
                    const bool wanted = (MODE == CHIRP_STFT_FWD) ? (2 * k <= n) : (k < n);
                    if (!wanted) continue;
                    const float2 zs = v[q * L::R + ct_bitrev(r, L::R)];
                    float2 z = cmul(make_float2(zs.y, zs.x), __ldg(a.chirp + k));       /* un-swap, post-chirp */
                    if constexpr (MODE == CHIRP_STFT_FWD) {
                        if (k == 0 || 2 * k == n) z.y = 0.0f;
                        const long long o = id * a.out_pitch + k;
                        if (a.out_kind == OUT_COMPLEX) reinterpret_cast<float2*>(a.out)[o] = z;
                        else if (a.out_kind == OUT_POWER) reinterpret_cast<float*>(a.out)[o] = z.x * z.x + z.y * z.y;
                        else reinterpret_cast<float*>(a.out)[o] = sqrtf(z.x * z.x + z.y * z.y);
                    } else if constexpr (MODE == CHIRP_STFT_INV) {
                        float val = z.x * invn;
                        if (a.win) val *= __ldg(a.win + k);
                        a.frames_out[id * n + k] = val;
                    } else {
                        a.cout[id * n + k] = a.inverse ? make_float2(z.x * invn, -z.y * invn) : z;
                    }
                }
        }
        team_sync<T>(team);
    }
}

}  // namespace vvb
