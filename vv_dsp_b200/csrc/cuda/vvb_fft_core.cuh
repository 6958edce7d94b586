/*
 * vvb_fft_core.cuh -- register/shared-memory FFT building blocks (sm_100a).
 *
 * Replaces the arithmetic of the reference's src/spectral/fft_kiss.c:27-74 (iterative
 * radix-2 with a float twiddle recurrence) with a Stockham autosort FFT:
 *   - a "team" of T = M/E threads owns one M-point complex transform, E points per
 *     thread held in registers;
 *   - each pass is a radix-R (R = 8/16/32) DFT done entirely in registers with
 *     compile-time twiddles, followed by one exchange through shared memory;
 *   - inter-pass twiddles come from tables computed in double on the host (no
 *     recurrence, so the error does not grow with N like the reference's does).
 * Real-input transforms of length N = 2M are done as one M-point complex transform
 * plus a split step (vvb_stft_kernels.cu), i.e. half the flops of the reference's
 * C2C-on-(x,0) (src/spectral/stft.c:83-90).
 *
 * Index algebra of one pass (radix R, Ns = product of the radices already done),
 * work item j in [0, M/R):
 *     v[r]  = in[j + r*M/R]                      r = 0..R-1
 *     v[r] *= exp(-2*pi*i * r*(j % Ns) / (Ns*R))
 *     v     = DFT_R(v)
 *     out[(j/Ns)*Ns*R + (j % Ns) + r*Ns] = v[r]
 * After the last pass `out` is in natural frequency order.
 */
#pragma once

#ifdef VVB_EMU
#include "cuda_emu.h"
#include <type_traits>
#define VVB_DEV inline __attribute__((always_inline))
#define VVB_CX constexpr
#define VVB_MAXNREG(n)
#else
#include <cuda_runtime.h>
#include <type_traits>
#define VVB_DEV __device__ __forceinline__
#define VVB_CX __host__ __device__ constexpr
#define VVB_MAXNREG(n) __maxnreg__(n)      /* explicit per-thread register cap (one CTA of W warps per SM) */
#endif

namespace vvb {

/* ---------------------------------------------------------------- compile-time trig */
constexpr double kPi = 3.141592653589793238462643383279502884;

constexpr double ct_sin_small(double x)   /* |x| <= pi/4 */
{
    double term = x, sum = x, x2 = x * x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double ct_cos_small(double x)
{
    double term = 1.0, sum = 1.0, x2 = x * x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
/* cos / sin of 2*pi*num/den for 0 <= num < den, octant-reduced so the series stays accurate */
constexpr double ct_cos2pi(int num, int den)
{
    int n8 = (8 * num) / den;                         /* octant 0..7 */
    double frac = (double)num / den;
    switch (n8) {
    case 0: return ct_cos_small(2 * kPi * frac);
    case 1: case 2: return -ct_sin_small(2 * kPi * (frac - 0.25));
    case 3: case 4: return -ct_cos_small(2 * kPi * (frac - 0.5));
    case 5: case 6: return ct_sin_small(2 * kPi * (frac - 0.75));
    default: return ct_cos_small(2 * kPi * (frac - 1.0));
    }
}
constexpr double ct_sin2pi(int num, int den)
{
    int n8 = (8 * num) / den;
    double frac = (double)num / den;
    switch (n8) {
    case 0: return ct_sin_small(2 * kPi * frac);
    case 1: case 2: return ct_cos_small(2 * kPi * (frac - 0.25));
    case 3: case 4: return -ct_sin_small(2 * kPi * (frac - 0.5));
    case 5: case 6: return -ct_cos_small(2 * kPi * (frac - 0.75));
    default: return ct_sin_small(2 * kPi * (frac - 1.0));
    }
}

template <int N, int I> struct TwC {
    static constexpr float c = (float)ct_cos2pi(I, N);
    static constexpr float s = (float)ct_sin2pi(I, N);
};

/* ------------------------------------------------------------------ complex helpers */
/* Complex numbers live in aligned register pairs and all arithmetic uses Blackwell's packed FP32
 * instructions (FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per issue slot).  The lane swaps,
 * scalar broadcasts and per-half negations written below as make_float2(...) fold into the
 * instructions' operand selectors (.LO_HI, .F32, -, .NP) -- verified in SASS -- so a complex add is
 * one instruction and a complex multiply two. */
VVB_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
VVB_DEV float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
VVB_DEV float2 cmul(float2 a, float2 w)            /* a * w */
{
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(w.y, w.y));       /* (a.y w.y, a.x w.y) */
    return __ffma2_rn(a, make_float2(w.x, w.x), make_float2(-t.x, t.y));
}
VVB_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
VVB_DEV float2 cswap(float2 a) { return make_float2(a.y, a.x); }
VVB_DEV float2 splat(float s) { return make_float2(s, s); }

/* x * exp(-2*pi*i*I/N), trivial factors resolved at compile time */
template <int N, int I> VVB_DEV float2 mul_w(float2 x)
{
    if constexpr (I == 0) {
        return x;
    } else if constexpr (4 * I == N) {            /* -i */
        return make_float2(x.y, -x.x);
    } else if constexpr (2 * I == N) {            /* -1 */
        return make_float2(-x.x, -x.y);
    } else if constexpr (8 * I == N) {            /* (1-i)/sqrt2 */
        constexpr float h = 0.70710678118654752440f;
        return make_float2((x.x + x.y) * h, (x.y - x.x) * h);
    } else if constexpr (8 * I == 3 * N) {        /* (-1-i)/sqrt2 */
        constexpr float h = 0.70710678118654752440f;
        return make_float2((x.y - x.x) * h, -(x.x + x.y) * h);
    } else {
        constexpr float c = TwC<N, I>::c, s = TwC<N, I>::s;   /* w = c - i s */
        return make_float2(x.x * c + x.y * s, x.y * c - x.x * s);
    }
}

/* ------------------------------------------- in-register DFT, radix-2 DIT, FMA-folded */
/* Decimation in time: X[k] = E[k] + W^k O[k], X[k+N/2] = E[k] - W^k O[k].  With the twiddle a
 * compile-time constant W = c - i s, the product is folded into the butterfly so a general
 * butterfly is 6 FMAs = 3 packed FFMA2 instead of 4 mul/fma + 4 add:
 *     |c| >= |s|:  b' = b * (1 - i s/c)       (1 FFMA2)    X = a +- c b'   (2 FFMA2)
 *     |c| <  |s|:  b' = b * (c/s - i)         (1 FFMA2)    X = a +- s b'   (2 FFMA2)
 * W = 1 and W = -i need 2 FADD2.  A 32-point DFT is 194 packed instructions (388 scalar ones in
 * the same DIT form, 456 in the decimation-in-frequency form that came first).  Input and output are in natural order; the
 * bit reversal of the recursion is only a renaming of registers. */
template <int N, int K> struct DitTw {       /* all evaluated by the host compiler: plain immediates in SASS */
    static constexpr double cd = ct_cos2pi(K, N), sd = ct_sin2pi(K, N);
    static constexpr bool cos_form = (cd < 0 ? -cd : cd) >= (sd < 0 ? -sd : sd);
    static constexpr float c = (float)cd, sn = (float)sd;
    static constexpr float tn = cos_form ? (float)(sd / cd) : 0.f;      /* tan */
    static constexpr float ct = cos_form ? 0.f : (float)(cd / sd);      /* cot */
};

template <int N, int K> VVB_DEV void dit_combine(float2 a, float2 b, float2& lo, float2& hi)
{
    if constexpr (K == 0) {
        lo = cadd(a, b); hi = csub(a, b);                                       /* 2 x FADD2 */
    } else if constexpr (4 * K == N) {            /* W = -i: W b = (b.y, -b.x) */
        lo = __fadd2_rn(a, make_float2(b.y, -b.x));
        hi = __fadd2_rn(a, make_float2(-b.y, b.x));
    } else {
        using TW = DitTw<N, K>;                   /* 3 x FFMA2 */
        if constexpr (TW::cos_form) {
            constexpr float tn = TW::tn, c = TW::c;
            const float2 p = __ffma2_rn(make_float2(tn, -tn), make_float2(b.y, b.x), b);     /* b (1 - i tn) */
            lo = __ffma2_rn(splat(c), p, a);
            hi = __ffma2_rn(splat(-c), p, a);
        } else {
            constexpr float ct = TW::ct, sn = TW::sn;
            const float2 p = __ffma2_rn(splat(ct), b, make_float2(b.y, -b.x));               /* b (ct - i) */
            lo = __ffma2_rn(splat(sn), p, a);
            hi = __ffma2_rn(splat(-sn), p, a);
        }
    }
}

template <int... Is> struct iseq {};
template <int N, int... Is> struct make_iseq : make_iseq<N - 1, N - 1, Is...> {};
template <int... Is> struct make_iseq<0, Is...> { using type = iseq<Is...>; };

/* DFT of in[OFF + STRIDE*i], i < N, into out[0..N) */
template <int N, int STRIDE, int OFF> struct Dit {
    template <int... Ks> VVB_DEV static void combine(const float2* e, const float2* o, float2* out, iseq<Ks...>)
    {
        (dit_combine<N, Ks>(e[Ks], o[Ks], out[Ks], out[Ks + N / 2]), ...);
    }
    VVB_DEV static void run(const float2* in, float2* out)
    {
        float2 e[N / 2], o[N / 2];
        Dit<N / 2, 2 * STRIDE, OFF>::run(in, e);
        Dit<N / 2, 2 * STRIDE, OFF + STRIDE>::run(in, o);
        combine(e, o, out, typename make_iseq<N / 2>::type{});
    }
};
template <int STRIDE, int OFF> struct Dit<1, STRIDE, OFF> {
    VVB_DEV static void run(const float2* in, float2* out) { out[0] = in[OFF]; }
};
/* 3-point DFT (odd leaf of the radix-12 pass of fft_size 480), forward sign:
 *     t = x1 + x2,  d = x1 - x2,  y0 = x0 + t,  m = x0 - t/2,  y1 = m - j (sqrt3/2) d,  y2 = m + j (sqrt3/2) d */
template <int STRIDE, int OFF> struct Dit<3, STRIDE, OFF> {
    VVB_DEV static void run(const float2* in, float2* out)
    {
        constexpr float s1 = TwC<3, 1>::s;                            /* sin 120 degrees */
        const float2 x0 = in[OFF], x1 = in[OFF + STRIDE], x2 = in[OFF + 2 * STRIDE];
        const float2 t = cadd(x1, x2), d = __fmul2_rn(splat(s1), csub(x1, x2));
        out[0] = cadd(x0, t);
        const float2 m = __ffma2_rn(splat(-0.5f), t, x0);
        out[1] = __fadd2_rn(m, make_float2(d.y, -d.x));               /* m - j d */
        out[2] = __fadd2_rn(m, make_float2(-d.y, d.x));               /* m + j d */
    }
};
/* 5-point DFT (the odd leaf of the radix-10 / radix-20 passes of fft_size 320 / 400 / 480 / 640), forward sign:
 *     t1 = x1 + x4, t2 = x2 + x3, t3 = x1 - x4, t4 = x2 - x3
 *     y0 = x0 + t1 + t2,  a = x0 + c1 t1 + c2 t2,  b = x0 + c2 t1 + c1 t2      (c1 = cos 72, c2 = cos 144 degrees)
 *     p = s1 t3 + s2 t4,  q = s2 t3 - s1 t4                                    (s1 = sin 72, s2 = sin 144 degrees)
 *     y1 = a - j p,  y4 = a + j p,  y2 = b - j q,  y3 = b + j q */
template <int STRIDE, int OFF> struct Dit<5, STRIDE, OFF> {
    VVB_DEV static void run(const float2* in, float2* out)
    {
        constexpr float c1 = TwC<5, 1>::c, c2 = TwC<5, 2>::c, s1 = TwC<5, 1>::s, s2 = TwC<5, 2>::s;
        const float2 x0 = in[OFF], x1 = in[OFF + STRIDE], x2 = in[OFF + 2 * STRIDE], x3 = in[OFF + 3 * STRIDE], x4 = in[OFF + 4 * STRIDE];
        const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
        out[0] = cadd(x0, cadd(t1, t2));
        const float2 a = __ffma2_rn(splat(c2), t2, __ffma2_rn(splat(c1), t1, x0));
        const float2 b = __ffma2_rn(splat(c1), t2, __ffma2_rn(splat(c2), t1, x0));
        const float2 p = __ffma2_rn(splat(s2), t4, __fmul2_rn(splat(s1), t3));
        const float2 q = __ffma2_rn(splat(-s1), t4, __fmul2_rn(splat(s2), t3));
        out[1] = __fadd2_rn(a, make_float2(p.y, -p.x));                /* a - j p */
        out[4] = __fadd2_rn(a, make_float2(-p.y, p.x));                /* a + j p */
        out[2] = __fadd2_rn(b, make_float2(q.y, -q.x));
        out[3] = __fadd2_rn(b, make_float2(-q.y, q.x));
    }
};
/* (M - k) mod M for 0 <= k < M: a mask for the power-of-two sizes (unchanged code), a select for the others */
template <int M> VVB_DEV int neg_mod(int k)
{
    if constexpr ((M & (M - 1)) == 0) return (M - k) & (M - 1);
    else return k == 0 ? 0 : M - k;
}

/* in-place (register renaming) DFT of v[O..O+N), natural order in and out */
template <int N, int O> VVB_DEV void fft_reg(float2* v)
{
    float2 out[N];
    Dit<N, 1, 0>::run(v + O, out);
#pragma unroll
    for (int i = 0; i < N; ++i) v[O + i] = out[i];
}

/* register slot of output r of an R-point fft_reg (identity: natural order) */
VVB_CX int ct_bitrev(int r, int) { return r; }

#ifndef VVB_SWIZZLE_884
#define VVB_SWIZZLE_884 1             /* one-warp 8.8.4 configuration: XOR swizzle of the exchange buffer instead of padding (A/B builds: 0) */
#endif
/* ------------------------------------------------------------ team configuration */
/* M complex points, E per thread, up to three passes with radices R1*R2*R3 == M */
template <int M_, int E_, int R1_, int R2_, int R3_ = 1> struct Cfg {
    static constexpr int M = M_, E = E_, T = M_ / E_;
    static constexpr int R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int NP = (R3_ > 1) ? 3 : 2;
    static constexpr int XBUF = M_ + M_ / R1_;               /* padded float2 per team buffer */
    static constexpr int TW2 = (R2_ - 1) * R1_;               /* pass-2 twiddles (float2) */
    static constexpr int TW3 = (R3_ > 1) ? (R3_ - 1) * R1_ * R2_ : 0;
    static constexpr int POST = M_ / 2 + 1;                   /* split-step twiddles */
    static_assert(R1_ * R2_ * R3_ == M_, "radices must multiply to M");
    static_assert(E_ % R1_ == 0 && E_ % R2_ == 0 && E_ % R3_ == 0, "E must be a multiple of every radix");
    static_assert(E_ >= 2 && (M_ / 2) % T == 0, "split step needs M/2 pairs divisible over the team");
    /* padded shared-memory position: one float2 of padding every R1 elements makes the
     * stride-R1 writes of pass 1 hit 16 distinct bank pairs per half warp.
     * The one-warp 8.8.4 configuration (fft_size 256 ISTFT) is swizzled instead: with the padding its pass-2 / pass-3 reads
     * (16 consecutive elements per half warp) spanned two padding steps and lanes 0 and 15 met in one bank pair -- ncu:
     * 2 x the ideal wavefronts on all 16 loads, a quarter of the kernel's shared-memory traffic at 89 % of that peak.
     * XOR of the bank-pair index with bits 4-6 and of its top bit with bit 6 of the element index keeps the stride-8 writes
     * of pass 1, the (64 a + b) writes of pass 2 and the consecutive reads all conflict-free. */
    VVB_DEV static int pad(int i)
    {
        if constexpr (VVB_SWIZZLE_884 && M_ == 256 && E_ == 8 && R1_ == 8 && R2_ == 8) return i ^ ((i >> 4) & 7) ^ (((i >> 6) & 1) << 3);
        else return i + i / R1_;
    }
};

/* team barrier: sub-warp and warp teams use __syncwarp (all teams of a warp run in
 * lockstep on the same control path), larger teams a named barrier per team */
template <int T> VVB_DEV void team_sync(int team)
{
    if constexpr (T <= 32) {
        __syncwarp();
    } else {
#ifdef VVB_EMU
        emu_named_barrier(1 + team, T);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(T) : "memory");
#endif
    }
}

/* One Stockham pass on the team's registers.
 *   FIRST: v was filled by the caller (no shared-memory read, Ns == 1, no twiddle)
 *   LAST : results stay in registers: X[j + r*NS] = v[q*R + bitrev(r)], j = t + T*q
 * tw: this pass's table, tw[(r-1)*NS + (j % NS)] = exp(-2 pi i r (j%NS) / (NS*R)). */
struct NoHook { VVB_DEV void operator()() const {} };

/* after_read: called once the team has finished READING xb in this pass (xb is free from then on
 * if the pass is the last one) -- used to start the next frame's asynchronous prefetch into xb. */
template <class C, int R, int NS, bool FIRST, bool LAST, class Hook = NoHook>
VVB_DEV void stockham_pass(float2 (&v)[C::E], float2* xb, const float2* tw, int t, int team, Hook after_read = Hook())
{
    constexpr int NQ = C::E / R;
    constexpr int STRIDE = C::M / R;
    if constexpr (!FIRST) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int j = t + C::T * q;
#pragma unroll
            for (int r = 0; r < R; ++r) v[q * R + r] = xb[C::pad(j + r * STRIDE)];
        }
        team_sync<C::T>(team);     /* everyone has read before anyone overwrites xb */
        after_read();
        if constexpr (NS > 1) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int jm = (t + C::T * q) % NS;
#pragma unroll
                for (int r = 1; r < R; ++r) v[q * R + r] = cmul(v[q * R + r], tw[(r - 1) * NS + jm]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
    if constexpr (!LAST) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int j = t + C::T * q;
            const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
            for (int r = 0; r < R; ++r) xb[C::pad(j0 + r * NS)] = v[q * R + ct_bitrev(r, R)];
        }
        team_sync<C::T>(team);
    }
}

/* All passes of an M-point forward DFT.  On entry v holds pass-1 input:
 *     v[q*R1 + r] = z[j + r*M/R1],  j = t + T*q
 * On exit it holds the spectrum of the last pass's items:
 *     Z[j + r*NSL] = v[q*RL + bitrev_RL(r)],  j = t + T*q,  NSL = M/RL  (RL = last radix). */
template <class C, class Hook = NoHook>
VVB_DEV void team_fft(float2 (&v)[C::E], float2* xb, const float2* tw2, const float2* tw3, int t, int team, Hook after_last_read = Hook())
{
    if constexpr (C::NP == 2) {
        stockham_pass<C, C::R1, 1, true, false>(v, xb, nullptr, t, team);
        stockham_pass<C, C::R2, C::R1, false, true>(v, xb, tw2, t, team, after_last_read);
    } else {
        stockham_pass<C, C::R1, 1, true, false>(v, xb, nullptr, t, team);
        stockham_pass<C, C::R2, C::R1, false, false>(v, xb, tw2, t, team);
        stockham_pass<C, C::R3, C::R1 * C::R2, false, true>(v, xb, tw3, t, team, after_last_read);
    }
}

/* 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, completion via groups */
VVB_DEV void cp_async8(float2* smem_dst, const float2* gsrc)
{
#ifdef VVB_EMU
    *smem_dst = *gsrc;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
VVB_DEV void cp_async_commit()
{
#ifndef VVB_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
VVB_DEV void cp_async_wait_group1()      /* all but the most recently committed group have landed */
{
#ifndef VVB_EMU
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
VVB_DEV void cp_async_wait_all()
{
#ifndef VVB_EMU
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

/* ---- TMA 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) completing on an mbarrier.
 * src, dst and bytes must be multiples of 16.  Issued by ONE thread. */
/* (emulator: the 8-byte barrier word holds {phase bit, expected arrivals, pending arrivals, transaction bytes}; a TMA
 * copy is a memcpy that works off the bytes announced by expect_tx, a phase completes when no arrival and no byte is
 * pending, and a wait yields to the other fibres until the phase flips) */
#ifdef VVB_EMU
struct EmuMbar { unsigned char phase, expected, pending, pad; int tx; };
static_assert(sizeof(EmuMbar) == 8, "mbarrier word");
inline void emu_mbar_check(EmuMbar* b)
{
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->expected; }
}
inline void emu_mbar_arrive(unsigned long long* bar)
{
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
    --b->pending;
    emu_mbar_check(b);
}
#endif
VVB_DEV void mbar_init(unsigned long long* bar, unsigned count)
{
#ifdef VVB_EMU
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
    b->phase = 0; b->expected = (unsigned char)count; b->pending = (unsigned char)count; b->pad = 0; b->tx = 0;
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
VVB_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
#ifndef VVB_EMU
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
#else
    EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);  /* one arrival + a transaction count that the copies work off */
    b->tx += (int)bytes;
    --b->pending;
    emu_mbar_check(b);
#endif
}
/* plain arrival (release semantics at CTA scope): signals "my earlier shared-memory writes / reads are done" */
VVB_DEV void mbar_arrive(unsigned long long* bar)
{
#ifdef VVB_EMU
    emu_mbar_arrive(bar);
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
#endif
}
VVB_DEV void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
#ifdef VVB_EMU
    memcpy(smem_dst, gsrc, bytes);
    {
        EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
        b->tx -= (int)bytes;
        emu_mbar_check(b);
    }
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
#endif
}
/* wait until the phase with the given parity has completed (a fresh barrier: parity 1 passes at once) */
VVB_DEV void mbar_wait(unsigned long long* bar, unsigned parity)
{
#ifdef VVB_EMU
    const EmuMbar* b = reinterpret_cast<const EmuMbar*>(bar);
    long spins = 0;
    while ((b->phase & 1u) == (parity & 1u)) {
        if (++spins > 50000000L) { fprintf(stderr, "vvb_emu: mbarrier wait never completes (block %u)\n", blockIdx.x); abort(); }
        emu_yield();
    }
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
#endif
}
/* order earlier generic-proxy accesses of shared memory before later async-proxy (TMA) writes */
VVB_DEV void fence_proxy_async()
{
#ifndef VVB_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

/* ---- inter-pass twiddles W_M^{r*t} computed instead of loaded (one-warp 32 x 32 transforms).
 * The 31 twiddles of a thread are powers of W_M^t; five of them (r = 1, 2, 4, 8, 16) are kept in
 * registers and every other power is the product of the bases of its set bits (at most 4 complex
 * multiplies deep, so the rounding error stays ~2e-7).  With packed FP32 math this costs fewer issue
 * slots than it saves: 31 LDS.64 (62 shared-memory wavefronts) per transform disappear, and the
 * kernels are bound by exactly those wavefronts. */
struct TwBase { float2 w[5]; };

template <int R, int BIT = 0, bool STARTED = false> VVB_DEV float2 tw_power(const TwBase& b, float2 acc = make_float2(1.f, 0.f))
{
    if constexpr ((R >> BIT) == 0) {
        return acc;
    } else if constexpr (((R >> BIT) & 1) == 0) {
        return tw_power<R, BIT + 1, STARTED>(b, acc);
    } else if constexpr (!STARTED) {
        return tw_power<R, BIT + 1, true>(b, b.w[BIT]);
    } else {
        return tw_power<R, BIT + 1, true>(b, cmul(acc, b.w[BIT]));
    }
}
template <int... Rs> VVB_DEV void apply_tw_powers(float2* v, const TwBase& b, iseq<Rs...>)
{
    ((v[Rs + 1] = cmul(v[Rs + 1], tw_power<Rs + 1>(b))), ...);
}

/* team_fft for Cfg<1024,32,32,32> with register twiddles; tw2 is only read once, for the bases */
template <class C> VVB_DEV TwBase load_tw_base(const float2* tw2_global, int t)
{
    static_assert(C::T == 32 && C::R1 == 32 && C::R2 == 32 && C::NP == 2, "32 x 32 one-warp transform");
    TwBase b;
#pragma unroll
    for (int j = 0; j < 5; ++j) b.w[j] = __ldg(tw2_global + ((1 << j) - 1) * 32 + t);   /* r = 2^j, jm = t */
    return b;
}
template <class C> VVB_DEV void team_fft_regtw(float2 (&v)[C::E], float2* xb, const TwBase& b, int t, int team)
{
    stockham_pass<C, 32, 1, true, false>(v, xb, nullptr, t, team);
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = xb[C::pad(t + r * 32)];
    team_sync<C::T>(team);
    apply_tw_powers(v, b, typename make_iseq<31>::type{});
    fft_reg<32, 0>(v);
}

/* Variant that keeps no register state between transforms: the five bases are re-read from the shared-memory
 * table (5 LDS.64 instead of 31) right where they are used, so they are live only during the twiddle phase. */
template <class C> VVB_DEV void team_fft_basetw(float2 (&v)[C::E], float2* xb, const float2* s_tw2, int t, int team)
{
    static_assert(C::T == 32 && C::R1 == 32 && C::R2 == 32 && C::NP == 2, "32 x 32 one-warp transform");
    stockham_pass<C, 32, 1, true, false>(v, xb, nullptr, t, team);
#pragma unroll
    for (int r = 0; r < 32; ++r) v[r] = xb[C::pad(t + r * 32)];
    team_sync<C::T>(team);
    TwBase b;
#pragma unroll
    for (int j = 0; j < 5; ++j) b.w[j] = s_tw2[((1 << j) - 1) * 32 + t];
    apply_tw_powers(v, b, typename make_iseq<31>::type{});
    fft_reg<32, 0>(v);
}

/* The same for any two-pass configuration (M = R1 * R2, T = M / E threads, E / R2 sub-transforms per thread in
 * pass 2).  Sub-transform q of thread t is column j = t + T q of the R1 x R2 decomposition and needs
 * W_M^{r j} = W_M^{r t} * W_M^{r T q}: powers of the per-thread base times a compile-time rotation. */
template <class C> VVB_DEV TwBase load_tw_base2(const float2* tw2_global, int t)
{
    static_assert(C::NP == 2 && C::R2 <= 32, "two-pass configuration");
    TwBase b;
#pragma unroll
    for (int j = 0; j < 5; ++j)
        b.w[j] = ((1 << j) < C::R2) ? __ldg(tw2_global + ((1 << j) - 1) * C::R1 + t) : make_float2(1.f, 0.f);   /* r = 2^j, jm = t */
    return b;
}
template <class C, int Q, int R> VVB_DEV float2 tw_power_q(const TwBase& b)
{
    if constexpr (Q == 0) {
        return tw_power<R>(b);
    } else {
        constexpr int K = (R * C::T * Q) % C::M;
        return cmul(tw_power<R>(b), make_float2(TwC<C::M, K>::c, -TwC<C::M, K>::s));
    }
}
template <class C, int Q, int... Rs> VVB_DEV void apply_tw_powers_q(float2* v, const TwBase& b, iseq<Rs...>)
{
    ((v[Rs + 1] = cmul(v[Rs + 1], tw_power_q<C, Q, Rs + 1>(b))), ...);
}
template <class C, int... Qs> VVB_DEV void apply_tw_all_q(float2* v, const TwBase& b, iseq<Qs...>)
{
    (apply_tw_powers_q<C, Qs>(v + Qs * C::R2, b, typename make_iseq<C::R2 - 1>::type{}), ...);
}
template <class C> VVB_DEV void team_fft_regtw2(float2 (&v)[C::E], float2* xb, const TwBase& b, int t, int team)
{
    constexpr int R = C::R2, NQ = C::E / R, STRIDE = C::M / R;
    stockham_pass<C, C::R1, 1, true, false>(v, xb, nullptr, t, team);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int r = 0; r < R; ++r) v[q * R + r] = xb[C::pad(t + C::T * q + r * STRIDE)];
    team_sync<C::T>(team);
    apply_tw_all_q<C>(v, b, typename make_iseq<NQ>::type{});
#pragma unroll
    for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
}

/* two-pass transform with the bases re-read from the shared-memory table at the point of use (no register state
 * between transforms); three-pass configurations fall through to the table version.  Measured: a gain only where
 * the kernel is bound by shared-memory wavefronts (marching ISTFT); the chirp-z and plan C2C kernels are not and
 * got slower with it (nfft=400 STFT 5.4 -> 7.7 ms), so they keep the table twiddles. */
template <class C> VVB_DEV void team_fft_auto(float2 (&v)[C::E], float2* xb, const float2* s_tw2, const float2* s_tw3, int t, int team)
{
    if constexpr (C::NP == 2 && C::R2 <= 32) {
        constexpr int R = C::R2, NQ = C::E / R, STRIDE = C::M / R;
        stockham_pass<C, C::R1, 1, true, false>(v, xb, nullptr, t, team);
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) v[q * R + r] = xb[C::pad(t + C::T * q + r * STRIDE)];
        team_sync<C::T>(team);
        TwBase b;
#pragma unroll
        for (int j = 0; j < 5; ++j) b.w[j] = ((1 << j) < R) ? s_tw2[((1 << j) - 1) * C::R1 + t] : make_float2(1.f, 0.f);
        apply_tw_all_q<C>(v, b, typename make_iseq<NQ>::type{});
#pragma unroll
        for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
    } else {
        team_fft<C>(v, xb, s_tw2, s_tw3, t, team);
    }
}

/* ---- three-pass configurations (M = R1 R2 R3, T a multiple of R1, T E / R3 = R1 R2): inter-pass twiddles computed.
 * Pass 2 (radix R2, Ns = R1): the twiddle index (t + T q) mod R1 = t mod R1 does not depend on the sub-transform q, so a
 * thread needs R2 - 1 twiddles in all: powers of ONE base, W_{R1 R2}^{t mod R1}.
 * Pass 3 (radix R3, Ns = R1 R2): the index is t + T q itself, W_M^{r (t + T q)} = W_M^{r t} x W_M^{r T q}: powers of one
 * per-thread base times compile-time rotations (as apply_tw_all_q does for the two-pass configurations).
 * The table version loads (R2 - 1) + (R3 - 1) E / R3 twiddles per thread and transform (35 LDS.64 at fft_size 4096, where
 * the kernels are bound by shared-memory wavefronts: two exchanges of the whole transform per frame); this one loads the
 * power-of-two bases (3 + 3 at radix 8) and spends packed FP32 instructions, of which there are plenty to spare. */
template <class C, int RADIX, int Q, int R> VVB_DEV float2 tw_power_rot(const TwBase& b)
{
    if constexpr (Q == 0) {
        return tw_power<R>(b);
    } else {
        constexpr int K = (R * C::T * Q) % C::M;
        return cmul(tw_power<R>(b), make_float2(TwC<C::M, K>::c, -TwC<C::M, K>::s));
    }
}
template <class C, int RADIX, int Q, int... Rs> VVB_DEV void apply_tw_rot_q(float2* v, const TwBase& b, iseq<Rs...>)
{
    ((v[Rs + 1] = cmul(v[Rs + 1], tw_power_rot<C, RADIX, Q, Rs + 1>(b))), ...);
}
template <class C, int RADIX, int... Qs> VVB_DEV void apply_tw_rot_all(float2* v, const TwBase& b, iseq<Qs...>)
{
    (apply_tw_rot_q<C, RADIX, Qs>(v + Qs * RADIX, b, typename make_iseq<RADIX - 1>::type{}), ...);
}
/* the same power for every sub-transform q (pass 2) */
template <int NQ, int RADIX, int R> VVB_DEV void apply_tw_same_one(float2* v, const TwBase& b)
{
    const float2 w = tw_power<R>(b);
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q * RADIX + R] = cmul(v[q * RADIX + R], w);
}
template <int NQ, int RADIX, int... Rs> VVB_DEV void apply_tw_same(float2* v, const TwBase& b, iseq<Rs...>)
{
    (apply_tw_same_one<NQ, RADIX, Rs + 1>(v, b), ...);
}
#ifndef VVB_TW3_HALF_EXCHANGE
#define VVB_TW3_HALF_EXCHANGE 1       /* 32.8.8 configuration: exchange only the partner warp's half between passes 2 and 3 */
#endif
template <class C, class Hook = NoHook>
VVB_DEV void team_fft_tw3(float2 (&v)[C::E], float2* xb, const float2* s_tw2, const float2* s_tw3, int t, int team, Hook after_last_read = Hook())
{
    static_assert(C::NP == 3 && C::T % C::R1 == 0 && C::T * (C::E / C::R3) == C::R1 * C::R2 && C::R2 <= 32 && C::R3 <= 32, "see above");
    stockham_pass<C, C::R1, 1, true, false>(v, xb, nullptr, t, team);
    constexpr bool HALFX = VVB_TW3_HALF_EXCHANGE && C::T == 64 && C::R1 == 32 && C::R2 == 8 && C::R3 == 8 && C::E == 32;
    {   /* pass 2 */
        constexpr int R = C::R2, NS = C::R1, NQ = C::E / R, STRIDE = C::M / R;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) v[q * R + r] = xb[C::pad(t + C::T * q + r * STRIDE)];
        team_sync<C::T>(team);
        TwBase b;
#pragma unroll
        for (int j = 0; j < 5; ++j) b.w[j] = ((1 << j) < R) ? s_tw2[((1 << j) - 1) * NS + (t % NS)] : make_float2(1.f, 0.f);
        apply_tw_same<NQ, R>(v, b, typename make_iseq<R - 1>::type{});
#pragma unroll
        for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
        if constexpr (!HALFX) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int j = t + C::T * q;
                const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
                for (int r = 0; r < R; ++r) xb[C::pad(j0 + r * NS)] = v[q * R + r];
            }
            team_sync<C::T>(team);
        }
    }
    if constexpr (HALFX) {
        /* pass 2 -> pass 3 of the 32.8.8 configuration (two warps per transform) moves data only between the SAME lane of
         * the two warps: the pass-3 input r' of sub-transform q' of thread (warp a', lane l) is output slot a' + 2 q' of
         * sub-transform r' / 2 of thread (warp r' mod 2, lane l).  Half of a thread's 32 inputs are therefore already in
         * its own registers; only the other half goes through shared memory (16 STS.64 + 16 LDS.64 instead of 32 + 32:
         * 128 of the ~960 shared-memory wavefronts per frame).  The register permutation depends on the warp, hence the
         * two (warp-uniform) code paths. */
        constexpr int R = 8, NQ = 4;
        const int a = t >> 5;
        float2 u[C::E];
        auto exchange = [&](auto AC) {
            constexpr int A = decltype(AC)::value;                     /* this thread's warp within the team */
#pragma unroll
            for (int q = 0; q < NQ; ++q)                               /* slots the partner needs: r = (1 - A) + 2 q' */
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq) {
                    const int r = (1 - A) + 2 * qq;
                    xb[C::pad((A + 2 * q) * 256 + (t & 31) + 32 * r)] = v[q * R + r];
                }
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq)                            /* own half: r' = A + 2 k  <-  v[k * 8 + A + 2 q'] */
#pragma unroll
                for (int k = 0; k < NQ; ++k) u[qq * R + A + 2 * k] = v[k * R + A + 2 * qq];
            team_sync<C::T>(team);
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq)                            /* partner's half: r' = (1 - A) + 2 k at its natural position */
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    const int rp = (1 - A) + 2 * k;
                    u[qq * R + rp] = xb[C::pad(t + C::T * qq + rp * 256)];
                }
        };
        if (a == 0) exchange(std::integral_constant<int, 0>{}); else exchange(std::integral_constant<int, 1>{});
        team_sync<C::T>(team);
        after_last_read();
#pragma unroll
        for (int i = 0; i < C::E; ++i) v[i] = u[i];
        constexpr int NS = C::R1 * C::R2;
        TwBase b;
#pragma unroll
        for (int j = 0; j < 5; ++j) b.w[j] = ((1 << j) < R) ? s_tw3[((1 << j) - 1) * NS + t] : make_float2(1.f, 0.f);
        apply_tw_rot_all<C, R>(v, b, typename make_iseq<NQ>::type{});
#pragma unroll
        for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
    } else {   /* pass 3 */
        constexpr int R = C::R3, NS = C::R1 * C::R2, NQ = C::E / R, STRIDE = C::M / R;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) v[q * R + r] = xb[C::pad(t + C::T * q + r * STRIDE)];
        team_sync<C::T>(team);
        after_last_read();
        TwBase b;
#pragma unroll
        for (int j = 0; j < 5; ++j) b.w[j] = ((1 << j) < R) ? s_tw3[((1 << j) - 1) * NS + t] : make_float2(1.f, 0.f);
        apply_tw_rot_all<C, R>(v, b, typename make_iseq<NQ>::type{});
#pragma unroll
        for (int q = 0; q < NQ; ++q) fft_reg<R, 0>(&v[q * R]);
    }
}

#ifndef VVB_TW3_COMPUTE
#define VVB_TW3_COMPUTE 1             /* marching kernels, three-pass configurations: 1 = computed inter-pass twiddles, 0 = tables */
#endif
/* marching kernels: the three-pass transform with computed twiddles where the configuration allows it */
template <class C, bool COMPUTE = true> VVB_DEV void team_fft_march(float2 (&v)[C::E], float2* xb, const float2* s_tw2, const float2* s_tw3, int t, int team)
{
    if constexpr (VVB_TW3_COMPUTE && COMPUTE && C::NP == 3 && C::T % C::R1 == 0 && C::T * (C::E / C::R3) == C::R1 * C::R2 && C::R2 <= 32 && C::R3 <= 32)
        team_fft_tw3<C>(v, xb, s_tw2, s_tw3, t, team);
    else
        team_fft<C>(v, xb, s_tw2, s_tw3, t, team);
}

template <class C> struct LastPass {
    static constexpr int R = (C::NP == 2) ? C::R2 : C::R3;
    static constexpr int NS = C::M / R;
    static constexpr int NQ = C::E / R;
};

/* Write the register-resident spectrum to the team buffer in natural order (padded). */
template <class C> VVB_DEV void team_store_natural(const float2 (&v)[C::E], float2* xb, int t)
{
    using L = LastPass<C>;
#pragma unroll
    for (int q = 0; q < L::NQ; ++q) {
        const int j = t + C::T * q;
#pragma unroll
        for (int r = 0; r < L::R; ++r) xb[C::pad(j + r * L::NS)] = v[q * L::R + ct_bitrev(r, L::R)];
    }
}

}  // namespace vvb
