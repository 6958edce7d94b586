/*
 * vvb_istft_ws.cuh -- warp-specialised marching ISTFT for one-warp 32 x 32 transforms (fft_size 2048).
 *
 * Same arithmetic, same frame order and same register-resident overlap-add as istft_march_kernel
 * (reference: C2C backward x 1/n, src/spectral/fft_kiss.c:27-74; synthesis window + overlap-add + norm,
 * src/spectral/stft.c:103-108; caller-side normalise, tools/dump_stft_roundtrip.c:50-54), but every frame is
 * worked on by TWO warps in a pipeline:
 *
 *   producer warp p   waits for the frame's half spectrum (one TMA bulk copy, issued one frame ahead),
 *                     merges it into the N/2-point complex spectrum, runs the first radix-32 pass in
 *                     registers and publishes the 1024 intermediate values in one of the pair's two
 *                     exchange buffers;
 *   consumer warp p   takes the exchange buffer (this IS the Stockham exchange between the two passes,
 *                     so the hand-over costs no additional shared-memory traffic), applies the inter-pass
 *                     twiddles, runs the second radix-32 pass, multiplies by the synthesis window and
 *                     adds into its register accumulator; finished hop-blocks leave with coalesced
 *                     64-bit stores.
 *
 * Why: ncu of istft_march_kernel (profiles/r01_ncu_full_v9_final_kernels.csv) shows no saturated unit --
 * FMA pipe 62 %, LSU wavefronts 68 %, issue 52 %, 1.33 `wait` stall cycles per issued instruction -- with
 * 2 warps per scheduler: it is bound by the dependent chain of one warp, and at 238 registers no third warp
 * fits.  Splitting the chain gives 16 warps per SM (4 per scheduler, 2 producers + 2 consumers each) inside
 * the same 64 K registers: producers give registers back (setmaxnreg.dec, no accumulator to hold), consumers
 * take them (setmaxnreg.inc, accumulator 64 + transform 64 registers).
 *
 * Buffers: a ring of THREE buffers per pair.  Frame i lives in buffer i mod 3 for its whole life: its half spectrum
 * lands there (TMA, issued TWO frames ahead -- ncu of the first version, one frame ahead, showed the producers
 * waiting for the copy 11 % of their time and the consumers for the producers 25 % of theirs), the producer reads it
 * in the merge step and, once every lane has its values, writes the pass-1 results over it; the consumer reads them
 * and hands the buffer back, whereupon the producer aims the copy of frame i + 3 at it.
 * Protocol per buffer: mbarriers stage (TMA transaction count), full (producer -> consumer) and empty (consumer ->
 * producer); one elected lane arrives after a __syncwarp, every lane waits.
 */
#pragma once
#include "vvb_stft_kernels.cuh"

namespace vvb {

#ifndef VVB_WS_PAIRMERGE
#define VVB_WS_PAIRMERGE 1            /* producer merge: 1 = bins k and M-k formed together, partner values by SHFL; 0 = every bin alone */
#endif
#ifndef VVB_WS_DIRECT
#define VVB_WS_DIRECT 0               /* producer reads the half spectrum straight from global memory into the merge registers
                                         (no TMA stage: the stage cost 64 shared-memory write + 64 read wavefronts per frame) */
#endif
/* register split of the two roles (8 producer + 8 consumer warps x 32 lanes x (P + C) = 64 K).  Same-box A/B, interleaved
 * launches, headline shape: 104 / 152 -> 1.995 ms, 96 / 160 -> 1.933, 112 / 144 -> 1.926, 120 / 136 -> 2.245 (on a
 * power-capped box at 1.68 GHz): the multiples of 16 win by 3 %; 96 / 160 also spills least at the other hops */
#ifndef VVB_WS_PRODUCER_REGS
#define VVB_WS_PRODUCER_REGS 96
#endif
#ifndef VVB_WS_CONSUMER_REGS
#define VVB_WS_CONSUMER_REGS 160
#endif

template <int REGS> VVB_DEV void setmaxnreg_dec()
{
#ifndef VVB_EMU
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}
template <int REGS> VVB_DEV void setmaxnreg_inc()
{
#ifndef VVB_EMU
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}

template <class C, int NPAIR> struct WsLayout {
    static constexpr int M = C::M, N = 2 * C::M, STG = C::M + 2, NBUF = 3, NBAR = 3 * NBUF;
    static constexpr size_t WSYN = 0;                                         /* float[N] */
    static constexpr size_t TWB = WSYN + sizeof(float) * N;                   /* float2[5][32]: W_M^(t 2^j) */
    static constexpr size_t BUF = TWB + sizeof(float2) * 5 * 32;              /* float2[NPAIR][NBUF][XBUF]: stage, then exchange */
    static constexpr size_t BAR = BUF + sizeof(float2) * NPAIR * NBUF * C::XBUF;   /* u64[NPAIR][NBAR]: stage | full | empty */
    static constexpr size_t TOTAL = BAR + 8 * NPAIR * NBAR;
    static_assert(BUF % 16 == 0 && (sizeof(float2) * C::XBUF) % 16 == 0 && C::XBUF >= STG + 1, "TMA destinations must be 16-byte aligned");
};

template <class C, int S, int NPAIR>
__global__ void __launch_bounds__(64 * NPAIR, 1) istft_ws_kernel(const InvArgs a)
{
    static_assert(C::T == 32 && C::E == 32 && C::R1 == 32 && C::R2 == 32 && C::NP == 2, "one-warp 32 x 32 transform");
    static_assert(NPAIR % 4 == 0, "producers and consumers are whole warpgroups");
    using TB = Tables<C>;
    using LY = WsLayout<C, NPAIR>;
    constexpr int M = C::M, N = 2 * M, E = C::E, T = 32;
    constexpr int PERIOD = E / S, HOP = 2 * T * S, EDGE = N - HOP, STG = LY::STG;
#ifdef VVB_EMU
    char* smem = reinterpret_cast<char*>(vvb_emu::g_dyn_smem);
#else
    extern __shared__ __align__(16) float smem_f[];
    char* smem = reinterpret_cast<char*>(smem_f);
#endif
    float* s_wsyn = reinterpret_cast<float*>(smem + LY::WSYN);
    float2* s_twb = reinterpret_cast<float2*>(smem + LY::TWB);
    float2* s_buf = reinterpret_cast<float2*>(smem + LY::BUF);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem + LY::BAR);

    const bool normalise = a.inv_norm != nullptr;
    copy_table(s_wsyn, a.tables + (normalise ? TB::WSYN_NORM : TB::WSYN), N);
    if (threadIdx.x < 5 * 32) {                                        /* bases r = 2^j of lane t: tw2[(r-1)*32 + t] */
        const int j = threadIdx.x / 32, t = threadIdx.x % 32;
        s_twb[j * 32 + t] = __ldg(reinterpret_cast<const float2*>(a.tables + TB::TW2) + ((1 << j) - 1) * 32 + t);
    }
    if (threadIdx.x < NPAIR) {
        unsigned long long* bar = s_bar + threadIdx.x * LY::NBAR;
#pragma unroll
        for (int i = 0; i < LY::NBAR; ++i) mbar_init(&bar[i], 1);
    }
    __syncthreads();

    const int warp = threadIdx.x / 32, t = threadIdx.x % 32;
    const bool producer = warp < NPAIR;
    const int pair = producer ? warp : warp - NPAIR;
    float2* buf0 = s_buf + (size_t)pair * LY::NBUF * C::XBUF;
    unsigned long long* bar_stage = s_bar + pair * LY::NBAR;
    unsigned long long* bar_full = bar_stage + LY::NBUF;
    unsigned long long* bar_empty = bar_stage + 2 * LY::NBUF;

    const int F = a.frames;
    /* the pair's range [g0, g1) of the flattened (signal, frame) list: computed inside each role, after its register
     * budget is set (computed up here the four 64-bit values were spilled across the role branch) */
    auto pair_range = [&](long long& g0, long long& g1) {
        const long long total = (long long)a.num_items * F;           /* num_items carries the batch */
        const long long nteams = (long long)gridDim.x * NPAIR;
        const long long quota = (total + nteams - 1) / nteams;
        g0 = ((long long)blockIdx.x * NPAIR + pair) * quota;
        g1 = min(total, g0 + quota);
    };
    unsigned it = 0;                                                   /* frames handed over so far (both roles count alike) */

    if (producer) {
        /* ================================================================ producer: stage -> merge -> pass 1 -> publish */
        setmaxnreg_dec<VVB_WS_PRODUCER_REGS>();
        long long g0, g1;
        pair_range(g0, g1);
        const float2 hw_t = __ldg(reinterpret_cast<const float2*>(a.tables + TB::POST) + t);
        const bool spec16 = (reinterpret_cast<uintptr_t>(a.spec) & 15) == 0;
        while (g0 < g1) {
            const int b = (int)(g0 / F);
            const int f_begin = (int)(g0 - (long long)b * F);
            const int f_end = (int)min((long long)F, (long long)f_begin + (g1 - g0));
            const int fr0 = f_begin - min(PERIOD - 1, f_begin);        /* halo frames re-synthesised */
            g0 += f_end - f_begin;
            /* aim the copy of `frame` (hand-over number i) at buffer i mod 3, once the consumer has given that buffer
             * back.  As in istft_march_kernel: ONE TMA bulk copy per frame, starting 8 bytes early on rows that are only
             * 8-byte aligned (the merge then reads at offset 1); per-lane 8-byte cp.async where that is impossible
             * (unaligned base, first / last row of the array).  Returns offset | (bulk << 1). */
            auto prefetch = [&](int frame, unsigned i) -> int {
                int code = 0;
                if (frame < f_end) {
                    const unsigned k = i % 3u;
                    float2* dst = buf0 + k * C::XBUF;
                    mbar_wait(&bar_empty[k], ((i / 3u) & 1u) ^ 1u);
                    const long long rowi = (long long)b * F + frame;
                    const float2* X = a.spec + rowi * a.spec_pitch;
                    const int mis = (int)((rowi * a.spec_pitch) & 1);
                    const bool last_row = (b == a.num_items - 1) && (frame == F - 1);
                    if (spec16 && !last_row && (rowi > 0 || mis == 0)) {
                        code = mis | 2;
                        if (t == 0) {
                            fence_proxy_async();
                            mbar_expect_tx(&bar_stage[k], STG * 8);
                            bulk_load(dst, X - mis, STG * 8, &bar_stage[k]);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < E; ++r) cp_async8(&dst[t + T * r], X + t + T * r);
                        if (t == 0) {
                            cp_async8(&dst[M], X + M);
                            mbar_arrive(&bar_stage[k]);                /* keeps the barrier's phase count = uses of the buffer */
                        }
                    }
                }
                cp_async_commit();                                     /* one group per call, empty or not */
                return code;
            };
            if constexpr (VVB_WS_DIRECT && VVB_WS_PAIRMERGE) {
                /* direct: the 33 loads of a frame are issued back to back and the warp sleeps on them; the other producer
                 * and the two consumers of the scheduler keep the pipes busy meanwhile, the ring of buffers absorbs the jitter */
#pragma unroll 1
                for (int frame = fr0; frame < f_end; ++frame) {
                    float2 v[E];
                    const unsigned k = it % 3u;
                    float2* buf = buf0 + k * C::XBUF;
                    const float2* X = a.spec + ((long long)b * F + frame) * a.spec_pitch;
                    float2 hw = hw_t;
#ifndef VVB_EMU
                    asm volatile("" : "+f"(hw.x), "+f"(hw.y));
#endif
                    march_merge_pairs<C, 1, true>(v, X, buf, t, hw);
                    fft_reg<32, 0>(v);                                 /* pass 1: column t, rows r -> v[r] */
                    mbar_wait(&bar_empty[k], ((it / 3u) & 1u) ^ 1u);   /* the consumer has read this buffer's previous frame */
#pragma unroll
                    for (int r = 0; r < 32; ++r) buf[C::pad(t * 32 + r)] = v[r];
                    __syncwarp();
                    if (t == 0) mbar_arrive(&bar_full[k]);
                    ++it;
                }
            } else {
            int code_cur = prefetch(fr0, it);
            int code_next = prefetch(fr0 + 1, it + 1);
#pragma unroll 1
            for (int frame = fr0; frame < f_end; ++frame) {
                float2 v[E];
                const unsigned k = it % 3u;
                float2* buf = buf0 + k * C::XBUF;
                if (code_cur & 2) mbar_wait(&bar_stage[k], (it / 3u) & 1u); else cp_async_wait_group1();
                __syncwarp();
                /* the 32 rotated merge twiddles are loop-invariant; hoisted they would cost 64 registers (spills at this
                 * register budget), so the base value is made opaque per frame and they are recomputed (2 FFMA2 each) */
                float2 hw = hw_t;
#ifndef VVB_EMU
                asm volatile("" : "+f"(hw.x), "+f"(hw.y));
#endif
                if constexpr (VVB_WS_PAIRMERGE) march_merge_pairs<C, 1>(v, buf + (code_cur & 1), buf, t, hw);
                else march_merge<C>(v, buf + (code_cur & 1), t, hw, typename make_iseq<E>::type{});
                __syncwarp();                                          /* every lane has its share of the staged X */
                code_cur = code_next;
                code_next = prefetch(frame + 2, it + 2);
                fft_reg<32, 0>(v);                                     /* pass 1: column t, rows r -> v[r] */
#pragma unroll
                for (int r = 0; r < 32; ++r) buf[C::pad(t * 32 + r)] = v[r];
                __syncwarp();
                if (t == 0) mbar_arrive(&bar_full[k]);
                ++it;
            }
            }
            cp_async_wait_all();                                       /* nothing in flight across pieces */
            __syncwarp();
        }
    } else {
        /* ================================================================ consumer: twiddle -> pass 2 -> window, overlap-add */
        setmaxnreg_inc<VVB_WS_CONSUMER_REGS>();
        long long g0, g1;
        pair_range(g0, g1);
        const float2* wsyn2 = reinterpret_cast<const float2*>(s_wsyn);
        while (g0 < g1) {
            const int b = (int)(g0 / F);
            const int f_begin = (int)(g0 - (long long)b * F);
            const int f_end = (int)min((long long)F, (long long)f_begin + (g1 - g0));
            const int emit_end = (f_end == F && a.tail_edge) ? f_end + PERIOD - 1 : f_end;   /* hop-blocks [f_begin, emit_end) are ours */
            const int emit_begin = max(f_begin, a.halo_frames);        /* a shard's halo frames: overlap only (see InvArgs) */
            const int fr0 = f_begin - min(PERIOD - 1, f_begin);
            float* yb = a.y + (long long)b * a.y_pitch;
            const bool y8 = (reinterpret_cast<uintptr_t>(yb) & 7) == 0;   /* 64-bit stores possible for this signal's row */
            g0 += f_end - f_begin;

            float2 acc[E];
#pragma unroll
            for (int i = 0; i < E; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int frame = fr0; frame < emit_end; ++frame) {
                if (frame < f_end) {                                   /* warp-uniform */
                    float2 v[E];
                    const unsigned k = it % 3u;
                    const float2* xb = buf0 + k * C::XBUF;
                    mbar_wait(&bar_full[k], (it / 3u) & 1u);
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = xb[C::pad(t + r * 32)];
                    __syncwarp();
                    if (t == 0) mbar_arrive(&bar_empty[k]);
                    ++it;
                    TwBase tb;
#pragma unroll
                    for (int j = 0; j < 5; ++j) tb.w[j] = s_twb[j * 32 + t];
                    apply_tw_powers(v, tb, typename make_iseq<31>::type{});
                    fft_reg<32, 0>(v);
                    /* v[r] is sample pair i = t + 32 r, stored (Im, Re) */
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        acc[r] = __ffma2_rn(make_float2(v[r].y, v[r].x), wsyn2[t + T * r], acc[r]);
                }
                if (frame >= emit_begin) {
                    const long long base = (long long)(frame - a.halo_frames) * HOP;
                    if (normalise && ((a.head_edge && frame < PERIOD - 1) || frame >= F)) {
                        /* edge block: undo the folded steady-state factor, apply this block's own 1/sum(w^2) */
                        const float* edge = (frame >= F) ? a.inv_norm + EDGE + HOP + (long long)(frame - F) * HOP
                                                         : a.inv_norm + (long long)frame * HOP;
                        const float* mid = a.tables + TB::MIDNORM;
#pragma unroll
                        for (int r = 0; r < S; ++r) {
                            const int c = 2 * (t + T * r);
                            acc[r].x *= __ldg(edge + c) * __ldg(mid + c);
                            acc[r].y *= __ldg(edge + c + 1) * __ldg(mid + c + 1);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < S; ++r) {
                        const long long tt = base + 2 * (t + T * r);
                        if (y8 && tt + 1 < a.n_out) *reinterpret_cast<float2*>(yb + tt) = acc[r];
                        else {                                             /* row not 8-byte aligned (odd pitch), or the last sample */
                            if (tt < a.n_out) yb[tt] = acc[r].x;
                            if (tt + 1 < a.n_out) yb[tt + 1] = acc[r].y;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < E - S; ++r) acc[r] = acc[r + S];
#pragma unroll
                for (int r = E - S; r < E; ++r) acc[r] = make_float2(0.f, 0.f);
            }
            if (f_end == F && a.tail_edge) {                           /* nothing covers [cov, n_out): zeros */
                const long long cov = (long long)(F - 1 - a.halo_frames) * HOP + N;
                for (long long tt = cov + t; tt < a.n_out; tt += T) yb[tt] = 0.f;
            }
        }
    }
}

}  // namespace vvb
