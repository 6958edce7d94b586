/*
 * cuda_emu.h -- TEST-ONLY execution emulator for the .cu kernel sources.
 *
 * Lets the unmodified kernel code in vv_dsp_b200/csrc/cuda/ be compiled with g++
 * (-DVVB_EMU) and run on the CPU, so index maps, barriers and edge cases can be
 * debugged in the build container, which has no GPU.  It is NOT a fallback: the
 * product library (libvvdsp_b200.so) is built by nvcc only and refuses to work
 * without a CUDA device; this header is only ever included when tests/emu builds
 * libvvdsp_b200_emu.so for the CPU test-suite.
 *
 * Model: CTAs run one after another; the threads of a CTA are ucontext fibers
 * scheduled round-robin.  A fiber runs until it reaches a barrier
 * (__syncthreads / __syncwarp / named bar.sync) and is resumed once the barrier's
 * generation has advanced.  A scheduling pass without progress = barrier
 * deadlock (divergent barrier) and aborts with a message.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>
#include <vector>
#include <algorithm>
using std::min;
using std::max;

struct uint3_emu { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict

namespace vvb_emu {

struct Barrier { unsigned expected = 0, count = 0, gen = 0; };

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    int state = 0;             /* 0 runnable, 1 waiting, 2 done */
    Barrier *wait_on = nullptr;
    unsigned wait_gen = 0;
    unsigned tid = 0;
};

struct Cta {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = -1;
    Barrier cta_bar;
    std::vector<Barrier> warp_bars;
    Barrier named[16];
    std::vector<float> shfl;   /* per-thread exchange slot */
    std::function<void()> body;
};

extern Cta *g_cta;
extern uint3_emu g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern char *g_dyn_smem;

inline void barrier_wait(Barrier &b, unsigned expected)
{
    Cta &c = *g_cta;
    Fiber &f = c.fibers[c.current];
    if (b.expected == 0) b.expected = expected;
    if (b.expected != expected) { fprintf(stderr, "vvb_emu: barrier used with different thread counts\n"); abort(); }
    unsigned my_gen = b.gen;
    if (++b.count == b.expected) { b.count = 0; b.expected = 0; b.gen++; return; }
    f.state = 1; f.wait_on = &b; f.wait_gen = my_gen;
    swapcontext(&f.ctx, &c.sched);
}

static void fiber_entry()
{
    Cta &c = *g_cta;
    c.body();
    c.fibers[c.current].state = 2;
    swapcontext(&c.fibers[c.current].ctx, &c.sched);
}

inline void run_cta(unsigned nthreads, const std::function<void()> &body)
{
    Cta c;
    g_cta = &c;
    c.body = body;
    c.fibers.resize(nthreads);
    c.warp_bars.resize((nthreads + 31) / 32);
    c.shfl.assign(nthreads, 0.f);
    const size_t STACK = 256 * 1024;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber &f = c.fibers[t];
        f.tid = t;
        f.stack = (char *)malloc(STACK);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK;
        f.ctx.uc_link = &c.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned done = 0;
    while (done < nthreads) {
        bool progress = false;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber &f = c.fibers[t];
            if (f.state == 2) continue;
            if (f.state == 1) {
                if (f.wait_on->gen == f.wait_gen) continue;
                f.state = 0;
            }
            c.current = (int)t;
            g_threadIdx.x = t % g_blockDim.x;
            g_threadIdx.y = (t / g_blockDim.x) % g_blockDim.y;
            g_threadIdx.z = t / (g_blockDim.x * g_blockDim.y);
            swapcontext(&c.sched, &f.ctx);
            progress = true;
            if (f.state == 2) ++done;
        }
        if (!progress) { fprintf(stderr, "vvb_emu: barrier deadlock (divergent barrier?) in block %u\n", g_blockIdx.x); abort(); }
    }
    for (auto &f : c.fibers) free(f.stack);
    g_cta = nullptr;
}

template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F &&body)
{
    g_gridDim = grid; g_blockDim = block;
    std::vector<char> sm(smem + 64);
    g_dyn_smem = sm.data() + ((64 - ((uintptr_t)sm.data() & 63)) & 63);
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = bz;
                run_cta(nthreads, body);
            }
}

}  // namespace vvb_emu

#define threadIdx (vvb_emu::g_threadIdx)
#define blockIdx (vvb_emu::g_blockIdx)
#define blockDim (vvb_emu::g_blockDim)
#define gridDim (vvb_emu::g_gridDim)

static inline unsigned emu_linear_tid() { return (unsigned)vvb_emu::g_cta->current; }
static inline void __syncthreads()
{
    vvb_emu::barrier_wait(vvb_emu::g_cta->cta_bar, (unsigned)vvb_emu::g_cta->fibers.size());
}
static inline void __syncwarp(unsigned = 0xffffffffu)
{
    unsigned t = emu_linear_tid(), n = (unsigned)vvb_emu::g_cta->fibers.size();
    unsigned w = t / 32, in_warp = (w * 32 + 32 <= n) ? 32 : n - w * 32;
    vvb_emu::barrier_wait(vvb_emu::g_cta->warp_bars[w], in_warp);
}
static inline void emu_named_barrier(int id, int nthreads)
{
    vvb_emu::barrier_wait(vvb_emu::g_cta->named[id], (unsigned)nthreads);
}
/* give the other fibres of the CTA a turn (spin-waits on shared-memory flags / mbarriers) */
static inline void emu_yield()
{
    vvb_emu::Cta &c = *vvb_emu::g_cta;
    vvb_emu::Fiber &f = c.fibers[c.current];
    swapcontext(&f.ctx, &c.sched);
}
static inline float __shfl_sync(unsigned, float v, int src, int width = 32)
{
    unsigned t = emu_linear_tid();
    vvb_emu::g_cta->shfl[t] = v;
    __syncwarp();
    unsigned base = (t / width) * width;
    float r = vvb_emu::g_cta->shfl[base + ((unsigned)src % width)];
    __syncwarp();
    return r;
}
static inline float __shfl_xor_sync(unsigned m, float v, int mask, int width = 32)
{
    return __shfl_sync(m, v, (int)((emu_linear_tid() % width) ^ (unsigned)mask), width);
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
/* packed FP32 pairs (sm_100 FFMA2 / FADD2 / FMUL2) */
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline void sincospi(double x, double *s, double *c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
