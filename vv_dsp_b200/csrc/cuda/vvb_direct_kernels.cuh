/*
 * vvb_direct_kernels.cuh -- direct-DFT kernels for sizes without a Stockham kernel
 * (non-powers of two, fft_size < 256 or > 8192), plus the stand-alone overlap-add
 * kernel they feed.  The reference serves the same sizes with its own O(n^2) path
 * (src/spectral/fft_kiss.c:76-92,115); here it runs on the GPU, one thread per output
 * value, double accumulators, twiddles from a host-computed table (exponent reduced
 * mod n, so there is no large-angle error).  Correctness path, not a throughput path.
 */
#pragma once
#include "vvb_stft_kernels.cuh"

namespace vvb {

struct DirFwdArgs {
    const float* x; long long x_pitch, n;
    int batch, frames, hop, nfft, pad_mode, out_kind;
    void* out; long long out_pitch;
    const float* win;        /* nfft */
    const float2* wtab;      /* nfft: (cos, -sin)(2 pi j / nfft) */
};

__global__ void stft_forward_direct_kernel(const DirFwdArgs a)
{
    const int bins = a.nfft / 2 + 1;
    const long long total = (long long)a.batch * a.frames * bins;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % bins);
        const long long bf = idx / bins;
        const int f = (int)(bf % a.frames);
        const long long b = bf / a.frames;
        const float* xs = a.x + b * a.x_pitch;
        long long start = (long long)f * a.hop;
        if (a.pad_mode == PAD_REFLECT) start -= a.nfft / 2;
        double sr = 0.0, si = 0.0;
        int ph = 0;
        for (int i = 0; i < a.nfft; ++i) {
            const float v = fetch_sample(xs, a.n, start + i, a.pad_mode) * a.win[i];
            const float2 w = a.wtab[ph];
            sr += (double)v * (double)w.x;
            si += (double)v * (double)w.y;
            ph += k; if (ph >= a.nfft) ph -= a.nfft;
        }
        float xr = (float)sr, xi = (float)si;
        if (k == 0 || 2 * k == a.nfft) xi = 0.0f;
        const long long o = bf * a.out_pitch + k;
        if (a.out_kind == OUT_COMPLEX) reinterpret_cast<float2*>(a.out)[o] = make_float2(xr, xi);
        else if (a.out_kind == OUT_POWER) reinterpret_cast<float*>(a.out)[o] = xr * xr + xi * xi;
        else reinterpret_cast<float*>(a.out)[o] = sqrtf(xr * xr + xi * xi);
    }
}

struct DirInvArgs {
    const float2* spec; long long spec_pitch;
    long long count;         /* frames in the flat list */
    int nfft;
    float* frames_out;       /* [count][nfft] */
    const float* win;        /* nfft, or nullptr for none (C2R) */
    const float2* wtab;
};

/* frames_out[f][i] = Re(IDFT_n(Hermitian extension of spec[f]))[i] / n * win[i] */
__global__ void stft_inverse_direct_kernel(const DirInvArgs a)
{
    const int n = a.nfft;
    const long long total = a.count * n;
    const float invn = 1.0f / (float)n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n);
        const long long f = idx / n;
        const float2* X = a.spec + f * a.spec_pitch;
        double acc = (double)X[0].x;
        if (n % 2 == 0 && n > 1) acc += ((i & 1) ? -1.0 : 1.0) * (double)X[n / 2].x;
        const int kmax = (n - 1) / 2;
        int ph = 0;
        for (int k = 1; k <= kmax; ++k) {
            ph += i; if (ph >= n) ph -= n;
            const float2 w = a.wtab[ph];               /* (cos, -sin) */
            acc += 2.0 * ((double)X[k].x * (double)w.x + (double)X[k].y * (double)w.y);
        }
        float v = (float)acc * invn;
        if (a.win) v *= a.win[i];
        a.frames_out[idx] = v;
    }
}

struct DirC2CArgs {
    const float2* in; float2* out;
    long long batch; int n, inverse;
    const float2* wtab;
};

__global__ void fft_c2c_direct_kernel(const DirC2CArgs a)
{
    const int n = a.n;
    const long long total = a.batch * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % n);
        const float2* x = a.in + (idx / n) * n;
        double sr = 0.0, si = 0.0;
        int ph = 0;
        for (int t = 0; t < n; ++t) {
            const float2 w = a.wtab[ph];
            const double wr = w.x, wi = a.inverse ? -(double)w.y : (double)w.y;
            sr += (double)x[t].x * wr - (double)x[t].y * wi;
            si += (double)x[t].x * wi + (double)x[t].y * wr;
            ph += k; if (ph >= n) ph -= n;
        }
        if (a.inverse) { sr /= n; si /= n; }
        a.out[idx] = make_float2((float)sr, (float)si);
    }
}

/* ------------------------------------------------------------------ four-step FFT plumbing (plan API, n = n1 n2 > 8192) */
/* out[b][c][r] = in[b][r][c] (x W_n^{+-r c} when twiddle_n != 0): the transposes of the four-step algorithm, the second
 * one fused with the inter-step twiddle (exponent reduced mod n in integers, angle in double: no large-angle error).
 * 32 x 32 tiles through shared memory so that both the loads and the stores are coalesced. */
struct FourStepArgs { const float2* in; float2* out; int rows, cols; long long batch; long long twiddle_n; int inverse; };

__global__ void __launch_bounds__(256) fourstep_transpose_kernel(const FourStepArgs a)
{
#ifdef VVB_EMU
    static float2 tile[32][33];                                        /* CTAs run one after another in the emulator */
#else
    __shared__ float2 tile[32][33];
#endif
    const int tiles_c = (a.cols + 31) / 32, tiles_r = (a.rows + 31) / 32;
    const long long per = (long long)tiles_c * tiles_r, total = per * a.batch;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            /* 32 x 8 threads */
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        const long long b = tl / per;
        const int tr = (int)((tl % per) / tiles_c), tc = (int)((tl % per) % tiles_c);
        const float2* src = a.in + b * (long long)a.rows * a.cols;
        float2* dst = a.out + b * (long long)a.rows * a.cols;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = tr * 32 + ty + 8 * i, c = tc * 32 + tx;
            if (r < a.rows && c < a.cols) tile[ty + 8 * i][tx] = src[(long long)r * a.cols + c];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tc * 32 + ty + 8 * i, r = tr * 32 + tx;
            if (r < a.rows && c < a.cols) {
                float2 v = tile[tx][ty + 8 * i];
                if (a.twiddle_n) {
                    const long long m = ((long long)r * c) % a.twiddle_n;
                    double sn, cs;
                    sincospi(2.0 * (double)m / (double)a.twiddle_n, &sn, &cs);
                    const float wr = (float)cs, wi = (float)(a.inverse ? sn : -sn);
                    v = make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
                }
                dst[(long long)c * a.rows + r] = v;
            }
        }
        __syncthreads();
    }
}

/* real <-> complex glue for R2C / C2R plans of four-step sizes */
__global__ void real_to_cpx_kernel(const float* in, float2* out, long long total)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_float2(in[i], 0.f);
}
__global__ void cpx_keep_half_kernel(const float2* in, float2* out, long long batch, int n)
{
    const int bins = n / 2 + 1;
    const long long total = batch * bins;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / bins; const int k = (int)(i % bins);
        out[i] = in[b * n + k];
    }
}
__global__ void hermitian_extend_kernel(const float2* half, float2* full, long long batch, int n)
{
    const int bins = n / 2 + 1;
    const long long total = batch * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n; const int k = (int)(i % n);
        float2 v = (k < bins) ? half[b * bins + k] : half[b * bins + (n - k)];
        if (k >= bins) v.y = -v.y;
        if (k == 0 || 2 * k == n) v.y = 0.f;                            /* Re(IDFT): DC / Nyquist imaginary parts drop out */
        full[i] = v;
    }
}
__global__ void cpx_real_part_kernel(const float2* in, float* out, long long total)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) out[i] = in[i].x;
}

struct OlaArgs {
    const float* frames_in;  /* [batch][frames][nfft] windowed synthesis frames */
    int batch, frames, hop, nfft;
    float* y; long long y_pitch, n_out;
    const float* inv_norm;   /* [head | mid | tail] as in InvArgs, or nullptr */
};

/* y[b][t] = (sum over frames covering t, ascending f) * inv_norm(t); one thread per sample */
__global__ void overlap_add_kernel(const OlaArgs a)
{
    const long long total = (long long)a.batch * a.n_out;
    const int edge = a.nfft - a.hop;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long t = idx % a.n_out, b = idx / a.n_out;
        long long f_lo = (t < a.nfft) ? 0 : (t - a.nfft) / a.hop + 1;
        long long f_hi = t / a.hop;
        if (f_hi > a.frames - 1) f_hi = a.frames - 1;
        float acc = 0.f;
        for (long long f = f_lo; f <= f_hi; ++f)
            acc += a.frames_in[(b * a.frames + f) * a.nfft + (t - f * a.hop)];
        float scale = 1.0f;
        if (a.inv_norm) {
            const long long tail0 = (long long)a.frames * a.hop;
            if (t >= tail0) scale = (t - tail0 < edge) ? a.inv_norm[edge + a.hop + (t - tail0)] : 0.f;
            else if (t < edge) scale = a.inv_norm[t];
            else scale = a.inv_norm[edge + (t % a.hop)];
        }
        a.y[b * a.y_pitch + t] = acc * scale;
    }
}

/* ------------------------------------------------------------------ Bluestein */
/* Sizes without a Stockham kernel, 32 <= n <= 2048, run as a chirp-z convolution on the power-of-two C2C kernel
 * (the reference itself serves them with an O(n^2) DFT, src/spectral/fft_kiss.c:76-92,115, and uses the same
 * identity on the CPU for its CZT, src/spectral/czt.c:126-160):
 *     X[k] = c[k] * sum_i (x[i] c[i]) conj(c)[k-i],   c[i] = exp(-j pi i^2 / n)   (i^2 taken mod 2n on the host)
 * i.e. pre-multiply by the chirp, zero-pad to M >= 2n-1, FFT_M, multiply by the precomputed FFT_M of the
 * wrapped conjugate chirp, inverse FFT_M, post-multiply by the chirp.  These kernels are the element-wise
 * stages around the two fft_c2c_kernel launches; `work` holds M complex values per transform. */
struct ChirpFwdArgs {          /* STFT analysis: frame gather + window + chirp */
    const float* x; long long x_pitch, n_sig;
    int frames, hop, nfft, pad_mode, M;
    long long g0, count;       /* flat (signal, frame) range [g0, g0 + count) */
    const float* win; const float2* chirp; float2* work;
    int out_kind; void* out; long long out_pitch;
};

__global__ void chirp_pre_stft_kernel(const ChirpFwdArgs a)
{
    const long long total = a.count * a.M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % a.M);
        float2 r = make_float2(0.f, 0.f);
        if (i < a.nfft) {
            const long long g = a.g0 + idx / a.M;
            const long long b = g / a.frames;
            const int f = (int)(g - b * a.frames);
            long long start = (long long)f * a.hop;
            if (a.pad_mode == PAD_REFLECT) start -= a.nfft / 2;
            const float v = fetch_sample(a.x + b * a.x_pitch, a.n_sig, start + i, a.pad_mode) * a.win[i];
            const float2 c = a.chirp[i];
            r = make_float2(v * c.x, v * c.y);
        }
        a.work[idx] = r;
    }
}

__global__ void chirp_mul_kernel(float2* work, const float2* bspec, long long count, int M)
{
    const long long total = count * M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
        work[idx] = cmul(work[idx], bspec[idx % M]);
}

__global__ void chirp_post_stft_kernel(const ChirpFwdArgs a)
{
    const int bins = a.nfft / 2 + 1;
    const long long total = a.count * bins;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % bins);
        const long long g = idx / bins;
        float2 z = cmul(a.work[g * a.M + k], a.chirp[k]);
        if (k == 0 || 2 * k == a.nfft) z.y = 0.0f;
        const long long o = (a.g0 + g) * a.out_pitch + k;
        if (a.out_kind == OUT_COMPLEX) reinterpret_cast<float2*>(a.out)[o] = z;
        else if (a.out_kind == OUT_POWER) reinterpret_cast<float*>(a.out)[o] = z.x * z.x + z.y * z.y;
        else reinterpret_cast<float*>(a.out)[o] = sqrtf(z.x * z.x + z.y * z.y);
    }
}

struct ChirpInvArgs {          /* STFT synthesis: Hermitian half spectrum -> windowed real frame */
    const float2* spec; long long spec_pitch;
    long long g0, count; int nfft, M;
    const float* win;          /* nullptr: no window (C2R) */
    const float2* chirp; float2* work;
    float* frames_out;         /* [..][nfft], indexed by the flat frame number */
};

/* Re(IDFT(X)) = Re(DFT(conj X)) / n: feed conj of the Hermitian extension through the forward transform */
__global__ void chirp_pre_spec_kernel(const ChirpInvArgs a)
{
    const long long total = a.count * a.M;
    const int n = a.nfft;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % a.M);
        float2 r = make_float2(0.f, 0.f);
        if (k < n) {
            const float2* X = a.spec + (a.g0 + idx / a.M) * a.spec_pitch;
            float2 v = (2 * k <= n) ? X[k] : make_float2(X[n - k].x, -X[n - k].y);
            if (k == 0 || 2 * k == n) v.y = 0.0f;
            r = cmul(make_float2(v.x, -v.y), a.chirp[k]);
        }
        a.work[idx] = r;
    }
}

__global__ void chirp_post_frames_kernel(const ChirpInvArgs a)
{
    const int n = a.nfft;
    const long long total = a.count * n;
    const float invn = 1.0f / (float)n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n);
        const long long g = idx / n;
        const float2 w = a.work[g * a.M + i], c = a.chirp[i];
        float v = (w.x * c.x - w.y * c.y) * invn;
        if (a.win) v *= a.win[i];
        a.frames_out[(a.g0 + g) * n + i] = v;
    }
}

struct ChirpC2CArgs { const float2* in; float2* out; long long count; int n, M, inverse; const float2* chirp; float2* work; };

__global__ void chirp_pre_c2c_kernel(const ChirpC2CArgs a)
{
    const long long total = a.count * a.M;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % a.M);
        float2 r = make_float2(0.f, 0.f);
        if (i < a.n) {
            float2 v = a.in[(idx / a.M) * a.n + i];
            if (a.inverse) v.y = -v.y;
            r = cmul(v, a.chirp[i]);
        }
        a.work[idx] = r;
    }
}

__global__ void chirp_post_c2c_kernel(const ChirpC2CArgs a)
{
    const long long total = a.count * a.n;
    const float invn = 1.0f / (float)a.n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % a.n);
        float2 z = cmul(a.work[(idx / a.n) * a.M + k], a.chirp[k]);
        if (a.inverse) z = make_float2(z.x * invn, -z.y * invn);
        a.out[idx] = z;
    }
}

/* ------------------------------------------------------------------ log-mel */
/* out[f][m] = logf(sum_k power[f][k] W[m][k] + eps), the reference's src/features/mel.c:204-245, with the
 * filterbank stored sparsely (each triangular filter is non-zero on one contiguous bin range).
 * A CTA takes a tile of 32 frames: the power rows are loaded coalesced (lanes along bins) and parked
 * TRANSPOSED in shared memory, Ps[bin][frame] with a row pitch of 33, so that in the reduction lane = frame
 * reads conflict-free while the weight is one broadcast load for the whole warp.  Warp w owns bands
 * m = w, w+8, ...; every band sum is accumulated by one thread in ascending bin order with a separate
 * multiply and add (no FMA contraction), i.e. bit-for-bit the reference's float32 sum.  Results are
 * transposed back through shared memory and stored coalesced.  HBM-bound: reads each power value once. */
struct MelArgs {
    const float* power; long long pitch; long long frames;
    int bins, n_mels;
    const int* meta;         /* lo[n_mels] | len[n_mels] | off[n_mels] */
    const float* w;
    float eps;
    float* out;              /* [frames][n_mels] */
};
constexpr int MEL_KC = 512;       /* bins per shared-memory chunk: 66 KB, three CTAs per SM */
constexpr int MEL_BG = 128;       /* bands per pass: 8 warps x 16 accumulators */

__global__ void __launch_bounds__(256) logmel_kernel(const MelArgs a)
{
#ifdef VVB_EMU
    float* Ps = reinterpret_cast<float*>(vvb_emu::g_dyn_smem);
#else
    extern __shared__ __align__(16) float Ps[];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntiles = (a.frames + 31) / 32;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long f0 = tile * 32;
        const int nf = (int)min((long long)32, a.frames - f0);
        for (int mg = 0; mg < a.n_mels; mg += MEL_BG) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
            for (int k0 = 0; k0 < a.bins; k0 += MEL_KC) {
                const int kc = min(MEL_KC, a.bins - k0);
                __syncthreads();                                   /* the previous chunk / output staging is consumed */
                for (int r = warp; r < 32; r += 8) {
                    const float* row = a.power + (f0 + r) * a.pitch + k0;
                    for (int kb = lane; kb < kc; kb += 32 * 8) {       /* 8 independent loads in flight per lane */
                        float tmp[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int k = kb + 32 * u;
                            tmp[u] = (r < nf && k < kc) ? __ldg(row + k) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int k = kb + 32 * u;
                            if (k < kc) Ps[k * 33 + r] = tmp[u];
                        }
                    }
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int m = mg + warp + 8 * j;
                    if (m < a.n_mels) {                            /* warp-uniform */
                        const int lo = __ldg(a.meta + m), len = __ldg(a.meta + a.n_mels + m), off = __ldg(a.meta + 2 * a.n_mels + m);
                        const int ka = max(lo, k0), kb = min(lo + len, k0 + kc);
                        for (int k = ka; k < kb; ++k)
                            acc[j] = __fadd_rn(acc[j], __fmul_rn(Ps[(k - k0) * 33 + lane], __ldg(a.w + off + (k - lo))));
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int m = mg + warp + 8 * j;
                if (m < a.n_mels) Ps[(m - mg) * 33 + lane] = logf(acc[j] + a.eps);
            }
            __syncthreads();
            const int nb = min(MEL_BG, a.n_mels - mg);
            for (int r = warp; r < nf; r += 8)
                for (int mm = lane; mm < nb; mm += 32) a.out[(f0 + r) * a.n_mels + mg + mm] = Ps[mm * 33 + r];
        }
    }
}

/* The same reduction fed by TMA, for the usual case of densely packed power rows (pitch == bins, base 16-byte
 * aligned).  R consecutive frames are one contiguous run of R*bins floats, so ONE bulk copy per tile brings
 * them into shared memory untransposed; because bins = nfft/2+1 is odd the row pitch is already conflict-free
 * for lane = frame.  `stages` tile buffers form a ring: thread 0 issues the copy for tile it+stages-1 right
 * after the CTA has finished tile it-1, so stages-1 tiles (>= 128 KB at the headline shape) are in flight per
 * SM while the warps reduce the current one.
 * Work split: a thread owns FPT frames (1: 16 warps per CTA, the default; 2: rows r and r + R/2, two independent
 * sums that share every weight, 8 warps) and one of the 16 / 32 / 64 / 128 band slots of its tile shape.  The host has cut every filter into groups of four taps and dealt whole filters to
 * slots, longest first (mel.c), so a slot is one flat list of groups: no per-band inner loop, nothing to
 * diverge on inside a warp, and the loads of group g+1 (descriptor, four weights, eight power values) are
 * issued before the sums of group g.  Summation order per band is unchanged: ascending bins, separate
 * multiply and add, taps beyond the filter's end predicated off. */
constexpr int MEL_HDR = 64;       /* bytes reserved for the mbarriers in front of the tile ring */
constexpr int MEL_TMA_MAX_MELS = 512;

struct MelGroup { int4 d; float4 w; float p0[4], p1[4]; };

template <int R, int FPT> __global__ void __launch_bounds__(256 * (2 / FPT)) logmel_tma_kernel(const MelArgs a, const int stages, const int stage_floats,
                                                                           const int n_groups)
{
    constexpr int HALF = R / FPT, SUB = 32 / HALF, NTHR = 256 * (2 / FPT), NSLOT = (NTHR / 32) * SUB;   /* FPT frames per thread */
    constexpr int VARIANT = (R == 32) ? 0 : (R == 16) ? 1 : (R == 8) ? 2 : 3;
    static_assert(NSLOT == (16 << VARIANT), "slot count must match the host-side tables");
#ifdef VVB_EMU
    unsigned char* mel_smem = reinterpret_cast<unsigned char*>(vvb_emu::g_dyn_smem);
#else
    extern __shared__ __align__(16) float smem[];
    unsigned char* mel_smem = reinterpret_cast<unsigned char*>(smem);
#endif
    /* shared memory: mbarriers | group descriptors | quad weights | tile ring | output staging */
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(mel_smem);
    int4* s_desc = reinterpret_cast<int4*>(mel_smem + MEL_HDR);
    float4* s_wq = reinterpret_cast<float4*>(s_desc + n_groups);
    float* ring = reinterpret_cast<float*>(s_wq + n_groups);
    float* sOut = ring + (size_t)stages * stage_floats;                 /* [R][n_mels | 1] */
    const int opitch = a.n_mels | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = lane & (HALF - 1), r1 = (FPT == 2) ? r0 + HALF : r0, slot = warp * SUB + lane / HALF;
    const long long ntiles = (a.frames + R - 1) / R;
    const long long tile_floats = (long long)R * a.bins;
    const int* hdr = a.meta + 3 * a.n_mels;
    const int* slot_ptr = a.meta + __ldg(hdr + VARIANT);
    const int g_begin = __ldg(slot_ptr + slot), g_end = __ldg(slot_ptr + slot + 1);

    auto issue = [&](long long it) {                                    /* thread 0 only */
        const long long tile = blockIdx.x + it * gridDim.x;
        if (tile >= ntiles) return;
        const long long nf = min((long long)R, a.frames - tile * R);
        const unsigned bytes = (unsigned)(nf * a.bins * 4);
        if (bytes & 15u) return;                                        /* ragged tail: plain loads below */
        const int s = (int)(it % stages);
        fence_proxy_async();
        mbar_expect_tx(&bars[s], bytes);
        bulk_load(ring + (size_t)s * stage_floats, a.power + tile * tile_floats, bytes, &bars[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        for (int it = 0; it < stages - 1; ++it) issue(it);
    }
    {   /* the tables of this tile shape, once per CTA (the first tiles are already in flight) */
        const int4* gdesc = reinterpret_cast<const int4*>(a.meta + __ldg(hdr + 4 + VARIANT));
        const float4* gwq = reinterpret_cast<const float4*>(a.w + __ldg(hdr + 8));
        for (int g = tid; g < n_groups; g += NTHR) {
            const int4 d = __ldg(gdesc + g);
            s_desc[g] = d;
            s_wq[g] = __ldg(gwq + d.y);                                 /* weights follow the slot order too */
        }
    }
    for (long long it = 0;; ++it) {
        const long long tile = blockIdx.x + it * gridDim.x;
        if (tile >= ntiles) break;
        __syncthreads();                                   /* tile it-1 and its output staging are consumed */
        if (tid == 0) issue(it + stages - 1);
        const int s = (int)(it % stages);
        const long long f0 = tile * R;
        const int nf = (int)min((long long)R, a.frames - f0);
        float* P = ring + (size_t)s * stage_floats;
        if (((unsigned)(nf * (long long)a.bins * 4) & 15u) == 0) {
            mbar_wait(&bars[s], (unsigned)((it / stages) & 1));
        } else {
            const float* src = a.power + tile * tile_floats;
            for (long long i = tid; i < (long long)nf * a.bins; i += NTHR) P[i] = __ldg(src + i);
            __syncthreads();
        }
        const float* P0 = P + (size_t)r0 * a.bins;
        const float* P1 = P + (size_t)r1 * a.bins;
        /* taps past the end of a run are read (they stay inside the shared-memory allocation: the output
         * staging follows the ring) but never added */
        auto fetch = [&](int g, MelGroup& q) {
            q.d = s_desc[g];
            q.w = s_wq[g];
#pragma unroll
            for (int u = 0; u < 4; ++u) { q.p0[u] = P0[q.d.x + u]; if constexpr (FPT == 2) q.p1[u] = P1[q.d.x + u]; else q.p1[u] = 0.f; }
        };
        float acc0 = 0.f, acc1 = 0.f;
        auto reduce = [&](const MelGroup& q) {
            const int taps = q.d.z & 7;
            const float wv[4] = {q.w.x, q.w.y, q.w.z, q.w.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float t0 = __fmul_rn(q.p0[u], wv[u]), t1 = __fmul_rn(q.p1[u], wv[u]);
                if (u < taps) { acc0 = __fadd_rn(acc0, t0); acc1 = __fadd_rn(acc1, t1); }
            }
            if (q.d.z & 8) {                                            /* last group of a filter */
                sOut[r0 * opitch + q.d.w] = acc0;
                if constexpr (FPT == 2) sOut[r1 * opitch + q.d.w] = acc1;
                acc0 = 0.f; acc1 = 0.f;
            }
        };
        MelGroup ga, gb;
        if (g_begin < g_end) fetch(g_begin, ga);
        for (int g = g_begin; g < g_end; g += 2) {                      /* ping-pong: group g+1 is in flight while g is summed */
            if (g + 1 < g_end) fetch(g + 1, gb);
            reduce(ga);
            if (g + 1 < g_end) {
                if (g + 2 < g_end) fetch(g + 2, ga);
                reduce(gb);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nf * a.n_mels; idx += NTHR) {
            const int rr = idx / a.n_mels, mm = idx - rr * a.n_mels;
            a.out[(f0 + rr) * a.n_mels + mm] = logf(sOut[rr * opitch + mm] + a.eps);
        }
    }
}

/* ------------------------------------------------------------------ MFCC (second half) */
/* out[f][k] = lifter[k] * sum_n logmel[f][n] * table[k][n], the reference's unnormalised DCT-II
 * (src/spectral/dct.c:21-30) truncated to n_coeffs and liftered (src/features/mel.c:283-295).  The cosine
 * table and lifter factors come from the host (same libm calls as the reference), the sum runs over n
 * ascending with a separate multiply and add, so the result is the reference's float32 value bit for bit.
 * A CTA takes 32 frames: the log-mel tile is parked in shared memory with an odd row pitch (lane = frame
 * reads conflict-free, the table entry is a broadcast); warp w owns coefficients w, w+8, ... */
struct MfccArgs {
    const float* logmel; long long frames;
    int n_mels, n_coeffs;
    const float* table;      /* [n_coeffs][n_mels] */
    const float* lifter;     /* [n_coeffs] */
    float* out;              /* [frames][n_coeffs] */
};

__global__ void __launch_bounds__(256) mfcc_kernel(const MfccArgs a)
{
#ifdef VVB_EMU
    float* ms = reinterpret_cast<float*>(vvb_emu::g_dyn_smem);
#else
    extern __shared__ __align__(16) float smem[];
    float* ms = smem;
#endif
    const int pitch = a.n_mels | 1, opitch = a.n_coeffs | 1;
    float* tile = ms;                                    /* [32][pitch] */
    float* tab = tile + 32 * pitch;                      /* [n_coeffs][n_mels] */
    float* so = tab + a.n_coeffs * a.n_mels;             /* [32][opitch] */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < a.n_coeffs * a.n_mels; i += 256) tab[i] = __ldg(a.table + i);
    const long long ntiles = (a.frames + 31) / 32;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long f0 = t * 32;
        const int nf = (int)min((long long)32, a.frames - f0);
        __syncthreads();
        for (int i = tid; i < nf * a.n_mels; i += 256) {
            const int r = i / a.n_mels, n = i - r * a.n_mels;
            tile[r * pitch + n] = __ldg(a.logmel + f0 * a.n_mels + i);
        }
        __syncthreads();
        if (lane < nf) {
            /* two coefficients per pass share every x[n] load: two independent ordered chains per thread */
            const float* x = tile + lane * pitch;
            for (int k = warp; k < a.n_coeffs; k += 16) {
                const int k2 = k + 8;
                const bool two = k2 < a.n_coeffs;
                const float* c0 = tab + k * a.n_mels;
                const float* c1 = tab + (two ? k2 : k) * a.n_mels;
                float s0 = 0.f, s1 = 0.f;
                for (int n = 0; n < a.n_mels; ++n) {
                    const float xv = x[n];
                    s0 = __fadd_rn(s0, __fmul_rn(xv, c0[n]));
                    s1 = __fadd_rn(s1, __fmul_rn(xv, c1[n]));
                }
                so[lane * opitch + k] = __fmul_rn(s0, __ldg(a.lifter + k));
                if (two) so[lane * opitch + k2] = __fmul_rn(s1, __ldg(a.lifter + k2));
            }
        }
        __syncthreads();
        for (int i = tid; i < nf * a.n_coeffs; i += 256) {
            const int r = i / a.n_coeffs, k = i - r * a.n_coeffs;
            a.out[f0 * a.n_coeffs + i] = so[r * opitch + k];
        }
    }
}

/* ------------------------------------------------------------------ PCM decode */
/* interleaved little-endian WAV samples -> planar float32, the reference's src/audio/wav.c:458-521:
 * (float)code * 2^-15 / 2^-23 / 2^-31 (int -> float rounds to nearest even like the C cast, the scale is a
 * power of two, so the result is exact to the bit); format -32 copies IEEE float32.  One thread per output
 * sample, consecutive threads along time: stores are coalesced, loads too for mono. */
struct PcmArgs {
    const unsigned char* in; int format; long long num_samples; int channels;
    float* out; long long pitch;
};

__global__ void pcm_to_planar_kernel(const PcmArgs a)
{
    const long long total = a.num_samples * a.channels;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long s = idx % a.num_samples;
        const int c = (int)(idx / a.num_samples);
        const long long i = s * a.channels + c;
        float v;
        if (a.format == -32) {
            v = reinterpret_cast<const float*>(a.in)[i];
        } else if (a.format == 16) {
            v = __fmul_rn((float)reinterpret_cast<const short*>(a.in)[i], 1.0f / 32768.0f);
        } else if (a.format == 24) {
            const unsigned char* b = a.in + 3 * i;
            int q = (int)b[0] | ((int)b[1] << 8) | ((int)b[2] << 16);
            if (q & 0x800000) q |= (int)0xFF000000;
            v = __fmul_rn((float)q, 1.0f / 8388608.0f);
        } else {
            v = __fmul_rn((float)reinterpret_cast<const int*>(a.in)[i], 1.0f / 2147483648.0f);
        }
        a.out[c * a.pitch + s] = v;
    }
}

}  // namespace vvb
