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

}  // namespace vvb
