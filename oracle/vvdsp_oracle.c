/*
 * vvdsp_oracle.c -- CPU restatement of vv-dsp's STFT/ISTFT/FFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product library (vv_dsp_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and there only as the checker / the CPU arm.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (a) bit-for-bit against the reference itself, compiled from its own
 *       sources into oracle/_ref/libvvdsp_ref.so (recipe: oracle/Makefile),
 *   (b) against sha256 digests / slices generated from that build and
 *       committed under tests/golden/ (script: tests/golden/make_golden.py),
 *   (c) against every known answer the reference's own tests hold for this
 *       path (framing goldens, Hann formula, impulse -> ones, ...).
 *
 * All arithmetic is IEEE float32 in exactly the reference's operation order.
 * Build with -std=c99 -ffp-contract=off and without -ffast-math / -march=native
 * so no FMA contraction changes the bits (oracle/Makefile does this).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef struct { float re, im; } orc_cpx;

enum { ORC_WIN_BOXCAR = 0, ORC_WIN_HANN = 1, ORC_WIN_HAMMING = 2 };
/* frame-count conventions, SURVEY.md section 8(a) last row */
enum { ORC_FRAMES_VALID = 0, ORC_FRAMES_SPECTROGRAM = 1, ORC_FRAMES_PADDED_TAIL = 2, ORC_FRAMES_CENTER = 3 };

/* include/vv_dsp/vv_dsp_math.h:11-27 -- pi is the double literal rounded to float */
#define ORC_PI_D 3.141592653589793238462643383279502884
static const float ORC_PI = (float)ORC_PI_D;
static const float ORC_TWO_PI = (float)(2.0 * ORC_PI_D);

/* ------------------------------------------------------------------ windows */

/* src/window/window.c:16-49: boxcar = 1; hann = 0.5 - 0.5 cosf(2pi/(N-1) * n);
 * hamming = 0.54 - 0.46 cosf(...); N == 1 -> 1.0; symmetric (denominator N-1). */
int orc_window(int type, size_t n, float *w)
{
    if (!w) return 1;
    if (n == 0) return 2;
    if (type == ORC_WIN_BOXCAR) {
        for (size_t i = 0; i < n; ++i) w[i] = 1.0f;
        return 0;
    }
    if (type != ORC_WIN_HANN && type != ORC_WIN_HAMMING) return 3;
    if (n == 1) { w[0] = 1.0f; return 0; }
    const float a0 = (type == ORC_WIN_HANN) ? 0.5f : 0.54f;
    const float a1 = (type == ORC_WIN_HANN) ? 0.5f : 0.46f;
    const float step = ORC_TWO_PI / (float)(n - 1);
    for (size_t i = 0; i < n; ++i) {
        float c = cosf(step * (float)i);
        w[i] = a0 - a1 * c;
    }
    return 0;
}

/* ---------------------------------------------------------------------- FFT */

static int orc_is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }

/* src/spectral/fft_kiss.c:27-74.  In-place radix-2 decimation in time:
 * bit-reversal swap pass, then for size = 2,4,..,n a per-block twiddle that is
 * advanced by the float recurrence w <- w * wp (this recurrence is where the
 * reference's large-N error comes from).  sign=+1 is
 * the forward transform exp(-j...), sign=-1 the backward one scaled by 1/n. */
static void orc_radix2(orc_cpx *a, size_t n, int sign)
{
    unsigned bits = 0;
    for (size_t t = n; t > 1; t >>= 1) ++bits;

    for (size_t i = 0; i < n; ++i) {
        size_t j = 0, x = i;
        for (unsigned b = 0; b < bits; ++b) { j = (j << 1) | (x & 1u); x >>= 1; }
        if (j > i) { orc_cpx t = a[i]; a[i] = a[j]; a[j] = t; }
    }

    for (size_t span = 2; span <= n; span <<= 1) {
        const size_t half = span >> 1;
        const float theta = (float)(-sign) * 2.0f * ORC_PI / (float)span;
        const float wpr = cosf(theta);
        const float wpi = sinf(theta);
        for (size_t base = 0; base < n; base += span) {
            float wr = 1.0f, wi = 0.0f;
            for (size_t k = 0; k < half; ++k) {
                orc_cpx *lo = a + base + k;
                orc_cpx *hi = lo + half;
                const float tr = wr * hi->re - wi * hi->im;
                const float ti = wr * hi->im + wi * hi->re;
                hi->re = lo->re - tr;
                hi->im = lo->im - ti;
                lo->re += tr;
                lo->im += ti;
                const float nr = wr * wpr - wi * wpi;
                const float ni = wr * wpi + wi * wpr;
                wr = nr; wi = ni;
            }
        }
    }

    if (sign < 0) {
        const float inv = 1.0f / (float)n;
        for (size_t i = 0; i < n; ++i) { a[i].re *= inv; a[i].im *= inv; }
    }
}

/* src/spectral/fft_kiss.c:76-92.  O(n^2) direct DFT, one cosf/sinf per term,
 * angle computed as (-sign)*2*pi*(float)(k*t)/n in float. */
static void orc_dft_direct(const orc_cpx *in, orc_cpx *out, size_t n, int sign)
{
    const float scale = (sign < 0) ? (1.0f / (float)n) : 1.0f;
    for (size_t k = 0; k < n; ++k) {
        float sr = 0, si = 0;
        for (size_t t = 0; t < n; ++t) {
            const float ang = (float)(-sign) * 2.0f * ORC_PI * (float)(k * t) / (float)n;
            const float c = cosf(ang), s = sinf(ang);
            sr += in[t].re * c - in[t].im * s;
            si += in[t].re * s + in[t].im * c;
        }
        out[k].re = sr * scale;
        out[k].im = si * scale;
    }
}

/* src/spectral/fft_kiss.c:105-118 (C2C branch of kiss_execute): power of two
 * -> copy then in-place radix-2; otherwise the direct DFT. dir: +1 / -1. */
int orc_fft_c2c(const orc_cpx *in, orc_cpx *out, size_t n, int dir)
{
    if (!in || !out) return 1;
    if (n == 0) return 2;
    if (dir != 1 && dir != -1) return 3;
    if (orc_is_pow2(n)) {
        if (out != in) memcpy(out, in, n * sizeof(orc_cpx));
        orc_radix2(out, n, dir);
    } else {
        orc_dft_direct(in, out, n, dir);
    }
    return 0;
}

/* src/spectral/fft_kiss.c:120-147 (R2C): C2C forward on (x,0), keep n/2+1
 * bins, force the Nyquist imaginary part to exactly 0 for even n. */
int orc_fft_r2c(const float *in, orc_cpx *out, size_t n)
{
    if (!in || !out) return 1;
    if (n == 0) return 2;
    orc_cpx *a = (orc_cpx *)malloc(n * sizeof(orc_cpx));
    orc_cpx *b = (orc_cpx *)malloc(n * sizeof(orc_cpx));
    if (!a || !b) { free(a); free(b); return 4; }
    for (size_t i = 0; i < n; ++i) { a[i].re = in[i]; a[i].im = 0.0f; }
    orc_fft_c2c(a, b, n, +1);
    const size_t nh = n / 2 + 1;
    memcpy(out, b, nh * sizeof(orc_cpx));
    if (n % 2 == 0 && nh > 1) out[nh - 1].im = 0.0f;
    free(a); free(b);
    return 0;
}

/* src/spectral/fft_kiss.c:149-174 (C2R): mirror-expand the Hermitian half to n
 * bins and ALWAYS run the direct DFT backward (even for powers of two); keep
 * the real part. */
int orc_fft_c2r(const orc_cpx *in, float *out, size_t n)
{
    if (!in || !out) return 1;
    if (n == 0) return 2;
    const size_t nh = n / 2 + 1;
    orc_cpx *full = (orc_cpx *)malloc(n * sizeof(orc_cpx));
    orc_cpx *time = (orc_cpx *)malloc(n * sizeof(orc_cpx));
    if (!full || !time) { free(full); free(time); return 4; }
    for (size_t k = 0; k < nh && k < n; ++k) full[k] = in[k];
    for (size_t k = nh; k < n; ++k) {
        const size_t m = n - k;
        if (m < nh && m > 0) { full[k].re = in[m].re; full[k].im = -in[m].im; }
        else { full[k].re = 0.0f; full[k].im = 0.0f; }
    }
    orc_dft_direct(full, time, n, -1);
    for (size_t i = 0; i < n; ++i) out[i] = time[i].re;
    free(full); free(time);
    return 0;
}

/* --------------------------------------------------------------------- STFT */

typedef struct orc_stft {
    size_t nfft, hop;
    float *win;      /* nfft window coefficients */
    float *tbuf;     /* nfft windowed samples */
    orc_cpx *cin;    /* nfft packed (x*w, 0) */
    orc_cpx *ctime;  /* nfft backward output */
} orc_stft;

/* src/spectral/stft.c:30-60: validation order NULL -> size -> window enum.
 * (The reference leaks two scratch buffers on its error paths; not reproduced.) */
int orc_stft_create(size_t nfft, size_t hop, int win, orc_stft **out)
{
    if (!out) return 1;
    *out = NULL;
    if (nfft == 0 || hop == 0 || hop > nfft) return 2;
    orc_stft *h = (orc_stft *)calloc(1, sizeof(*h));
    if (!h) return 4;
    h->nfft = nfft; h->hop = hop;
    h->win = (float *)malloc(nfft * sizeof(float));
    h->tbuf = (float *)malloc(nfft * sizeof(float));
    h->cin = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    h->ctime = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    int st = (h->win && h->tbuf && h->cin && h->ctime) ? orc_window(win, nfft, h->win) : 4;
    if (st != 0) {
        free(h->win); free(h->tbuf); free(h->cin); free(h->ctime); free(h);
        return st;
    }
    *out = h;
    return 0;
}

void orc_stft_destroy(orc_stft *h)
{
    if (!h) return;
    free(h->win); free(h->tbuf); free(h->cin); free(h->ctime); free(h);
}

/* src/spectral/stft.c:74-92 + src/core/vv_dsp_vectorized_math_fallback.c:13-29:
 * t = in*w (float), pack (t,0), C2C forward unscaled, all nfft bins out. */
int orc_stft_process(orc_stft *h, const float *in, orc_cpx *out)
{
    if (!h || !in || !out) return 1;
    const size_t n = h->nfft;
    for (size_t i = 0; i < n; ++i) h->tbuf[i] = in[i] * h->win[i];
    for (size_t i = 0; i < n; ++i) { h->cin[i].re = h->tbuf[i]; h->cin[i].im = 0.0f; }
    return orc_fft_c2c(h->cin, out, n, +1);
}

/* src/spectral/stft.c:95-110: C2C backward (x 1/n), v = Re * w,
 * out_add += v, norm_add += w*w when given. */
int orc_stft_reconstruct(orc_stft *h, const orc_cpx *in, float *out_add, float *norm_add)
{
    if (!h || !in || !out_add) return 1;
    const size_t n = h->nfft;
    int st = orc_fft_c2c(in, h->ctime, n, -1);
    if (st) return st;
    for (size_t i = 0; i < n; ++i) {
        const float w = h->win[i];
        const float v = h->ctime[i].re * w;
        out_add[i] += v;
        if (norm_add) norm_add[i] += w * w;
    }
    return 0;
}

/* ------------------------------------------------------------------ framing */

/* src/core/framing.c:58-69 (valid / centred) and src/spectral/stft.c:119
 * (spectrogram) and tests/spectral_tests.c:101 (padded tail). */
size_t orc_num_frames(size_t n, size_t nfft, size_t hop, int convention)
{
    if (hop == 0) return 0;
    switch (convention) {
    case ORC_FRAMES_VALID:          return n < nfft ? 0 : 1 + (n - nfft) / hop;
    case ORC_FRAMES_SPECTROGRAM:    return n < nfft ? 1 : 1 + (n - nfft + hop) / hop;
    case ORC_FRAMES_PADDED_TAIL: {  /* frames while start + nfft <= n + (nfft - hop) */
        size_t f = 0;
        if (n + (nfft - hop) < nfft) return 0;
        f = 1 + (n + (nfft - hop) - nfft) / hop;
        return f;
    }
    case ORC_FRAMES_CENTER:         return (n + hop - 1) / hop;
    default: return 0;
    }
}

/* src/core/framing.c:21-56: edge-inclusive ("symmetric") reflection,
 * -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2, multiple reflections folded. */
static size_t orc_reflect(long idx, size_t n)
{
    const long len = (long)n;
    if (n == 0) return 0;
    if (idx < 0) {
        long a = -idx - 1;
        if (a >= len) {
            const long period = 2 * len;
            a %= period;
            if (a >= len) a = period - 1 - a;
        }
        return (size_t)a;
    }
    if (idx >= len) {
        long r = len - 1 - (idx - len);
        if (r < 0) {
            r = -r - 1;
            if (r >= len) {
                const long period = 2 * len;
                r %= period;
                if (r >= len) r = period - 1 - r;
            }
        }
        if (r < 0) r = 0;
        if (r > len - 1) r = len - 1;
        return (size_t)r;
    }
    return (size_t)idx;
}

/* src/core/framing.c:71-121 */
int orc_fetch_frame(const float *x, size_t n, float *frame, size_t flen, size_t hop,
                    size_t index, int center, const float *window)
{
    if (!x || !frame) return 1;
    if (n == 0 || flen == 0 || hop == 0) return 2;
    long start = (long)(index * hop);
    if (center) start -= (long)(flen / 2);
    for (size_t i = 0; i < flen; ++i) {
        const long s = start + (long)i;
        float v;
        if (center) v = x[orc_reflect(s, n)];
        else v = (s < 0 || s >= (long)n) ? 0.0f : x[s];
        frame[i] = window ? v * window[i] : v;
    }
    return 0;
}

/* src/core/framing.c:123-148 */
int orc_overlap_add(const float *frame, float *out, size_t out_len, size_t flen, size_t hop, size_t index)
{
    if (!frame || !out) return 1;
    if (out_len == 0 || flen == 0 || hop == 0) return 2;
    const size_t start = index * hop;
    for (size_t i = 0; i < flen; ++i)
        if (start + i < out_len) out[start + i] += frame[i];
    return 0;
}

/* ------------------------------------------------- whole-signal conveniences */

/* src/spectral/stft.c:112-144: frames by the SPECTROGRAM convention, trailing
 * zero pad, magnitude sqrtf(re^2+im^2) for all nfft bins, row-major. */
int orc_stft_spectrogram(orc_stft *h, const float *x, size_t n, float *out_mag, size_t *out_frames)
{
    if (!h || !x || !out_mag || !out_frames) return 1;
    const size_t nfft = h->nfft, hop = h->hop;
    const size_t frames = orc_num_frames(n, nfft, hop, ORC_FRAMES_SPECTROGRAM);
    *out_frames = frames;
    orc_cpx *spec = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    float *frame = (float *)malloc(nfft * sizeof(float));
    if (!spec || !frame) { free(spec); free(frame); return 4; }
    for (size_t f = 0; f < frames; ++f) {
        for (size_t i = 0; i < nfft; ++i) {
            const size_t s = f * hop + i;
            frame[i] = s < n ? x[s] : 0.0f;
        }
        orc_stft_process(h, frame, spec);
        for (size_t k = 0; k < nfft; ++k)
            out_mag[f * nfft + k] = sqrtf(spec[k].re * spec[k].re + spec[k].im * spec[k].im);
    }
    free(spec); free(frame);
    return 0;
}

/* Analysis of a whole signal, the way every reference caller loops
 * (tools/dump_stft_roundtrip.c:44-45, tests/spectral_tests.c:101-108): frame f
 * = x[f*hop .. f*hop+nfft) with zeros past n (center=0) or the centred
 * reflect-padded gather of vv_dsp_fetch_frame (center convention), then
 * vv_dsp_stft_process.  Output: half spectra [frames][nfft/2+1] (bins 0..nfft/2
 * of the full C2C output) when half != 0, else all nfft bins. */
int orc_stft_forward(orc_stft *h, const float *x, size_t n, int convention, int half,
                     orc_cpx *out, size_t frames)
{
    if (!h || !x || !out) return 1;
    const size_t nfft = h->nfft, hop = h->hop;
    const size_t bins = half ? nfft / 2 + 1 : nfft;
    orc_cpx *spec = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    float *frame = (float *)malloc(nfft * sizeof(float));
    if (!spec || !frame) { free(spec); free(frame); return 4; }
    for (size_t f = 0; f < frames; ++f) {
        if (n == 0) memset(frame, 0, nfft * sizeof(float));
        else orc_fetch_frame(x, n, frame, nfft, hop, f, convention == ORC_FRAMES_CENTER, NULL);
        orc_stft_process(h, frame, spec);
        memcpy(out + f * bins, spec, bins * sizeof(orc_cpx));
    }
    free(spec); free(frame);
    return 0;
}

/* "Power" as consumed by include/vv_dsp/features/mel.h:74-77: re^2+im^2 of
 * the process() output, bins 0..nfft/2 (Power spectrum). */
int orc_stft_power(orc_stft *h, const float *x, size_t n, int convention, float *out, size_t frames)
{
    if (!h || !x || !out) return 1;
    const size_t nfft = h->nfft, hop = h->hop, bins = nfft / 2 + 1;
    orc_cpx *spec = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    float *frame = (float *)malloc(nfft * sizeof(float));
    if (!spec || !frame) { free(spec); free(frame); return 4; }
    for (size_t f = 0; f < frames; ++f) {
        if (n == 0) memset(frame, 0, nfft * sizeof(float));
        else orc_fetch_frame(x, n, frame, nfft, hop, f, convention == ORC_FRAMES_CENTER, NULL);
        orc_stft_process(h, frame, spec);
        for (size_t k = 0; k < bins; ++k)
            out[f * bins + k] = spec[k].re * spec[k].re + spec[k].im * spec[k].im;
    }
    free(spec); free(frame);
    return 0;
}

/* Synthesis of a whole signal exactly as tools/dump_stft_roundtrip.c:44-54:
 * zeroed recon/norm of length n_out (+nfft slack so padded-tail frames fit,
 * tests/spectral_tests.c:97-99), reconstruct every frame at f*hop with the
 * norm accumulator, then y = norm > 1e-12 ? recon/norm : 0.  Input is the
 * full nfft-bin spectrum per frame (half == 0) or the half spectrum, which is
 * mirror-expanded X[nfft-k] = conj(X[k]) first (the layout the batched GPU
 * entry points use).  normalise == 0 returns the raw overlap-add sum. */
int orc_stft_istft(orc_stft *h, const orc_cpx *spec, size_t frames, int half,
                   float *y, size_t n_out, int normalise)
{
    if (!h || !spec || !y) return 1;
    const size_t nfft = h->nfft, hop = h->hop;
    const size_t bins = half ? nfft / 2 + 1 : nfft;
    const size_t span = (frames ? (frames - 1) * hop + nfft : 0);
    const size_t len = (span > n_out ? span : n_out);
    float *recon = (float *)calloc(len ? len : 1, sizeof(float));
    float *norm = (float *)calloc(len ? len : 1, sizeof(float));
    orc_cpx *full = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    if (!recon || !norm || !full) { free(recon); free(norm); free(full); return 4; }
    for (size_t f = 0; f < frames; ++f) {
        const orc_cpx *src = spec + f * bins;
        if (half) {
            for (size_t k = 0; k < bins; ++k) full[k] = src[k];
            for (size_t k = bins; k < nfft; ++k) { full[k].re = src[nfft - k].re; full[k].im = -src[nfft - k].im; }
            src = full;
        }
        orc_stft_reconstruct(h, src, recon + f * hop, norm + f * hop);
    }
    for (size_t i = 0; i < n_out; ++i) {
        if (normalise) y[i] = norm[i] > 1e-12f ? recon[i] / norm[i] : 0.0f;
        else y[i] = recon[i];
    }
    free(recon); free(norm); free(full);
    return 0;
}

/* Round trip used as the CPU baseline loop: process ->
 * reconstruct(norm) per valid frame, then normalise; never materialises spectra. */
int orc_stft_roundtrip(orc_stft *h, const float *x, size_t n, float *y)
{
    if (!h || !x || !y) return 1;
    const size_t nfft = h->nfft, hop = h->hop;
    float *recon = (float *)calloc(n ? n : 1, sizeof(float));
    float *norm = (float *)calloc(n ? n : 1, sizeof(float));
    orc_cpx *spec = (orc_cpx *)malloc(nfft * sizeof(orc_cpx));
    if (!recon || !norm || !spec) { free(recon); free(norm); free(spec); return 4; }
    for (size_t f = 0; f * hop + nfft <= n; ++f) {
        orc_stft_process(h, x + f * hop, spec);
        orc_stft_reconstruct(h, spec, recon + f * hop, norm + f * hop);
    }
    for (size_t i = 0; i < n; ++i) y[i] = norm[i] > 1e-12f ? recon[i] / norm[i] : 0.0f;
    free(recon); free(norm); free(spec);
    return 0;
}

/* ---------------------------------------------- mel filterbank / log-mel */

/* src/features/mel.c:14-29 -- HTK mel scale in float32: 2595 log10f(1 + hz/700), inverse with powf */
float orc_hz_to_mel(float hz) { return hz < 0.0f ? 0.0f : 2595.0f * log10f(1.0f + hz / 700.0f); }
float orc_mel_to_hz(float mel) { return mel < 0.0f ? 0.0f : 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); }

/* src/features/mel.c:51-62: first index whose value is not below v (lower bound) */
static size_t orc_lower_bound(const float *a, size_t n, float v)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* src/features/mel.c:66-185: dense [n_mels][n_fft/2+1] triangular filters; band edges are n_mels+2
 * points equally spaced in mel (start + step*i in float32, :35-46) mapped back to Hz; rising slope over
 * bins [left_idx, center_idx), falling over [center_idx, right_idx); each filter divided by its sum.
 * Status codes as the reference: NULL -> 1, bad sizes / n_mels >= bins -> 2, fmax > sr/2 or a variant
 * other than HTK -> 3.  weights must hold n_mels*(n_fft/2+1) floats. */
int orc_mel_filterbank(size_t n_fft, size_t n_mels, float sr, float fmin, float fmax, int variant, float *weights)
{
    if (!weights) return 1;
    if (n_fft == 0 || n_mels == 0 || sr <= 0.0f || fmin < 0.0f || fmax <= fmin) return 2;
    if (fmax > sr / 2.0f) return 3;
    if (variant != 0) return 3;
    const size_t bins = n_fft / 2 + 1, npts = n_mels + 2;
    if (n_mels >= bins) return 2;
    float *hz = (float *)malloc(npts * sizeof(float));
    float *freq = (float *)malloc(bins * sizeof(float));
    if (!hz || !freq) { free(hz); free(freq); return 4; }
    memset(weights, 0, n_mels * bins * sizeof(float));
    const float m0 = orc_hz_to_mel(fmin), m1 = orc_hz_to_mel(fmax);
    const float step = (m1 - m0) / (float)(npts - 1);
    for (size_t i = 0; i < npts; ++i) hz[i] = orc_mel_to_hz(m0 + step * (float)i);
    for (size_t k = 0; k < bins; ++k) freq[k] = (float)k * sr / (float)n_fft;
    for (size_t m = 0; m < n_mels; ++m) {
        const float left = hz[m], center = hz[m + 1], right = hz[m + 2];
        const size_t li = orc_lower_bound(freq, bins, left), ci = orc_lower_bound(freq, bins, center),
                     ri = orc_lower_bound(freq, bins, right);
        float *w = weights + m * bins;
        for (size_t k = li; k < ci && k < bins; ++k) w[k] = (freq[k] - left) / (center - left);
        for (size_t k = ci; k < ri && k < bins; ++k) w[k] = (right - freq[k]) / (right - center);
        float sum = 0.0f;
        for (size_t k = 0; k < bins; ++k) sum += w[k];
        if (sum > 0.0f) for (size_t k = 0; k < bins; ++k) w[k] /= sum;
    }
    free(hz); free(freq);
    return 0;
}

/* src/features/mel.c:204-245: out[f][m] = logf(sum_k power[f][k]*W[m][k] + eps), k ascending in float32 */
int orc_log_mel(const float *power, size_t frames, size_t bins, const float *weights, size_t n_mels, float eps, float *out)
{
    if (!power || !weights || !out) return 1;
    if (frames == 0 || bins == 0 || n_mels == 0) return 2;
    if (eps < 0.0f) return 3;
    for (size_t f = 0; f < frames; ++f)
        for (size_t m = 0; m < n_mels; ++m) {
            float e = 0.0f;
            for (size_t k = 0; k < bins; ++k) e += power[f * bins + k] * weights[m * bins + k];
            out[f * n_mels + m] = logf(e + eps);
        }
    return 0;
}

/* src/features/mel.c:249-310 on top of the naive DCT-II of src/spectral/dct.c:21-30:
 * c[k] = sum_n x[n]*cosf(pi*(n+0.5)*k/N) (angle and sum in float32, n ascending, unnormalised), the first
 * n_coeffs kept, then c[i] *= 1 + (L/2)*sinf(pi*i/L) for i >= 1 when L > 0.  dct_type: only DCT-II (2). */
int orc_mfcc(const float *log_mel, size_t frames, size_t n_mels, size_t n_coeffs, int dct_type, float lifter, float *out)
{
    const float pi = (float)3.141592653589793238462643383279502884;
    if (!log_mel || !out) return 1;
    if (frames == 0 || n_mels == 0 || n_coeffs == 0) return 2;
    if (n_coeffs > n_mels) return 2;
    if (dct_type != 2) return 3;
    if (lifter < 0.0f) return 3;
    for (size_t f = 0; f < frames; ++f) {
        const float *x = log_mel + f * n_mels;
        float *c = out + f * n_coeffs;
        for (size_t k = 0; k < n_coeffs; ++k) {
            float sum = 0;
            for (size_t n = 0; n < n_mels; ++n) {
                float ang = pi * ((float)n + 0.5f) * (float)k / (float)n_mels;
                sum += x[n] * cosf(ang);
            }
            c[k] = sum;
        }
        if (lifter > 0.0f)
            for (size_t i = 1; i < n_coeffs; ++i) {
                float factor = 1.0f + (lifter / 2.0f) * sinf((float)3.14159265358979323846 * (float)i / lifter);
                c[i] *= factor;
            }
    }
    return 0;
}

/* src/audio/wav.c:458-521: interleaved little-endian samples -> planar float32 [channels][pitch].
 * format 16 / 24 / 32 = signed PCM scaled by 1/2^15, 1/2^23, 1/2^31; format -32 = IEEE float32 copied. */
int orc_pcm_to_planar(const void *interleaved, int format, size_t num_samples, size_t channels, float *planar, size_t pitch)
{
    const unsigned char *b = (const unsigned char *)interleaved;
    if (!interleaved || !planar) return 1;
    if (num_samples == 0 || channels == 0 || pitch < num_samples) return 2;
    if (format != 16 && format != 24 && format != 32 && format != -32) return 3;
    for (size_t s = 0; s < num_samples; ++s)
        for (size_t c = 0; c < channels; ++c) {
            const size_t i = s * channels + c;
            float v;
            if (format == -32) { memcpy(&v, b + 4 * i, 4); }
            else if (format == 16) { int16_t q; memcpy(&q, b + 2 * i, 2); v = (float)q * (float)(1.0 / 32768.0); }
            else if (format == 24) {
                int32_t q = (int32_t)b[3 * i] | ((int32_t)b[3 * i + 1] << 8) | ((int32_t)b[3 * i + 2] << 16);
                if (q & 0x800000) q |= (int32_t)0xFF000000;
                v = (float)q * (float)(1.0 / 8388608.0);
            } else { int32_t q; memcpy(&q, b + 4 * i, 4); v = (float)q * (float)(1.0 / 2147483648.0); }
            planar[c * pitch + s] = v;
        }
    return 0;
}
