/*
 * reference_style_callers.c -- C callers written the way vv-dsp's own tests and tools use the library, compiled
 * against THIS repo's include/ and linked with libvvdsp_b200.so, run on the GPU by tests/test_c_callers.py.
 * (Own code, not a copy: the known answers it checks are the reference's, each cited.)
 *
 *   framing goldens            reference tests/framing_tests.c:17-45,56-73,85-102,130-150
 *   Hann N=8 / N=17 formula    reference tests/window_tests.c:118-131 (through VV_DSP_COS / VV_DSP_TWO_PI)
 *   impulse -> all-ones FFT    reference tests/fft_backend_tests.c:76-99, tests/spectral_tests.c:22-31
 *   zero frame -> zero bins    reference tests/gtest/test_stft.cpp:400-407
 *   per-frame STFT round trip  reference tools/dump_stft_roundtrip.c:44-54, tolerance of tests/gtest/test_stft.cpp:452-522
 *   status codes               reference src/spectral/stft.c:30-72, src/spectral/fft.c:15-107
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vv_dsp/vv_dsp.h"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

static void framing_goldens(void)
{
    const vv_dsp_real sig[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    vv_dsp_real frame[4], out[8];
    size_t i;
    CHECK(vv_dsp_get_num_frames(8, 4, 2, 0) == 3);
    CHECK(vv_dsp_get_num_frames(8, 4, 2, 1) == 4);
    CHECK(vv_dsp_get_num_frames(3, 4, 2, 0) == 0);
    CHECK(vv_dsp_get_num_frames(8, 4, 0, 0) == 0);
    CHECK(vv_dsp_fetch_frame(sig, 8, frame, 4, 2, 1, 0, NULL) == VV_DSP_OK);
    CHECK(frame[0] == 3 && frame[1] == 4 && frame[2] == 5 && frame[3] == 6);
    CHECK(vv_dsp_fetch_frame(sig, 8, frame, 4, 2, 3, 0, NULL) == VV_DSP_OK);     /* runs off the end: zero padded */
    CHECK(frame[0] == 7 && frame[1] == 8 && frame[2] == 0 && frame[3] == 0);
    CHECK(vv_dsp_fetch_frame(sig, 8, frame, 4, 2, 0, 1, NULL) == VV_DSP_OK);     /* centred: edge-inclusive reflection */
    CHECK(frame[0] == 2 && frame[1] == 1 && frame[2] == 1 && frame[3] == 2);
    memset(out, 0, sizeof(out));
    for (i = 0; i < 3; ++i) {
        CHECK(vv_dsp_fetch_frame(sig, 8, frame, 4, 2, i, 0, NULL) == VV_DSP_OK);
        CHECK(vv_dsp_overlap_add(frame, out, 8, 4, 2, i) == VV_DSP_OK);
    }
    {
        const vv_dsp_real want[8] = {1, 2, 6, 8, 10, 12, 7, 8};
        for (i = 0; i < 8; ++i) CHECK(out[i] == want[i]);
    }
    CHECK(vv_dsp_fetch_frame(NULL, 8, frame, 4, 2, 0, 0, NULL) == VV_DSP_ERROR_NULL_POINTER);
}

static void window_formula(void)
{
    static const size_t sizes[2] = {8, 17};
    size_t s, n;
    for (s = 0; s < 2; ++s) {
        const size_t N = sizes[s];
        vv_dsp_real w[17], h[17];
        CHECK(vv_dsp_window_hann(N, w) == VV_DSP_OK);
        CHECK(vv_dsp_window_hamming(N, h) == VV_DSP_OK);
        for (n = 0; n < N; ++n) {
            const vv_dsp_real c = VV_DSP_COS((vv_dsp_real)(VV_DSP_TWO_PI) * (vv_dsp_real)n / (vv_dsp_real)(N - 1));
            CHECK(fabsf(w[n] - ((vv_dsp_real)0.5 - (vv_dsp_real)0.5 * c)) <= 1e-6f);
            CHECK(fabsf(h[n] - ((vv_dsp_real)0.54 - (vv_dsp_real)0.46 * c)) <= 1e-6f);
        }
        CHECK(w[0] == 0.0f || fabsf(w[0]) < 1e-7f);
    }
    {
        vv_dsp_real one[1];
        CHECK(vv_dsp_window_hann(1, one) == VV_DSP_OK && one[0] == 1.0f);
        CHECK(vv_dsp_window_boxcar(0, one) == VV_DSP_ERROR_INVALID_SIZE || vv_dsp_window_boxcar(0, one) == VV_DSP_ERROR_NULL_POINTER
              || vv_dsp_window_boxcar(0, one) == VV_DSP_OK);
    }
}

static void fft_known_answers(void)
{
    static const size_t sizes[4] = {8, 16, 100, 1024};
    size_t s, k;
    for (s = 0; s < 4; ++s) {
        const size_t n = sizes[s];
        vv_dsp_fft_plan* p = NULL;
        vv_dsp_cpx* x = (vv_dsp_cpx*)calloc(n, sizeof(vv_dsp_cpx));
        vv_dsp_cpx* X = (vv_dsp_cpx*)calloc(n, sizeof(vv_dsp_cpx));
        x[0].re = 1.0f;                                                     /* unit impulse -> flat spectrum of ones */
        CHECK(vv_dsp_fft_make_plan(n, VV_DSP_FFT_C2C, VV_DSP_FFT_FORWARD, &p) == VV_DSP_OK);
        CHECK(vv_dsp_fft_execute(p, x, X) == VV_DSP_OK);
        for (k = 0; k < n; ++k) CHECK(fabsf(X[k].re - 1.0f) < 1e-5f && fabsf(X[k].im) < 1e-5f);
        CHECK(vv_dsp_fft_destroy(p) == VV_DSP_OK);
        free(x); free(X);
    }
    CHECK(vv_dsp_fft_destroy(NULL) == VV_DSP_OK);                               /* reference fft.c:103 */
    {
        vv_dsp_fft_plan* p = (vv_dsp_fft_plan*)1;
        CHECK(vv_dsp_fft_make_plan(0, VV_DSP_FFT_C2C, VV_DSP_FFT_FORWARD, &p) == VV_DSP_ERROR_INVALID_SIZE && p == NULL);
        CHECK(vv_dsp_fft_make_plan(8, (vv_dsp_fft_type)7, VV_DSP_FFT_FORWARD, &p) == VV_DSP_ERROR_OUT_OF_RANGE);
        CHECK(vv_dsp_fft_set_backend(VV_DSP_FFT_BACKEND_FFTW) == VV_DSP_ERROR_UNSUPPORTED);
        CHECK(vv_dsp_fft_set_backend((vv_dsp_fft_backend)9) == VV_DSP_ERROR_OUT_OF_RANGE);
        CHECK(vv_dsp_fft_get_backend() == VV_DSP_FFT_BACKEND_KISS);
    }
}

static float urand(unsigned* state)
{
    *state = *state * 1664525u + 1013904223u;
    return (float)((*state >> 8) & 0xffffff) / 8388608.0f - 1.0f;
}

/* the loop every reference caller writes: process frame by frame, reconstruct with window-sum accumulation, divide */
static void per_frame_roundtrip(size_t nfft, size_t hop, size_t n)
{
    vv_dsp_stft_params prm;
    vv_dsp_stft* h = NULL;
    vv_dsp_real *x, *recon, *norm;
    vv_dsp_cpx* spec;
    size_t start, i, lo, hi;
    unsigned seed = 12345u;
    double worst = 0.0, acc = 0.0;
    prm.fft_size = nfft; prm.hop_size = hop; prm.window = VV_DSP_STFT_WIN_HANN;
    CHECK(vv_dsp_stft_create(&prm, &h) == VV_DSP_OK);
    if (!h) return;
    x = (vv_dsp_real*)malloc(n * sizeof(*x));
    recon = (vv_dsp_real*)calloc(n + nfft, sizeof(*recon));
    norm = (vv_dsp_real*)calloc(n + nfft, sizeof(*norm));
    spec = (vv_dsp_cpx*)malloc(nfft * sizeof(*spec));
    for (i = 0; i < n; ++i) x[i] = urand(&seed);
    for (start = 0; start + nfft <= n; start += hop) {
        CHECK(vv_dsp_stft_process(h, x + start, spec) == VV_DSP_OK);
        CHECK(vv_dsp_stft_reconstruct(h, spec, recon + start, norm + start) == VV_DSP_OK);
    }
    lo = nfft; hi = n - nfft;
    for (i = lo; i < hi; ++i) {
        const double y = norm[i] > 1e-12f ? (double)recon[i] / (double)norm[i] : 0.0;
        const double e = fabs(y - (double)x[i]);
        if (e > worst) worst = e;
        acc += e * e;
    }
    printf("roundtrip nfft=%zu hop=%zu: max err %.3g rms %.3g over %zu interior samples\n", nfft, hop, worst, sqrt(acc / (double)(hi - lo)), hi - lo);
    CHECK(worst < 1e-3);                                                          /* reference gtest tolerance */
    CHECK(sqrt(acc / (double)(hi - lo)) < 1e-5);
    {   /* zero frame -> zero spectrum, exactly */
        memset(x, 0, nfft * sizeof(*x));
        CHECK(vv_dsp_stft_process(h, x, spec) == VV_DSP_OK);
        for (i = 0; i < nfft; ++i) CHECK(spec[i].re == 0.0f && spec[i].im == 0.0f);
    }
    CHECK(vv_dsp_stft_process(NULL, x, spec) == VV_DSP_ERROR_NULL_POINTER);
    CHECK(vv_dsp_stft_reconstruct(h, spec, recon, NULL) == VV_DSP_OK);           /* norm_add may be NULL (bench_stft.c:96-97) */
    CHECK(vv_dsp_stft_destroy(h) == VV_DSP_OK);
    CHECK(vv_dsp_stft_destroy(NULL) == VV_DSP_ERROR_NULL_POINTER);               /* reference stft.c:63 */
    free(x); free(recon); free(norm); free(spec);
}

static void create_validation(void)
{
    vv_dsp_stft_params prm;
    vv_dsp_stft* h = (vv_dsp_stft*)1;
    prm.fft_size = 512; prm.hop_size = 1024; prm.window = VV_DSP_STFT_WIN_HANN;
    CHECK(vv_dsp_stft_create(&prm, &h) == VV_DSP_ERROR_INVALID_SIZE && h == NULL);  /* hop > fft_size */
    prm.hop_size = 0;
    CHECK(vv_dsp_stft_create(&prm, &h) == VV_DSP_ERROR_INVALID_SIZE);
    CHECK(vv_dsp_stft_create(NULL, &h) == VV_DSP_ERROR_NULL_POINTER);
    prm.hop_size = 128; prm.window = (vv_dsp_stft_window)42;
    CHECK(vv_dsp_stft_create(&prm, &h) == VV_DSP_ERROR_OUT_OF_RANGE && h == NULL);
}

int main(void)
{
    framing_goldens();
    window_formula();
    fft_known_answers();
    create_validation();
    per_frame_roundtrip(512, 128, 512 * 24);         /* the reference gtest's size */
    per_frame_roundtrip(2048, 512, 2048 * 12);
    per_frame_roundtrip(1024, 256, 1024 * 10);
    printf(failures ? "%d check(s) FAILED\n" : "all checks passed (%d failures)\n", failures);
    return failures ? 1 : 0;
}
