/*
 * perframe_latency.c -- what a LITERAL drop-in costs: existing vv-dsp callers (reference bench/bench_stft.c:80-98,
 * tools/dump_stft_roundtrip.c:44-47) call vv_dsp_stft_process / vv_dsp_stft_reconstruct once per frame with host
 * pointers.  This program times exactly that loop.  It only uses the reference's public API, so the same source is
 * linked once against libvvdsp_b200.so (every call = H2D + kernel + D2H + sync) and once against the reference
 * compiled for the CPU (oracle/_ref/libvvdsp_ref.so); tests/test_c_callers.py prints both.
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "vv_dsp/vv_dsp_types.h"
#include "vv_dsp/spectral/stft.h"

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char** argv)
{
    const size_t nfft = argc > 1 ? (size_t)atol(argv[1]) : 2048, hop = nfft / 4;
    const size_t frames = argc > 2 ? (size_t)atol(argv[2]) : 400, n = nfft + (frames - 1) * hop;
    vv_dsp_stft_params prm;
    vv_dsp_stft* h = NULL;
    vv_dsp_real *x, *recon, *norm;
    vv_dsp_cpx* spec;
    size_t f, i;
    double t0, t1, t2;
    prm.fft_size = nfft; prm.hop_size = hop; prm.window = VV_DSP_STFT_WIN_HANN;
    if (vv_dsp_stft_create(&prm, &h) != VV_DSP_OK) { printf("{\"error\": \"create failed\"}\n"); return 1; }
    x = (vv_dsp_real*)malloc(n * sizeof(*x));
    recon = (vv_dsp_real*)calloc(n, sizeof(*recon));
    norm = (vv_dsp_real*)calloc(n, sizeof(*norm));
    spec = (vv_dsp_cpx*)malloc(frames * nfft * sizeof(*spec));
    for (i = 0; i < n; ++i) x[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 32768.0f - 1.0f;
    for (f = 0; f < 8; ++f) if (vv_dsp_stft_process(h, x + f * hop, spec) != VV_DSP_OK) return 2;   /* warm-up */
    t0 = now_us();
    for (f = 0; f < frames; ++f)
        if (vv_dsp_stft_process(h, x + f * hop, spec + f * nfft) != VV_DSP_OK) return 2;
    t1 = now_us();
    for (f = 0; f < frames; ++f)
        if (vv_dsp_stft_reconstruct(h, spec + f * nfft, recon + f * hop, norm + f * hop) != VV_DSP_OK) return 3;
    t2 = now_us();
    {
        double worst = 0.0;
        for (i = nfft; i + nfft < n; ++i) {
            const double y = norm[i] > 1e-12f ? recon[i] / norm[i] : 0.0, e = y > x[i] ? y - x[i] : x[i] - y;
            if (e > worst) worst = e;
        }
        printf("{\"fft_size\": %zu, \"hop\": %zu, \"frames\": %zu, \"process_us_per_frame\": %.2f, \"reconstruct_us_per_frame\": %.2f, "
               "\"roundtrip_max_err\": %.3g}\n", nfft, hop, frames, (t1 - t0) / frames, (t2 - t1) / frames, worst);
    }
    (void)vv_dsp_stft_destroy(h);
    free(x); free(recon); free(norm); free(spec);
    return 0;
}
