/*
 * driver.c -- whole-batch CPU loops over the per-frame STFT API, for timing the
 * CPU baseline and for bulk parity runs.  TEST / BENCH INFRASTRUCTURE ONLY
 * (see the header of vvdsp_oracle.c for who may use oracle/).
 *
 * Compiled twice by oracle/Makefile:
 *   -DDRV_USE_REF : against the UNMODIFIED reference sources under
 *                   /root/reference (exports ref_*; lives in oracle/_ref/)
 *   (default)     : against the restatement vvdsp_oracle.c (exports orcdrv_*)
 *
 * The loops are the reference callers' own loops:
 *   round trip : tools/dump_stft_roundtrip.c:44-54
 *   forward    : process() on x + f*hop for every full frame, then re^2+im^2
 *                of bins 0..nfft/2
 * Signals are partitioned across threads with ONE handle per thread, because a
 * handle owns mutable scratch (src/spectral/stft.c:13-18).
 */
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef DRV_USE_REF
#include "vv_dsp/spectral/stft.h"
typedef vv_dsp_stft drv_handle;
typedef vv_dsp_cpx drv_cpx;
#define DRV(name) ref_##name
static int drv_create(size_t nfft, size_t hop, int win, drv_handle **h)
{
    vv_dsp_stft_params p; p.fft_size = nfft; p.hop_size = hop; p.window = (vv_dsp_stft_window)win;
    return (int)vv_dsp_stft_create(&p, h);
}
#define drv_destroy(h) ((void)vv_dsp_stft_destroy(h))
#define drv_process(h, in, out) ((int)vv_dsp_stft_process(h, in, out))
#define drv_reconstruct(h, in, o, nrm) ((int)vv_dsp_stft_reconstruct(h, in, o, nrm))
#else
typedef struct orc_stft drv_handle;
typedef struct { float re, im; } drv_cpx;
int orc_stft_create(size_t, size_t, int, drv_handle **);
void orc_stft_destroy(drv_handle *);
int orc_stft_process(drv_handle *, const float *, drv_cpx *);
int orc_stft_reconstruct(drv_handle *, const drv_cpx *, float *, float *);
#define DRV(name) orcdrv_##name
#define drv_create orc_stft_create
#define drv_destroy orc_stft_destroy
#define drv_process orc_stft_process
#define drv_reconstruct orc_stft_reconstruct
#endif

typedef struct {
    const float *x; size_t b0, b1, n, pitch, nfft, hop; int win;
    float *y;       /* round trip: [B][pitch] output, or NULL -> thread scratch */
    float *power;   /* forward: [B][frames][nfft/2+1] output, or NULL -> scratch */
    drv_cpx *spec;  /* forward complex: [B][frames][nfft/2+1] or NULL */
    int mode;       /* 0 round trip, 1 forward power, 2 forward complex half */
    int status;
} drv_job;

static void *drv_worker(void *arg)
{
    drv_job *j = (drv_job *)arg;
    const size_t nfft = j->nfft, hop = j->hop, n = j->n, bins = nfft / 2 + 1;
    const size_t frames = n < nfft ? 0 : 1 + (n - nfft) / hop;
    drv_handle *h = NULL;
    j->status = drv_create(nfft, hop, j->win, &h);
    if (j->status) return NULL;
    drv_cpx *spec = (drv_cpx *)malloc(nfft * sizeof(drv_cpx));
    float *recon = (float *)malloc((n ? n : 1) * sizeof(float));
    float *norm = (float *)malloc((n ? n : 1) * sizeof(float));
    float *scratch = (float *)malloc((n > bins ? n : bins) * sizeof(float));
    if (!spec || !recon || !norm || !scratch) { j->status = 4; goto done; }
    for (size_t b = j->b0; b < j->b1; ++b) {
        const float *x = j->x + b * j->pitch;
        if (j->mode == 0) {
            float *y = j->y ? j->y + b * j->pitch : scratch;
            memset(recon, 0, n * sizeof(float));
            memset(norm, 0, n * sizeof(float));
            for (size_t f = 0; f * hop + nfft <= n; ++f) {
                if (drv_process(h, x + f * hop, spec)) { j->status = 4; goto done; }
                if (drv_reconstruct(h, spec, recon + f * hop, norm + f * hop)) { j->status = 4; goto done; }
            }
            for (size_t i = 0; i < n; ++i) y[i] = norm[i] > 1e-12f ? recon[i] / norm[i] : 0.0f;
        } else {
            for (size_t f = 0; f < frames; ++f) {
                if (drv_process(h, x + f * hop, spec)) { j->status = 4; goto done; }
                if (j->mode == 1) {
                    float *p = j->power ? j->power + (b * frames + f) * bins : scratch;
                    for (size_t k = 0; k < bins; ++k) p[k] = spec[k].re * spec[k].re + spec[k].im * spec[k].im;
                } else if (j->spec) {
                    memcpy(j->spec + (b * frames + f) * bins, spec, bins * sizeof(drv_cpx));
                }
            }
        }
    }
done:
    free(spec); free(recon); free(norm); free(scratch);
    drv_destroy(h);
    return NULL;
}

static int drv_run(drv_job proto, size_t batch, int threads)
{
    if (threads < 1) threads = 1;
    if ((size_t)threads > batch) threads = (int)(batch ? batch : 1);
    pthread_t *tid = (pthread_t *)malloc((size_t)threads * sizeof(pthread_t));
    drv_job *jobs = (drv_job *)malloc((size_t)threads * sizeof(drv_job));
    if (!tid || !jobs) { free(tid); free(jobs); return 4; }
    for (int t = 0; t < threads; ++t) {
        jobs[t] = proto;
        jobs[t].b0 = batch * (size_t)t / (size_t)threads;
        jobs[t].b1 = batch * (size_t)(t + 1) / (size_t)threads;
        jobs[t].status = 0;
        if (threads == 1) drv_worker(&jobs[t]);
        else pthread_create(&tid[t], NULL, drv_worker, &jobs[t]);
    }
    int st = 0;
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(tid[t], NULL);
        if (jobs[t].status) st = jobs[t].status;
    }
    free(tid); free(jobs);
    return st;
}

/* STFT -> ISTFT -> normalise for `batch` signals of n samples (row pitch in
 * samples).  y may be NULL (timing only). */
int DRV(batch_roundtrip)(const float *x, size_t batch, size_t n, size_t pitch, size_t nfft, size_t hop,
                         int win, float *y, int threads)
{
    drv_job j; memset(&j, 0, sizeof(j));
    j.x = x; j.n = n; j.pitch = pitch; j.nfft = nfft; j.hop = hop; j.win = win; j.y = y; j.mode = 0;
    return drv_run(j, batch, threads);
}

/* STFT -> |X|^2 over bins 0..nfft/2, valid frames only.  power may be NULL. */
int DRV(batch_power)(const float *x, size_t batch, size_t n, size_t pitch, size_t nfft, size_t hop,
                     int win, float *power, int threads)
{
    drv_job j; memset(&j, 0, sizeof(j));
    j.x = x; j.n = n; j.pitch = pitch; j.nfft = nfft; j.hop = hop; j.win = win; j.power = power; j.mode = 1;
    return drv_run(j, batch, threads);
}

/* STFT -> complex half spectra [batch][frames][nfft/2+1], valid frames only. */
int DRV(batch_forward)(const float *x, size_t batch, size_t n, size_t pitch, size_t nfft, size_t hop,
                       int win, drv_cpx *spec, int threads)
{
    drv_job j; memset(&j, 0, sizeof(j));
    j.x = x; j.n = n; j.pitch = pitch; j.nfft = nfft; j.hop = hop; j.win = win; j.spec = spec; j.mode = 2;
    return drv_run(j, batch, threads);
}
