/*
 * window.c -- boxcar / Hann / Hamming tables for the STFT handle (host, C99).
 *
 * Follows the reference's definition exactly (src/window/window.c:16-49 with
 * include/vv_dsp/vv_dsp_math.h:21-27): symmetric windows, denominator N-1, the
 * step 2*pi/(N-1) formed in float32 and the cosine taken with cosf, so the table
 * is bit-identical to the reference's on the same libm.  N == 1 yields 1.0;
 * out == NULL -> NULL_POINTER, N == 0 -> INVALID_SIZE (window.c:9-14).
 */
#include <math.h>
#include "vv_dsp/window.h"

static vv_dsp_status check_args(size_t count, const vv_dsp_real* out)
{
    if (out == NULL) return VV_DSP_ERROR_NULL_POINTER;
    if (count == 0) return VV_DSP_ERROR_INVALID_SIZE;
    return VV_DSP_OK;
}

/* generalised cosine window a0 - a1*cos(2 pi n/(N-1)) */
static vv_dsp_status raised_cosine(size_t count, vv_dsp_real* out, float a0, float a1)
{
    vv_dsp_status st = check_args(count, out);
    if (st != VV_DSP_OK) return st;
    if (count == 1) {
        out[0] = 1.0f;
        return VV_DSP_OK;
    }
    {
        const float two_pi = (float)(2.0 * 3.141592653589793238462643383279502884);
        const float step = two_pi / (float)(count - 1);
        size_t i;
        for (i = 0; i < count; ++i) out[i] = a0 - a1 * cosf(step * (float)i);
    }
    return VV_DSP_OK;
}

vv_dsp_status vv_dsp_window_boxcar(size_t N, vv_dsp_real* out)
{
    size_t i;
    vv_dsp_status st = check_args(N, out);
    if (st != VV_DSP_OK) return st;
    for (i = 0; i < N; ++i) out[i] = 1.0f;
    return VV_DSP_OK;
}

vv_dsp_status vv_dsp_window_hann(size_t N, vv_dsp_real* out) { return raised_cosine(N, out, 0.5f, 0.5f); }

vv_dsp_status vv_dsp_window_hamming(size_t N, vv_dsp_real* out) { return raised_cosine(N, out, 0.54f, 0.46f); }
