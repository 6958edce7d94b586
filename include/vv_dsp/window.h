/*
 * vv_dsp/window.h -- the three window generators reachable from vv_dsp_stft_window.
 *
 * Same declarations as the reference's include/vv_dsp/window.h:51,66,81.  Host C,
 * float32, the reference's exact formula (src/window/window.c:16-49): symmetric
 * (denominator N-1), cosf in float, N == 1 -> 1.0; so tables are bit-identical.
 * The other 11 reference windows are out of scope (not selectable by an STFT handle).
 */
#ifndef VV_DSP_WINDOW_H
#define VV_DSP_WINDOW_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

vv_dsp_status vv_dsp_window_boxcar(size_t N, vv_dsp_real* out);
vv_dsp_status vv_dsp_window_hann(size_t N, vv_dsp_real* out);
vv_dsp_status vv_dsp_window_hamming(size_t N, vv_dsp_real* out);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_WINDOW_H */
