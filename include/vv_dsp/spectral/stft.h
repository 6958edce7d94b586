/*
 * vv_dsp/spectral/stft.h -- STFT handle API of the drop-in boundary.
 *
 * Same declarations as the reference's include/vv_dsp/spectral/stft.h:12-56,
 * implemented by vv_dsp_b200/csrc/host/stft.c.  All data pointers are caller-owned
 * HOST memory, every call is synchronous (returns after the device->host copy).
 *
 *  create       fft_size == 0, hop_size == 0 or hop > fft_size -> INVALID_SIZE; bad
 *               window enum -> OUT_OF_RANGE; *out is NULLed first (reference stft.c:30-60)
 *  process      out[k], k < fft_size = unscaled forward DFT of in[i]*w[i]
 *               (reference stft.c:74-92).  The device computes bins 0..fft_size/2 with
 *               a real-input FFT; bins above are the conjugate mirror.
 *  reconstruct  v = Re(backward DFT(in) / fft_size) * w;  out_add[i] += v[i];
 *               norm_add[i] += w[i]^2 when non-NULL (reference stft.c:95-110).
 *  spectrogram  frames = n < fft ? 1 : 1 + (n - fft + hop)/hop, trailing zero pad,
 *               magnitude of all fft_size bins, row-major (reference stft.c:112-144).
 *
 * A handle is not safe for concurrent calls (same rule as the reference, which keeps
 * scratch in the handle, stft.c:13-18); distinct handles are independent.
 */
#ifndef VV_DSP_SPECTRAL_STFT_H
#define VV_DSP_SPECTRAL_STFT_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"
#include "vv_dsp/spectral/fft.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vv_dsp_stft vv_dsp_stft;

typedef enum vv_dsp_stft_window {
    VV_DSP_STFT_WIN_BOXCAR = 0,
    VV_DSP_STFT_WIN_HANN = 1,
    VV_DSP_STFT_WIN_HAMMING = 2
} vv_dsp_stft_window;

typedef struct vv_dsp_stft_params {
    size_t fft_size;
    size_t hop_size;
    vv_dsp_stft_window window;
} vv_dsp_stft_params;

VV_DSP_NODISCARD vv_dsp_status vv_dsp_stft_create(const vv_dsp_stft_params* params, vv_dsp_stft** out);
vv_dsp_status vv_dsp_stft_destroy(vv_dsp_stft* h);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_stft_process(vv_dsp_stft* h, const vv_dsp_real* in, vv_dsp_cpx* out);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_stft_reconstruct(vv_dsp_stft* h, const vv_dsp_cpx* in,
                                                       vv_dsp_real* out_add, vv_dsp_real* norm_add);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_stft_spectrogram(vv_dsp_stft* h, const vv_dsp_real* signal, size_t n,
                                                       vv_dsp_real* out_mag, size_t* out_frames);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_SPECTRAL_STFT_H */
