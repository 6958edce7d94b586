/*
 * vv_dsp/features/mel.h -- the mel front end that consumes the STFT power output
 * (SURVEY.md section 8f, rank 2).  Same declarations as the reference's
 * include/vv_dsp/features/mel.h:12-66 for the entry points listed here; the MFCC / DCT part
 * of that header is out of scope.
 *
 *  hz_to_mel / mel_to_hz     HTK scale in float32: 2595 log10f(1 + hz/700)   (src/features/mel.c:14-29)
 *  mel_filterbank_create     dense [n_mels][n_fft/2+1] triangular filters, each divided by its
 *                            sum; HTK only (SLANEY -> OUT_OF_RANGE like the reference);
 *                            host code, bit-identical to the reference     (mel.c:66-185)
 *  compute_log_mel_spectrogram  out[f][m] = logf(sum_k power[f][k] W[m][k] + eps), host
 *                            pointers, synchronous; computed on the GPU    (mel.c:204-245)
 */
#ifndef VV_DSP_FEATURES_MEL_H
#define VV_DSP_FEATURES_MEL_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vv_dsp_mel_variant {
    VV_DSP_MEL_VARIANT_HTK = 0,
    VV_DSP_MEL_VARIANT_SLANEY = 1
} vv_dsp_mel_variant;

vv_dsp_real vv_dsp_hz_to_mel(vv_dsp_real hz);
vv_dsp_real vv_dsp_mel_to_hz(vv_dsp_real mel);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_mel_filterbank_create(size_t n_fft, size_t n_mels, vv_dsp_real sample_rate,
                                                            vv_dsp_real fmin, vv_dsp_real fmax, vv_dsp_mel_variant variant,
                                                            vv_dsp_real** out_filterbank_weights, size_t* out_num_filters,
                                                            size_t* out_filter_len);
void vv_dsp_mel_filterbank_free(vv_dsp_real* filterbank_weights, size_t n_mels);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_compute_log_mel_spectrogram(const vv_dsp_real* power_spectrogram, size_t num_frames,
                                                                  size_t n_fft_bins, const vv_dsp_real* filterbank_weights,
                                                                  size_t n_mels, vv_dsp_real log_epsilon,
                                                                  vv_dsp_real* out_log_mel_spectrogram);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_FEATURES_MEL_H */
