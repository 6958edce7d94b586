/*
 * vv_dsp/features/mel.h -- the mel front end that consumes the STFT power output
 * (SURVEY.md section 8f, rank 2).  Same declarations as the reference's
 * include/vv_dsp/features/mel.h:12-168.
 *
 *  hz_to_mel / mel_to_hz     HTK scale in float32: 2595 log10f(1 + hz/700)   (src/features/mel.c:14-29)
 *  mel_filterbank_create     dense [n_mels][n_fft/2+1] triangular filters, each divided by its
 *                            sum; HTK only (SLANEY -> OUT_OF_RANGE like the reference);
 *                            host code, bit-identical to the reference     (mel.c:66-185)
 *  compute_log_mel_spectrogram  out[f][m] = logf(sum_k power[f][k] W[m][k] + eps), host
 *                            pointers, synchronous; computed on the GPU    (mel.c:204-245)
 *  mfcc                      unnormalised DCT-II of each log-mel frame (the reference's naive
 *                            src/spectral/dct.c:21-30), first num_mfcc_coeffs kept, then
 *                            c[i] *= 1 + (L/2) sinf(pi i / L) for i >= 1 when L > 0; only
 *                            VV_DSP_DCT_II is accepted; computed on the GPU (mel.c:249-310)
 *  mfcc_init / process / destroy  plan = filterbank + tables resident on the device;
 *                            process = power -> log-mel -> MFCC, host pointers (mel.c:333-461)
 */
#ifndef VV_DSP_FEATURES_MEL_H
#define VV_DSP_FEATURES_MEL_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"
#include "vv_dsp/spectral/dct.h"

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

VV_DSP_NODISCARD vv_dsp_status vv_dsp_mfcc(const vv_dsp_real* log_mel_spectrogram, size_t num_frames, size_t n_mels,
                                           size_t num_mfcc_coeffs, vv_dsp_dct_type dct_type, vv_dsp_real lifter_coeff,
                                           vv_dsp_real* out_mfcc_coeffs);

typedef struct vv_dsp_mfcc_plan vv_dsp_mfcc_plan;

VV_DSP_NODISCARD vv_dsp_status vv_dsp_mfcc_init(size_t n_fft, size_t n_mels, size_t num_mfcc_coeffs, vv_dsp_real sample_rate,
                                                vv_dsp_real fmin, vv_dsp_real fmax, vv_dsp_mel_variant variant,
                                                vv_dsp_dct_type dct_type, vv_dsp_real lifter_coeff, vv_dsp_real log_epsilon,
                                                vv_dsp_mfcc_plan** out_plan);
VV_DSP_NODISCARD vv_dsp_status vv_dsp_mfcc_process(const vv_dsp_mfcc_plan* plan, const vv_dsp_real* power_spectrogram,
                                                   size_t num_frames, vv_dsp_real* out_mfcc_coeffs);
vv_dsp_status vv_dsp_mfcc_destroy(vv_dsp_mfcc_plan* plan);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_FEATURES_MEL_H */
