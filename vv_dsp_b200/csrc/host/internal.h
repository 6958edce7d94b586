/* internal.h -- helpers shared between the C99 host translation units (not part of the ABI) */
#ifndef VVDSP_B200_HOST_INTERNAL_H
#define VVDSP_B200_HOST_INTERNAL_H
#include <stddef.h>

/* device-resident sparse mel filterbank: meta = lo | len | off per band, packed non-zero weights */
/* d_fw / d_fseg (may be NULL): lane schedules of the fused STFT -> log-mel kernel (f_segments segments per lane, power row of
 * f_prow floats, f_unit quads per segment), see mel.c build_fused_tables */
typedef struct mel_device { int* d_meta; float* d_w; size_t n_groups; float* d_fw; int* d_fseg; size_t f_segments, f_prow, f_unit; } mel_device;
int vvdsp_internal_mel_device_build(const float* dense_weights, size_t n_mels, size_t bins, void* stream, mel_device* md);
void vvdsp_internal_mel_device_free(mel_device* md);
/* log-mel of densely packed power rows on the device (the four-tap group kernels) */
int vvdsp_internal_logmel(const mel_device* md, const float* d_power, size_t frames, size_t bins, size_t n_mels, float eps,
                          float* d_out, void* stream);

/* MFCC tables on the device: cosine table [n_coeffs][n_mels] and lifter factors [n_coeffs] */
typedef struct mfcc_device { float* d_table; float* d_lifter; size_t n_mels, n_coeffs; float lifter; } mfcc_device;
int vvdsp_internal_mfcc_device_build(size_t n_mels, size_t n_coeffs, float lifter, void* stream, mfcc_device* md);
void vvdsp_internal_mfcc_device_free(mfcc_device* md);

#endif
