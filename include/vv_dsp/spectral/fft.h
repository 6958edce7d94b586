/*
 * vv_dsp/spectral/fft.h -- FFT plan API of the drop-in boundary.
 *
 * Same declarations as the reference's include/vv_dsp/spectral/fft.h
 * (enums :34-38,57-61,143-156; functions :71-124,190-252), implemented by
 * vv_dsp_b200/csrc/host/fft.c on top of the CUDA engine.
 *
 * Conventions kept from the reference (include/vv_dsp/spectral/fft.h:174-176,
 * src/spectral/fft_kiss.c:45,69-73,84): forward = sum x[t] exp(-j 2 pi k t / n),
 * unscaled; backward is scaled by 1/n.  C2C: cpx[n] -> cpx[n] (in may alias out);
 * R2C: real[n] -> cpx[n/2+1] with the Nyquist imaginary part forced to 0 for even
 * n; C2R: cpx[n/2+1] -> real[n].  Any n >= 1 is accepted: powers of two 128..8192
 * run the register Stockham kernels, powers of two up to 2^26 four-step plans built
 * on them, every other n in [32, 4096] a fused chirp-z (Bluestein) kernel, other n up
 * to 2^22 the same chirp-z transform on four-step plans, and the rest (n < 32,
 * anything larger) a direct-DFT kernel like the reference's own O(n^2) path
 * (src/spectral/fft_kiss.c:76-92,115); nothing ever runs on the CPU.
 *
 * Backend ids: the B200 engine answers to id 0 (VV_DSP_FFT_BACKEND_KISS, the
 * default every caller uses).  FFTW / FFTS are "not compiled in": selecting them
 * returns VV_DSP_ERROR_UNSUPPORTED exactly as the reference does in that case
 * (src/spectral/fft.c:21-23).
 */
#ifndef VV_DSP_SPECTRAL_FFT_H
#define VV_DSP_SPECTRAL_FFT_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vv_dsp_fft_backend {
    VV_DSP_FFT_BACKEND_KISS = 0,
    VV_DSP_FFT_BACKEND_FFTW = 1,
    VV_DSP_FFT_BACKEND_FFTS = 2
} vv_dsp_fft_backend;

typedef enum vv_dsp_fftw_flag {
    VV_DSP_FFTW_ESTIMATE = 0,
    VV_DSP_FFTW_MEASURE = 1,
    VV_DSP_FFTW_PATIENT = 2
} vv_dsp_fftw_flag;

typedef enum vv_dsp_fft_dir {
    VV_DSP_FFT_FORWARD = +1,
    VV_DSP_FFT_BACKWARD = -1
} vv_dsp_fft_dir;

typedef enum vv_dsp_fft_type {
    VV_DSP_FFT_C2C = 0,
    VV_DSP_FFT_R2C = 1,
    VV_DSP_FFT_C2R = 2
} vv_dsp_fft_type;

typedef struct vv_dsp_fft_plan vv_dsp_fft_plan;

VV_DSP_NODISCARD vv_dsp_status vv_dsp_fft_set_backend(vv_dsp_fft_backend backend);
vv_dsp_fft_backend vv_dsp_fft_get_backend(void);
int vv_dsp_fft_is_backend_available(vv_dsp_fft_backend backend);
VV_DSP_NODISCARD vv_dsp_status vv_dsp_fft_set_fftw_flag(vv_dsp_fftw_flag flag);
VV_DSP_NODISCARD vv_dsp_status vv_dsp_fft_flush_fftw_cache(void);

VV_DSP_NODISCARD vv_dsp_status vv_dsp_fft_make_plan(size_t n, vv_dsp_fft_type type, vv_dsp_fft_dir dir,
                                                    vv_dsp_fft_plan** out_plan);
VV_DSP_NODISCARD vv_dsp_status vv_dsp_fft_execute(const vv_dsp_fft_plan* plan, const void* in, void* out);
vv_dsp_status vv_dsp_fft_destroy(vv_dsp_fft_plan* plan);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_SPECTRAL_FFT_H */
