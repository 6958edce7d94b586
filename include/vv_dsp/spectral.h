/* vv_dsp/spectral.h -- umbrella for the spectral entry points this library provides
 * (reference include/vv_dsp/spectral.h; DCT/CZT/Hilbert/utils are out of scope). */
#ifndef VV_DSP_SPECTRAL_H
#define VV_DSP_SPECTRAL_H
#include "vv_dsp/spectral/fft.h"
#include "vv_dsp/spectral/stft.h"
#endif
