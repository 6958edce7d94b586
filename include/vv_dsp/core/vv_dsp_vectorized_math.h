/*
 * vv_dsp/core/vv_dsp_vectorized_math.h -- only the entry the STFT path calls.
 * Reference: include/vv_dsp/core/vv_dsp_vectorized_math.h:38-43,
 * src/core/vv_dsp_vectorized_math_fallback.c:13-29 (NULL or n == 0 -> NULL_POINTER).
 */
#ifndef VV_DSP_VECTORIZED_MATH_H
#define VV_DSP_VECTORIZED_MATH_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

VV_DSP_NODISCARD vv_dsp_status vv_dsp_vectorized_window_apply(const vv_dsp_real* in, const vv_dsp_real* window,
                                                              vv_dsp_real* out, size_t n);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_VECTORIZED_MATH_H */
