/*
 * vv_dsp/vv_dsp_math.h -- math constants and type-matched libm wrappers of the drop-in boundary.
 *
 * Same macro names and values as the reference's include/vv_dsp/vv_dsp_math.h:11-49 (PI / TWO_PI in double and in
 * vv_dsp_real, VV_DSP_SIN ... VV_DSP_ATAN2 mapped to the float libm functions), so callers written against the
 * reference (e.g. its tests/window_tests.c:17,26) compile unchanged.  vv_dsp_real is float in this library
 * (vv_dsp_types.h); the reference's optional double build and its C++ math_approx overrides are out of scope.
 */
#ifndef VV_DSP_MATH_H
#define VV_DSP_MATH_H

#include <math.h>
#include "vv_dsp/vv_dsp_types.h"

#if defined(VV_DSP_USE_DOUBLE)
#error "vv-dsp_b200 is a float32 library: VV_DSP_USE_DOUBLE is not supported"
#endif

#ifndef VV_DSP_PI_D
#define VV_DSP_PI_D 3.141592653589793238462643383279502884
#endif
#ifndef VV_DSP_PI
#define VV_DSP_PI ((vv_dsp_real)VV_DSP_PI_D)
#endif
#ifndef VV_DSP_TWO_PI_D
#define VV_DSP_TWO_PI_D (2.0 * VV_DSP_PI_D)
#endif
#ifndef VV_DSP_TWO_PI
#define VV_DSP_TWO_PI ((vv_dsp_real)VV_DSP_TWO_PI_D)
#endif

#define VV_DSP_SIN(x) sinf(x)
#define VV_DSP_COS(x) cosf(x)
#define VV_DSP_EXP(x) expf(x)
#define VV_DSP_SQRT(x) sqrtf(x)
#define VV_DSP_LOG(x) logf(x)
#define VV_DSP_TAN(x) tanf(x)
#define VV_DSP_POW(x, y) powf(x, y)
#define VV_DSP_ATAN2(y, x) atan2f(y, x)

#endif /* VV_DSP_MATH_H */
