/*
 * vv_dsp/vv_dsp_types.h -- ABI types of the drop-in boundary (vv-dsp_b200).
 *
 * Replaces, with identical names, values and layouts, the reference's
 * include/vv_dsp/vv_dsp_types.h:
 *   vv_dsp_real   (float unless VV_DSP_USE_DOUBLE)        reference :70-74
 *   vv_dsp_cpx    {re, im}, interleaved, sizeof == 2*real  reference :88-91,150
 *   vv_dsp_status 0..6                                     reference :120-128
 * The B200 library is float32 only: building it with VV_DSP_USE_DOUBLE is an error.
 * Unlike the reference header (reference :143-150) this one also compiles as C11+.
 */
#ifndef VV_DSP_TYPES_H
#define VV_DSP_TYPES_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef VV_DSP_INLINE
#  if defined(_MSC_VER)
#    define VV_DSP_INLINE __inline
#  else
#    define VV_DSP_INLINE inline
#  endif
#endif

#ifndef VV_DSP_NODISCARD
#  if defined(__GNUC__) || defined(__clang__)
#    define VV_DSP_NODISCARD __attribute__((warn_unused_result))
#  else
#    define VV_DSP_NODISCARD
#  endif
#endif

#ifdef VV_DSP_USE_DOUBLE
#  error "vv-dsp_b200 computes in float32 only (vv_dsp_real = float)"
#endif
typedef float vv_dsp_real;

typedef struct vv_dsp_cpx {
    vv_dsp_real re;
    vv_dsp_real im;
} vv_dsp_cpx;

static VV_DSP_INLINE vv_dsp_cpx vv_dsp_cpx_make(vv_dsp_real re, vv_dsp_real im)
{
    vv_dsp_cpx z;
    z.re = re;
    z.im = im;
    return z;
}

typedef enum vv_dsp_status {
    VV_DSP_OK = 0,
    VV_DSP_ERROR_NULL_POINTER = 1,
    VV_DSP_ERROR_INVALID_SIZE = 2,
    VV_DSP_ERROR_OUT_OF_RANGE = 3,
    VV_DSP_ERROR_INTERNAL = 4,     /* also: any CUDA runtime failure */
    VV_DSP_ERROR_NAN_INF = 5,
    VV_DSP_ERROR_UNSUPPORTED = 6   /* also: no CUDA device; backend not compiled in */
} vv_dsp_status;

/* typedef-array trick in every language mode, so C99, C11+ and C++ all accept it */
#define VV_DSP_STATIC_ASSERT(cond, msg) typedef char vv_dsp_static_assert_##msg[(cond) ? 1 : -1]
VV_DSP_STATIC_ASSERT(sizeof(vv_dsp_cpx) == sizeof(vv_dsp_real) * 2, cpx_size_must_be_2x_real);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_TYPES_H */
