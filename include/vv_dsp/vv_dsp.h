/* vv_dsp/vv_dsp.h -- umbrella (reference include/vv_dsp/vv_dsp.h), hot-path subset + B200 extension. */
#ifndef VV_DSP_H
#define VV_DSP_H
#include "vv_dsp/vv_dsp_types.h"
#include "vv_dsp/vv_dsp_math.h"
#include "vv_dsp/core.h"
#include "vv_dsp/window.h"
#include "vv_dsp/spectral.h"
#include "vv_dsp/features/mel.h"
#include "vv_dsp/b200.h"
#endif
