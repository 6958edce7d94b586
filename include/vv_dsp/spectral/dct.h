/*
 * vv_dsp/spectral/dct.h -- only the enums of the reference's include/vv_dsp/spectral/dct.h:12-23, which the
 * MFCC entry points of vv_dsp/features/mel.h take as arguments.  The DCT plan API itself
 * (vv_dsp_dct_make_plan / execute / forward / inverse, src/spectral/dct.c) is out of scope (SURVEY.md
 * section 2 row 11); the DCT-II that MFCC needs runs inside the MFCC kernel.
 */
#ifndef VV_DSP_SPECTRAL_DCT_H
#define VV_DSP_SPECTRAL_DCT_H

typedef enum vv_dsp_dct_type {
    VV_DSP_DCT_II = 2,
    VV_DSP_DCT_III = 3,
    VV_DSP_DCT_IV = 4
} vv_dsp_dct_type;

typedef enum vv_dsp_dct_dir {
    VV_DSP_DCT_FORWARD = +1,
    VV_DSP_DCT_BACKWARD = -1
} vv_dsp_dct_dir;

#endif /* VV_DSP_SPECTRAL_DCT_H */
