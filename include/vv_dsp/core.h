/*
 * vv_dsp/core.h -- framing helpers of the STFT path (host C, index contract only).
 * Same declarations as the reference's include/vv_dsp/core.h:475-529; behaviour per
 * src/core/framing.c:58-148 (see vv_dsp_b200/csrc/host/framing.c).
 */
#ifndef VV_DSP_CORE_H
#define VV_DSP_CORE_H

#include <stddef.h>
#include "vv_dsp/vv_dsp_types.h"
#include "vv_dsp/core/vv_dsp_vectorized_math.h"

#ifdef __cplusplus
extern "C" {
#endif

/* center == 0: n < frame_len ? 0 : 1 + (n - frame_len)/hop;  center != 0: ceil(n/hop);  hop == 0 -> 0 */
size_t vv_dsp_get_num_frames(size_t signal_len, size_t frame_len, size_t hop_len, int center);

/* center == 0: start = index*hop, zeros outside the signal;  center != 0: start = index*hop - frame_len/2,
 * edge-inclusive reflection (-1 -> 0, -2 -> 1, n -> n-1); optional multiply by window[i]. */
vv_dsp_status vv_dsp_fetch_frame(const vv_dsp_real* signal, size_t signal_len, vv_dsp_real* frame_buffer,
                                 size_t frame_len, size_t hop_len, size_t frame_index, int center,
                                 const vv_dsp_real* window);

/* output[index*hop + i] += frame[i] for positions < output_len (overflow silently dropped) */
vv_dsp_status vv_dsp_overlap_add(const vv_dsp_real* frame, vv_dsp_real* output_signal, size_t output_len,
                                 size_t frame_len, size_t hop_len, size_t frame_index);

#ifdef __cplusplus
}
#endif

#endif /* VV_DSP_CORE_H */
