/*
 * framing.c -- the index / padding contract of the STFT path on the host (C99):
 * frame counts, single-frame gather, single-frame overlap-add, window multiply.
 * Behaviour per the reference's src/core/framing.c:58-148 and
 * src/core/vv_dsp_vectorized_math_fallback.c:13-29.  These are the per-frame helpers
 * existing callers use (bench/bench_stft.c:86-92); the batched GPU entry points
 * implement the same index rules inside the kernels (csrc/cuda/vvb_stft_kernels.cuh).
 */
#include "vv_dsp/core.h"

size_t vv_dsp_get_num_frames(size_t signal_len, size_t frame_len, size_t hop_len, int center)
{
    if (hop_len == 0) return 0;
    if (center) return (signal_len + hop_len - 1) / hop_len;
    return signal_len < frame_len ? 0 : 1 + (signal_len - frame_len) / hop_len;
}

/* ... 1 0 | 0 1 ... n-1 | n-1 n-2 ...  (edge samples repeated), any distance outside */
static size_t mirror_into(long long idx, long long n)
{
    const long long period = 2 * n;
    long long m = idx % period;
    if (m < 0) m += period;
    return (size_t)(m < n ? m : period - 1 - m);
}

vv_dsp_status vv_dsp_fetch_frame(const vv_dsp_real* signal, size_t signal_len, vv_dsp_real* frame_buffer,
                                 size_t frame_len, size_t hop_len, size_t frame_index, int center,
                                 const vv_dsp_real* window)
{
    long long first;
    size_t i;
    if (!signal || !frame_buffer) return VV_DSP_ERROR_NULL_POINTER;
    if (signal_len == 0 || frame_len == 0 || hop_len == 0) return VV_DSP_ERROR_INVALID_SIZE;
    first = (long long)(frame_index * hop_len);
    if (center) first -= (long long)(frame_len / 2);
    for (i = 0; i < frame_len; ++i) {
        const long long pos = first + (long long)i;
        vv_dsp_real v;
        if (center) v = signal[mirror_into(pos, (long long)signal_len)];
        else v = (pos >= 0 && pos < (long long)signal_len) ? signal[pos] : 0.0f;
        frame_buffer[i] = window ? v * window[i] : v;
    }
    return VV_DSP_OK;
}

vv_dsp_status vv_dsp_overlap_add(const vv_dsp_real* frame, vv_dsp_real* output_signal, size_t output_len,
                                 size_t frame_len, size_t hop_len, size_t frame_index)
{
    size_t i, base;
    if (!frame || !output_signal) return VV_DSP_ERROR_NULL_POINTER;
    if (output_len == 0 || frame_len == 0 || hop_len == 0) return VV_DSP_ERROR_INVALID_SIZE;
    base = frame_index * hop_len;
    for (i = 0; i < frame_len && base + i < output_len; ++i) output_signal[base + i] += frame[i];
    return VV_DSP_OK;
}

vv_dsp_status vv_dsp_vectorized_window_apply(const vv_dsp_real* in, const vv_dsp_real* window, vv_dsp_real* out, size_t n)
{
    size_t i;
    if (!in || !window || !out || n == 0) return VV_DSP_ERROR_NULL_POINTER;
    for (i = 0; i < n; ++i) out[i] = in[i] * window[i];
    return VV_DSP_OK;
}
