#!/usr/bin/env python
"""STFT -> log-mel at the sizes of the generic forward kernel (fft_size 256 ... 1024 and the speech framings): the one-kernel path
(band sums of a CTA's frames in shared memory, mel_phase_cta) against the chained power + log-mel kernels (VVB_MEL_UNFUSED=1) and
the power kernel alone; device-resident signals, CUDA events on the library's stream.  One JSON line per shape."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft, mel_filterbank  # noqa: E402

dev = torch.device("cuda:0")
s = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B = 1024
SHAPES = ((400, 160, 16000.0, 160_000, 80), (512, 160, 16000.0, 160_000, 80), (512, 128, 16000.0, 160_000, 40),
          (640, 160, 16000.0, 160_000, 80), (256, 64, 8000.0, 160_000, 26), (1024, 256, 48000.0, 480_000, 80))
only = [int(v) for v in os.environ.get("LOGMEL_SMALL_NFFT", "").split(",") if v]        # e.g. "400" for an ncu capture
fused_only = bool(os.environ.get("LOGMEL_SMALL_FUSED_ONLY"))                              # 3 warm-up launches + 1: ncu --launch-skip 3
for nfft, hop, sr, n, n_mels in SHAPES:
    if only and nfft not in only:
        continue
    x = torch.rand((B, n), device=dev) * 2 - 1
    st, w = mel_filterbank(nfft, n_mels, sr, 0.0, sr / 2)
    F = 1 + (n - nfft) // hop
    out = torch.empty((B, F, n_mels), device=dev)
    power = torch.empty((B, F, nfft // 2 + 1), device=dev)
    with Stft(nfft, hop, "hann") as h:
        h.set_stream(s.cuda_stream)
        ms_fused = timed(lambda: h.batch_logmel(x, w, 1e-10, out=out), 1 if fused_only else 10)
        if fused_only:
            print(json.dumps({"nfft": nfft, "hop": hop, "fused_ms": ms_fused}))
            continue
        a = out.clone()
        os.environ["VVB_MEL_UNFUSED"] = "1"
        try:
            ms_chain = timed(lambda: h.batch_logmel(x, w, 1e-10, out=out))
        finally:
            del os.environ["VVB_MEL_UNFUSED"]
        same = bool(torch.equal(a, out))
        ms_fused2 = timed(lambda: h.batch_logmel(x, w, 1e-10, out=out))
        ms_pow = timed(lambda: h.batch_forward(x, out=power, kind="power"))
    os.environ["VVB_MEL_UNIT4"] = "1"          # read when a handle builds its lane schedules: four quads per segment as in the marching kernel
    try:
        with Stft(nfft, hop, "hann") as h4:
            h4.set_stream(s.cuda_stream)
            ms_unit4 = timed(lambda: h4.batch_logmel(x, w, 1e-10, out=out))
            same = same and bool(torch.equal(a, out))
    finally:
        del os.environ["VVB_MEL_UNIT4"]
    alg = (B * n * 4 + B * F * n_mels * 4) / 1e9
    print(json.dumps({"workload": f"STFT->log-mel, {B} x {n} samples, nfft={nfft} hop={hop}, {n_mels} mels", "fused_ms": round(ms_fused, 4),
                      "fused_again_ms": round(ms_fused2, 4), "fused_unit4_ms": round(ms_unit4, 4), "chained_ms": round(ms_chain, 4), "power_kernel_ms": round(ms_pow, 4), "bit_identical": same,
                      "Msamples_per_s": round(B * n / ms_fused / 1e3, 1), "algorithmic_GB": round(alg, 4), "GBps": round(alg / ms_fused * 1e3, 1)}))
    del x, out, power, a
