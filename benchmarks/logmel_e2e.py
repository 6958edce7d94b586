#!/usr/bin/env python
"""Host signals -> log-mel rows in host memory through vv_dsp_stft_batch_logmel (fused kernel, fft_size 2048 / hop 512 / 80 bands),
pinned buffers, wall clock: chunks of VVB_STAGE_TARGET_BYTES flow upload -> kernel -> download on two streams.  One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft, mel_filterbank  # noqa: E402

B, n, nfft, hop, n_mels = 1024, 480_000, 2048, 512, 80
F = 1 + (n - nfft) // hop
x = torch.empty((B, n), dtype=torch.float32).pin_memory()
x.uniform_(-1, 1)
out = torch.empty((B, F, n_mels), dtype=torch.float32).pin_memory()
st, w = mel_filterbank(nfft, n_mels, 48000.0, 0.0, 24000.0)
xn, on = x.numpy(), out.numpy()
def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts


x16 = torch.empty((B, n), dtype=torch.int16).pin_memory()
x16.copy_((x * 32767.0).round().to(torch.int16))
x16n = x16.numpy()
with Stft(nfft, hop, "hann") as h:
    t_f = timed(lambda: h.batch_logmel(xn, w, 1e-10, out=on))
    t_16 = timed(lambda: h.batch_logmel_pcm(x16n, w, 16, 1e-10, out=on))
    t_f2 = timed(lambda: h.batch_logmel(xn, w, 1e-10, out=on))
for name, ts, bps in (("host float32 signals -> log-mel in host memory (vv_dsp_stft_batch_logmel)", t_f, 4),
                      ("host 16-bit PCM -> log-mel in host memory (vv_dsp_stft_batch_logmel_pcm)", t_16, 2),
                      ("host float32 signals again, after the PCM calls", t_f2, 4)):
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"workload": f"{name}, {B} x {n} samples, nfft={nfft} hop={hop}, {n_mels} mels",
                      "stage_target_bytes": os.environ.get("VVB_STAGE_TARGET_BYTES", "default (96 MB)"), "ms_median": round(ms, 3),
                      "ms_each": [round(t, 2) for t in ts], "Msamples_per_s": round(B * n / ms / 1e3, 1),
                      "h2d_GB": B * n * bps / 1e9, "d2h_GB": B * F * n_mels * 4 / 1e9}))
