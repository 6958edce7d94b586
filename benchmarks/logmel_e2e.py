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
with Stft(nfft, hop, "hann") as h:
    h.batch_logmel(xn, w, 1e-10, out=on)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        h.batch_logmel(xn, w, 1e-10, out=on)
    dt = (time.perf_counter() - t0) / 4
print(json.dumps({"workload": f"host signals -> log-mel in host memory, {B} x {n} samples, nfft={nfft} hop={hop}, {n_mels} mels",
                  "stage_target_bytes": os.environ.get("VVB_STAGE_TARGET_BYTES", "default (96 MB)"), "ms": dt * 1e3,
                  "Msamples_per_s": B * n / dt / 1e6, "h2d_GB": B * n * 4 / 1e9, "d2h_GB": B * F * n_mels * 4 / 1e9}))
