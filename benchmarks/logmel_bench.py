#!/usr/bin/env python
"""STFT -> power -> mel -> log at the headline shape (1024 x 10 s @ 48 kHz, nfft=2048 hop=512, 80 mels),
device-resident, CUDA-event timed.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft, mel_filterbank  # noqa: E402

B, n, nfft, hop, n_mels = 1024, 480_000, 2048, 512, 80
dev = torch.device("cuda", 0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
x = torch.rand((B, n), device=dev) * 2 - 1
st, w = mel_filterbank(nfft, n_mels, 48000.0, 0.0, 24000.0)
F = 1 + (n - nfft) // hop
out = torch.empty((B, F, n_mels), device=dev)
with Stft(nfft, hop, "hann") as h:
    h.set_stream(s.cuda_stream)
    for _ in range(2):
        h.batch_logmel(x, w, 1e-10, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5):
        h.batch_logmel(x, w, 1e-10, out=out)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    pw = torch.empty((B, F, nfft // 2 + 1), device=dev)
    e0.record(s)
    for _ in range(5):
        h.batch_forward(x, "power", "valid", out=pw)
    e1.record(s); torch.cuda.synchronize()
    ms_pow = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": f"STFT->log-mel, {B} x {n} samples, nfft={nfft} hop={hop}, {n_mels} mels", "ms": ms,
                  "Msamples_per_s": B * n / ms / 1e3, "power_kernel_ms": ms_pow, "logmel_and_overhead_ms": ms - ms_pow,
                  "path": "power kernel + log-mel kernel through a 768 MB device scratch" if os.environ.get("VVB_MEL_UNFUSED") else "one fused kernel (no power spectrogram in HBM)",
                  "algorithmic_GB": (B * n * 4 + B * F * n_mels * 4) / 1e9, "GBps": (B * n * 4 + B * F * n_mels * 4) / ms / 1e6}))
# MFCC tail: same chain + DCT-II (13 coefficients, lifter 22)
out2 = torch.empty((B, F, 13), device=dev)
with Stft(nfft, hop, "hann") as h:
    h.set_stream(s.cuda_stream)
    for _ in range(2):
        h.batch_mfcc(x, w, 13, lifter=22.0, out=out2)
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(5):
        h.batch_mfcc(x, w, 13, lifter=22.0, out=out2)
    e1.record(s); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": f"STFT->MFCC, {B} x {n} samples, nfft={nfft} hop={hop}, {n_mels} mels, 13 coefficients", "ms": ms2,
                  "Msamples_per_s": B * n / ms2 / 1e3, "mfcc_stage_ms": ms2 - ms}))

# the speech front end: 16 kHz, 25 ms / 10 ms frames (fft_size 400, hop 160: mixed-radix kernels), 80 bands
B3, n3, nfft3, hop3 = 1024, 160_000, 400, 160
x3 = torch.rand((B3, n3), device=dev) * 2 - 1
st3, w3 = mel_filterbank(nfft3, n_mels, 16000.0, 0.0, 8000.0)
F3 = 1 + (n3 - nfft3) // hop3
out3 = torch.empty((B3, F3, n_mels), device=dev)
with Stft(nfft3, hop3, "hann") as h:
    h.set_stream(s.cuda_stream)
    for _ in range(2):
        h.batch_logmel(x3, w3, 1e-10, out=out3)
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(5):
        h.batch_logmel(x3, w3, 1e-10, out=out3)
    e1.record(s); torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": f"STFT->log-mel, {B3} x {n3} samples (10 s at 16 kHz), nfft={nfft3} hop={hop3}, {n_mels} mels", "ms": ms3,
                  "Msamples_per_s": B3 * n3 / ms3 / 1e3, "path": "power kernel + log-mel kernel through a device scratch" if os.environ.get("VVB_MEL_UNFUSED") else "one fused kernel (mixed-radix generic kernel, band sums per warp)"}))
