#!/usr/bin/env python
"""Non-power-of-two fft_size and hops without a marching kernel, device-resident, CUDA-event timed: the speech-standard 25 ms /
10 ms framing at 16 kHz (nfft=400, hop=160: mixed-radix Stockham kernels; VVB_NO_MIXED_RADIX=1: chirp-z; VVB_NO_BLUESTEIN=1:
the direct O(n^2) kernels), a chirp-z size, power-of-two sizes with odd hops, a size beyond one fused kernel.  One JSON line each."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft  # noqa: E402

dev = torch.device("cuda", 0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)


def run(nfft, hop, B, n, label, reps):
    x = torch.rand((B, n), device=dev) * 2 - 1
    with Stft(nfft, hop, "hann") as h:
        h.set_stream(s.cuda_stream)
        F = h.num_frames(n)
        spec = torch.empty((B, F, nfft // 2 + 1), device=dev, dtype=torch.complex64)
        y = torch.empty((B, n), device=dev)
        h.batch_forward(x, "complex", "valid", out=spec); h.batch_inverse(spec, n, True, out=y)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(s)
        for _ in range(reps):
            h.batch_forward(x, "complex", "valid", out=spec)
        e[1].record(s)
        for _ in range(reps):
            h.batch_inverse(spec, n, True, out=y)
        e[2].record(s); torch.cuda.synchronize()
        fwd, inv = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
        lo, hi = nfft, n - nfft
        err = float(torch.linalg.norm(y[:, lo:hi] - x[:, lo:hi]) / torch.linalg.norm(x[:, lo:hi]))
    print(json.dumps({"path": label, "nfft": nfft, "hop": hop, "batch": B, "n": n, "frames": F, "stft_ms": fwd, "istft_ms": inv,
                      "Msamples_per_s": B * n / (fwd + inv) / 1e3, "roundtrip_rel_l2": err}))


if __name__ == "__main__":
    label = "direct" if os.environ.get("VVB_NO_BLUESTEIN") else "bluestein"
    small = label == "direct"
    mixed = not small and not os.environ.get("VVB_NO_MIXED_RADIX")
    run(400, 160, 16 if small else 1024, 160_000, "mixed-radix Stockham (Cfg200)" if mixed else label, 2 if small else 5)
    if mixed:
        run(480, 160, 1024, 160_000, "mixed-radix Stockham (Cfg240)", 5)
        run(640, 160, 1024, 160_000, "mixed-radix Stockham (Cfg320)", 5)
    run(1000, 250, 16 if small else 1024, 160_000, label, 2 if small else 5)
    if not small:
        # power-of-two sizes with hops that have no marching / pair kernel (generic forward kernel + slot overlap-add)
        for nfft, hop in ((512, 160), (1024, 160), (2048, 300)):
            run(nfft, hop, 1024, 160_000, "pow2, odd hop (generic kernels)", 5)
        run(512, 128, 1024, 160_000, "pow2, hop N/4 (for comparison)", 5)
        # chirp-z beyond one fused kernel: four-step plans underneath
        run(6000, 1500, 256, 160_000, "bluestein on four-step plans", 3)
