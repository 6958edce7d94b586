#!/usr/bin/env python
"""Device-resident sweep over BASELINE configs 4 and 5 (not the headline: that is bench.py).

    python benchmarks/sweep.py [--quick] > gpurun_out/sweep.jsonl

For every (nfft, hop = nfft/4, batch, n) it times STFT->complex, STFT->power and ISTFT
(normalised) with CUDA events and reports Msamples/s plus the fraction of the roofline
max(bytes / HBM peak, 5 N log2 N flops / 74.5 TF) of SURVEY.md section 8(d)."""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft  # noqa: E402

HBM = 6546.9e9
FP32 = 74.5e12


def timeit(fn, stream, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--nfft", type=int, nargs="*", default=None, help="restrict the sweep to these sizes")
    ap.add_argument("--batch", type=int, nargs="*", default=None, help="batch sizes (default 64 512 4096; --quick 64 1024)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    cases = []
    for nfft in (args.nfft or (256, 512, 1024, 2048, 4096, 8192)):
        for batch in (args.batch or ((64, 1024) if args.quick else (64, 512, 4096))):
            cases.append((nfft, nfft // 4, batch, 480_000, "config5"))
    if not args.nfft or 4096 in args.nfft:
        cases.append((4096, 1024, 1, 172_800_000 if not args.quick else 17_280_000, "config4 single stream"))
    for nfft, hop, B, n, tag in cases:
        F = 1 + (n - nfft) // hop
        bins = nfft // 2 + 1
        if B * F * bins * 12 + B * n * 8 > 150e9:      # spectra + power + signals + output must fit 180 GB
            continue
        x = torch.rand((B, n), device=dev) * 2 - 1
        spec = torch.empty((B, F, bins), device=dev, dtype=torch.complex64)
        pw = torch.empty((B, F, bins), device=dev, dtype=torch.float32)
        y = torch.empty((B, n), device=dev)
        with Stft(nfft, hop, "hann") as h:
            h.set_stream(stream.cuda_stream)
            reps = 3 if B * n > 1e9 else 10
            t_c = timeit(lambda: h.batch_forward(x, "complex", "valid", out=spec), stream, reps)
            t_p = timeit(lambda: h.batch_forward(x, "power", "valid", out=pw), stream, reps)
            t_i = timeit(lambda: h.batch_inverse(spec, n, True, out=y), stream, reps)
            lo, hi = nfft, n - nfft
            err = float(torch.linalg.vector_norm((y - x)[:, lo:hi].double()) / torch.linalg.vector_norm(x[:, lo:hi].double()))
        flops = B * F * 5 * nfft * math.log2(nfft)
        b_c = B * (4 * n + 8 * F * bins); b_p = B * (4 * n + 4 * F * bins)
        roof = lambda b: max(b / HBM, flops / FP32)
        print(json.dumps({"tag": tag, "nfft": nfft, "hop": hop, "batch": B, "n": n, "frames": F,
                          "stft_complex_ms": t_c * 1e3, "stft_power_ms": t_p * 1e3, "istft_ms": t_i * 1e3,
                          "stft_complex_Msps": B * n / t_c / 1e6, "stft_power_Msps": B * n / t_p / 1e6, "istft_Msps": B * n / t_i / 1e6,
                          "roofline_frac_complex": roof(b_c) / t_c, "roofline_frac_power": roof(b_p) / t_p, "roofline_frac_istft": roof(b_c) / t_i,
                          "roundtrip_rel_l2": err}), flush=True)
        del x, spec, pw, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
