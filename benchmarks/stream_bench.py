#!/usr/bin/env python
"""BASELINE config 4: ONE 1-hour 48 kHz stream (172.8 M samples), nfft=4096 hop=1024 Hann, sharded by frame range over
1 / 2 / 4 / 8 GPUs by the C library's multi-device handle (vv_dsp_stft_stream_*, include/vv_dsp/b200.h).

    python benchmarks/stream_bench.py [--gpus 1 2 4 8] [--seconds 3600] [--no-graph]

ONE process drives all GPUs: per step and device one halo-gather kernel (peer-to-peer loads of the two nfft-hop sample
halos), the fused STFT kernel and the fused normalised ISTFT kernel, replayed as one CUDA graph per device.  Timed on
the devices (CUDA events on every device's stream, slowest device).  Prints one JSON line per GPU count: ms per step,
Msamples/s, strong-scaling efficiency against the 1-GPU time of the same run, fraction of the aggregate HBM roofline,
and whether the sharded result is bit-identical to the 1-GPU result."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import StftStream  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, nargs="*", default=None)
    ap.add_argument("--seconds", type=int, default=3600)
    ap.add_argument("--nfft", type=int, default=4096)
    ap.add_argument("--hop", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.no_graph:
        os.environ["VVB_STREAM_NO_GRAPH"] = "1"
    avail = torch.cuda.device_count()
    counts = args.gpus or [g for g in (1, 2, 4, 8) if g <= avail]
    n, nfft, hop = 48000 * args.seconds, args.nfft, args.hop
    frames, bins = 1 + (n - nfft) // hop, nfft // 2 + 1
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    rng = np.random.default_rng(4)
    x = np.tile(rng.uniform(-1, 1, 48000 * 30).astype(np.float32), args.seconds // 30 + 1)[:n]
    gbytes = 2 * (4 * n + 8 * frames * bins) / 1e9
    ms1, y1 = None, None
    for g in counts:
        with StftStream(nfft, hop, n, list(range(g))) as s:
            s.upload(x)
            ms = s.time_roundtrip(3, args.steps)
            y = s.download()
        if ms1 is None:
            ms1, y1 = ms, y
        lo, hi = nfft, n - nfft
        print(json.dumps({
            "workload": f"config4: single {args.seconds}-s 48 kHz stream, nfft={nfft} hop={hop}, frame-range sharded, peer-to-peer halos, "
                        f"{'plain enqueue' if args.no_graph else 'one CUDA graph per device and step'}",
            "n_gpus": g, "samples": n, "frames": frames, "steps": args.steps, "ms_per_step": ms, "value": n / ms / 1e3, "unit": "Msamples/s",
            "strong_scaling_efficiency_vs_first": ms1 / ms * counts[0] / g, "algorithmic_GB_per_step": gbytes,
            "hbm_frac_aggregate": gbytes / (ms * 1e-3) / (hbm * g), "halo_bytes_per_boundary_and_direction": (nfft - hop) * 4,
            "bit_identical_to_first": bool(np.array_equal(y, y1)),
            "roundtrip_rel_l2": float(np.linalg.norm((y[lo:hi] - x[lo:hi]).astype(np.float64)) / np.linalg.norm(x[lo:hi].astype(np.float64)))}),
            flush=True)


if __name__ == "__main__":
    main()
