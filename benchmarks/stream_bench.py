#!/usr/bin/env python
"""BASELINE config 4: ONE 1-hour 48 kHz stream (172.8 M samples), nfft=4096 hop=1024 Hann, sharded by
frame range over N GPUs with an (nfft-hop)-sample halo exchanged over NVLink (NCCL point-to-point).

    python benchmarks/stream_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/stream_bench.py

Each rank holds only its own span of the stream in HBM (synthetic, generated on device).  A step =
stream_stft (halo recv + fused STFT) + stream_istft (fused normalised ISTFT + tail send/add), timed
with CUDA events, max over ranks.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft, sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=3600)
    ap.add_argument("--nfft", type=int, default=4096)
    ap.add_argument("--hop", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29541")
    os.environ["NCCL_DEBUG"] = os.environ.get("VVB_NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, nfft, hop = 48000 * args.seconds, args.nfft, args.hop
    frames = 1 + (n - nfft) // hop
    s0, s1 = sharding.owned_samples(n, nfft, hop, world, rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    w = torch.hann_window(nfft, periodic=False, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    with Stft(nfft, hop, "hann") as h:
        h.set_stream(stream.cuda_stream)
        plan = sharding.StreamPlan(h, n, w)
        plan.x_owned.copy_(torch.rand(s1 - s0, device=dev, generator=g) * 2 - 1)
        x = plan.x_owned

        def step():
            return plan.istft(plan.stft())

        for _ in range(3):
            y = step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            y = step()
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo, hi = max(s0, nfft), min(s1, n - nfft)
        err = torch.tensor([float(torch.linalg.vector_norm((y[lo - s0: hi - s0] - x[lo - s0: hi - s0]).double())
                                  / torch.linalg.vector_norm(x[lo - s0: hi - s0].double()))], device=dev, dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
    if rank == 0:
        bins = nfft // 2 + 1
        gbytes = 2 * (4 * n + 8 * frames * bins) / 1e9
        print(json.dumps({"workload": f"config4: single {args.seconds}-s 48 kHz stream, nfft={nfft} hop={hop}, frame-range sharded, NVLink halo",
                          "n_gpus": world, "samples": n, "frames": frames, "ms_per_step": float(ms), "value": n / float(ms) / 1e3, "unit": "Msamples/s",
                          "algorithmic_GB_per_step": gbytes, "aggregate_GBps": gbytes / (float(ms) * 1e-3),
                          "halo_bytes_per_boundary": (nfft - hop) * 4, "roundtrip_rel_l2_max": float(err)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
