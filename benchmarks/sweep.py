#!/usr/bin/env python
"""BASELINE config 5: nfft sweep 256-8192 x batch 64-8192 signals, SHARDED BY SIGNAL across the ranks, vs roofline.

    python benchmarks/sweep.py [--quick] > gpurun_out/sweep_n1.jsonl                          # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port 29533 benchmarks/sweep.py > gpurun_out/sweep_nN.jsonl               # N GPUs

One process per GPU.  For every (nfft, hop = nfft/4, total batch B) the batch is split into contiguous ranges of
signals, one per rank (vv_dsp_b200.sharding.shard_batch: no collective on the data path), and
STFT->complex, STFT->power and ISTFT (normalised) are timed on every rank with CUDA events; the cell's time is the
MAX over ranks (all-reduce of one float64 after the timed region).  Rank 0 prints one JSON line per cell with
whole-job Msamples/s and the fraction of the roofline max(bytes / (N x HBM), 5 N log2 N flops / (N x FP32)) of
SURVEY.md section 8(d).  Inputs are synthetic (uniform noise generated on the device) and, except at the smallest
batches, larger than L2; every repetition reads the whole batch, so nothing is served from cache across repetitions
for cells above ~126 MB per GPU (smaller cells are marked "l2_resident": their numbers are L2, not HBM, numbers)."""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Stft  # noqa: E402
from vv_dsp_b200.sharding import shard_batch  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]) * 1e9
    except Exception:
        return 6650e9


HBM = peaks()
FP32 = 74.5e12          # nominal 148 SMs x 128 lanes x 2 x 1.965 GHz (the library's probe measures 75.9 with scalar FFMA)


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
    ap.add_argument("--batch", type=int, nargs="*", default=None, help="TOTAL batch sizes (default 64 512 1024 4096 8192)")
    ap.add_argument("--n", type=int, default=480_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    n = args.n
    batches = args.batch or ((64, 1024) if args.quick else (64, 512, 1024, 4096, 8192))
    for nfft in (args.nfft or (256, 512, 1024, 2048, 4096, 8192)):
        hop = nfft // 4
        F, bins = 1 + (n - nfft) // hop, nfft // 2 + 1
        with Stft(nfft, hop, "hann") as h:
            h.set_stream(stream.cuda_stream)
            for Btot in batches:
                b0, b1 = shard_batch(Btot, world, rank)
                B = b1 - b0
                per_gpu = max(1, -(-Btot // world)) * (F * bins * 12 + n * 8)
                if per_gpu > 150e9 or B == 0:      # spectra + power + signals + output must fit 180 GB
                    continue
                g = torch.Generator(device=dev).manual_seed(1000 + b0)
                x = torch.rand((max(B, 1), n), device=dev, generator=g) * 2 - 1
                spec = torch.empty((max(B, 1), F, bins), device=dev, dtype=torch.complex64)
                pw = torch.empty((max(B, 1), F, bins), device=dev, dtype=torch.float32)
                y = torch.empty((max(B, 1), n), device=dev)
                reps = 3 if B * n > 1e9 else 10
                t = [timeit(lambda: h.batch_forward(x, "complex", "valid", out=spec), stream, reps),
                     timeit(lambda: h.batch_forward(x, "power", "valid", out=pw), stream, reps),
                     timeit(lambda: h.batch_inverse(spec, n, True, out=y), stream, reps)]
                lo, hi = nfft, n - nfft
                err = float(torch.linalg.vector_norm((y - x)[:, lo:hi].double()) / torch.linalg.vector_norm(x[:, lo:hi].double()))
                tt = torch.tensor(t + [err], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_c, t_p, t_i, err = [float(v) for v in tt]
                if rank == 0:
                    flops = Btot * F * 5 * nfft * math.log2(nfft)
                    b_c = Btot * (4 * n + 8 * F * bins); b_p = Btot * (4 * n + 4 * F * bins)
                    roof = lambda b: max(b / (HBM * world), flops / (FP32 * world))
                    print(json.dumps({
                        "tag": "config5", "n_gpus": world, "nfft": nfft, "hop": hop, "batch_total": Btot, "batch_per_gpu": B, "n": n, "frames": F,
                        "stft_complex_ms": t_c * 1e3, "stft_power_ms": t_p * 1e3, "istft_ms": t_i * 1e3,
                        "stft_complex_Msps": Btot * n / t_c / 1e6, "stft_power_Msps": Btot * n / t_p / 1e6, "istft_Msps": Btot * n / t_i / 1e6,
                        "roofline_frac_complex": roof(b_c) / t_c, "roofline_frac_power": roof(b_p) / t_p, "roofline_frac_istft": roof(b_c) / t_i,
                        "l2_resident": bool((b_c / world) < 126e6), "roundtrip_rel_l2": err}), flush=True)
                del x, spec, pw, y
                torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
