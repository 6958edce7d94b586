#!/usr/bin/env python
"""What the host link gives: pinned H2D alone, D2H alone, and both directions at once (1.97 GB each way, the bytes the
e2e leg of bench.py moves per step and GPU), CUDA-event timed -- for ONE GPU or for N ranks AT THE SAME TIME.

    python benchmarks/pcie_probe.py                                                        # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port 29544 benchmarks/pcie_probe.py [--affinity]                      # N GPUs concurrently

Under torchrun every rank drives its own GPU; the ranks start each measurement together (gloo barrier) and the slowest
rank's time counts, so "both_aggregate_GBps_each_way" is the floor of the N-GPU e2e leg.  --affinity pins every rank
to its own 1/N-th of the host cores before it allocates its pinned buffers (first touch), to see whether placement
matters on this host.  One JSON line on rank 0."""
import argparse
import json
import os

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--affinity", action="store_true")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if args.affinity:
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // world)
    os.sched_setaffinity(0, set(cores[rank * per:(rank + 1) * per]) or set(cores))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")

n = 1024 * 480_000
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
h_in.fill_(1.0); h_out.fill_(0.0)
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def h2d():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)


def d2h():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)


gb = n * 4 / 1e9
t1, t2, t3 = timed(h2d, args.reps), timed(d2h, args.reps), timed(both, args.reps)
if rank == 0:
    print(json.dumps({"n_gpus": world, "affinity": bool(args.affinity), "host_cores": len(os.sched_getaffinity(0)) if not args.affinity else None,
                      "bytes_each_way_per_gpu": n * 4, "h2d_alone_ms": t1, "h2d_aggregate_GBps": world * gb / t1 * 1e3,
                      "d2h_alone_ms": t2, "d2h_aggregate_GBps": world * gb / t2 * 1e3,
                      "both_ms": t3, "both_aggregate_GBps_each_way": world * gb / t3 * 1e3,
                      "e2e_floor_Msamples_per_s": world * n / t3 / 1e3}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
