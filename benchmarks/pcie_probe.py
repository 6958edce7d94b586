#!/usr/bin/env python
"""What the host link gives: pinned H2D alone, D2H alone, and both directions at once (2 GB each), CUDA-event timed.
The e2e leg of bench.py moves 1.97 GB each way per step; this is its floor.  One JSON line."""
import json

import torch

n = 1024 * 480_000
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


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
t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"bytes_each_way": n * 4, "h2d_alone_ms": t1, "h2d_GBps": gb / t1 * 1e3, "d2h_alone_ms": t2, "d2h_GBps": gb / t2 * 1e3,
                  "both_ms": t3, "both_GBps_each_way": gb / t3 * 1e3}))
