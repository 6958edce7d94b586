#!/usr/bin/env python
"""Same-box A/B of library variants (benchmarks/build_variant.py): every variant is loaded into ONE process
and the kernels are timed interleaved, round after round, so clock / thermal drift hits all variants alike.

    python benchmarks/ab_kernels.py [--nfft 2048] [--hop 512] [--batch 1024] [--rounds 5] lib1.so lib2.so ...

Prints one JSON line per variant: median and min ms of STFT->complex, STFT->power and ISTFT (normalised)."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import Library, Stft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfft", type=int, default=2048)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--n", type=int, default=480_000)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--kinds", default="complex,power,inverse")
    ap.add_argument("--warm", type=int, default=200, help="warm-up launches before timing (lower it under ncu)")
    ap.add_argument("libs", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    B, n, nfft, hop = args.batch, args.n, args.nfft, args.hop
    F, bins = 1 + (n - nfft) // hop, nfft // 2 + 1
    x = torch.rand((B, n), device=dev) * 2 - 1
    spec = torch.empty((B, F, bins), device=dev, dtype=torch.complex64)
    pw = torch.empty((B, F, bins), device=dev, dtype=torch.float32)
    y = torch.empty((B, n), device=dev)
    hs, envs = [], []
    for spec_ in args.libs:                          # "path" or "path@VAR=1,VAR2=x": environment switches read by the library per call
        path, _, env = spec_.partition("@")
        h = Stft(nfft, hop, "hann", lib=Library(os.path.abspath(path)))
        h.set_stream(stream.cuda_stream)
        hs.append(h)
        envs.append(dict(kv.split("=", 1) for kv in env.split(",")) if env else {})
    all_vars = sorted({k for e in envs for k in e})

    def with_env(i, fn):
        for k in all_vars:
            if k in envs[i]:
                os.environ[k] = envs[i][k]
            else:
                os.environ.pop(k, None)
        fn(hs[i])
    kinds = args.kinds.split(",")
    fns = {"complex": lambda h: h.batch_forward(x, "complex", "valid", out=spec),
           "power": lambda h: h.batch_forward(x, "power", "valid", out=pw),
           "inverse": lambda h: h.batch_inverse(spec, n, True, out=y)}
    with_env(0, fns["complex"])
    res = {(i, k): [] for i in range(len(hs)) for k in kinds}
    errs = {}
    import random
    rnd = random.Random(0)
    for i, h in enumerate(hs):                      # round-trip check per variant + warm-up
        for k in kinds:
            with_env(i, fns[k])
        if "inverse" in kinds:
            torch.cuda.synchronize()
            errs[i] = float(torch.linalg.vector_norm((y - x)[:, nfft:-nfft].double()) / torch.linalg.vector_norm(x[:, nfft:-nfft].double()))
    for _ in range(args.warm):                      # ~0.4 s of load so the clocks have settled
        with_env(0, fns[kinds[0]])
    torch.cuda.synchronize()
    # single launches, variants in random order inside every repetition: drift hits all variants alike
    order = list(range(len(hs)))
    probe = next((h.lib for h in hs if hasattr(h.lib.dll, "vv_dsp_b200_sm_clock_mhz")), None)
    mhz = {k: [] for k in kinds}
    for k in kinds:
        n_samples = args.rounds * args.reps
        evs = []
        for r in range(n_samples + 5):
            rnd.shuffle(order)
            for i in order:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                with_env(i, fns[k])
                e1.record(stream)
                if r >= 5:
                    evs.append((i, e0, e1))
            if probe is not None and r % 10 == 9:          # SM clock while the queue is still full (syncs the stream)
                mhz[k].append(probe.sm_clock_mhz(stream.cuda_stream))
        torch.cuda.synchronize()
        for i, e0, e1 in evs:
            res[(i, k)].append(e0.elapsed_time(e1))
    for i, path in enumerate(args.libs):
        line = {"lib": os.path.basename(path), "nfft": nfft, "hop": hop, "batch": B, "n": n, "roundtrip_rel_l2": errs.get(i)}
        for k in kinds:
            line[k + "_ms_median"] = round(statistics.median(res[(i, k)]), 4)
            line[k + "_ms_mean"] = round(statistics.fmean(res[(i, k)]), 4)
            line[k + "_ms_min"] = round(min(res[(i, k)]), 4)
            if mhz[k]:
                line[k + "_sm_mhz"] = [round(min(mhz[k])), round(statistics.median(mhz[k])), round(max(mhz[k]))]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
