#!/usr/bin/env python
"""Per-frame latency of a LITERAL drop-in: existing vv-dsp callers call vv_dsp_stft_process /
vv_dsp_stft_reconstruct once per frame with host pointers (reference bench/bench_stft.c:80-98).  tests/c/perframe_latency.c
times that loop; it is built twice from the same source -- against libvvdsp_b200.so (every call = H2D + kernel + D2H +
stream sync) and against the reference compiled for the CPU (oracle/_ref/libvvdsp_ref.so) -- and run for several sizes.

    python benchmarks/perframe_latency.py > gpurun_out/perframe_latency.jsonl

One JSON line per (library, fft_size).  The batched entry points (include/vv_dsp/b200.h) are the throughput path; this
figure tells a maintainer what happens if nothing but the link line changes."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
SRC = os.path.join(ROOT, "tests", "c", "perframe_latency.c")


def build(out, libdir, libname):
    subprocess.run(["gcc", "-std=gnu99", "-O2", "-I" + INC, SRC, "-o", out, "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir, "-lm"], check=True)
    return out


def main():
    tmp = tempfile.mkdtemp()
    libs = [("vv-dsp_b200 (B200, per-frame API)", build(os.path.join(tmp, "lat_b200"), os.path.join(ROOT, "vv_dsp_b200", "lib"), "vvdsp_b200"))]
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref, "libvvdsp_ref.so")):
        libs.append(("reference (CPU, one core)", build(os.path.join(tmp, "lat_ref"), ref, "vvdsp_ref")))
    for name, exe in libs:
        for nfft in (256, 512, 1024, 2048, 4096, 8192):
            r = subprocess.run([exe, str(nfft), "400"], capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                print(json.dumps({"library": name, "fft_size": nfft, "error": (r.stdout + r.stderr)[-300:]}), flush=True)
                continue
            row = json.loads(r.stdout)
            row["library"] = name
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    sys.exit(main())
