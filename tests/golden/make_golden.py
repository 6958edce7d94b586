#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libvvdsp_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed):
  voicebank_aka_sa_pcm16.npz  PCM16 samples of /root/reference/voicebank/_a'ka'sa.wav
                              (BASELINE config 1 input; decoded the way src/audio/wav.c:471-483
                              does it: float = int16 / 32768)
  golden_cases.json           per case: parameters, seeds, sha256 of the reference's float32
                              output bytes, and a few leading values for eyeballing
  golden_slices.npz           small slices of reference outputs (spectra rows, round-trip
                              segments) so GPU parity can be checked against the reference
                              itself on the GPU box, where /root/reference does not exist
The oracle restatement must reproduce every digest bit-for-bit (tests/test_oracle.py).
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import Reference, num_frames  # noqa: E402

WAV = "/root/reference/voicebank/_a'ka'sa.wav"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def read_pcm16_mono(path):
    d = open(path, "rb").read()
    assert d[:4] == b"RIFF" and d[8:12] == b"WAVE"
    pos = 12
    fmt = None
    while pos < len(d):
        cid, sz = d[pos:pos + 4], struct.unpack("<I", d[pos + 4:pos + 8])[0]
        body = d[pos + 8:pos + 8 + sz]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            assert fmt and fmt[0] == 1 and fmt[1] == 1 and fmt[5] == 16, fmt
            return np.frombuffer(body, dtype="<i2").copy(), fmt[2]
        pos += 8 + sz + (sz & 1)
    raise RuntimeError("no data chunk")


def signal(seed, n):
    """uniform(-1,1) float32, numpy PCG64 (stream is version-stable)."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(np.float32)


def main():
    ref = Reference()
    cases = []
    slices = {}

    pcm, sr = read_pcm16_mono(WAV)
    np.savez_compressed(os.path.join(HERE, "voicebank_aka_sa_pcm16.npz"), pcm=pcm, sample_rate=sr)
    wav = (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)

    # --- windows
    for kind in ("boxcar", "hann", "hamming"):
        for n in (1, 2, 8, 17, 64, 1024, 2048, 4096):
            st, w = ref.window(kind, n)
            cases.append(dict(kind="window", window=kind, n=n, status=st, sha256=sha(w), head=[float(v) for v in w[:4]]))

    # --- FFT plan API
    for n in (1, 2, 4, 8, 16, 64, 100, 200, 256, 1024, 2048, 4096):
        rng = np.random.default_rng(1000 + n)
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        xr = rng.uniform(-1, 1, n).astype(np.float32)
        f = ref.fft_c2c(x, +1)
        b = ref.fft_c2c(x, -1)
        r2c = ref.fft_r2c(xr)
        entry = dict(kind="fft", n=n, seed=1000 + n, c2c_fwd=sha(f), c2c_bwd=sha(b), r2c=sha(r2c),
                     head=[[float(v.real), float(v.imag)] for v in f[:2]])
        if n <= 256:
            entry["c2r"] = sha(ref.fft_c2r(r2c, n))
        cases.append(entry)

    # --- STFT / ISTFT on seeded noise, all four frame conventions
    stft_cases = [
        (2048, 512, "hann", 20000, 11), (2048, 512, "hamming", 9000, 12), (1024, 256, "hann", 12000, 13),
        (512, 128, "hann", 4096, 2), (256, 64, "boxcar", 3000, 14), (64, 32, "hann", 256, 15),
        (4096, 1024, "hann", 30000, 16), (8192, 2048, "hann", 40000, 17), (2048, 300, "hann", 10000, 18),
        (128, 128, "hann", 1000, 19), (16, 8, "hamming", 200, 20), (2, 1, "boxcar", 40, 21),
    ]
    for nfft, hop, win, n, seed in stft_cases:
        x = signal(seed, n)
        e = dict(kind="stft", nfft=nfft, hop=hop, window=win, n=n, seed=seed, frames={}, sha256={})
        for conv in ("valid", "spectrogram", "padded_tail", "center"):
            s = ref.stft(x, nfft, hop, win, convention=conv)
            e["frames"][conv] = int(s.shape[0])
            e["sha256"]["stft_" + conv] = sha(s)
            if conv == "valid":
                sv = s
        e["sha256"]["roundtrip"] = sha(ref.batch_roundtrip(x[None, :], nfft, hop, win)[0])
        e["sha256"]["power"] = sha(ref.batch_power(x[None, :], nfft, hop, win)[0])
        e["sha256"]["spectrogram_mag"] = sha(ref.spectrogram(x, nfft, hop, win))
        y = ref.istft(sv, nfft, hop, n, win)
        e["sha256"]["istft_half_valid"] = sha(y)
        cases.append(e)
        key = f"n{nfft}_h{hop}_{win}_s{seed}"
        if sv.shape[0]:
            slices[key + "_stft_row0"] = sv[0]
            slices[key + "_stft_rowlast"] = sv[-1]
        slices[key + "_istft_mid"] = y[n // 2: n // 2 + 256]

    # --- BASELINE config 1: the voicebank WAV, nfft=1024 hop=256 Hann
    nfft, hop = 1024, 256
    s = ref.batch_forward(wav[None, :], nfft, hop)[0]
    y = ref.batch_roundtrip(wav[None, :], nfft, hop)[0]
    cases.append(dict(kind="config1", nfft=nfft, hop=hop, window="hann", n=int(wav.size), frames=int(s.shape[0]),
                      sha256=dict(input=sha(wav), stft_valid=sha(s), roundtrip=sha(y),
                                  power=sha(ref.batch_power(wav[None, :], nfft, hop)[0]))))
    for f in (0, 310, 620):
        slices[f"config1_stft_row{f}"] = s[f]
    slices["config1_roundtrip_80000"] = y[80000:81024]

    json.dump(dict(generator="tests/golden/make_golden.py", reference="/root/reference (crlotwhite/vv-dsp)",
                   cases=cases), open(os.path.join(HERE, "golden_cases.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_slices.npz"), **slices)
    print(f"{len(cases)} cases, {len(slices)} slices")


if __name__ == "__main__":
    main()
