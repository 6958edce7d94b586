"""Shared helpers for the tests (input generators, tolerances)."""
import hashlib

import numpy as np

# north-star tolerance for float32 spectra: |X - Xref| <= ATOL*max|Xref| + RTOL*|Xref|
RTOL = 5e-5
ATOL = 5e-5
# north-star tolerance for ISTFT(STFT(x)) on the interior [nfft, n-nfft): relative L2
ROUNDTRIP_REL_L2 = 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def noise(seed, n):
    """uniform(-1,1) float32 -- same generator as tests/golden/make_golden.py"""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(np.float32)


def spectra_close(x, ref, rtol=RTOL, atol=ATOL):
    """Returns (ok, worst budget fraction). Budget per bin = atol*max|ref| + rtol*|ref| (per frame row)."""
    x = np.asarray(x)
    ref = np.asarray(ref)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if ref.size == 0:
        return True, 0.0
    mx = np.abs(ref).max(axis=-1, keepdims=True)
    budget = atol * mx + rtol * np.abs(ref)
    err = np.abs(x.astype(np.complex128) - ref.astype(np.complex128)) if np.iscomplexobj(ref) else np.abs(
        x.astype(np.float64) - ref.astype(np.float64))
    zero = budget == 0
    frac = np.where(zero, np.where(err == 0, 0.0, np.inf), err / np.where(zero, 1, budget))
    return bool((frac <= 1.0).all()), float(frac.max())


def rel_l2(y, x):
    y = np.asarray(y, np.float64)
    x = np.asarray(x, np.float64)
    d = np.linalg.norm(x)
    return float(np.linalg.norm(y - x) / d) if d > 0 else float(np.linalg.norm(y - x))


def stft_truth_f64(x, w, nfft, hop, frames):
    """float64 STFT truth (valid frames) for accuracy-vs-truth reporting."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    idx = np.arange(nfft)[None, :] + hop * np.arange(frames)[:, None]
    return np.fft.rfft(x[idx] * w[None, :], axis=-1)
