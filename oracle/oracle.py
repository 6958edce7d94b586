"""ctypes access to the CPU checkers under oracle/.  TEST / BENCH INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package vv_dsp_b200 never does.

Two libraries:
  * ``libvvdsp_oracle.so``      -- the restatement (oracle/vvdsp_oracle.c), class ``Oracle``
  * ``_ref/libvvdsp_ref.so``    -- the unmodified reference compiled from /root/reference
                                   (oracle/Makefile target ``ref``), class ``Reference``
Both expose the same numpy-level methods so tests can run one against the other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libvvdsp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvvdsp_ref.so")

WIN = {"boxcar": 0, "hann": 1, "hamming": 2}
CONV = {"valid": 0, "spectrogram": 1, "padded_tail": 2, "center": 3}

_f32p = C.POINTER(C.c_float)
_sz = C.c_size_t


def build(force: bool = False) -> None:
    """Compile the restatement and, where /root/reference exists, the reference .so."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("vvdsp_oracle.c", "driver.c"))
    ):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src/spectral") and (force or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex64)


def num_frames(n: int, nfft: int, hop: int, convention: str = "valid") -> int:
    """Pure-python statement of the four frame-count conventions."""
    if hop == 0:
        return 0
    if convention == "valid":
        return 0 if n < nfft else 1 + (n - nfft) // hop
    if convention == "spectrogram":
        return 1 if n < nfft else 1 + (n - nfft + hop) // hop
    if convention == "padded_tail":
        return n // hop if hop <= nfft else 0
    if convention == "center":
        return (n + hop - 1) // hop
    raise ValueError(convention)


class _Base:
    """Batch-loop entry points shared by both libraries (oracle/driver.c)."""

    lib: C.CDLL
    prefix: str

    def _drv(self, name):
        fn = getattr(self.lib, f"{self.prefix}_{name}")
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, _sz, _sz, _sz, _sz, _sz, C.c_int, C.c_void_p, C.c_int]
        return fn

    def batch_roundtrip(self, x, nfft, hop, win="hann", threads=1, want_output=True):
        x = _f32(x)
        B, n = x.shape
        y = np.empty_like(x) if want_output else None
        st = self._drv("batch_roundtrip")(_p(x), B, n, n, nfft, hop, WIN[win], _p(y) if want_output else None, threads)
        assert st == 0, st
        return y

    def batch_power(self, x, nfft, hop, win="hann", threads=1, want_output=True):
        x = _f32(x)
        B, n = x.shape
        F = num_frames(n, nfft, hop)
        out = np.empty((B, F, nfft // 2 + 1), np.float32) if want_output else None
        st = self._drv("batch_power")(_p(x), B, n, n, nfft, hop, WIN[win], _p(out) if want_output else None, threads)
        assert st == 0, st
        return out

    def batch_forward(self, x, nfft, hop, win="hann", threads=1):
        x = _f32(x)
        B, n = x.shape
        F = num_frames(n, nfft, hop)
        out = np.empty((B, F, nfft // 2 + 1), np.complex64)
        st = self._drv("batch_forward")(_p(x), B, n, n, nfft, hop, WIN[win], _p(out), threads)
        assert st == 0, st
        return out


class Oracle(_Base):
    """The restatement, vvdsp_oracle.c."""

    prefix = "orcdrv"

    def __init__(self):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_num_frames.restype = _sz
        L.orc_num_frames.argtypes = [_sz, _sz, _sz, C.c_int]
        for name in ("orc_window", "orc_fft_c2c", "orc_fft_r2c", "orc_fft_c2r", "orc_stft_create",
                     "orc_stft_process", "orc_stft_reconstruct", "orc_fetch_frame", "orc_overlap_add",
                     "orc_stft_spectrogram", "orc_stft_forward", "orc_stft_power", "orc_stft_istft",
                     "orc_stft_roundtrip"):
            getattr(L, name).restype = C.c_int
        L.orc_stft_destroy.restype = None

    # --- elementary pieces
    def window(self, kind, n):
        w = np.empty(max(n, 1), np.float32)
        st = self.lib.orc_window(C.c_int(WIN[kind] if isinstance(kind, str) else kind), _sz(n), _p(w))
        return st, w[:n]

    def fft_c2c(self, x, direction=+1):
        x = _c64(x)
        out = np.empty_like(x)
        st = self.lib.orc_fft_c2c(_p(x), _p(out), _sz(x.size), C.c_int(direction))
        assert st == 0, st
        return out

    def fft_r2c(self, x):
        x = _f32(x)
        out = np.empty(x.size // 2 + 1, np.complex64)
        assert self.lib.orc_fft_r2c(_p(x), _p(out), _sz(x.size)) == 0
        return out

    def fft_c2r(self, X, n):
        X = _c64(X)
        out = np.empty(n, np.float32)
        assert self.lib.orc_fft_c2r(_p(X), _p(out), _sz(n)) == 0
        return out

    def num_frames(self, n, nfft, hop, convention="valid"):
        return int(self.lib.orc_num_frames(n, nfft, hop, CONV[convention]))

    def fetch_frame(self, x, flen, hop, index, center=False, window=None):
        x = _f32(x)
        out = np.zeros(flen, np.float32)
        w = _f32(window) if window is not None else None
        st = self.lib.orc_fetch_frame(_p(x), _sz(x.size), _p(out), _sz(flen), _sz(hop), _sz(index),
                                      C.c_int(int(center)), _p(w) if w is not None else None)
        return st, out

    def overlap_add(self, frame, out, hop, index):
        frame = _f32(frame)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        return self.lib.orc_overlap_add(_p(frame), _p(out), _sz(out.size), _sz(frame.size), _sz(hop), _sz(index))

    # --- handle-based
    def _handle(self, nfft, hop, win):
        h = C.c_void_p()
        st = self.lib.orc_stft_create(_sz(nfft), _sz(hop), C.c_int(WIN[win] if isinstance(win, str) else win), C.byref(h))
        return st, h

    def create_status(self, nfft, hop, win):
        st, h = self._handle(nfft, hop, win)
        if st == 0:
            self.lib.orc_stft_destroy(h)
        return st

    def process(self, frame, nfft, hop, win="hann"):
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        frame = _f32(frame)
        out = np.empty(nfft, np.complex64)
        assert self.lib.orc_stft_process(h, _p(frame), _p(out)) == 0
        self.lib.orc_stft_destroy(h)
        return out

    def reconstruct(self, spec, nfft, hop, win="hann"):
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        spec = _c64(spec)
        out = np.zeros(nfft, np.float32)
        norm = np.zeros(nfft, np.float32)
        assert self.lib.orc_stft_reconstruct(h, _p(spec), _p(out), _p(norm)) == 0
        self.lib.orc_stft_destroy(h)
        return out, norm

    def stft(self, x, nfft, hop, win="hann", convention="valid", half=True):
        x = _f32(x)
        F = num_frames(x.size, nfft, hop, convention)
        bins = nfft // 2 + 1 if half else nfft
        out = np.empty((F, bins), np.complex64)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.orc_stft_forward(h, _p(x), _sz(x.size), C.c_int(CONV[convention]), C.c_int(int(half)),
                                         _p(out), _sz(F)) == 0
        self.lib.orc_stft_destroy(h)
        return out

    def power(self, x, nfft, hop, win="hann", convention="valid"):
        x = _f32(x)
        F = num_frames(x.size, nfft, hop, convention)
        out = np.empty((F, nfft // 2 + 1), np.float32)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.orc_stft_power(h, _p(x), _sz(x.size), C.c_int(CONV[convention]), _p(out), _sz(F)) == 0
        self.lib.orc_stft_destroy(h)
        return out

    def spectrogram(self, x, nfft, hop, win="hann"):
        x = _f32(x)
        F = num_frames(x.size, nfft, hop, "spectrogram")
        out = np.empty((F, nfft), np.float32)
        frames = _sz(0)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.orc_stft_spectrogram(h, _p(x), _sz(x.size), _p(out), C.byref(frames)) == 0
        self.lib.orc_stft_destroy(h)
        assert frames.value == F
        return out

    def istft(self, spec, nfft, hop, n_out, win="hann", half=True, normalise=True):
        spec = _c64(spec)
        F = spec.shape[0]
        y = np.empty(n_out, np.float32)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.orc_stft_istft(h, _p(spec), _sz(F), C.c_int(int(half)), _p(y), _sz(n_out),
                                       C.c_int(int(normalise))) == 0
        self.lib.orc_stft_destroy(h)
        return y

    def roundtrip(self, x, nfft, hop, win="hann"):
        x = _f32(x)
        y = np.empty_like(x)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.orc_stft_roundtrip(h, _p(x), _sz(x.size), _p(y)) == 0
        self.lib.orc_stft_destroy(h)
        return y

    # --- mel
    def mel_filterbank(self, n_fft, n_mels, sr, fmin, fmax, variant=0):
        bins = n_fft // 2 + 1
        w = np.zeros((max(n_mels, 1), bins), np.float32)
        self.lib.orc_mel_filterbank.restype = C.c_int
        st = self.lib.orc_mel_filterbank(_sz(n_fft), _sz(n_mels), C.c_float(sr), C.c_float(fmin), C.c_float(fmax),
                                         C.c_int(variant), _p(w))
        return st, w[:n_mels]

    def log_mel(self, power, weights, eps):
        power, weights = _f32(power), _f32(weights)
        out = np.empty((power.shape[0], weights.shape[0]), np.float32)
        self.lib.orc_log_mel.restype = C.c_int
        st = self.lib.orc_log_mel(_p(power), _sz(power.shape[0]), _sz(power.shape[1]), _p(weights), _sz(weights.shape[0]),
                                  C.c_float(eps), _p(out))
        assert st == 0, st
        return out

    def pcm_to_planar(self, raw, fmt, channels):
        """raw: bytes-like interleaved little-endian samples; fmt 16 / 24 / 32 (PCM) or -32 (float32)"""
        raw = np.frombuffer(bytes(raw), np.uint8)
        n = raw.size // (abs(fmt) // 8) // channels
        out = np.zeros((channels, n), np.float32)
        self.lib.orc_pcm_to_planar.restype = C.c_int
        st = self.lib.orc_pcm_to_planar(_p(raw), C.c_int(fmt), _sz(n), _sz(channels), _p(out), _sz(n))
        assert st == 0, st
        return out

    def mfcc(self, log_mel, n_coeffs, lifter=0.0, dct_type=2):
        log_mel = _f32(log_mel)
        out = np.empty((log_mel.shape[0], max(n_coeffs, 1)), np.float32)
        self.lib.orc_mfcc.restype = C.c_int
        st = self.lib.orc_mfcc(_p(log_mel), _sz(log_mel.shape[0]), _sz(log_mel.shape[1]), _sz(n_coeffs), C.c_int(dct_type),
                               C.c_float(lifter), _p(out))
        return st, out[:, :n_coeffs]


class _Params(C.Structure):
    _fields_ = [("fft_size", _sz), ("hop_size", _sz), ("window", C.c_int)]


class Reference(_Base):
    """The unmodified reference library (vv_dsp_* API), when oracle/_ref was built."""

    prefix = "ref"

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(REF_SO)

    def __init__(self):
        build()
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.vv_dsp_get_num_frames.restype = _sz
        L.vv_dsp_get_num_frames.argtypes = [_sz, _sz, _sz, C.c_int]

    def window(self, kind, n):
        w = np.empty(max(n, 1), np.float32)
        fn = getattr(self.lib, f"vv_dsp_window_{kind}")
        st = fn(_sz(n), _p(w))
        return st, w[:n]

    def _plan(self, n, ftype, direction):
        plan = C.c_void_p()
        st = self.lib.vv_dsp_fft_make_plan(_sz(n), C.c_int(ftype), C.c_int(direction), C.byref(plan))
        assert st == 0, st
        return plan

    def fft_c2c(self, x, direction=+1):
        x = _c64(x)
        out = np.empty_like(x)
        plan = self._plan(x.size, 0, direction)
        assert self.lib.vv_dsp_fft_execute(plan, _p(x), _p(out)) == 0
        self.lib.vv_dsp_fft_destroy(plan)
        return out

    def fft_r2c(self, x):
        x = _f32(x)
        out = np.empty(x.size // 2 + 1, np.complex64)
        plan = self._plan(x.size, 1, +1)
        assert self.lib.vv_dsp_fft_execute(plan, _p(x), _p(out)) == 0
        self.lib.vv_dsp_fft_destroy(plan)
        return out

    def fft_c2r(self, X, n):
        X = _c64(X)
        out = np.empty(n, np.float32)
        plan = self._plan(n, 2, -1)
        assert self.lib.vv_dsp_fft_execute(plan, _p(X), _p(out)) == 0
        self.lib.vv_dsp_fft_destroy(plan)
        return out

    def num_frames(self, n, nfft, hop, convention="valid"):
        if convention == "valid":
            return int(self.lib.vv_dsp_get_num_frames(n, nfft, hop, 0))
        if convention == "center":
            return int(self.lib.vv_dsp_get_num_frames(n, nfft, hop, 1))
        raise ValueError("reference has no function for this convention")

    def fetch_frame(self, x, flen, hop, index, center=False, window=None):
        x = _f32(x)
        out = np.zeros(flen, np.float32)
        w = _f32(window) if window is not None else None
        st = self.lib.vv_dsp_fetch_frame(_p(x), _sz(x.size), _p(out), _sz(flen), _sz(hop), _sz(index),
                                         C.c_int(int(center)), _p(w) if w is not None else None)
        return st, out

    def overlap_add(self, frame, out, hop, index):
        frame = _f32(frame)
        return self.lib.vv_dsp_overlap_add(_p(frame), _p(out), _sz(out.size), _sz(frame.size), _sz(hop), _sz(index))

    def _handle(self, nfft, hop, win):
        h = C.c_void_p()
        p = _Params(nfft, hop, WIN[win] if isinstance(win, str) else win)
        st = self.lib.vv_dsp_stft_create(C.byref(p), C.byref(h))
        return st, h

    def create_status(self, nfft, hop, win):
        st, h = self._handle(nfft, hop, win)
        if st == 0:
            self.lib.vv_dsp_stft_destroy(h)
        return st

    def process(self, frame, nfft, hop, win="hann"):
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        frame = _f32(frame)
        out = np.empty(nfft, np.complex64)
        assert self.lib.vv_dsp_stft_process(h, _p(frame), _p(out)) == 0
        self.lib.vv_dsp_stft_destroy(h)
        return out

    def reconstruct(self, spec, nfft, hop, win="hann"):
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        spec = _c64(spec)
        out = np.zeros(nfft, np.float32)
        norm = np.zeros(nfft, np.float32)
        assert self.lib.vv_dsp_stft_reconstruct(h, _p(spec), _p(out), _p(norm)) == 0
        self.lib.vv_dsp_stft_destroy(h)
        return out, norm

    def spectrogram(self, x, nfft, hop, win="hann"):
        x = _f32(x)
        F = num_frames(x.size, nfft, hop, "spectrogram")
        out = np.empty((F, nfft), np.float32)
        frames = _sz(0)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        assert self.lib.vv_dsp_stft_spectrogram(h, _p(x), _sz(x.size), _p(out), C.byref(frames)) == 0
        self.lib.vv_dsp_stft_destroy(h)
        assert frames.value == F
        return out

    def stft(self, x, nfft, hop, win="hann", convention="valid", half=True):
        """Per-frame python loop over the reference API (small inputs only)."""
        x = _f32(x)
        F = num_frames(x.size, nfft, hop, convention)
        bins = nfft // 2 + 1 if half else nfft
        out = np.empty((F, bins), np.complex64)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        frame = np.zeros(nfft, np.float32)
        spec = np.empty(nfft, np.complex64)
        for f in range(F):
            assert self.lib.vv_dsp_fetch_frame(_p(x), _sz(x.size), _p(frame), _sz(nfft), _sz(hop), _sz(f),
                                               C.c_int(int(convention == "center")), None) == 0
            assert self.lib.vv_dsp_stft_process(h, _p(frame), _p(spec)) == 0
            out[f] = spec[:bins]
        self.lib.vv_dsp_stft_destroy(h)
        return out

    def mel_filterbank(self, n_fft, n_mels, sr, fmin, fmax, variant=0):
        w = C.POINTER(C.c_float)()
        nf, fl = _sz(0), _sz(0)
        st = self.lib.vv_dsp_mel_filterbank_create(_sz(n_fft), _sz(n_mels), C.c_float(sr), C.c_float(fmin), C.c_float(fmax),
                                                   C.c_int(variant), C.byref(w), C.byref(nf), C.byref(fl))
        if st != 0:
            return st, None
        out = np.ctypeslib.as_array(w, shape=(nf.value, fl.value)).copy()
        self.lib.vv_dsp_mel_filterbank_free(w, nf)
        return st, out

    def log_mel(self, power, weights, eps):
        power, weights = _f32(power), _f32(weights)
        out = np.empty((power.shape[0], weights.shape[0]), np.float32)
        st = self.lib.vv_dsp_compute_log_mel_spectrogram(_p(power), _sz(power.shape[0]), _sz(power.shape[1]), _p(weights),
                                                         _sz(weights.shape[0]), C.c_float(eps), _p(out))
        assert st == 0, st
        return out

    def mfcc(self, log_mel, n_coeffs, lifter=0.0, dct_type=2):
        log_mel = _f32(log_mel)
        out = np.empty((log_mel.shape[0], max(n_coeffs, 1)), np.float32)
        st = self.lib.vv_dsp_mfcc(_p(log_mel), _sz(log_mel.shape[0]), _sz(log_mel.shape[1]), _sz(n_coeffs), C.c_int(dct_type),
                                  C.c_float(lifter), _p(out))
        return st, out[:, :n_coeffs]

    def wav_read(self, path):
        """vv_dsp_wav_read (src/audio/wav.c:288-370) -> float32 [channels, samples]"""
        class Info(C.Structure):
            _fields_ = [("num_samples", _sz), ("num_channels", C.c_int), ("sample_rate", C.c_double), ("bit_depth", C.c_int),
                        ("is_float", C.c_int)]
        buf = C.POINTER(C.POINTER(C.c_float))()
        info = Info()
        st = self.lib.vv_dsp_wav_read(str(path).encode(), C.byref(buf), C.byref(info))
        assert st == 0, st
        out = np.stack([np.ctypeslib.as_array(buf[c], shape=(info.num_samples,)).copy() for c in range(info.num_channels)])
        self.lib.vv_dsp_wav_free_buffer(C.byref(buf), info.num_channels)
        return out

    def mfcc_plan_process(self, power, n_fft, n_mels, n_coeffs, sr, fmin, fmax, lifter, eps):
        """vv_dsp_mfcc_init -> vv_dsp_mfcc_process -> vv_dsp_mfcc_destroy (src/features/mel.c:333-461)"""
        power = _f32(power)
        plan = C.c_void_p()
        st = self.lib.vv_dsp_mfcc_init(_sz(n_fft), _sz(n_mels), _sz(n_coeffs), C.c_float(sr), C.c_float(fmin), C.c_float(fmax),
                                       C.c_int(0), C.c_int(2), C.c_float(lifter), C.c_float(eps), C.byref(plan))
        if st != 0:
            return st, None
        out = np.empty((power.shape[0], n_coeffs), np.float32)
        st = self.lib.vv_dsp_mfcc_process(plan, _p(power), _sz(power.shape[0]), _p(out))
        self.lib.vv_dsp_mfcc_destroy(plan)
        return st, out

    def istft(self, spec, nfft, hop, n_out, win="hann", half=True, normalise=True):
        """Per-frame python loop: reconstruct at f*hop, then the caller-side divide."""
        spec = _c64(spec)
        F = spec.shape[0]
        span = (F - 1) * hop + nfft if F else 0
        L = max(span, n_out, 1)
        recon = np.zeros(L, np.float32)
        norm = np.zeros(L, np.float32)
        st, h = self._handle(nfft, hop, win)
        assert st == 0
        full = np.empty(nfft, np.complex64)
        for f in range(F):
            if half:
                full[: nfft // 2 + 1] = spec[f]
                full[nfft // 2 + 1:] = np.conj(spec[f][1: nfft - nfft // 2][::-1])
            else:
                full[:] = spec[f]
            o = recon[f * hop: f * hop + nfft]
            nn = norm[f * hop: f * hop + nfft]
            assert self.lib.vv_dsp_stft_reconstruct(h, _p(full), _p(o), _p(nn)) == 0
        self.lib.vv_dsp_stft_destroy(h)
        if not normalise:
            return recon[:n_out].copy()
        y = np.zeros(n_out, np.float32)
        m = norm[:n_out] > np.float32(1e-12)
        y[m] = recon[:n_out][m] / norm[:n_out][m]
        return y
