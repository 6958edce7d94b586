"""CPU tests (-m "not gpu"): the product library builds, loads, exports every symbol that
include/*.h declares, fails LOUDLY (no CPU fallback) when there is no CUDA device, and its
pure-host helpers (windows, framing) are bit-exact against the oracle.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product():
    from vv_dsp_b200 import Library, build
    return Library(build.build())


def _declared_symbols():
    names = set()
    for d, _, fs in os.walk(os.path.join(ROOT, "include")):
        for f in fs:
            src = open(os.path.join(d, f)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(vv_dsp_[a-z0-9_]+|vvb_[a-z0-9_]+)\s*\(", src))
    names -= {"vv_dsp_cpx_make", "vv_dsp_static_assert_"}       # static inline / macro
    return names


def test_exports_every_declared_symbol(product):
    out = subprocess.run(["nm", "-D", "--defined-only", product.path], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = _declared_symbols()
    assert len(declared) > 50
    missing = sorted(declared - exported)
    assert not missing, missing
    # the reference's internal backend symbols must NOT be exported
    assert not [s for s in exported if s.startswith("vv_dsp_fft_backend_")]


def test_library_is_sm100a_only(product):
    out = subprocess.run(["cuobjdump", "-lelf", product.path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_device_fails_loudly(product):
    """Without a GPU the library must refuse (UNSUPPORTED), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vv_dsp_b200.api import StftParams, VvDspError, Stft, FftPlan
    h = C.c_void_p()
    p = StftParams(2048, 512, 1)
    assert product.vv_dsp_stft_create(C.byref(p), C.byref(h)) == 6
    assert not h.value and "CUDA" in product.last_error()
    with pytest.raises(VvDspError):
        Stft(256, 64, "hann", lib=product)
    with pytest.raises(VvDspError):
        FftPlan(16, lib=product)
    assert product.kernel_launches() == 0


def test_missing_library_is_an_error(tmp_path):
    from vv_dsp_b200 import Library
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Library(str(tmp_path / "libvvdsp_b200.so"))


def test_validation_precedes_device_use(product):
    import parity_cases as pc
    import torch
    if not torch.cuda.is_available():
        # argument errors are reported with the reference's codes even when no device exists
        from vv_dsp_b200.api import StftParams
        h = C.c_void_p()
        for (nfft, hop, win), want in (((0, 1, 1), 2), ((8, 0, 1), 2), ((8, 9, 1), 2), ((8, 4, 7), 3)):
            assert product.vv_dsp_stft_create(C.byref(StftParams(nfft, hop, win)), C.byref(h)) == want
        assert product.vv_dsp_stft_create(None, C.byref(h)) == 1
        plan = C.c_void_p()
        assert product.vv_dsp_fft_make_plan(0, 0, 1, C.byref(plan)) == 2
        assert product.vv_dsp_fft_make_plan(8, 5, 1, C.byref(plan)) == 3
        assert product.vv_dsp_fft_set_backend(1) == 6 and product.vv_dsp_fft_set_backend(3) == 3
    else:
        pc.check_status_codes(product)


def test_host_windows_bit_exact(product, oracle):
    from vv_dsp_b200 import window
    for kind in ("boxcar", "hann", "hamming"):
        for n in (1, 2, 8, 17, 1024, 2048, 8192):
            st, w = window(kind, n, lib=product)
            so, wo = oracle.window(kind, n)
            assert st == so == 0 and w.tobytes() == wo.tobytes(), (kind, n)
    assert window("hann", 0, lib=product)[0] == 2
    assert product.vv_dsp_window_hann(8, None) == 1


def test_host_framing_bit_exact(product, oracle):
    from vv_dsp_b200 import fetch_frame, get_num_frames, overlap_add
    for args in ((1024, 256, 128, 0), (1024, 256, 128, 1), (100, 256, 128, 0), (100, 256, 128, 1), (1024, 256, 0, 0)):
        assert get_num_frames(*args, lib=product) == oracle.num_frames(args[0], args[1], args[2], "center" if args[3] else "valid")
    rng = np.random.default_rng(9)
    for n, flen, hop in ((10, 4, 2), (37, 16, 5), (5, 32, 8), (3, 64, 16)):
        x = rng.standard_normal(n).astype(np.float32)
        w = rng.standard_normal(flen).astype(np.float32)
        for idx in range(0, 9):
            for center in (0, 1):
                for win in (None, w):
                    a = fetch_frame(x, flen, hop, idx, center, win, lib=product)
                    b = oracle.fetch_frame(x, flen, hop, idx, bool(center), win)
                    assert a[0] == b[0] and a[1].tobytes() == b[1].tobytes(), (n, flen, hop, idx, center)
    out = np.zeros(8, np.float32)
    for i, fr in enumerate([[1, 2, 3, 4], [3, 4, 5, 6], [5, 6, 7, 8]]):
        assert overlap_add(np.array(fr, np.float32), out, 2, i, lib=product) == 0
    assert out.tolist() == [1, 2, 6, 8, 10, 12, 7, 8]                 # tests/framing_tests.c:159-192
    assert product.vv_dsp_overlap_add(None, None, 1, 1, 1, 0) == 1
    assert fetch_frame(np.zeros(4, np.float32), 0, 2, 0, lib=product)[0] == 2
    a = np.arange(5, dtype=np.float32); o = np.empty(5, np.float32)
    assert product.vv_dsp_vectorized_window_apply(a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                                  o.ctypes.data_as(C.c_void_p), 5) == 0
    assert o.tolist() == [0, 1, 4, 9, 16]
    assert product.vv_dsp_vectorized_window_apply(a.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                                  o.ctypes.data_as(C.c_void_p), 0) == 1
