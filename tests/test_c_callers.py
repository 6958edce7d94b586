"""C callers written the way the reference's own tests / tools use vv-dsp (tests/c/*.c), compiled with gcc against
this repo's include/ and linked with the product library: the drop-in boundary exercised from C, not through ctypes.

  -m "not gpu"  : the callers compile and link (headers declaration-compatible, every symbol exported)
  -m gpu        : they run on the B200 -- known answers of the reference's tests, the per-frame round trip of
                  tools/dump_stft_roundtrip.c:44-54 -- and the per-frame latency of a literal drop-in is printed next
                  to the CPU reference's (oracle/_ref), SURVEY.md section 7 / VERDICT round 1 item 7."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "vv_dsp_b200", "lib")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _build(src, out, libdir, libname, std="c99"):
    cmd = ["gcc", f"-std={std}", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + INC, os.path.join(HERE, "c", src), "-o", out,
           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    from vv_dsp_b200 import build
    build.build()
    d = tmp_path_factory.mktemp("c_callers")
    return {"callers": _build("reference_style_callers.c", str(d / "callers"), LIBDIR, "vvdsp_b200"),
            "latency": _build("perframe_latency.c", str(d / "latency_b200"), LIBDIR, "vvdsp_b200", std="gnu99"),
            "dir": d}


def test_c_callers_compile_and_link(binaries):
    assert os.path.exists(binaries["callers"]) and os.path.exists(binaries["latency"])


def test_reference_header_users_compile():
    """a caller that includes only what the reference's callers include, math macros included (window_tests.c:17,26)"""
    src = ('#include "vv_dsp/vv_dsp.h"\n'
           'int main(void) { vv_dsp_real w[8]; vv_dsp_real c = VV_DSP_COS((vv_dsp_real)(VV_DSP_TWO_PI) / 7);\n'
           '  return vv_dsp_window_hann(8, w) == VV_DSP_OK && c < 1 && VV_DSP_PI > 3 && VV_DSP_SQRT(4.0f) == 2.0f ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + INC, "-fsyntax-only", "-x", "c", "-"], input=src,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_c_callers_run_on_the_gpu(binaries):
    r = subprocess.run([binaries["callers"]], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_per_frame_latency_of_a_literal_drop_in(binaries):
    rows = {}
    for nfft in (512, 2048, 8192):
        r = subprocess.run([binaries["latency"], str(nfft), "300"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        rows[f"b200_{nfft}"] = json.loads(r.stdout)
        assert rows[f"b200_{nfft}"]["roundtrip_max_err"] < 1e-3
    if os.path.exists(os.path.join(REFDIR, "libvvdsp_ref.so")):
        ref = _build("perframe_latency.c", str(binaries["dir"] / "latency_ref"), REFDIR, "vvdsp_ref", std="gnu99")
        for nfft in (512, 2048, 8192):
            r = subprocess.run([ref, str(nfft), "300"], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            rows[f"cpu_reference_{nfft}"] = json.loads(r.stdout)
    print(json.dumps(rows, indent=1))
