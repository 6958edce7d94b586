"""A/B builds: python benchmarks/build_variant.py NAME "-DVVB_X=1 ..." tu1.cu [tu2.cu ...]

Recompiles only the named translation units of vv_dsp_b200/csrc/cuda with the extra nvcc flags and
links them with the standard objects into vv_dsp_b200/lib/libvvdsp_b200_NAME.so (select it at run
time with VVDSP_B200_LIB=<path>).  Every kernel family lives in its own translation unit, so a
variant changes the code of exactly the kernels it names."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vv_dsp_b200 import build as b  # noqa: E402


def main():
    name, flags, tus = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
    b.build()
    objdir = os.path.join(b.PKG, "lib", "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for o in sorted(os.listdir(b.OBJ)):
        if not o.endswith(".o"):
            continue
        src = o[:-2]
        if src in tus:
            out = os.path.join(objdir, o)
            cmd = [b.NVCC, *b.ARCH, "-std=c++17", "-O3", "-lineinfo", *flags, "-Xptxas", "-warn-spills", "-Xcompiler", "-fPIC",
                   "-I" + b.INC, "-c", os.path.join(b.CUDA_DIR, src), "-o", out]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            objs.append(out)
        else:
            objs.append(os.path.join(b.OBJ, o))
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip():
            print(f"[{src}]\n{out}")
        if p.returncode != 0:
            raise SystemExit(f"{src}: nvcc failed")
    lib = os.path.join(b.PKG, "lib", f"libvvdsp_b200_{name}.so")
    subprocess.run([b.NVCC, *b.ARCH, "-shared", "-cudart", "static", "-o", lib, *objs, "-lm"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
