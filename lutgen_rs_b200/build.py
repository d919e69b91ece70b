"""Builds liblutgen_cuda.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m lutgen_rs_b200.build [--force]

The shared object lands next to this file so it travels to the GPU box with the
repository snapshot. cudart is linked statically: the library needs nothing but
the NVIDIA driver at run time (no torch, no toolkit).

Four translation units: api.cu (the C ABI and most kernels), blur.cu (the GaussianBlurRemapper
kernels, compiled with --fmad=false: see csrc/blur_launch.h), png.cu (the PNG writer) and png_decode.cu (the PNG reader).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblutgen_cuda.so")
OBJ_DIR = os.path.join(HERE, "_obj")
# (source, extra nvcc flags)
UNITS = [("api.cu", []), ("blur.cu", ["--fmad=false"]), ("png.cu", []), ("png_decode.cu", [])]
SOURCES = [os.path.join(CSRC, u) for u, _ in UNITS]

COMMON_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "--expt-relaxed-constexpr",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]


def _deps():
    files = list(SOURCES)
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h", ".inc", ".cu")):
                files.append(os.path.join(d, f))
    return files


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build liblutgen_cuda.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs, objs = [], []
    for unit, flags in UNITS:
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        cmd = [nvcc, *COMMON_FLAGS, *flags, *extra, "-ccbin", "/usr/bin/g++", "-c", "-o", obj, os.path.join(CSRC, unit)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, *LINK_FLAGS, "-ccbin", "/usr/bin/g++", "-o", OUT, *objs])
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
