"""Builds liblutgen_cuda.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m lutgen_rs_b200.build [--force]

The shared object lands next to this file so it travels to the GPU box with the
repository snapshot. cudart is linked statically: the library needs nothing but
the NVIDIA driver at run time (no torch, no toolkit).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblutgen_cuda.so")
SOURCES = [os.path.join(CSRC, "api.cu")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-cudart", "static",
    "--expt-relaxed-constexpr",
]


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
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-ccbin", "/usr/bin/g++", "-o", OUT, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
