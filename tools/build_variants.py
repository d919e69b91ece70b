#!/usr/bin/env python3
"""Builds experiment variants of liblutgen_cuda.so (api.cu recompiled with extra -D flags, the other objects reused)
as tools/_variant_<name>.so -- git-ignored, but they travel to the GPU box -- so that ONE gpurun call can time them all
(tools/time_variants.py; the library to load is chosen with LUTGEN_CUDA_LIB).

    python tools/build_variants.py name=-DFOO=1,-DBAR=2 other=-DBAZ ...
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lutgen_rs_b200 import build as B  # noqa: E402


def one(spec):
    name, _, flags = spec.partition("=")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(B.OBJ_DIR, f"api_{name}.o")
    out = os.path.join(ROOT, "tools", f"_variant_{name}.so")
    nvcc = "/usr/local/cuda/bin/nvcc"
    log = subprocess.run([nvcc, *B.COMMON_FLAGS, "-Xptxas=-v", *flags, "-ccbin", "/usr/bin/g++", "-c", "-o", obj, os.path.join(B.CSRC, "api.cu")],
                         capture_output=True, text=True)
    if log.returncode:
        return name, "FAILED\n" + log.stderr[-2000:]
    others = [os.path.join(B.OBJ_DIR, u.replace(".cu", ".o")) for u, _ in B.UNITS if u != "api.cu"]
    subprocess.check_call([nvcc, *B.LINK_FLAGS, "-ccbin", "/usr/bin/g++", "-o", out, obj, *others])
    # registers / spills of the headline instantiation
    lines = log.stderr.splitlines()
    info = ""
    for i, l in enumerate(lines):
        if "k_remap_warpILi0ELi0ELi4ELi0ELi1" in l and "Compiling" in l:
            info = " | ".join(x.strip() for x in lines[i + 1:i + 4])
            break
    return name, info


if __name__ == "__main__":
    B.build()
    with ThreadPoolExecutor(4) as ex:
        for name, info in ex.map(one, sys.argv[1:]):
            print(name, "::", info)
