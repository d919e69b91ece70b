"""Host build of the palette-extraction kernel bodies (tests/extract_sim.cpp over
lutgen_rs_b200/csrc/extract_core.h). TEST INFRASTRUCTURE: the product path never loads this."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "extract_sim.cpp")
CORE = os.path.join(os.path.dirname(HERE), "lutgen_rs_b200", "csrc", "extract_core.h")
OUT = os.path.join(HERE, "_build", "libextract_sim.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", OUT, SRC])
    return OUT


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
        _lib.lext_sim_points.restype = u32
        _lib.lext_sim_points.argtypes = [vp, sz, u64, vp, vp, vp]
        _lib.lext_sim_kmeans.restype = u32
        _lib.lext_sim_kmeans.argtypes = [u32, vp, vp, vp, u32, u32, u64, vp]
    return _lib


def extract_palette(rgb, color_count, iterations=16, seed=0):
    """The kernel bodies run on the host in an order scrambled by `seed`; the two Oklab conversions by the oracle."""
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    npix = rgb.shape[0]
    cap = min(npix, 1 << 24)
    keys, weights, lst = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty((cap, 3), np.uint8)
    n = lib().lext_sim_points(rgb.ctypes.data, npix, seed, keys.ctypes.data, weights.ctypes.data, lst.ctypes.data)
    lab = np.ascontiguousarray(O.srgb_to_oklab(lst[:n]), np.float32)
    out_lab = np.zeros((256, 3), np.float32)
    k = lib().lext_sim_kmeans(n, keys.ctypes.data, weights.ctypes.data, lab.ctypes.data, int(color_count), int(iterations), seed, out_lab.ctypes.data)
    return O.oklab_to_srgb(out_lab[:k]) if k else np.zeros((0, 3), np.uint8)
