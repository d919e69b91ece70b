"""Host build of the PNG decoder core (tests/png_dec_sim.cpp over lutgen_rs_b200/csrc/png_decode_core.h).
TEST INFRASTRUCTURE: the product path never loads this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "png_dec_sim.cpp")
CORE = os.path.join(os.path.dirname(HERE), "lutgen_rs_b200", "csrc", "png_decode_core.h")
OUT = os.path.join(HERE, "_build", "libpng_dec_sim.so")
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
        _lib.lpngd_sim_inflate.restype = C.c_int
        _lib.lpngd_sim_inflate.argtypes = [C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
        _lib.lpngd_sim_decode.restype = C.c_int
        _lib.lpngd_sim_decode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_uint)]
    return _lib


def inflate(z, cap):
    src = np.frombuffer(z, np.uint8)
    dst = np.empty(max(cap, 1), np.uint8)
    n = C.c_ulonglong(0)
    rc = lib().lpngd_sim_inflate(src.ctypes.data, len(z), dst.ctypes.data, cap, C.byref(n))
    return rc, dst[:n.value].tobytes()


def decode(data, out_channels=0):
    """PNG bytes -> (status, (h, w, c) uint8 array or None, path) with path 1 = pieces, 2 = serial."""
    src = np.frombuffer(data, np.uint8)
    info = (C.c_uint * 6)()
    cap = 1 << 28
    # width / height are in the IHDR at a fixed place: size the output from them
    if len(data) >= 24:
        w = int.from_bytes(data[16:20], "big")
        h = int.from_bytes(data[20:24], "big")
        cap = max(16, min(cap, w * h * 4))
    out = np.empty(cap, np.uint8)
    rc = lib().lpngd_sim_decode(src.ctypes.data, len(data), int(out_channels), out.ctypes.data, cap, info)
    if rc != 0:
        return rc, None, 0
    w, h, oc = info[0], info[1], info[4]
    return 0, out[: w * h * oc].reshape(h, w, oc).copy(), int(info[5])
