"""Host build of the PNG encoder core (tests/png_sim.cpp over lutgen_rs_b200/csrc/png_core.h) and an
independent PNG reader for the tests. TEST INFRASTRUCTURE: the product path never loads this."""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "png_sim.cpp")
CORE = os.path.join(os.path.dirname(HERE), "lutgen_rs_b200", "csrc", "png_core.h")
OUT = os.path.join(HERE, "_build", "libpng_sim.so")
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
        _lib.lpng_sim_bound.restype = C.c_ulonglong
        _lib.lpng_sim_bound.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.lpng_sim_encode.restype = C.c_int
        _lib.lpng_sim_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong),
                                         C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.lpng_sim_shared_bytes.restype = C.c_uint32
        _lib.lpng_sim_set_order.argtypes = [C.c_int]
    return _lib


def set_order(mode):
    """Order in which the threads of each phase run on the host: 0 ascending, 1 descending, 2 scrambled per phase."""
    lib().lpng_sim_set_order(int(mode))


def encode(img):
    """img: (h, w, c) uint8, c in 1..4. Returns (png bytes, per-row filter types, pieces stored uncompressed)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, c = img.shape
    cap = lib().lpng_sim_bound(w, h, c)
    out = np.empty(cap, np.uint8)
    n = C.c_ulonglong(0)
    stored = C.c_uint32(0)
    filters = np.empty(h, np.uint8)
    rc = lib().lpng_sim_encode(img.ctypes.data, w, h, c, out.ctypes.data, cap, C.byref(n), filters.ctypes.data, C.byref(stored))
    assert rc == 0
    return out[:n.value].tobytes(), filters, stored.value


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def decode(data):
    """A PNG reader written from the specification (8-bit, non-interlaced, colour types 0/2/4/6): checks the
    signature, every chunk's CRC, the zlib stream (zlib.decompress verifies Adler-32) and undoes the filters.
    Returns (pixels (h, w, c), filter types, number of IDAT chunks)."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "signature"
    pos, idat, ihdr, nidat, ended = 8, [], None, 0, False
    while pos < len(data):
        ln, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        (crc,) = struct.unpack(">I", data[pos + 8 + ln:pos + 12 + ln])
        assert zlib.crc32(typ + body) == crc, f"CRC of {typ} chunk at {pos}"
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
            nidat += 1
        elif typ == b"IEND":
            ended = True
            assert pos + 12 + ln == len(data), "bytes after IEND"
        pos += 12 + ln
    assert ended and ihdr is not None
    w, h, depth, ctype, comp, flt, lace = ihdr
    assert (depth, comp, flt, lace) == (8, 0, 0, 0)
    c = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    raw = zlib.decompress(b"".join(idat))
    rb = w * c
    assert len(raw) == h * (rb + 1), (len(raw), h * (rb + 1))
    rows = np.frombuffer(raw, np.uint8).reshape(h, rb + 1)
    filters = rows[:, 0].copy()
    res = rows[:, 1:].astype(np.int32)
    out = np.zeros((h, rb), np.int32)
    prev = np.zeros(rb, np.int32)
    for y in range(h):
        ft, line = int(filters[y]), res[y]
        if ft == 0:
            cur = line.copy()
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft == 1:  # each of the c interleaved channels is a running sum
            cur = (np.cumsum(line.reshape(-1, c), axis=0) & 255).reshape(-1)
        else:
            cur = np.zeros(rb, np.int32)
            for x in range(rb):
                a = cur[x - c] if x >= c else 0
                b = prev[x]
                cc = prev[x - c] if x >= c else 0
                pred = ((a + b) >> 1) if ft == 3 else _paeth(int(a), int(b), int(cc))
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
        prev = cur
    return out.astype(np.uint8).reshape(h, w, c), filters, nidat
