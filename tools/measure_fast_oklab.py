#!/usr/bin/env python3
"""Exhaustive error of the fast f32 sRGB -> Oklab conversion (cbrt2: MUFU seed + first-order correction) against the
reference's f32 arithmetic (correctly rounded cube root), over all 2^24 colours: per component and as a vector norm.
The domain is finite and the kernels deterministic, so these maxima are bounds, not estimates -- they are what the
near-tie tolerance (WT_TIE_SLOPE) and the box padding stand on.

    python tools/measure_fast_oklab.py            # on the GPU box
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lutgen_rs_b200 import _native  # noqa: E402

_native.init(0)
L = _native.lib()
worst = np.zeros(3)
worst_norm = 0.0
worst_rel_l = 0.0
step = 1 << 22
for start in range(0, 1 << 24, step):
    i = np.arange(start, start + step, dtype=np.uint32)
    rgb = np.stack([i & 255, (i >> 8) & 255, i >> 16], 1).astype(np.uint8)
    fast = np.empty((step, 3), np.float32)
    exact = np.empty((step, 3), np.float32)
    _native.check(L.lutgen_cuda_srgb_to_oklab_fast(C.c_void_p(rgb.ctypes.data), step, C.c_void_p(fast.ctypes.data)))
    _native.check(L.lutgen_cuda_srgb_to_oklab(C.c_void_p(rgb.ctypes.data), step, C.c_void_p(exact.ctypes.data)))
    d = fast.astype(np.float64) - exact.astype(np.float64)
    worst = np.maximum(worst, np.abs(d).max(0))
    worst_norm = max(worst_norm, float(np.sqrt((d * d).sum(1)).max()))
    worst_rel_l = max(worst_rel_l, float((np.abs(d[:, 0]) / np.maximum(exact[:, 0].astype(np.float64), 1e-3)).max()))
print("max |fast - exact| per component (L, a, b):", " ".join(f"{x:.3e}" for x in worst))
print(f"max |fast - exact| as a vector: {worst_norm:.3e}")
print(f"max relative error of L (L >= 1e-3): {worst_rel_l:.3e}")
