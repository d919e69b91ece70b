#!/usr/bin/env python3
"""Decode timings of the PNG reader: a level-16 CLUT written by this library (piece path) and by PIL (serial path).

    python tools/time_png_decode.py [own|foreign|both] [level]
"""
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

from lutgen_rs_b200 import interpolation, png  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
level = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
lut = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).par_generate_lut(level)
files = {}
if what in ("own", "both"):
    files["own writer (piece path)"] = png.encode(lut)
if what in ("foreign", "both"):
    bio = io.BytesIO()
    Image.fromarray(lut).save(bio, "PNG", compress_level=1)
    files["PIL level 1 (serial path)"] = bio.getvalue()
for name, data in files.items():
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        out = png.decode(data, 3)
        ts.append(time.perf_counter() - t)
    assert (out == lut).all()
    t = time.perf_counter()
    ref = np.array(Image.open(io.BytesIO(data)))
    tp = time.perf_counter() - t
    print(f"{name}: {len(data)} bytes -> {lut.nbytes} bytes of pixels: {' '.join('%.2f ms' % (x * 1e3) for x in ts)}; PIL decodes the same file in {tp * 1e3:.1f} ms")
