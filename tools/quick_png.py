#!/usr/bin/env python3
"""Shortest possible timing of the PNG writer on the level-16 headline CLUT (device resident, CUDA events)."""
import json
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lutgen_rs_b200 import device, interpolation, png  # noqa: E402

device.init(0)
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
lut = torch.empty((4096, 4096, 3), dtype=torch.uint8, device="cuda")
device.generate_lut_rows_(r, 16, lut.view(-1), 0, 4096)
out = torch.empty(png.bound(4096, 4096, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    v = device.encode_png_(lut, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    v = device.encode_png_(lut, out)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 10:.3f} ms per encode, {v.numel()} bytes, crc32 {zlib.crc32(v.cpu().numpy().tobytes()):08x}")
