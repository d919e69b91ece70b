#!/usr/bin/env python3
"""Two PNG encodes of the level-16 headline CLUT on the device (for ncu: profile the second one).

    ncu --set full --import-source on --clock-control none -k regex:k_png_deflate -s 1 -c 1 -o gpurun_out/png python tools/prof_png.py
"""
import json
import os
import sys

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
torch.cuda.synchronize()
print(v.numel())
