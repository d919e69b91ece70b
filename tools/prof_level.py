#!/usr/bin/env python3
"""python tools/prof_level.py LEVEL [reps]: device-resident Gaussian RBF CLUT of one level, per-call times."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from lutgen_rs_b200 import device, interpolation
level = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
device.init(0)
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
side = level ** 3
lut = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); device.generate_lut_rows_(r, level, lut, 0, side); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("level", level, "ms:", " ".join(f"{t:.4f}" for t in ts))
