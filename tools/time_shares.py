#!/usr/bin/env python3
"""How the headline kernel's time falls with the share of the CLUT one GPU computes (what strong scaling over N GPUs
leaves per GPU): level-16 Gaussian CLUT, rows [0, side/f) for f = 1, 2, 4 ... with and without an L2 flush before.

    python tools/time_shares.py [reps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import device, interpolation  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    device.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    level = 16
    side = level ** 3
    lut = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        device.generate_lut_rows_(r, level, lut, 0, side)
    torch.cuda.synchronize()
    for f in (1, 2, 4, 8, 16, 32, 64, 256):
        rows = side // f
        out = []
        for flush in (True, False):
            for r0 in (0, (side // 2 // rows) * rows if f > 1 else 0):
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
                for a, b in evs:
                    if flush:
                        flush_buf.fill_(1)
                    a.record()
                    device.generate_lut_rows_(r, level, lut, r0, r0 + rows)
                    b.record()
                torch.cuda.synchronize()
                t = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
                out.append(f"{'flush' if flush else 'warm '} rows@{r0:4d}: {np.median(t):7.1f} us (min {t.min():6.1f})")
        print(f"1/{f:<3d} ({rows:4d} rows, {rows * side // 128:6d} tiles): " + " | ".join(out), flush=True)
    # an empty-ish launch: event-to-event time of the smallest possible share
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        b.record()
    torch.cuda.synchronize()
    print("back-to-back events:", np.median([a.elapsed_time(b) for a, b in evs]) * 1e3, "us")


if __name__ == "__main__":
    main()
