#!/usr/bin/env python3
"""Host-side cost of what one rank of an N-process job does per end-to-end step (bench.py's e2e at N > 1), on one
GPU: build the remapper, generate part 0 of N of a level-16 CLUT into page-locked host memory, drop the remapper.

    python tools/time_e2e_part.py [reps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import device, interpolation  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    device.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    side = 4096
    img = torch.empty((side, side, 3), dtype=torch.uint8).pin_memory().numpy()
    for parts in (1, 2, 4, 8):
        tc = tg = td = 0.0
        for i in range(reps + 5):
            t0 = time.perf_counter()
            r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
            t1 = time.perf_counter()
            r.generate_lut_part(16, img, 0, parts)
            t2 = time.perf_counter()
            del r
            t3 = time.perf_counter()
            if i >= 5:
                tc += t1 - t0
                tg += t2 - t1
                td += t3 - t2
        print(f"part 0 of {parts}: construct {tc / reps * 1e6:7.1f} us, generate + copy {tg / reps * 1e6:7.1f} us "
              f"({3 * side * side / parts / 1e6:.1f} MB to the host), drop {td / reps * 1e6:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
