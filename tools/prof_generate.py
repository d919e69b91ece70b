#!/usr/bin/env python3
"""Short driver for profiling: a few device-resident CLUT generations / applies and nothing else.

    python tools/prof_generate.py [generate|apply|shepard256|nn|sampling] [reps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import device, interpolation  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "generate"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    device.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    if what == "apply":
        lut = torch.from_numpy(interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).generate_lut(8)).cuda().reshape(-1)
        src = torch.randint(0, 256, (8, 2160, 3840, 4), dtype=torch.uint8, device="cuda")
        frames = torch.empty_like(src)
        for _ in range(reps):
            frames.copy_(src)
            ev0.record()
            device.apply_hald_(frames, lut, 8)
            ev1.record()
            torch.cuda.synchronize()
            times.append(ev0.elapsed_time(ev1))
    else:
        level = 16
        if what == "shepard256":
            from conftest import synthetic_palette
            r = interpolation.ShepardRemapper(synthetic_palette(256, 1), 4.0, 16, 1.0, False)
        elif what == "nn":
            r = interpolation.NearestNeighborRemapper(pal, 1.0, False)
        elif what == "blur":
            r = None  # a fresh remapper per repetition: the blurred volume is cached per handle
        elif what == "sampling":
            from conftest import synthetic_palette
            r = interpolation.GaussianSamplingRemapper(synthetic_palette(256, 1), 0.0, 20.0, 512, 1.0, 42080085, False)
            level = 12
        else:
            r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
        side = level ** 3
        lut = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
        for _ in range(reps):
            if what == "blur":
                r = interpolation.GaussianBlurRemapper(pal, 8.0, 1.0, False)
            ev0.record()
            device.generate_lut_rows_(r, level, lut, 0, side)
            ev1.record()
            torch.cuda.synchronize()
            times.append(ev0.elapsed_time(ev1))
    print(what, "ms per rep:", " ".join(f"{t:.3f}" for t in times))
    if os.environ.get("LUTGEN_CUDA_STATS"):
        import ctypes
        from lutgen_rs_b200 import _native
        st = np.zeros(128, np.uint32)
        _native.check(_native.lib().lutgen_cuda_debug_stats(ctypes.c_void_p(st.ctypes.data), 1))
        tiles = max(1, int(st[:64].sum()))
        print("tiles", tiles, "(all reps)")
        print("  |UNCERTAIN| histogram:", {i: int(v) for i, v in enumerate(st[:64]) if v})
        print("  |IN| histogram:", {i: int(v) for i, v in enumerate(st[64:126]) if v})
        print("  near-ties settled by the lane itself (f64 distances only):", int(st[126]))
        print("  colours evaluated entirely in the reference's f64 arithmetic:", int(st[127]))


if __name__ == "__main__":
    main()
