#!/usr/bin/env python3
"""Times the PNG writer on the device (level-16 Gaussian CLUT, a photo-like 4K RGBA frame) and against PIL
(zlib level 1) on the host. Development aid; bench.py reports the same figures in `other_configs`."""
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lutgen_rs_b200 import device, interpolation, png  # noqa: E402


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


def main():
    device.init(0)
    pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    lut = torch.empty((4096, 4096, 3), dtype=torch.uint8, device="cuda")
    device.generate_lut_rows_(r, 16, lut.view(-1), 0, 4096)
    out = torch.empty(png.bound(4096, 4096, 3), dtype=torch.uint8, device="cuda")
    ms, view = timed(lambda: device.encode_png_(lut, out))
    print(f"level-16 CLUT: {ms:.3f} ms per encode on the device, {view.numel()} bytes ({view.numel() / lut.numel():.3f} of raw), "
          f"{lut.numel() / ms / 1e6:.1f} GB/s of pixels")
    ref = view.cpu().numpy().tobytes()
    for split in ("1", "2", "4", "8"):
        os.environ["LUTGEN_CUDA_PNG_SPLIT"] = split
        ms, view = timed(lambda: device.encode_png_(lut, out), 6)
        print(f"  LUTGEN_CUDA_PNG_SPLIT={split}: {ms:.3f} ms, same bytes: {view.cpu().numpy().tobytes() == ref}")
    del os.environ["LUTGEN_CUDA_PNG_SPLIT"]
    t = time.perf_counter()
    data = png.generate_lut_png(r, 16)
    t1 = time.perf_counter() - t
    t = time.perf_counter()
    data = png.generate_lut_png(r, 16)
    t2 = time.perf_counter() - t
    print(f"generate_lut_png(16) host to host: first {t1 * 1e3:.2f} ms, second {t2 * 1e3:.2f} ms, {len(data)} bytes")
    host = lut.cpu().numpy()
    from PIL import Image
    t = time.perf_counter()
    b = io.BytesIO()
    Image.fromarray(host).save(b, "PNG", compress_level=1)
    print(f"PIL compress_level=1 on the host: {(time.perf_counter() - t) * 1e3:.0f} ms, {len(b.getvalue())} bytes")
    assert (np.array(Image.open(io.BytesIO(data))) == host).all()
    h, w = 2160, 3840
    yy, xx = np.mgrid[0:h, 0:w]
    rng = np.random.default_rng(9)
    base = np.stack([128 + 100 * np.sin(xx / 90.0 + yy / 130.0), 128 + 90 * np.cos(xx / 70.0), 128 + 80 * np.sin(yy / 50.0)], -1)
    frame = np.clip(base + rng.normal(0, 2, (h, w, 3)), 0, 255).astype(np.uint8)
    frame = np.concatenate([frame, np.full((h, w, 1), 255, np.uint8)], -1)
    ft = torch.from_numpy(frame).cuda()
    out2 = torch.empty(png.bound(w, h, 4), dtype=torch.uint8, device="cuda")
    ms, view = timed(lambda: device.encode_png_(ft, out2))
    print(f"4K RGBA frame: {ms:.3f} ms per encode, {view.numel()} bytes ({view.numel() / ft.numel():.3f} of raw)")


if __name__ == "__main__":
    main()
