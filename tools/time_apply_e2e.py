#!/usr/bin/env python3
"""Host-to-host time of identity.correct_images on four pinned 4K RGBA frames (bench.py's apply `e2e`) for several
chunk sizes of the upload / kernel / download pipeline (LUTGEN_CUDA_APPLY_CHUNK, pixels; 0 = whole frames).

    python tools/time_apply_e2e.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import device, identity, interpolation  # noqa: E402

device.init(0)
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
lut = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).generate_lut(8)
nframes, h, w = 4, 2160, 3840
pinned = torch.empty((nframes, h, w, 4), dtype=torch.uint8).pin_memory()
src = torch.randint(0, 256, (nframes, h, w, 4), dtype=torch.uint8)
frames = [pinned[i].numpy() for i in range(nframes)]
want = None
for chunk in (0, 1 << 19, 1 << 20, 1 << 21, 1 << 22, None):
    if chunk is None:
        os.environ.pop("LUTGEN_CUDA_APPLY_CHUNK", None)
    else:
        os.environ["LUTGEN_CUDA_APPLY_CHUNK"] = str(chunk)
    ts = []
    for rep in range(8):
        pinned.copy_(src)
        t = time.perf_counter()
        identity.correct_images(frames, lut, 8)
        ts.append(time.perf_counter() - t)
    out = pinned.numpy().copy()
    if want is None:
        want = out
    ts = sorted(ts[2:])
    print(f"chunk {chunk if chunk is not None else 'default'}: median {ts[len(ts) // 2] * 1e3:.3f} ms, min {ts[0] * 1e3:.3f} ms, "
          f"{nframes * h * w / ts[len(ts) // 2] / 1e9:.2f} Gpix/s, same bytes as whole frames: {bool((out == want).all())}")
# one frame through lutgen_cuda_apply_hald
one = frames[0]
for chunk in (0, None):
    if chunk is None:
        os.environ.pop("LUTGEN_CUDA_APPLY_CHUNK", None)
    else:
        os.environ["LUTGEN_CUDA_APPLY_CHUNK"] = str(chunk)
    ts = []
    for rep in range(8):
        pinned[0].copy_(src[0])
        t = time.perf_counter()
        identity.correct_image(one, lut)
        ts.append(time.perf_counter() - t)
    ts = sorted(ts[2:])
    print(f"one 4K frame, chunk {chunk if chunk is not None else 'default'}: median {ts[len(ts) // 2] * 1e3:.3f} ms")
