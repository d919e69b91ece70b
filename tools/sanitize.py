#!/usr/bin/env python3
"""Small driver for compute-sanitizer: one launch of every kernel family at small sizes, host-pointer
API only (no torch, so the tool does not wade through torch's start-up). No result checking here --
that is tests/; this only gives memcheck / racecheck / initcheck something to look at.

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py

(The GPU pool this round ran on has compute-sanitizer closed -- gpurun answers "closed on this pool" -- so the
driver has only been run bare there; it is kept for a box where the tool is available.)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import synthetic_palette  # noqa: E402
from lutgen_rs_b200 import _native, extract, identity, interpolation, png  # noqa: E402


def main():
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    _native.init(0)
    before = _native.launch_count()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (97, 131, 4), dtype=np.uint8)

    # warp-tile kernel: small palette, every functor, batch 1 and batch 4 hand-outs, odd level (image tails)
    lut = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).par_generate_lut(8)
    interpolation.GaussianRemapper(pal, 128.0, 16, 0.7, True).par_generate_lut(5)
    interpolation.ShepardRemapper(pal, 4.0, 8, 1.0, False).par_generate_lut(6)
    interpolation.LinearRemapper(pal, 4, 1.0, False).par_generate_lut(4)
    interpolation.GaussianRemapper(pal, 128.0, 0, 1.0, False).par_generate_lut(4)
    interpolation.NearestNeighborRemapper(pal, 1.0, False).par_generate_lut(8)
    interpolation.NearestNeighborRemapper(pal, 1.3, True).par_generate_lut(5)
    # two-level kernel (33..256 colours), CTA fallback (> 256 colours), exact all-palette route
    interpolation.ShepardRemapper(synthetic_palette(64, 1), 4.0, 16, 1.0, False).par_generate_lut(6)
    interpolation.GaussianRemapper(synthetic_palette(256, 1), 128.0, 16, 1.0, False).par_generate_lut(6)
    interpolation.GaussianRemapper(synthetic_palette(300, 1), 128.0, 16, 1.0, False).par_generate_lut(4)
    interpolation.ShepardRemapper(synthetic_palette(40, 1), 4.0, 0, 1.0, False).par_generate_lut(4)
    # arbitrary image front end and the apply kernel
    im = img.copy()
    interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).remap_image(im)
    im = img.copy()
    interpolation.NearestNeighborRemapper(pal, 1.0, False).remap_image(im)
    im = img.copy()
    identity.correct_image(im, lut)
    identity.generate(5)
    # sampling (Philox normals, cube table, gather) and the blur remapper (both hint variants)
    s = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, 48, 1.0, 42080085, False)
    s.par_generate_lut(4)
    im = img.copy()
    s.remap_image(im)
    b = interpolation.GaussianBlurRemapper(pal, 3.0, 1.0, False)
    b.par_generate_lut(6)
    b.generate_lut(5)
    interpolation.GaussianBlurRemapper(pal, 8.0, 0.7, True).par_generate_lut(4)
    # PNG writer: dynamic-Huffman pieces (CLUT), stored pieces (noise), several pieces, grey and RGBA
    png.encode(lut)
    png.encode(img)
    png.encode(np.zeros((3, 9000, 1), np.uint8))
    png.generate_lut_png(interpolation.NearestNeighborRemapper(pal, 1.0, False), 6)
    png.correct_image_png(img, lut)
    extract.extract_palette(img[:, :, :3], 8, 4)
    print(f"sanitize driver done: {_native.launch_count() - before} kernel launches")


if __name__ == "__main__":
    main()
