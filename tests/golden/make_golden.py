#!/usr/bin/env python3
"""Regenerate the committed golden fixtures from the read-only reference tree.

Run HERE (the build container, where /root/reference exists); the GPU box only
ever sees the committed outputs:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  palettes.json              the four built-in palettes the path's configs use
                             (crates/palettes/palettes.toml), as [[r,g,b],...]
  gruvbox_dark_hald8.npz     docs/assets/gruvbox-dark-hald-clut.png decoded to
                             raw RGB8 (512x512x3). Written by the reference's own
                             doctest crates/lib/src/lib.rs:51-58:
                             GaussianRemapper::new(GruvboxDark, 96.0, 0, 1.0, false)
                               .par_generate_lut(8)
  catppuccin_mocha_hald10.npz docs/assets/catppuccin-mocha-hald-clut.png (1000x1000x3);
                             CLI defaults of the time: Gaussian RBF shape 128,
                             nearest 16, lum 1.0, preserve false, level 10
  nord_hald10.npz            docs/assets/nord-hald-clut.png, same parameters

  gruvbox-dark-hald-clut.png the same docs/assets file as it is (277 KB): a foreign, single-stream PNG for the
                             PNG reader's tests (tests/test_gpu_png_decode.py), whose expected pixels are the
                             .npz above (PIL's decode of it)

These three images are the ONLY numeric pins the reference holds for the RBF
path; everything else is pinned by the KAT in
crates/cli/src/main.rs:1192-1226 (restated in tests/test_oracle.py).
"""
import json
import os
import shutil
import tomllib

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def hex_to_rgb(h):
    h = h.lstrip("#")
    return [int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16)]


def main():
    with open(os.path.join(REF, "crates/palettes/palettes.toml"), "rb") as f:
        toml = tomllib.load(f)
    names = ["catppuccin_mocha", "gruvbox_dark", "nord", "carburetor"]
    palettes = {n: [hex_to_rgb(c) for c in toml[n]] for n in names}
    with open(os.path.join(HERE, "palettes.json"), "w") as f:
        json.dump(palettes, f, indent=0, separators=(",", ":"))

    assets = {
        "gruvbox_dark_hald8": ("docs/assets/gruvbox-dark-hald-clut.png", 512),
        "catppuccin_mocha_hald10": ("docs/assets/catppuccin-mocha-hald-clut.png", 1000),
        "nord_hald10": ("docs/assets/nord-hald-clut.png", 1000),
    }
    for out, (rel, side) in assets.items():
        img = Image.open(os.path.join(REF, rel)).convert("RGB")
        arr = np.asarray(img, dtype=np.uint8)
        assert arr.shape == (side, side, 3), (out, arr.shape)
        np.savez_compressed(os.path.join(HERE, out + ".npz"), rgb=arr)
        print(out, arr.shape, "written")


def copy_png_fixture():
    shutil.copyfile(os.path.join(REF, "docs/assets/gruvbox-dark-hald-clut.png"), os.path.join(HERE, "gruvbox-dark-hald-clut.png"))


if __name__ == "__main__":
    copy_png_fixture()
    main()
