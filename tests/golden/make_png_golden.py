#!/usr/bin/env python3
"""Writes tests/golden/png_files.json and png_identity_level3.png: digests of the PNG files the encoder core
(lutgen_rs_b200/csrc/png_core.h, run on the host through tests/png_sim.cpp) writes for four small images. They
pin the FORMAT the encoder produces from one round to the next (filter choice, tokenisation, Huffman codes,
header layout, chunking): a change of any of these must be deliberate and come with regenerated digests. The CUDA
path is required to write the same bytes as the host instantiation (tests/test_gpu_png.py), so the digests pin it
too. Run from tests/:  python golden/make_png_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import png_sim as P  # noqa: E402
from conftest import synthetic_palette  # noqa: E402


def images():
    rng = np.random.default_rng(77)
    return {
        "identity_level3": O.identity(3),
        "gaussian_level4_12colours": O.Remapper(O.GAUSSIAN, synthetic_palette(12, seed=5), param=128.0, nearest=8).generate_lut(4),
        "noise_20x30_rgba": rng.integers(0, 256, (20, 30, 4), dtype=np.uint8),
        "flat_150x200_grey": np.full((150, 200, 1), 9, np.uint8),
    }


def main():
    cases = {}
    for name, img in images().items():
        data = P.encode(img)[0]
        cases[name] = {"shape": list(img.shape), "pixels_sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
                       "png_bytes": len(data), "png_sha256": hashlib.sha256(data).hexdigest()}
    with open(os.path.join(HERE, "png_files.json"), "w") as f:
        f.write(json.dumps(cases, indent=1) + "\n")
    with open(os.path.join(HERE, "png_identity_level3.png"), "wb") as f:
        f.write(P.encode(images()["identity_level3"])[0])
    print(json.dumps(cases, indent=1))


if __name__ == "__main__":
    main()
