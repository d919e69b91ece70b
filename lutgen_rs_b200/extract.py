"""Palette extraction next to the path: the role of `quantette`'s
`PalettePipeline::try_from(&image).colorspace(Oklab).quantize_method(kmeans()).palette_size(n).palette_par()` in
`lutgen extract` (crates/cli/src/main.rs:772-815), done by CUDA kernels through lutgen_cuda_extract_palette.

Not quantette's palette (that crate is not vendored under the reference and nothing there pins its output): a
deterministic k-means in Oklab, specified in csrc/extract_core.h and bit exact with its restatement in oracle/.
No CPU route.
"""
import ctypes as C

import numpy as np

from . import _native
from .identity import _ptr

DEFAULT_ITERATIONS = 16


def extract_palette(image, color_count, iterations=DEFAULT_ITERATIONS):
    """Up to `color_count` (1..=256) colours for an RGB8 image ((h, w, 3) or (n, 3) uint8): a (k, 3) uint8 array,
    k <= color_count (fewer when the image has fewer distinct colours)."""
    a = np.asarray(image)
    if a.dtype != np.uint8 or a.ndim not in (2, 3) or a.shape[-1] != 3:
        raise TypeError("image must be a (h, w, 3) or (n, 3) uint8 array (RgbImage)")
    a = np.ascontiguousarray(a).reshape(-1, 3)
    color_count, iterations = int(color_count), int(iterations)
    if color_count < 0 or iterations < 0:
        raise OverflowError("color_count and iterations are unsigned")
    out = np.empty((max(color_count, 1), 3), np.uint8)
    n = C.c_uint32(0)
    _native.check(_native.lib().lutgen_cuda_extract_palette(_ptr(a), a.shape[0], color_count, iterations, _ptr(out), C.byref(n)))
    return out[:n.value].copy()
