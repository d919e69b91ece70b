"""Mirror of `lutgen::identity` (crates/lib/src/identity.rs) over the CUDA C ABI.

Images are numpy uint8 arrays: an `RgbImage` is (h, w, 3), an `RgbaImage` is
(h, w, 4) (any leading shape is accepted as long as the last axis is the
channel and the array is C-contiguous). In-place functions mutate their
argument, as the `&mut RgbaImage` reference functions do.
"""
import ctypes as C

import numpy as np

from . import _native


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _as_lut(hald_clut):
    lut = np.asarray(hald_clut)
    if lut.dtype != np.uint8 or lut.ndim != 3 or lut.shape[2] != 3:
        raise TypeError("hald_clut must be a (n, n, 3) uint8 array (RgbImage)")
    return np.ascontiguousarray(lut)


def _check_lut_level(lut, level):
    """The reference's `get_pixel` indexes out of bounds and panics when the CLUT is smaller than the level asks for
    (identity.rs:48-51); the library would read past the host buffer instead. Shared by every entry that takes a level."""
    level = int(level)
    n = level ** 3
    if 2 <= level <= 33 and lut.shape[:2] != (n, n):
        raise _native.LutgenPanic(f"hald clut is {lut.shape[1]}x{lut.shape[0]}, level {level} needs {n}x{n}")


def _check_rgba(image):
    if not isinstance(image, np.ndarray) or image.dtype != np.uint8 or image.shape[-1] != 4:
        raise TypeError("image must be a (..., 4) uint8 numpy array (RgbaImage)")
    if not image.flags.c_contiguous or not image.flags.writeable:
        raise TypeError("image must be C-contiguous and writeable (it is modified in place)")


def generate(level):
    """identity::generate(level) -> RgbImage   (identity.rs:7-34)"""
    level = int(level)
    if not 0 <= level <= 255:
        raise OverflowError("level is a u8")
    L = _native.lib()
    n = level ** 3
    out = np.empty((n, n, 3), np.uint8) if 2 <= level <= 33 else np.empty((1, 1, 3), np.uint8)
    _native.check(L.lutgen_cuda_identity(level, _ptr(out)))
    return out


def detect_level(hald_clut):
    """identity::detect_level(&RgbImage) -> u8   (identity.rs:85-99); LutgenPanic where the reference asserts."""
    lut = np.asarray(hald_clut)
    h, w = lut.shape[0], lut.shape[1]
    L = _native.load()
    lvl = C.c_uint8(0)
    _native.check(L.lutgen_cuda_detect_level(w, h, C.byref(lvl)))
    return int(lvl.value)


def correct_pixel(input, hald_clut, level):
    """identity::correct_pixel(&[u8;3], &RgbImage, level) -> [u8;3]   (identity.rs:40-52)"""
    return correct_pixels(np.asarray(input, np.uint8).reshape(1, 3), hald_clut, level)[0]


def correct_pixels(colors, hald_clut, level):
    """correct_pixel over an (n, 3) array of colours in one launch."""
    lut = _as_lut(hald_clut)
    _check_lut_level(lut, level)
    cols = np.ascontiguousarray(colors, np.uint8).reshape(-1, 3)
    out = np.empty_like(cols)
    L = _native.lib()
    _native.check(L.lutgen_cuda_correct_pixels(_ptr(cols), cols.shape[0], _ptr(lut), int(level), _ptr(out)))
    return out


def correct_image_with_level(image, hald_clut, level):
    """identity::correct_image_with_level(&mut RgbaImage, &RgbImage, level)   (identity.rs:71-78)"""
    _check_rgba(image)
    lut = _as_lut(hald_clut)
    _check_lut_level(lut, level)
    L = _native.lib()
    _native.check(L.lutgen_cuda_apply_hald(_ptr(image), image.size // 4, _ptr(lut), int(level)))


def correct_image(image, hald_clut):
    """identity::correct_image(&mut RgbaImage, &RgbImage)   (identity.rs:62-65)"""
    level = detect_level(hald_clut)
    correct_image_with_level(image, hald_clut, level)


def correct_images(images, hald_clut, level=None):
    """One CLUT over many RGBA images (the CLI's frame loops, crates/cli/src/main.rs:852-886)."""
    lut = _as_lut(hald_clut)
    if level is None:
        level = detect_level(lut)
    _check_lut_level(lut, level)
    for im in images:
        _check_rgba(im)
    n = len(images)
    if n == 0:
        return
    ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in images])
    sizes = (C.c_size_t * n)(*[im.size // 4 for im in images])
    L = _native.lib()
    _native.check(L.lutgen_cuda_apply_hald_batch(ptrs, sizes, n, _ptr(lut), int(level)))
