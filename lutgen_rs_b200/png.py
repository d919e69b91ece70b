"""PNG writer around the path: what `RgbImage::save(path)` / `RgbaImage::save(path)`
do for the reference's callers (crates/cli/src/main.rs:767, 811, 839 `lut.save(&path)`; 862-871
`image.save(&path)`), with the encoding done by the CUDA kernels of csrc/png.cu through the C ABI
(lutgen_cuda_encode_png*). Lossless: any PNG reader returns the pixels that went in. No CPU route.
"""
import ctypes as C
import threading

import numpy as np

from . import _native
from .identity import _ptr


def _as_image(image):
    a = np.asarray(image)
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise TypeError("image must be a uint8 array of shape (h, w) or (h, w, channels)")
    if a.ndim == 2:
        a = a[:, :, None]
    if not 1 <= a.shape[2] <= 4:
        raise TypeError("1 to 4 channels (grey, grey+alpha, RGB, RGBA)")
    return np.ascontiguousarray(a)


_stage_lock = threading.Lock()
_stage = None  # (pointer, capacity, uint8 view): page-locked landing zone for the file bytes, grown on demand


def _staging(nbytes):
    """A page-locked buffer of at least nbytes (lutgen_cuda_host_alloc): the download of the file is then one
    DMA instead of a staged copy into pageable memory. Call with _stage_lock held."""
    global _stage
    if _stage is None or _stage[1] < nbytes:
        L = _native.lib()
        if _stage is not None:
            L.lutgen_cuda_host_free(_stage[0])
            _stage = None
        p = C.c_void_p()
        _native.check(L.lutgen_cuda_host_alloc(nbytes, C.byref(p)))
        _stage = (p, nbytes, np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value)))
    return _stage[2]


def bound(width, height, channels):
    """Largest file the encoder can produce for such an image (lutgen_cuda_png_bound)."""
    n = C.c_size_t(0)
    _native.check(_native.load().lutgen_cuda_png_bound(int(width), int(height), int(channels), C.byref(n)))
    return n.value


def encode(image):
    """PNG file bytes of a (h, w[, c]) uint8 image (`image::ImageBuffer::save` with a .png path, minus the file)."""
    a = _as_image(image)
    h, w, c = a.shape
    n = C.c_size_t(0)
    with _stage_lock:
        out = _staging(bound(w, h, c))
        _native.check(_native.lib().lutgen_cuda_encode_png(_ptr(a), w, h, c, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].tobytes()


def save(image, path):
    """`image.save(path)` for a .png path."""
    data = encode(image)
    with open(path, "wb") as f:
        f.write(data)


def generate_lut_png(remapper, level, abort=None):
    """`remapper.par_generate_lut(level)` followed by `lut.save(path)` (crates/cli/src/main.rs:756-768) with the
    CLUT never leaving the device: returns the PNG bytes, or None if `abort` was seen set."""
    from .interpolation import _abort_ptr
    level = int(level)
    if not 0 <= level <= 255:
        raise OverflowError("level is a u8")
    side = level ** 3 if 2 <= level <= 33 else 1
    n = C.c_size_t(0)
    with _stage_lock:
        out = _staging(bound(side, side, 3))
        aborted = _native.check(_native.lib().lutgen_cuda_generate_lut_png(remapper._h, level, _ptr(out), out.size, C.byref(n), _abort_ptr(abort)))
        return None if aborted else out[:n.value].tobytes()


def correct_image_png(image, hald_clut, level=None):
    """`identity::correct_image(&mut image, &lut)` followed by `image.save(path)` (what `lutgen apply` does with a
    still image, crates/cli/src/main.rs:852-871) with the frame staying on the device in between: returns the PNG
    bytes of the corrected (h, w, 4) uint8 image; `image` itself is left as it is."""
    from .identity import _as_lut, _check_rgba, detect_level
    lut = _as_lut(hald_clut)
    if level is None:
        level = detect_level(lut)
    _check_rgba(image)
    if image.ndim != 3:
        raise TypeError("image must have shape (h, w, 4)")
    a = np.ascontiguousarray(image)
    h, w = a.shape[0], a.shape[1]
    n = C.c_size_t(0)
    with _stage_lock:
        out = _staging(bound(w, h, 4))
        _native.check(_native.lib().lutgen_cuda_apply_hald_png(_ptr(a), w, h, _ptr(lut), int(level), _ptr(out), out.size, C.byref(n)))
        return out[:n.value].tobytes()
