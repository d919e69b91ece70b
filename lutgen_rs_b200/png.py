"""PNG files around the path: what `RgbImage::save(path)` / `RgbaImage::save(path)` do for the
reference's callers (crates/cli/src/main.rs:767, 811, 839 `lut.save(&path)`; 862-871 `image.save(&path)`) and what
`load_image` (main.rs:602-613: image::open + to_rgba8 / to_rgb8) does for the PNG files they read, with encoding and
decoding done by the CUDA kernels of csrc/png.cu and csrc/png_decode.cu through the C ABI (lutgen_cuda_encode_png*,
lutgen_cuda_decode_png*). Lossless both ways. No CPU route.
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


def info(data):
    """(width, height, channels, has_alpha) of a PNG file's bytes without decoding it (lutgen_cuda_png_info)."""
    buf = np.frombuffer(data, np.uint8)
    w, h, c, a = C.c_uint32(0), C.c_uint32(0), C.c_int32(0), C.c_int32(0)
    _native.check(_native.load().lutgen_cuda_png_info(_ptr(buf), buf.size, C.byref(w), C.byref(h), C.byref(c), C.byref(a)))
    return w.value, h.value, c.value, bool(a.value)


def decode(data, channels=0):
    """PNG file bytes -> (h, w, c) uint8 array, decoded on the device (`image::open(path)` for a PNG followed by
    to_rgb8() for channels=3, to_rgba8() for channels=4; channels=0 keeps the file's own layout). Raises
    LutgenError for flavours the decoder leaves to the caller (16-bit, interlaced) and for damaged files."""
    buf = np.frombuffer(data, np.uint8)
    w, h, c, _ = info(data)
    oc = int(channels) if channels else c
    out = np.empty((h, w, oc), np.uint8)
    _native.check(_native.lib().lutgen_cuda_decode_png(_ptr(buf), buf.size, int(channels), _ptr(out), out.size))
    return out


def load(path, channels=0):
    """`image::open(path)` for a .png path."""
    with open(path, "rb") as f:
        return decode(f.read(), channels)


def correct_png_file(image_png, hald_clut_png):
    """`lutgen apply --hald-clut clut.png image.png`, file bytes to file bytes (crates/cli/src/main.rs:832-871): both PNGs
    are decoded, the image corrected and re-encoded on the device; only compressed bytes cross PCIe. Returns the PNG
    bytes of the corrected RGBA image."""
    a = np.frombuffer(image_png, np.uint8)
    b = np.frombuffer(hald_clut_png, np.uint8)
    w, h, _, _ = info(image_png)
    n = C.c_size_t(0)
    with _stage_lock:
        out = _staging(bound(w, h, 4))
        _native.check(_native.lib().lutgen_cuda_apply_hald_png_file(_ptr(a), a.size, _ptr(b), b.size, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].tobytes()
