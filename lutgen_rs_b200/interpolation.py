"""Mirror of `lutgen::interpolation` + the `GenerateLut` blanket impl
(crates/lib/src/interpolation/*.rs, crates/lib/src/lib.rs:113-171) over the CUDA C ABI.

Same constructor argument order and meaning as the Rust types. `remap_*`
mutate the (h, w, 4) uint8 image in place; `generate_lut*` return a new
(level^3, level^3, 3) uint8 array, or None when interrupted, like the
reference's `Option<RgbImage>`. The sequential and `par_` spellings are the same
GPU call (rayon has no meaning here); both are kept so callers do not change.
"""
import ctypes as C

import numpy as np

from . import _native
from .identity import _check_rgba, _ptr


class AbortFlag:
    """Stand-in for the reference's `Arc<AtomicBool>`: one byte another thread may set."""

    def __init__(self, value=False):
        self._b = C.c_uint8(1 if value else 0)

    def set(self, value=True):
        self._b.value = 1 if value else 0

    def load(self):
        return bool(self._b.value)

    @property
    def _as_parameter_(self):
        return C.addressof(self._b)

    def _addr(self):
        return C.c_void_p(C.addressof(self._b))


def _abort_ptr(abort):
    if abort is None:
        return None
    if isinstance(abort, AbortFlag):
        return abort._addr()
    raise TypeError("abort must be an AbortFlag")


class InterpolatedRemapper:
    """`InterpolatedRemapper` + `GenerateLut` (interpolation/mod.rs:23-61, lib.rs:113-171)."""

    _kind = None

    def _create(self, palette, param=0.0, nearest=0, lum_factor=1.0, preserve=False, mean=0.0, std_dev=0.0,
                iterations=0, seed=0):
        pal = np.ascontiguousarray(palette, np.uint8).reshape(-1, 3)
        for name, v in (("nearest", nearest), ("iterations", iterations), ("seed", seed)):
            if int(v) < 0:
                raise OverflowError(f"{name} is unsigned")
        self._palette = pal
        desc = _native.RemapperDesc(self._kind, 1 if preserve else 0, pal.ctypes.data, pal.shape[0], float(param),
                                    int(nearest), float(lum_factor), float(mean), float(std_dev), int(iterations),
                                    int(seed) & 0xFFFFFFFFFFFFFFFF)
        L = _native.lib()
        h = C.c_void_p()
        _native.check(L.lutgen_cuda_remapper_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.iterations = int(iterations)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _native.load().lutgen_cuda_remapper_destroy(h)
            except Exception:
                pass

    # -- InterpolatedRemapper ------------------------------------------------
    def remap_pixel(self, pixel):
        """remap_pixel(&self, &mut Rgba<u8>) (mod.rs:25): `pixel` is a writable uint8[4] array."""
        px = np.asarray(pixel)
        if px.dtype != np.uint8 or px.shape != (4,) or not px.flags.writeable or not px.flags.c_contiguous:
            raise TypeError("pixel must be a writable contiguous uint8[4] array")
        _native.check(_native.lib().lutgen_cuda_remap_image(self._h, _ptr(px), 1, None))

    def remap_image(self, image):
        _check_rgba(image)
        _native.check(_native.lib().lutgen_cuda_remap_image(self._h, _ptr(image), image.size // 4, None))

    def remap_image_with_interrupt(self, image, abort):
        _check_rgba(image)
        _native.check(_native.lib().lutgen_cuda_remap_image(self._h, _ptr(image), image.size // 4, _abort_ptr(abort)))

    par_remap_image = remap_image
    par_remap_image_with_interrupt = remap_image_with_interrupt

    # -- GenerateLut ---------------------------------------------------------
    # The sequential and the par_* spellings (lib.rs:113-132) bind two entry points: they only differ
    # for GaussianBlurRemapper, whose sequential variant carries its hint across rows.
    def generate_lut(self, level, out=None):
        return self._generate(level, None, out, sequential=True)

    def generate_lut_with_interrupt(self, level, abort, out=None):
        return self._generate(level, abort, out, sequential=True)

    def par_generate_lut(self, level, out=None):
        return self._generate(level, None, out, sequential=False)

    def par_generate_lut_with_interrupt(self, level, abort, out=None):
        return self._generate(level, abort, out, sequential=False)

    def generate_lut_part(self, level, out, part_index, part_count, abort=None):
        """One process per GPU, CLUT wanted on the host (lutgen_cuda_generate_lut_part): this rank's share of the row
        groups, generated on its GPU and copied to their places in `out`, the WHOLE (n, n, 3) image -- e.g. a shared-memory
        buffer all ranks have mapped. Complete when every rank has returned."""
        level = int(level)
        n = level ** 3 if 2 <= level <= 33 else 1
        if out.dtype != np.uint8 or out.size != 3 * n * n or not out.flags.c_contiguous:
            raise TypeError("out must be a contiguous uint8 array of 3*level^6 bytes")
        aborted = _native.check(_native.lib().lutgen_cuda_generate_lut_part(self._h, level, _ptr(out), int(part_index), int(part_count),
                                                                           _abort_ptr(abort)))
        return None if aborted else out

    @staticmethod
    def part_rows(level, part_index, part_count):
        """The row ranges [(r0, r1), ...] of the (n, n, 3) image that generate_lut_part(level, out, part_index, part_count)
        writes (without an abort flag): row groups of 16 * level rows dealt to the parts in turn. For callers that place
        `out` themselves -- e.g. touch their own rows of a shared buffer first, so that those pages are allocated on
        the memory node next to their GPU."""
        level, part_index, part_count = int(level), int(part_index), int(part_count)
        n = level ** 3
        group = 16 * level
        ngroups = (n + group - 1) // group
        chunks = min(ngroups, part_count * (2 if ngroups >= 2 * part_count else 1))
        rows = []
        for c in range(part_index, chunks, part_count):
            rows.append((min(n, ngroups * c // chunks * group), min(n, ngroups * (c + 1) // chunks * group)))
        return rows

    def _generate(self, level, abort, out, sequential):
        """`out` (optional, not in the reference): a preallocated (n, n, 3) uint8 array to fill,
        e.g. page-locked memory so the device-to-host copy is a single DMA."""
        level = int(level)
        if not 0 <= level <= 255:
            raise OverflowError("level is a u8")
        n = level ** 3 if 2 <= level <= 33 else 1
        if out is None:
            out = np.empty((n, n, 3), np.uint8)
        elif out.dtype != np.uint8 or out.size != 3 * n * n or not out.flags.c_contiguous:
            raise TypeError("out must be a contiguous uint8 array of 3*level^6 bytes")
        L = _native.lib()
        fn = L.lutgen_cuda_generate_lut_sequential if sequential else L.lutgen_cuda_generate_lut
        aborted = _native.check(fn(self._h, level, _ptr(out), _abort_ptr(abort)))
        return None if aborted else out


class GaussianRemapper(InterpolatedRemapper):
    """GaussianRemapper::new(palette, shape, nearest, lum_factor, preserve_lum)  (rbf.rs:171-180)"""
    _kind = _native.GAUSSIAN

    def __init__(self, palette, shape, nearest, lum_factor, preserve_lum):
        self._create(palette, param=shape, nearest=nearest, lum_factor=lum_factor, preserve=preserve_lum)


class ShepardRemapper(InterpolatedRemapper):
    """ShepardRemapper::new(palette, power, nearest, lum_factor, preserve_lum)  (rbf.rs:161-169)"""
    _kind = _native.SHEPARD

    def __init__(self, palette, power, nearest, lum_factor, preserve_lum):
        self._create(palette, param=power, nearest=nearest, lum_factor=lum_factor, preserve=preserve_lum)


class LinearRemapper(InterpolatedRemapper):
    """LinearRemapper::new(palette, nearest, lum_factor, preserve_lum)  (rbf.rs:153-159)"""
    _kind = _native.LINEAR

    def __init__(self, palette, nearest, lum_factor, preserve_lum):
        self._create(palette, nearest=nearest, lum_factor=lum_factor, preserve=preserve_lum)


class NearestNeighborRemapper(InterpolatedRemapper):
    """NearestNeighborRemapper::new(palette, lum_factor, preserve)  (nearest_neighbor.rs:19-37)"""
    _kind = _native.NEAREST

    def __init__(self, palette, lum_factor, preserve):
        self._create(palette, lum_factor=lum_factor, preserve=preserve)


class GaussianSamplingRemapper(InterpolatedRemapper):
    """GaussianSamplingRemapper::new(palette, mean, std_dev, iterations, lum_factor, seed, preserve)
    (gaussian_sample.rs:27-46)"""
    _kind = _native.SAMPLING

    def __init__(self, palette, mean, std_dev, iterations, lum_factor, seed, preserve):
        self._create(palette, mean=mean, std_dev=std_dev, iterations=iterations, lum_factor=lum_factor, seed=seed,
                     preserve=preserve)

    def noise(self):
        """The iterations x 4 normals every pixel sees (the library's own Philox stream)."""
        out = np.empty((self.iterations, 4), np.float64)
        _native.check(_native.lib().lutgen_cuda_sampling_get_noise(self._h, _ptr(out)))
        return out

    def set_noise(self, noise):
        """Test hook: replace the noise table (lets the CPU oracle share it)."""
        nz = np.ascontiguousarray(noise, np.float64).reshape(self.iterations, 4)
        _native.check(_native.lib().lutgen_cuda_sampling_set_noise(self._h, _ptr(nz)))


class GaussianBlurRemapper(InterpolatedRemapper):
    """GaussianBlurRemapper::new(palette, radius, lum_factor, preserve)  (gaussian_blur.rs:34-53).

    The CLI's and Studio's default algorithm. Like the reference type it implements `GenerateLut`
    only (gaussian_blur.rs:512-541): the `remap_*` methods of `InterpolatedRemapper` do not exist on
    it. `par_generate_lut*` follow `par_generate_lut_inner` (the nearest-neighbour hint restarts per
    (r, g) row; what the CLI and Studio call), `generate_lut*` follow `generate_lut_inner` (the hint
    carries on from row to row).
    """
    _kind = _native.BLUR

    def __init__(self, palette, radius, lum_factor, preserve):
        self._create(palette, param=radius, lum_factor=lum_factor, preserve=preserve)

    def _no_remap(self, *a, **k):
        raise AttributeError("GaussianBlurRemapper implements GenerateLut only, as in the reference")

    remap_pixel = remap_image = remap_image_with_interrupt = _no_remap
    par_remap_image = par_remap_image_with_interrupt = _no_remap
