"""ctypes binding of the CPU oracle (oracle/lutgen_oracle.c).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. Never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liblutgen_oracle.so")

GAUSSIAN, SHEPARD, LINEAR, NEAREST, SAMPLING, BLUR = range(6)


class Desc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("preserve", C.c_int32),
        ("palette", C.c_void_p),
        ("palette_len", C.c_uint64),
        ("param", C.c_double),
        ("nearest", C.c_uint64),
        ("lum_factor", C.c_double),
        ("mean", C.c_double),
        ("std_dev", C.c_double),
        ("iterations", C.c_uint64),
        ("seed", C.c_uint64),
    ]


def build(force=False):
    src = os.path.join(ORACLE_DIR, "lutgen_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(ORACLE_SO)
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        vp, u8, u32, sz, i32 = C.c_void_p, C.c_uint8, C.c_uint32, C.c_size_t, C.c_int
        L.lutgen_oracle_srgb_to_oklab.argtypes = [vp, sz, vp]
        L.lutgen_oracle_oklab_to_srgb.argtypes = [vp, sz, vp]
        L.lutgen_oracle_srgb_decode_table.argtypes = [vp]
        L.lutgen_oracle_cbrtf.argtypes = [C.c_float]
        L.lutgen_oracle_cbrtf.restype = C.c_float
        L.lutgen_oracle_identity.argtypes = [u8, vp]
        L.lutgen_oracle_correct_pixel.argtypes = [vp, vp, u8, vp]
        L.lutgen_oracle_correct_image.argtypes = [vp, sz, vp, u8, i32]
        L.lutgen_oracle_detect_level.argtypes = [u32, u32]
        L.lutgen_oracle_detect_level.restype = i32
        L.lutgen_oracle_remapper_create.argtypes = [C.POINTER(Desc)]
        L.lutgen_oracle_remapper_create.restype = vp
        L.lutgen_oracle_remapper_destroy.argtypes = [vp]
        L.lutgen_oracle_sampling_set_noise.argtypes = [vp, vp]
        L.lutgen_oracle_sampling_get_noise.argtypes = [vp, vp]
        L.lutgen_oracle_remap_image.argtypes = [vp, vp, sz, i32]
        L.lutgen_oracle_generate_lut_rows.argtypes = [vp, u8, vp, u32, u32, i32]
        L.lutgen_oracle_generate_lut.argtypes = [vp, u8, vp, i32]
        L.lutgen_oracle_max_threads.restype = i32
        L.lutgen_oracle_blur_generate.argtypes = [vp, u8, vp, i32, i32]
        L.lutgen_oracle_blur_generate.restype = i32
        L.lutgen_oracle_extract_palette.argtypes = [vp, sz, u32, u32, vp]
        L.lutgen_oracle_extract_palette.restype = u32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def max_threads():
    return int(lib().lutgen_oracle_max_threads())


def identity(level):
    n = level ** 3
    out = np.empty((n, n, 3), np.uint8)
    lib().lutgen_oracle_identity(level, _ptr(out))
    return out


def detect_level(w, h):
    return int(lib().lutgen_oracle_detect_level(w, h))


def correct_pixel(rgb, lut, level):
    i = np.ascontiguousarray(rgb, np.uint8)
    o = np.empty(3, np.uint8)
    lut = np.ascontiguousarray(lut, np.uint8)
    lib().lutgen_oracle_correct_pixel(_ptr(i), _ptr(lut), level, _ptr(o))
    return o


def correct_image(rgba, lut, level, threads=1):
    """Returns a corrected copy of `rgba` (..., 4) uint8."""
    out = np.ascontiguousarray(rgba, np.uint8).copy()
    lut = np.ascontiguousarray(lut, np.uint8)
    lib().lutgen_oracle_correct_image(_ptr(out), out.size // 4, _ptr(lut), level, threads)
    return out


def extract_palette(rgb, color_count, iterations=16):
    """The oracle's restatement of the deterministic k-means of lutgen_rs_b200/csrc/extract_core.h (parity with the
    reference's quantette dependency is unpinned, see there). rgb: (..., 3) uint8. Returns (k, 3) uint8."""
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    out = np.zeros((256, 3), np.uint8)
    k = lib().lutgen_oracle_extract_palette(_ptr(rgb), rgb.shape[0], int(color_count), int(iterations), _ptr(out))
    return out[:k].copy()


def srgb_to_oklab(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8).reshape(-1, 3)
    out = np.empty((rgb.shape[0], 3), np.float32)
    lib().lutgen_oracle_srgb_to_oklab(_ptr(rgb), rgb.shape[0], _ptr(out))
    return out


def oklab_to_srgb(lab):
    lab = np.ascontiguousarray(lab, np.float32).reshape(-1, 3)
    out = np.empty((lab.shape[0], 3), np.uint8)
    lib().lutgen_oracle_oklab_to_srgb(_ptr(lab), lab.shape[0], _ptr(out))
    return out


def srgb_decode_table():
    out = np.empty(256, np.float32)
    lib().lutgen_oracle_srgb_decode_table(_ptr(out))
    return out


class Remapper:
    """One of the reference's remappers, restated on the CPU."""

    def __init__(self, kind, palette, param=0.0, nearest=0, lum_factor=1.0, preserve=False,
                 mean=0.0, std_dev=0.0, iterations=0, seed=0):
        self._pal = np.ascontiguousarray(palette, np.uint8).reshape(-1, 3)
        d = Desc(kind, int(bool(preserve)), self._pal.ctypes.data, self._pal.shape[0], float(param),
                 int(nearest), float(lum_factor), float(mean), float(std_dev), int(iterations), int(seed))
        self.kind = kind
        self.iterations = int(iterations)
        self._h = lib().lutgen_oracle_remapper_create(C.byref(d))
        if not self._h:
            raise ValueError("oracle: invalid remapper description")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().lutgen_oracle_remapper_destroy(h)

    def set_noise(self, noise):
        noise = np.ascontiguousarray(noise, np.float64).reshape(self.iterations, 4)
        assert lib().lutgen_oracle_sampling_set_noise(self._h, _ptr(noise)) == 0

    def get_noise(self):
        out = np.empty((self.iterations, 4), np.float64)
        assert lib().lutgen_oracle_sampling_get_noise(self._h, _ptr(out)) == 0
        return out

    def remap_image(self, rgba, threads=None):
        out = np.ascontiguousarray(rgba, np.uint8).copy()
        lib().lutgen_oracle_remap_image(self._h, _ptr(out), out.size // 4, threads or max_threads())
        return out

    def generate_lut(self, level, threads=None, rows=None, sequential_hints=False):
        n = level ** 3
        out = np.zeros((n, n, 3), np.uint8)
        if self.kind == BLUR:
            rc = lib().lutgen_oracle_blur_generate(self._h, level, _ptr(out), int(bool(sequential_hints)), threads or max_threads())
            if rc != 0:
                raise ValueError("oracle: GaussianBlurRemapper with an empty palette panics")
            return out
        r0, r1 = rows if rows is not None else (0, n)
        lib().lutgen_oracle_generate_lut_rows(self._h, level, _ptr(out), r0, r1, threads or max_threads())
        return out
