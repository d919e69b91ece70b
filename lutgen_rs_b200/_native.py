"""ctypes loader for liblutgen_cuda.so (the C ABI in include/lutgen_cuda.h).

There is no fallback: if the shared object is missing and cannot be built, or
there is no sm_100 GPU, the error propagates to the caller.
"""
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# LUTGEN_CUDA_LIB: an experiment build to load instead (tools/build_variants.py); never set by the product or the tests
SO_PATH = os.environ.get("LUTGEN_CUDA_LIB") or os.path.join(HERE, "liblutgen_cuda.so")

OK, ABORTED = 0, 1
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED = -1, -2, -3, -4
GAUSSIAN, SHEPARD, LINEAR, NEAREST, SAMPLING, BLUR = range(6)

# Every symbol include/lutgen_cuda.h declares (tests check the .so exports them all).
EXPORTS = [
    "lutgen_cuda_init", "lutgen_cuda_shutdown", "lutgen_cuda_device_count", "lutgen_cuda_last_error",
    "lutgen_cuda_version", "lutgen_cuda_launch_count", "lutgen_cuda_host_alloc", "lutgen_cuda_host_free",
    "lutgen_cuda_identity", "lutgen_cuda_identity_dev", "lutgen_cuda_detect_level",
    "lutgen_cuda_apply_hald", "lutgen_cuda_apply_hald_dev", "lutgen_cuda_apply_hald_batch",
    "lutgen_cuda_correct_pixels", "lutgen_cuda_remapper_create", "lutgen_cuda_remapper_destroy",
    "lutgen_cuda_generate_lut", "lutgen_cuda_generate_lut_rows_dev", "lutgen_cuda_remap_image",
    "lutgen_cuda_remap_image_dev", "lutgen_cuda_sampling_get_noise", "lutgen_cuda_sampling_set_noise",
    "lutgen_cuda_srgb_to_oklab", "lutgen_cuda_oklab_to_srgb", "lutgen_cuda_measure_peaks",
    "lutgen_cuda_srgb_to_oklab_fast", "lutgen_cuda_debug_stats", "lutgen_cuda_debug_exchange_probe_dev", "lutgen_cuda_debug_timeline",
    "lutgen_cuda_generate_lut_rows_multi_dev", "lutgen_cuda_device_alloc", "lutgen_cuda_device_free",
    "lutgen_cuda_ipc_export", "lutgen_cuda_ipc_import", "lutgen_cuda_ipc_close",
    "lutgen_cuda_generate_lut_shard_multi_dev", "lutgen_cuda_peer_barrier_dev", "lutgen_cuda_generate_lut_shard_mc_dev",
    "lutgen_cuda_generate_lut_exchange_dev", "lutgen_cuda_init_devices", "lutgen_cuda_device_count_in_use",
    "lutgen_cuda_generate_lut_part",
    "lutgen_cuda_generate_lut_sequential",
    "lutgen_cuda_png_bound", "lutgen_cuda_encode_png", "lutgen_cuda_encode_png_dev", "lutgen_cuda_generate_lut_png",
    "lutgen_cuda_apply_hald_png", "lutgen_cuda_extract_palette",
    "lutgen_cuda_png_info", "lutgen_cuda_decode_png", "lutgen_cuda_decode_png_dev", "lutgen_cuda_apply_hald_png_file",
]


class LutgenError(RuntimeError):
    """A CUDA / environment failure reported by the native library."""


class LutgenPanic(LutgenError):
    """Raised where the Rust reference panics (invalid CLUT, bad level, ...)."""


class RemapperDesc(C.Structure):
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


_lib = None
_lock = threading.Lock()
_initialised = False


def load():
    """dlopen the library (building it in-tree first if a toolchain is present)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        try:
            if not os.environ.get("LUTGEN_CUDA_LIB") and _build.needs_build():
                _build.build()
        except Exception as e:  # no nvcc on this machine: use the prebuilt .so if any
            if not os.path.exists(SO_PATH):
                raise ImportError(
                    f"liblutgen_cuda.so is missing and could not be built ({e}); "
                    "this package has no CPU fallback") from e
        L = C.CDLL(SO_PATH)
        vp, u8, u32, u64, sz, i32 = C.c_void_p, C.c_uint8, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int
        sig = {
            "lutgen_cuda_init": ([i32], i32),
            "lutgen_cuda_shutdown": ([], i32),
            "lutgen_cuda_device_count": ([C.POINTER(i32)], i32),
            "lutgen_cuda_last_error": ([], C.c_char_p),
            "lutgen_cuda_version": ([], C.c_char_p),
            "lutgen_cuda_launch_count": ([], u64),
            "lutgen_cuda_host_alloc": ([sz, C.POINTER(vp)], i32),
            "lutgen_cuda_host_free": ([vp], i32),
            "lutgen_cuda_identity": ([u8, vp], i32),
            "lutgen_cuda_identity_dev": ([u8, vp, vp], i32),
            "lutgen_cuda_detect_level": ([u32, u32, C.POINTER(u8)], i32),
            "lutgen_cuda_apply_hald": ([vp, sz, vp, u8], i32),
            "lutgen_cuda_apply_hald_dev": ([vp, sz, vp, u8, vp], i32),
            "lutgen_cuda_apply_hald_batch": ([C.POINTER(vp), C.POINTER(sz), sz, vp, u8], i32),
            "lutgen_cuda_correct_pixels": ([vp, sz, vp, u8, vp], i32),
            "lutgen_cuda_remapper_create": ([C.POINTER(RemapperDesc), C.POINTER(vp)], i32),
            "lutgen_cuda_remapper_destroy": ([vp], i32),
            "lutgen_cuda_generate_lut": ([vp, u8, vp, vp], i32),
            "lutgen_cuda_generate_lut_sequential": ([vp, u8, vp, vp], i32),
            "lutgen_cuda_generate_lut_rows_dev": ([vp, u8, vp, u32, u32, vp], i32),
            "lutgen_cuda_remap_image": ([vp, vp, sz, vp], i32),
            "lutgen_cuda_remap_image_dev": ([vp, vp, sz, vp], i32),
            "lutgen_cuda_sampling_get_noise": ([vp, vp], i32),
            "lutgen_cuda_sampling_set_noise": ([vp, vp], i32),
            "lutgen_cuda_srgb_to_oklab": ([vp, sz, vp], i32),
            "lutgen_cuda_oklab_to_srgb": ([vp, sz, vp], i32),
            "lutgen_cuda_srgb_to_oklab_fast": ([vp, sz, vp], i32),
            "lutgen_cuda_debug_stats": ([vp, i32], i32),
            "lutgen_cuda_debug_timeline": ([vp, i32], i32),
            "lutgen_cuda_debug_exchange_probe_dev": ([u8, C.POINTER(vp), i32, vp, u32, u32, vp, i32, vp], i32),
            "lutgen_cuda_measure_peaks": ([C.POINTER(C.c_double), C.POINTER(C.c_double)], i32),
            "lutgen_cuda_generate_lut_rows_multi_dev": ([vp, u8, C.POINTER(vp), i32, u32, u32, vp], i32),
            "lutgen_cuda_generate_lut_shard_multi_dev": ([vp, u8, C.POINTER(vp), i32, u32, u32, vp], i32),
            "lutgen_cuda_generate_lut_shard_mc_dev": ([vp, u8, C.POINTER(vp), i32, vp, u32, u32, vp], i32),
            "lutgen_cuda_generate_lut_exchange_dev": ([vp, u8, C.POINTER(vp), i32, vp, u32, u32, vp, vp], i32),
            "lutgen_cuda_init_devices": ([C.POINTER(i32), i32], i32),
            "lutgen_cuda_generate_lut_part": ([vp, u8, vp, u32, u32, vp], i32),
            "lutgen_cuda_device_count_in_use": ([C.POINTER(i32)], i32),
            "lutgen_cuda_peer_barrier_dev": ([vp, i32, i32, u32, vp], i32),
            "lutgen_cuda_device_alloc": ([sz, C.POINTER(vp)], i32),
            "lutgen_cuda_device_free": ([vp], i32),
            "lutgen_cuda_ipc_export": ([vp, vp], i32),
            "lutgen_cuda_ipc_import": ([vp, C.POINTER(vp)], i32),
            "lutgen_cuda_ipc_close": ([vp], i32),
            "lutgen_cuda_png_bound": ([u32, u32, i32, C.POINTER(sz)], i32),
            "lutgen_cuda_encode_png": ([vp, u32, u32, i32, vp, sz, C.POINTER(sz)], i32),
            "lutgen_cuda_encode_png_dev": ([vp, u32, u32, i32, vp, sz, C.POINTER(sz), vp], i32),
            "lutgen_cuda_generate_lut_png": ([vp, u8, vp, sz, C.POINTER(sz), vp], i32),
            "lutgen_cuda_png_info": ([vp, sz, C.POINTER(u32), C.POINTER(u32), C.POINTER(i32), C.POINTER(i32)], i32),
            "lutgen_cuda_decode_png": ([vp, sz, i32, vp, sz], i32),
            "lutgen_cuda_decode_png_dev": ([vp, sz, i32, vp, sz, vp], i32),
            "lutgen_cuda_apply_hald_png_file": ([vp, sz, vp, sz, vp, sz, C.POINTER(sz)], i32),
            "lutgen_cuda_apply_hald_png": ([vp, u32, u32, vp, u8, vp, sz, C.POINTER(sz)], i32),
            "lutgen_cuda_extract_palette": ([vp, sz, u32, u32, vp, C.POINTER(u32)], i32),
        }
        for name, (args, res) in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
        return _lib


def last_error():
    return load().lutgen_cuda_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a status code to Python: returns True when aborted, raises on error."""
    if rc == OK:
        return False
    if rc == ABORTED:
        return True
    msg = last_error()
    if rc == ERR_INVALID:
        raise LutgenPanic(msg)
    raise LutgenError(f"lutgen_cuda error {rc}: {msg}")


def init(device=-1):
    """Bind this process to one GPU. Called lazily by every compute entry point."""
    global _initialised
    L = load()
    if not _initialised:
        check(L.lutgen_cuda_init(int(device)))
        _initialised = True
    return L


def init_devices(devices):
    """One process, several GPUs (lutgen_cuda_init_devices): generate_lut and correct_images shard their work over
    them; devices[0] is the primary device every other entry point uses."""
    global _initialised
    L = load()
    arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    check(L.lutgen_cuda_init_devices(arr, len(devices)))
    _initialised = True
    return L


def shutdown():
    """Release every device (remappers must have been dropped first)."""
    global _initialised
    check(load().lutgen_cuda_shutdown())
    _initialised = False


def lib():
    return init()


def launch_count():
    return int(load().lutgen_cuda_launch_count())
