"""lutgen_rs_b200 -- B200 (sm_100a) implementation of lutgen's Hald-CLUT hot path.

Python host mirror of the Rust crate `lutgen` (crates/lib of ozwaldorf/lutgen-rs):

    from lutgen_rs_b200 import identity, interpolation
    lut = interpolation.GaussianRemapper(palette, 128.0, 16, 1.0, False).par_generate_lut(8)
    identity.correct_image(rgba_image, lut)

Everything routes through the C ABI in include/lutgen_cuda.h (liblutgen_cuda.so,
hand-written CUDA kernels). There is no CPU fallback; without an sm_100 GPU every
compute call raises. `lutgen_rs_b200.device` holds the device-resident / multi-GPU
entry points (torch tensors, torch.distributed over NCCL).
"""
from . import _native, extract, identity, interpolation, png
from ._native import LutgenError, LutgenPanic
from .interpolation import (AbortFlag, GaussianBlurRemapper, GaussianRemapper, GaussianSamplingRemapper, InterpolatedRemapper,
                            LinearRemapper, NearestNeighborRemapper, ShepardRemapper)

__all__ = [
    "identity", "interpolation", "png", "extract", "LutgenError", "LutgenPanic", "AbortFlag", "InterpolatedRemapper",
    "GaussianRemapper", "GaussianBlurRemapper", "ShepardRemapper", "LinearRemapper", "NearestNeighborRemapper",
    "GaussianSamplingRemapper",
]
