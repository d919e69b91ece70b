"""Device-resident and multi-GPU entry points (torch tensors in, torch tensors out).

PyTorch is plumbing here: it owns device memory, streams and the NCCL process
group. All arithmetic still happens in liblutgen_cuda.so, reached through the
`_dev` functions of the C ABI with raw device pointers.

Multi-GPU model: one process per GPU.
  * LUT generation shards the level^3 x level^3 CLUT by ROWS; every rank writes
    its rows at their place in a full-size buffer and one in-place NCCL
    all-gather over NVLink assembles the CLUT on every rank.
  * LUT application / direct remap shard FRAMES; no collective is needed.
"""
import os

import torch

from . import _native


def init(local_rank=None):
    """Bind torch and the native library to this process's GPU."""
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _native.init(local_rank)
    return local_rank


def _stream(stream):
    s = torch.cuda.current_stream() if stream is None else stream
    return s.cuda_stream


def _check_u8(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous CUDA uint8 tensor")


def identity_(out, level, stream=None):
    """identity::generate into a device tensor of 3*level^6 bytes."""
    _check_u8(out, "out")
    assert out.numel() == 3 * int(level) ** 6
    _native.check(_native.lib().lutgen_cuda_identity_dev(int(level), out.data_ptr(), _stream(stream)))
    return out


def generate_lut_rows_(remapper, level, lut, row_begin, row_end, stream=None):
    """Rows [row_begin, row_end) of remapper's level-`level` CLUT into `lut` (full-size device image)."""
    _check_u8(lut, "lut")
    assert lut.numel() >= 3 * int(level) ** 6
    _native.check(_native.lib().lutgen_cuda_generate_lut_rows_dev(
        remapper._h, int(level), lut.data_ptr(), int(row_begin), int(row_end), _stream(stream)))
    return lut


def shard_rows(side, world, rank):
    """Contiguous row shard of rank `rank`: ceil(side/world) rows each, the last ranks may be short/empty."""
    per = -(-side // world)
    return min(rank * per, side), min((rank + 1) * per, side), per


def sharded_rows_all_gather(side, compute_rows, out=None, group=None, device="cuda"):
    """Row-sharded generation + in-place all-gather, independent of what fills the rows.

    `compute_rows(flat, r0, r1)` writes rows [r0, r1) of the side x side RGB8 image at their place in
    `flat` (a uint8 tensor of padded_rows * side * 3 bytes). Every rank owns ceil(side/world) rows (the
    last ranks may own fewer or none; the buffer is padded so all shards have equal size) and one
    all_gather_into_tensor, whose output IS that buffer, leaves the whole image on every rank.
    Returns the (side, side, 3) view. The GPU path passes the CUDA kernel launch as `compute_rows`;
    the CPU (gloo) tests pass the oracle, so the sharding arithmetic is tested without a GPU.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    r0, r1, per = shard_rows(side, world, rank)
    nbytes = per * world * side * 3
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if not (out.dtype == torch.uint8 and out.is_contiguous() and out.numel() >= nbytes):
        raise TypeError(f"out must be a contiguous uint8 tensor of at least {nbytes} bytes")
    flat = out.view(-1)[:nbytes]
    compute_rows(flat, r0, r1)
    if world > 1:
        mine = flat[rank * per * side * 3:(rank + 1) * per * side * 3]
        dist.all_gather_into_tensor(flat, mine, group=group)
    return flat[: side * side * 3].view(side, side, 3)


def generate_lut(remapper, level, group=None, stream=None, out=None):
    """GenerateLut::par_generate_lut on the device(s): returns a (side, side, 3) uint8 CUDA tensor.

    With a process group of G ranks each rank computes side/G rows and an in-place
    all-gather (NCCL, NVLink) leaves the complete CLUT on every rank.
    """
    level = int(level)
    side = level ** 3
    if out is not None:
        _check_u8(out, "out")

    def compute_rows(flat, r0, r1):
        generate_lut_rows_(remapper, level, flat, r0, r1, stream)

    return sharded_rows_all_gather(side, compute_rows, out=out, group=group, device="cuda")


def apply_hald_(rgba, lut, level, stream=None):
    """identity::correct_image_with_level on device tensors, in place. `rgba`: (..., 4) uint8."""
    _check_u8(rgba, "rgba")
    _check_u8(lut, "lut")
    assert rgba.shape[-1] == 4 and lut.numel() == 3 * int(level) ** 6
    _native.check(_native.lib().lutgen_cuda_apply_hald_dev(rgba.data_ptr(), rgba.numel() // 4, lut.data_ptr(), int(level),
                                                           _stream(stream)))
    return rgba


def remap_image_(remapper, rgba, stream=None):
    """InterpolatedRemapper::par_remap_image on a device tensor, in place."""
    _check_u8(rgba, "rgba")
    assert rgba.shape[-1] == 4
    _native.check(_native.lib().lutgen_cuda_remap_image_dev(remapper._h, rgba.data_ptr(), rgba.numel() // 4, _stream(stream)))
    return rgba


def shard_frames(nframes, world, rank):
    """Contiguous frame shard for batched apply / remap (no exchange afterwards)."""
    per = -(-nframes // world)
    return min(rank * per, nframes), min((rank + 1) * per, nframes)
