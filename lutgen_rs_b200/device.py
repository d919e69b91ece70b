"""Device-resident and multi-GPU entry points (torch tensors in, torch tensors out).

PyTorch is plumbing here: it owns device memory, streams and the NCCL process
group. All arithmetic still happens in liblutgen_cuda.so, reached through the
`_dev` functions of the C ABI with raw device pointers.

Multi-GPU model: one process per GPU.
  * LUT generation shards the level^3 x level^3 CLUT by ROWS; every rank writes
    its rows at their place in a full-size buffer and one in-place NCCL
    all-gather over NVLink assembles the CLUT on every rank.
    `generate_lut_fused` is the fused form: every rank's kernel writes its rows straight into ALL
    ranks' buffers (CUDA IPC mappings, stores over NVLink while it computes) and a barrier replaces
    the collective.
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

    Stream contract: the kernels AND the collective run on `stream` (torch's current stream when None); the returned
    tensor is ordered with that stream like any torch result.
    """
    level = int(level)
    side = level ** 3
    if out is not None:
        _check_u8(out, "out")

    def compute_rows(flat, r0, r1):
        generate_lut_rows_(remapper, level, flat, r0, r1, None)   # torch's current stream: the one the collective uses

    if stream is None:
        return sharded_rows_all_gather(side, compute_rows, out=out, group=group, device="cuda")
    with torch.cuda.stream(stream):
        return sharded_rows_all_gather(side, compute_rows, out=out, group=group, device="cuda")


class _RawCuda:
    """A raw device allocation as a zero-copy torch tensor (via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerLut:
    """One CLUT-sized buffer per rank, each mapped into every other rank's process (CUDA IPC), in two
    copies used alternately (a rank may still be reading step i's CLUT when another rank starts
    writing step i + 1's), plus a flag array per rank for the peer-memory barrier. Collective: every
    rank of `group` must construct it at the same time."""

    MAX_WORLD = 8  # 1 + LUTGEN_MAX_PEERS destinations per launch
    FLAG_BYTES = 256   # 16 u32 used: slots 0..7 written by the peers, word 8 the owner's step counter (lutgen_cuda.h)

    def __init__(self, nbytes, group=None):
        import ctypes as C

        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > self.MAX_WORLD:
            raise ValueError(f"at most {self.MAX_WORLD} ranks")
        self.nbytes = int(nbytes)
        self._lut_bytes = (self.nbytes + 255) // 256 * 256   # flags follow the CLUT, 256-byte aligned
        L = _native.lib()
        self._own, self._ptrs, self._imported, self._mc, self._symm = [], [], [], [], []
        total = self._lut_bytes + self.FLAG_BYTES
        use_symm = not os.environ.get("LUTGEN_NO_SYMM_MEM")
        for _ in range(2):
            got = None
            if use_symm:
                # torch symmetric memory: peer mappings AND one NVSwitch multicast address for all of them
                try:
                    import torch.distributed._symmetric_memory as symm_mem
                    t = symm_mem.empty(total, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
                    hdl = symm_mem.rendezvous(t, group=group if group is not None else dist.group.WORLD)
                    got = (t, hdl)
                except Exception:  # not available here: CUDA IPC below
                    got = None
                # all ranks must take the same route
                okt = torch.tensor([1 if got is not None else 0], device="cuda")
                dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)
                if int(okt.item()) == 0:
                    got, use_symm = None, False
            if got is not None:
                t, hdl = got
                t.zero_()
                self._symm.append(got)
                self._own.append(t.data_ptr())
                self._ptrs.append([int(q) for q in hdl.buffer_ptrs])
                self._mc.append(int(hdl.multicast_ptr or 0))
                continue
            p = C.c_void_p()
            _native.check(L.lutgen_cuda_device_alloc(total, C.byref(p)))
            self._own.append(p.value)
            self._mc.append(0)
            torch.as_tensor(_RawCuda(p.value + self._lut_bytes, self.FLAG_BYTES), device="cuda").zero_()
            handle = (C.c_uint8 * 64)()
            _native.check(L.lutgen_cuda_ipc_export(C.c_void_p(p.value), C.cast(handle, C.c_void_p)))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().numpy().reshape(self.world, 64)
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(p.value)
                    continue
                h = (C.c_uint8 * 64)(*allh[r].tolist())
                q = C.c_void_p()
                _native.check(L.lutgen_cuda_ipc_import(C.cast(h, C.c_void_p), C.byref(q)))
                self._imported.append(q.value)
                ptrs.append(q.value)
            self._ptrs.append(ptrs)
        self.tensors = [torch.as_tensor(_RawCuda(p, self.nbytes), device="cuda") for p in self._own]
        # flag arrays of copy 0 throughout, device tables of their addresses by rank: bytes 0..63 belong to the
        # exchange kernel's own barrier (lutgen_cuda_generate_lut_exchange_dev), bytes 128.. to barrier() below
        self._flag_table = torch.tensor([q + self._lut_bytes for q in self._ptrs[0]], dtype=torch.int64, device="cuda")
        self._flag_table_host = torch.tensor([q + self._lut_bytes + 128 for q in self._ptrs[0]], dtype=torch.int64, device="cuda")
        self._epoch = 0
        self._turn = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def next(self):
        """(tensor of this rank's buffer, ctypes array of all ranks' pointers -- own first --, multicast
        address of the same buffers or 0) for the next step."""
        import ctypes as C
        i = self._turn
        self._turn ^= 1
        ptrs = self._ptrs[i]
        order = [ptrs[self.rank]] + [ptrs[r] for r in range(self.world) if r != self.rank]
        return self.tensors[i], (C.c_void_p * len(order))(*order), self._mc[i]

    def barrier(self, stream=None):
        """All ranks' earlier writes into the peer buffers are complete and visible after this."""
        self._epoch += 1
        _native.check(_native.lib().lutgen_cuda_peer_barrier_dev(self._flag_table_host.data_ptr(), self.world, self.rank,
                                                                  self._epoch & 0xFFFFFFFF, _stream(stream)))

    def close(self):
        import ctypes as C
        L = _native.lib()
        self.tensors = []
        self._flag_table = self._flag_table_host = None
        for q in self._imported:
            L.lutgen_cuda_ipc_close(C.c_void_p(q))
        self._imported = []
        if not self._symm:
            for p in self._own:
                L.lutgen_cuda_device_free(C.c_void_p(p))
        self._own, self._symm = [], []


_peer_luts = {}


def generate_lut_fused(remapper, level, group=None, stream=None):
    """Sharded generation with the exchange fused into the kernel: each rank computes its share of
    the tiles (every world-th hand-out: balanced) and stores the results into EVERY rank's buffer over
    NVLink while it computes -- once, to an NVSwitch multicast address, when torch's symmetric memory
    provides one (lutgen_cuda_generate_lut_shard_mc_dev), else to each peer mapping (CUDA IPC); a
    barrier through flags in peer memory then makes the whole CLUT valid on every rank. Returns a (side, side, 3) uint8 CUDA
    tensor that stays valid until the second next call. Falls back to `generate_lut` (all-gather)
    for remappers the fused entry point does not take, and for a single rank."""
    import torch.distributed as dist
    level = int(level)
    side = level ** 3
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    # measured (tools/time_config3.py, 8 GPUs): with 256 colours the two-level kernel's per-tile 24-byte stores make
    # the fused form slightly slower than the collective (0.47 vs 0.44 ms), so palettes beyond the small-palette
    # kernel's 32 colours take the all-gather
    if (world == 1 or world > PeerLut.MAX_WORLD or remapper._kind in (_native.SAMPLING, _native.BLUR)
            or len(remapper._palette) > 32):
        return generate_lut(remapper, level, group=group, stream=stream)
    key = (side, id(group))
    peers = _peer_luts.get(key)
    if peers is None:
        peers = _peer_luts[key] = PeerLut(3 * side * side, group)
    mine, ptrs, mc = peers.next()
    # one launch: the kernel's last thread block signals and waits for the peers (no barrier launch, no collective)
    _native.check(_native.lib().lutgen_cuda_generate_lut_exchange_dev(remapper._h, level, ptrs, len(ptrs), mc or None, peers.rank, world,
                                                                     peers._flag_table.data_ptr(), _stream(stream)))
    return mine.view(side, side, 3)


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


def encode_png_(image, out=None, stream=None):
    """PNG file bytes of a device image ((h, w, c) uint8 tensor, c in 1..4) into a device buffer, without the
    pixels or the file visiting the host (lutgen_cuda_encode_png_dev). Returns a view of `out` of the file's
    length; `out` (optional) must hold lutgen_rs_b200.png.bound(w, h, c) bytes."""
    import ctypes as C

    from . import png
    _check_u8(image, "image")
    if image.dim() == 2:
        image = image.unsqueeze(-1)
    h, w, c = image.shape
    need = png.bound(w, h, c)
    if out is None:
        out = torch.empty(need, dtype=torch.uint8, device=image.device)
    _check_u8(out, "out")
    n = C.c_size_t(0)
    _native.check(_native.lib().lutgen_cuda_encode_png_dev(image.data_ptr(), w, h, c, out.data_ptr(), out.numel(), C.byref(n), _stream(stream)))
    return out[:n.value]


def shard_frames(nframes, world, rank):
    """Contiguous frame shard for batched apply / remap (no exchange afterwards)."""
    per = -(-nframes // world)
    return min(rank * per, nframes), min((rank + 1) * per, nframes)
