#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/dbg_fused.py : fused row-sharded generation (peer stores over
NVLink + barrier) against the all-gather path and against one GPU, with timings."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lutgen_rs_b200 import device, interpolation  # noqa: E402
from conftest import synthetic_palette  # noqa: E402

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
device.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
ok = True
for name, make, levels in (("gauss26", lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), (16, 9, 7)),
                           ("nn26", lambda: interpolation.NearestNeighborRemapper(pal, 1.0, False), (16, 5)),
                           ("shep100", lambda: interpolation.ShepardRemapper(synthetic_palette(100, 1), 4.0, 16, 1.0, False), (12,)),
                           ("gauss300", lambda: interpolation.GaussianRemapper(synthetic_palette(300, 2), 128.0, 16, 1.0, False), (8,))):
    for level in levels:
        r = make()
        side = level ** 3
        single = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
        device.generate_lut_rows_(r, level, single, 0, side)
        ag = device.generate_lut(r, level).clone()
        for rep in range(3):
            fused = device.generate_lut_fused(r, level)
            torch.cuda.synchronize()
            same = bool(torch.equal(fused.reshape(-1), single)) and bool(torch.equal(ag.reshape(-1)[: 3 * side * side], single))
            ok &= same
            if not same:
                print(f"rank {rank}: {name} level {level} rep {rep}: MISMATCH fused/single {int((fused.reshape(-1) != single).sum())} ag/single {int((ag.reshape(-1)[:3*side*side] != single).sum())}", flush=True)
if rank == 0:
    print("fused == all-gather == single GPU:", ok, flush=True)
# timing, level 16 Gaussian
r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
side = 4096
out = torch.empty(-(-side // world) * world * side * 3, dtype=torch.uint8, device="cuda")
for fn, label in ((lambda: device.generate_lut(r, 16, out=out), "all-gather"), (lambda: device.generate_lut_fused(r, 16), "fused")):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label}: {t.item():.4f} ms per level-16 CLUT on {world} GPUs", flush=True)
# components
import ctypes as C
from lutgen_rs_b200 import _native
peers = device._peer_luts[(side, id(None))]
r0, r1, _ = device.shard_rows(side, world, rank)
def rows_multi():
    mine, ptrs, mc = peers.next()
    _native.check(_native.lib().lutgen_cuda_generate_lut_rows_multi_dev(r._h, 16, ptrs, len(ptrs), int(r0), int(r1), torch.cuda.current_stream().cuda_stream))
def shard_multi():
    mine, ptrs, mc = peers.next()
    _native.check(_native.lib().lutgen_cuda_generate_lut_shard_multi_dev(r._h, 16, ptrs, len(ptrs), rank, world, torch.cuda.current_stream().cuda_stream))
def shard_mc():
    mine, ptrs, mc = peers.next()
    _native.check(_native.lib().lutgen_cuda_generate_lut_shard_mc_dev(r._h, 16, ptrs, len(ptrs), mc or None, rank, world, torch.cuda.current_stream().cuda_stream))
def rows_local():
    device.generate_lut_rows_(r, 16, out, r0, r1)
for fn, label in ((rows_local, "rows, local only"), (rows_multi, "rows, all destinations (no barrier)"), (shard_multi, "interleaved shard, all destinations (no barrier)"), (shard_mc, "interleaved shard, multicast (no barrier)"),
                  (lambda: dist.barrier(), "dist.barrier"), (lambda: peers.barrier(), "peer-memory barrier")):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label}: {t.item():.4f} ms", flush=True)
dist.destroy_process_group()
