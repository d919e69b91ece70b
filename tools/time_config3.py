#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/time_config3.py : BASELINE.json config 3 -- level-16 ShepardRemapper,
256 synthetic colours, nearest 16 -- sharded over N GPUs, fused (kernel stores into every rank's CLUT +
peer barrier) and unfused (row shards + NCCL all-gather); checks both against rank 0's single-GPU CLUT."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import synthetic_palette  # noqa: E402
from lutgen_rs_b200 import device, interpolation  # noqa: E402

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
device.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
r = interpolation.ShepardRemapper(synthetic_palette(256, 1), 4.0, 16, 1.0, False)
side = 4096
out = torch.empty(-(-side // world) * world * side * 3, dtype=torch.uint8, device="cuda")
single = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
device.generate_lut_rows_(r, 16, single, 0, side)
fused = device.generate_lut_fused(r, 16)
ag = device.generate_lut(r, 16, out=out)
torch.cuda.synchronize()
same = bool(torch.equal(fused.reshape(-1), single)) and bool(torch.equal(ag.reshape(-1), single))
t_same = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(t_same, op=dist.ReduceOp.MIN)
for fn, label in ((lambda: device.generate_lut(r, 16, out=out), "row shards + all-gather"), (lambda: device.generate_lut_fused(r, 16), "fused")):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"config 3 on {world} GPUs, {label}: {t.item():.4f} ms per CLUT", flush=True)
if rank == 0:
    print("fused == all-gather == single GPU:", bool(int(t_same.item())), flush=True)
dist.destroy_process_group()
