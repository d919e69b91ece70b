#!/usr/bin/env python3
"""Where does the time of one sharded generation step go? (run under torchrun, N ranks)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lutgen_rs_b200 import device, interpolation  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
device.init(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pal = np.array(json.load(open(os.path.join(ROOT, "tests/golden/palettes.json")))["catppuccin_mocha"], np.uint8)
r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
side = 4096
r0, r1, per = device.shard_rows(side, world, rank)
lut = torch.empty(per * world * side * 3, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
mine = lut[rank * per * side * 3:(rank + 1) * per * side * 3]
for _ in range(3):
    device.generate_lut_rows_(r, 16, lut, r0, r1)
    dist.all_gather_into_tensor(lut, mine)
torch.cuda.synchronize()
rows = []
for it in range(8):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    flush.fill_(1)
    dist.barrier()
    e[0].record()
    device.generate_lut_rows_(r, 16, lut, r0, r1)
    e[1].record()
    dist.all_gather_into_tensor(lut, mine)
    e[2].record()
    device.generate_lut_rows_(r, 16, lut, r0, r1)
    e[3].record()
    torch.cuda.synchronize()
    rows.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
print(f"rank {rank}: kernel(cold L2) / all-gather / kernel(warm) ms:", " | ".join(f"{a:.3f} {b:.3f} {c:.3f}" for a, b, c in rows), flush=True)

# bench-style: all steps enqueued back to back, no host sync in between
stream = torch.cuda.current_stream()
for label, with_gather in (("kernel+gather", True), ("kernel only", False), ("kernel+gather again", True)):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in evs:
        flush.fill_(1)
        a.record(stream)
        device.generate_lut_rows_(r, 16, lut, r0, r1)
        if with_gather:
            dist.all_gather_into_tensor(lut, mine)
        b.record(stream)
    torch.cuda.synchronize()
    print(f"rank {rank} {label}:", " ".join(f"{a.elapsed_time(b):.3f}" for a, b in evs), flush=True)
ref = torch.from_numpy(r.generate_lut(16)).cuda().reshape(-1)
print(f"rank {rank}: assembled CLUT equals the single-GPU CLUT: {bool(torch.equal(ref, lut[:ref.numel()]))}", flush=True)
dist.destroy_process_group()
