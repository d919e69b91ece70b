#!/usr/bin/env python3
"""torchrun --nproc-per-node N tests/tools/multi_check.py

One CLUT shared by N GPUs, three ways -- the exchange kernel (every rank's tiles stored into every rank's buffer by the
kernel itself, multicast or peer mappings, barrier inside the kernel), row shards + NCCL all-gather, and ONE GPU --
must be the same bytes on every rank: whole tiles (16-byte multicast pieces), cubes whose side is no multiple of the
tile (byte stores to the peer mappings), entries settled by the f64 tie resolver, NearestNeighbor and RBF kinds.
Prints `MULTI_CHECK_OK` from rank 0 on success; exit code 1 otherwise."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import synthetic_palette  # noqa: E402
from lutgen_rs_b200 import device, interpolation  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
    cases = (("gauss26", lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), (16, 9, 7, 2)),
             ("gauss26-preserve-lum", lambda: interpolation.GaussianRemapper(pal, 96.0, 8, 1.3, True), (8,)),
             ("shepard26", lambda: interpolation.ShepardRemapper(pal, 4.0, 16, 1.0, False), (12,)),
             ("nn26", lambda: interpolation.NearestNeighborRemapper(pal, 1.0, False), (16, 5)),
             ("shepard100", lambda: interpolation.ShepardRemapper(synthetic_palette(100, 1), 4.0, 16, 1.0, False), (12,)),
             ("gauss300", lambda: interpolation.GaussianRemapper(synthetic_palette(300, 2), 128.0, 16, 1.0, False), (8,)))
    bad = 0
    for name, make, levels in cases:
        for level in levels:
            r = make()
            side = level ** 3
            single = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
            device.generate_lut_rows_(r, level, single, 0, side)
            ag = device.generate_lut(r, level).clone().reshape(-1)[: 3 * side * side]
            # the exchange kernel's two forms (strips sent whole by the warp that completes them / tiles staged and sent
            # by their own warp) and the library's own choice between them; LUTGEN_CUDA_STRIPS is read at every launch
            for form in (None, "1", "0"):
                os.environ.pop("LUTGEN_CUDA_STRIPS", None)
                if form is not None:
                    os.environ["LUTGEN_CUDA_STRIPS"] = form
                for rep in range(3):   # both buffers of the pair, and the first again
                    fused = device.generate_lut_fused(r, level)
                    torch.cuda.synchronize()
                    nf, na = int((fused.reshape(-1) != single).sum()), int((ag != single).sum())
                    if nf or na:
                        bad += 1
                        print(f"rank {rank}: {name} level {level} strips={form} rep {rep}: {nf} bytes differ (exchange kernel), {na} (all-gather)", flush=True)
            os.environ.pop("LUTGEN_CUDA_STRIPS", None)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0 and int(t.item()) == 0:
        peers = next(iter(device._peer_luts.values()), None)
        route = "multicast" if (peers and peers._mc[0]) else ("symmetric-memory peers" if (peers and peers._symm) else "CUDA IPC peers")
        print(f"MULTI_CHECK_OK world={dist.get_world_size()} route={route}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
