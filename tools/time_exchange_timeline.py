#!/usr/bin/env python3
"""Where a fused multi-GPU step spends its time (torchrun, one rank per GPU, LUTGEN_CUDA_STATS=1 set here): per rank,
from the GPU's nanosecond clock, when the last warp ran out of tiles, when the last warp's stores had landed on the
peers, and when the exchange barrier was passed -- all relative to the first thread block's start.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/time_exchange_timeline.py
"""
import ctypes as C
import json
import os
import sys

os.environ["LUTGEN_CUDA_STATS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lutgen_rs_b200 import _native, device, interpolation  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device.init(local)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    level = 16
    L = _native.lib()
    tl = np.zeros(4, np.uint32)

    def timeline(reset):
        _native.check(L.lutgen_cuda_debug_timeline(C.c_void_p(tl.ctypes.data), reset))
        return tl.copy()

    for name, env in (("strips", {"LUTGEN_CUDA_STRIPS": "1"}), ("staged hand-outs of 4", {"LUTGEN_CUDA_STRIPS": "0", "LUTGEN_CUDA_BATCH": "4"}),
                      ("staged single tiles", {"LUTGEN_CUDA_STRIPS": "0", "LUTGEN_CUDA_BATCH": "1"})):
        for k in ("LUTGEN_CUDA_STRIPS", "LUTGEN_CUDA_BATCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            device.generate_lut_fused(r, level)
        rows = []
        for _ in range(5):
            torch.cuda.synchronize()
            timeline(1)
            dist.barrier()
            torch.cuda.synchronize()
            device.generate_lut_fused(r, level)
            t = timeline(0)
            d = [(int(t[i]) - int(t[0])) & 0xffffffff for i in (1, 2, 3)]
            rows.append(d + [int(t[0])])
        mine = torch.tensor(rows, dtype=torch.int64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        if rank == 0:
            print(f"== {name}: per step, per rank: out of tiles / stores landed / barrier passed (us after the rank's first block started); start skew (us)")
            for s in range(5):
                starts = [int(a[s][3]) for a in allr]
                s0 = min(starts)
                print("  step", s, " ".join(f"[{a[s][0] / 1e3:5.1f} {a[s][1] / 1e3:5.1f} {a[s][2] / 1e3:5.1f} +{((st - s0) & 0xffffffff) / 1e3:4.1f}]" for a, st in zip(allr, starts)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
