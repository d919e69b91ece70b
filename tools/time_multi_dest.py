#!/usr/bin/env python3
"""One GPU, the multi-destination kernel variant: shard 0 of N of a level-16 CLUT written to `nd` buffers on the same
GPU, strip exchange against staged hand-outs. Separates what the variants cost in the kernel from what NVLink costs.

    python tools/time_multi_dest.py [shards] [destinations] [reps]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import _native, device, interpolation  # noqa: E402


def main():
    shards = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    nd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    device.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    level, side = 16, 4096
    bufs = [torch.zeros(3 * side * side, dtype=torch.uint8, device="cuda") for _ in range(nd)]
    ptrs = (C.c_void_p * nd)(*[b.data_ptr() for b in bufs])
    ref = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
    device.generate_lut_rows_(r, level, ref, 0, side)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    L = _native.lib()

    def run(shard):
        _native.check(L.lutgen_cuda_generate_lut_shard_multi_dev(r._h, level, ptrs, nd, shard, shards, None))

    variants = [("strips", {"LUTGEN_CUDA_STRIPS": "1"}), ("strips, hand-outs of 4", {"LUTGEN_CUDA_STRIPS": "1", "LUTGEN_CUDA_BATCH": "4"}), ("staged hand-outs of 4", {"LUTGEN_CUDA_STRIPS": "0", "LUTGEN_CUDA_BATCH": "4"}),
                ("staged single tiles", {"LUTGEN_CUDA_STRIPS": "0"})]
    for name, env in variants:
        for k in ("LUTGEN_CUDA_STRIPS", "LUTGEN_CUDA_BATCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for b in bufs:
            b.zero_()
        for s in range(shards):
            run(s)
        torch.cuda.synchronize()
        same = all(torch.equal(b, ref) for b in bufs)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            flush_buf.fill_(1)
            a.record()
            run(0)
            b.record()
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
        print(f"{name:26s}: shard 0 of {shards} into {nd} buffers: {np.median(t):7.1f} us (min {t.min():7.1f}) | all shards give the one-GPU CLUT: {same}", flush=True)


if __name__ == "__main__":
    main()
