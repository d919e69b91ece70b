#!/usr/bin/env python3
"""One GPU: a level-16 Gaussian CLUT generated as `count` interleaved shards (as each rank of a
multi-GPU job would run one of them), for timing a shard's launches in isolation."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lutgen_rs_b200 import _native, device, interpolation  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
device.init(0)
pal = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json")))["catppuccin_mocha"], np.uint8)
r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
lut = torch.zeros(3 * 4096 * 4096, dtype=torch.uint8, device="cuda")
ptrs = (C.c_void_p * 1)(lut.data_ptr())
L = _native.lib()
s = torch.cuda.current_stream().cuda_stream
for rep in range(reps):
    for idx in range(count):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _native.check(L.lutgen_cuda_generate_lut_shard_multi_dev(r._h, 16, ptrs, 1, idx, count, s))
        b.record()
        torch.cuda.synchronize()
        print(f"rep {rep} shard {idx}/{count}: {a.elapsed_time(b):.4f} ms")
whole = torch.from_numpy(r.generate_lut(16)).cuda().reshape(-1)
print("shards tile the CLUT:", bool(torch.equal(lut, whole)))
