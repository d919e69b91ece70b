#!/usr/bin/env python3
"""Multi-GPU generation with the exchange fused into the kernel, timed per variant (torchrun, one rank per GPU):
strip exchange (whole strips sent by the warp that completes them) against staged hand-outs (96-byte rows), each
with and without an L2 flush between steps, and the share-only kernel for comparison. Every variant's CLUT is
compared with the one-GPU CLUT.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/time_exchange.py [reps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lutgen_rs_b200 import _native, device, interpolation  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device.init(local)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    level = 16
    side = level ** 3
    ref = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
    device.generate_lut_rows_(r, level, ref, 0, side)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def timed(fn, flush):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            if flush:
                flush_buf.fill_(1)
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in evs])), float(np.median([a.elapsed_time(b) for a, b in evs]))],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def fused():
        return device.generate_lut_fused(r, level)

    variants = [("the library's choice", {}), ("strips", {"LUTGEN_CUDA_STRIPS": "1"}), ("strips, hand-outs of 4", {"LUTGEN_CUDA_STRIPS": "1", "LUTGEN_CUDA_BATCH": "4"}), ("staged hand-outs of 4", {"LUTGEN_CUDA_STRIPS": "0", "LUTGEN_CUDA_BATCH": "4"}),
                ("staged single tiles", {"LUTGEN_CUDA_STRIPS": "0"})]
    for name, env in variants:
        for k in ("LUTGEN_CUDA_STRIPS", "LUTGEN_CUDA_BATCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            fused()
        torch.cuda.synchronize()
        got = fused()
        torch.cuda.synchronize()
        same = torch.tensor([int(torch.equal(got.reshape(-1), ref))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        tf, tw = timed(fused, True), timed(fused, False)
        if rank == 0:
            print(f"{name:26s}: flushed {tf[0] * 1e3:7.1f} us (median {tf[1] * 1e3:7.1f}) | back to back {tw[0] * 1e3:7.1f} us (median {tw[1] * 1e3:7.1f}) | "
                  f"equals the one-GPU CLUT: {bool(same.item())}", flush=True)
    for k in ("LUTGEN_CUDA_STRIPS", "LUTGEN_CUDA_BATCH"):
        os.environ.pop(k, None)
    # the exchange without the arithmetic: what this GPU's share costs to send (+ the barrier), by store pattern and route
    peers = next(iter(device._peer_luts.values()))
    L = _native.lib()

    def probe(pattern, use_mc):
        def fn():
            mine, ptrs, mc = peers.next()
            _native.check(L.lutgen_cuda_debug_exchange_probe_dev(level, ptrs, len(ptrs), (mc or None) if use_mc else None, peers.rank, world,
                                                                 peers._flag_table.data_ptr(), pattern, None))
        return fn

    for t in peers.tensors:
        t.copy_(ref)
    torch.cuda.synchronize()
    dist.barrier()
    for pattern, pname in ((0, "whole strips (3072-byte runs)"), (1, "hand-outs of 4 (96-byte rows)")):
        for use_mc in ((True, False) if peers._mc[0] else (False,)):
            fn = probe(pattern, use_mc)
            for _ in range(3):
                fn()
            tf, tw = timed(fn, True), timed(fn, False)
            ok = torch.tensor([int(all(torch.equal(t, ref) for t in peers.tensors))], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"exchange only, {pname:30s} {'multicast' if use_mc else 'peer stores':11s}: flushed {tf[0] * 1e3:7.1f} us | back to back {tw[0] * 1e3:7.1f} us "
                      f"| buffers intact: {bool(ok.item())}", flush=True)
    # the share alone (contiguous rows; no exchange, no barrier)
    r0, r1, _ = device.shard_rows(side, world, rank)
    out = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
    tk = timed(lambda: device.generate_lut_rows_(r, level, out, r0, r1), True)
    # the exchange alone: NCCL all-gather of the CLUT
    shard = out[: 3 * side * side // world]
    ta = timed(lambda: dist.all_gather_into_tensor(out, shard), True)
    if rank == 0:
        print(f"share-only kernel (rows)  : flushed {tk[0] * 1e3:7.1f} us;  NCCL all-gather of the CLUT alone: {ta[0] * 1e3:7.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
