"""CPU-only (gloo, world_size > 1): the multi-GPU host logic -- row shards of a CLUT assembled by an
in-place all-gather, frame shards of a batch -- with the CPU oracle standing in for the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, level, results_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from lutgen_rs_b200 import device
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pal = np.array([[17, 17, 27], [205, 214, 244], [243, 139, 168], [166, 227, 161], [137, 180, 250], [249, 226, 175]], np.uint8)
    orc = O.Remapper(O.GAUSSIAN, pal, param=128.0, nearest=3, lum_factor=1.0, preserve=False)
    side = level ** 3
    calls = []

    def compute_rows(flat, r0, r1):
        calls.append((r0, r1))
        if r1 > r0:
            rows = orc.generate_lut(level, threads=1, rows=(r0, r1))
            flat[r0 * side * 3:r1 * side * 3] = torch.from_numpy(rows[r0:r1].reshape(-1))

    lut = device.sharded_rows_all_gather(side, compute_rows, device="cpu")
    whole = orc.generate_lut(level, threads=1)
    ok = bool((lut.numpy() == whole).all())
    # frame sharding: contiguous, disjoint, covering
    a, b = device.shard_frames(1024, world, rank)
    spans = [None] * world
    dist.all_gather_object(spans, (a, b))
    cover = sorted(spans)
    ok_frames = cover[0][0] == 0 and cover[-1][1] == 1024 and all(x[1] == y[0] for x, y in zip(cover[:-1], cover[1:]))
    np.save(os.path.join(results_dir, f"r{rank}.npy"), np.array([ok, ok_frames, calls[0][0], calls[0][1]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,level", [(2, 4), (3, 5), (2, 3)])
def test_row_sharded_generation_assembles_on_every_rank(tmp_path, world, level):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, level, str(tmp_path)), nprocs=world, join=True)
    side = level ** 3
    per = -(-side // world)
    for rank in range(world):
        ok, ok_frames, r0, r1 = np.load(tmp_path / f"r{rank}.npy").tolist()
        assert ok, f"rank {rank}: assembled CLUT differs from the single-process CLUT"
        assert ok_frames
        assert (r0, r1) == (min(rank * per, side), min((rank + 1) * per, side))


def test_shard_helpers():
    from lutgen_rs_b200 import device
    for side in (8, 125, 4096):
        for world in (1, 2, 3, 8):
            rows = [device.shard_rows(side, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == side
            assert all(a[1] == b[0] for a, b in zip(rows[:-1], rows[1:]))
            assert all(r[1] - r[0] <= r[2] for r in rows)
    assert device.shard_frames(10, 4, 3) == (9, 10)
    assert device.shard_frames(2, 4, 3) == (2, 2)


def test_part_rows_partition_the_image():
    """InterpolatedRemapper.part_rows (what generate_lut_part writes per rank): disjoint, whole row groups of 16 * level
    rows, covering every row once whatever the number of parts (more parts than row groups: the last ones get nothing)."""
    from lutgen_rs_b200.interpolation import InterpolatedRemapper
    for level in (2, 3, 8, 16, 33):
        n, group = level ** 3, 16 * level
        for parts in (1, 2, 3, 8, 64):
            seen = np.zeros(n, int)
            for p in range(parts):
                for r0, r1 in InterpolatedRemapper.part_rows(level, p, parts):
                    assert r0 % group == 0 and (r1 % group == 0 or r1 == n) and r0 < r1
                    seen[r0:r1] += 1
            assert (seen == 1).all(), (level, parts)
