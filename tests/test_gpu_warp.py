"""The warp-tile kernels (kernels_warp.cuh): k_remap_warp for palettes of at most 32 colours and the
two-level k_remap_warp_big for 33..256 colours. Every front end (CLUT entries, the sRGB-cube table,
RGBA images), every remapper kind, shard independence, and the dispatch boundaries where the CTA-tile
kernel (kernels_fast.cuh) takes over. NearestNeighbor must be bit exact against the CPU oracle; the
float remappers within +-1 LSB (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import random_rgba, synthetic_palette

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import interpolation  # noqa: E402


def _diff(a, b):
    return np.abs(a.astype(np.int16) - b.astype(np.int16))


def _close(got, want, frac=5e-3):
    d = _diff(got, want)
    assert d.max() <= 1, (int(d.max()), int((d > 1).sum()))
    assert (d > 0).mean() < frac, float((d > 0).mean())


@pytest.mark.parametrize("n", [33, 64, 100, 256])
@pytest.mark.parametrize("level", [6, 9])
def test_big_palette_nearest_neighbor_bit_exact(n, level):
    pal = synthetic_palette(n, seed=n)
    got = interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(level)
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).generate_lut(level)
    assert (got == want).all(), int((got != want).sum())


@pytest.mark.parametrize("n,nearest", [(33, 16), (64, 1), (100, 32), (256, 16), (200, 8)])
def test_big_palette_gaussian_and_shepard_within_1lsb(n, nearest):
    pal = synthetic_palette(n, seed=100 + n)
    for level in (7, 10):
        got = interpolation.GaussianRemapper(pal, 128.0, nearest, 1.0, False).generate_lut(level)
        want = O.Remapper(O.GAUSSIAN, pal, param=128.0, nearest=nearest).generate_lut(level)
        _close(got, want)
    got = interpolation.ShepardRemapper(pal, 4.0, nearest, 0.8, True).generate_lut(8)
    want = O.Remapper(O.SHEPARD, pal, param=4.0, nearest=nearest, lum_factor=0.8, preserve=True).generate_lut(8)
    _close(got, want)
    got = interpolation.LinearRemapper(pal, nearest, 1.0, False).generate_lut(6)
    want = O.Remapper(O.LINEAR, pal, nearest=nearest).generate_lut(6)
    _close(got, want)


def test_big_palette_nn_preserve_and_lum():
    pal = synthetic_palette(90, seed=9)
    got = interpolation.NearestNeighborRemapper(pal, 0.7, True).generate_lut(8)
    want = O.Remapper(O.NEAREST, pal, lum_factor=0.7, preserve=True).generate_lut(8)
    assert (got == want).all()


@pytest.mark.parametrize("n", [20, 150])
def test_image_front_end_small_sizes_and_tails(n, monkeypatch):
    """RGBA front end below the colour-cube threshold: coherent and random pixels, pixel counts that
    leave ragged warp tiles and ragged CTA tiles, alpha untouched."""
    pal = synthetic_palette(n, seed=3 * n)
    rng = np.random.default_rng(n)
    for npix in (1, 31, 128, 129, 1024, 1025, 5000, 70001):
        coherent = np.zeros((npix, 4), np.uint8)
        base = rng.integers(0, 200, 3)
        coherent[:, :3] = base + rng.integers(0, 12, (npix, 3))
        coherent[:, 3] = rng.integers(0, 256, npix)
        for img in (coherent, random_rgba(npix, seed=npix)):
            img = img.reshape(1, npix, 4)
            got = img.copy()
            interpolation.NearestNeighborRemapper(pal, 1.0, False).remap_image(got)
            assert (got == O.Remapper(O.NEAREST, pal, lum_factor=1.0).remap_image(img)).all()
            got = img.copy()
            interpolation.GaussianRemapper(pal, 96.0, 8, 1.0, False).remap_image(got)
            want = O.Remapper(O.GAUSSIAN, pal, param=96.0, nearest=8).remap_image(img)
            assert _diff(got, want).max() <= 1
            assert (got[..., 3] == img[..., 3]).all()


@pytest.mark.parametrize("n", [26, 120])
def test_warp_and_cta_kernels_agree(palettes, monkeypatch, n):
    """The same CLUT from the warp-tile kernels and from the CTA-tile kernel they replaced."""
    pal = palettes["catppuccin_mocha"] if n == 26 else synthetic_palette(n, seed=5)
    for make, exact in ((lambda: interpolation.NearestNeighborRemapper(pal, 1.0, False), True),
                        (lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), False),
                        (lambda: interpolation.ShepardRemapper(pal, 6.0, 12, 1.0, False), False)):
        monkeypatch.setenv("LUTGEN_CUDA_NO_WARP", "1")
        cta = make().generate_lut(11)
        monkeypatch.delenv("LUTGEN_CUDA_NO_WARP")
        warp = make().generate_lut(11)
        d = _diff(warp, cta)
        assert d.max() <= (0 if exact else 1)


def test_dispatch_boundaries_fall_back_to_the_cta_kernel():
    """257 colours, nearest = 0 or > 32 on a big palette, and Gaussian shapes outside the expanded
    distance's comfortable range all leave the warp kernels; results stay within tolerance."""
    for n, nearest, shape in ((257, 16, 128.0), (100, 0, 128.0), (100, 40, 128.0), (20, 8, 5000.0), (20, 8, 1e-5)):
        pal = synthetic_palette(n, seed=n + nearest)
        got = interpolation.GaussianRemapper(pal, shape, nearest, 1.0, False).generate_lut(6)
        want = O.Remapper(O.GAUSSIAN, pal, param=shape, nearest=nearest).generate_lut(6)
        assert _diff(got, want).max() <= 1


@pytest.mark.parametrize("n", [26, 77])
def test_row_shards_are_byte_identical(palettes, n):
    """Rows generated shard by shard equal the single launch (multi-GPU unit), for both kernels,
    with cuts that fall inside warp tiles and CTA tiles."""
    import torch
    from lutgen_rs_b200 import device
    device.init(0)
    pal = palettes["catppuccin_mocha"] if n == 26 else synthetic_palette(n, seed=1)
    for level, cuts in ((9, [0, 1, 5, 333, 334, 729]), (12, [0, 100, 1727, 1728])):
        for r in (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), interpolation.NearestNeighborRemapper(pal, 1.0, False)):
            whole = torch.from_numpy(r.generate_lut(level)).reshape(-1)
            lut = torch.zeros(3 * level ** 6, dtype=torch.uint8, device="cuda")
            for a, b in zip(cuts[:-1], cuts[1:]):
                device.generate_lut_rows_(r, level, lut, a, b)
            torch.cuda.synchronize()
            assert torch.equal(lut.cpu(), whole)


def test_palette_colours_pass_through_with_big_palettes():
    """rbf.rs:66-68 `palette.contains`: a CLUT entry that equals a palette colour keeps its value,
    also when the palette goes through the two-level kernel (level 16: every colour is an entry)."""
    pal = synthetic_palette(60, seed=60)
    lut = interpolation.ShepardRemapper(pal, 4.0, 16, 1.0, False).generate_lut(16).reshape(-1, 3)
    idx = pal[:, 0].astype(np.int64) + 256 * pal[:, 1].astype(np.int64) + 65536 * pal[:, 2].astype(np.int64)
    assert (lut[idx] == pal).all()


@pytest.mark.parametrize("n,nearest,level", [(26, 16, 16), (26, 16, 9), (120, 16, 8), (300, 16, 6), (40, 0, 6)])
def test_interleaved_shards_tile_the_clut(palettes, n, nearest, level):
    """lutgen_cuda_generate_lut_shard_multi_dev: shard i of c takes every c-th tile hand-out of the
    warp-tile kernels (or the shard's rows on the other kernels); the c launches together must write
    every entry exactly as one launch does. One GPU stands in for the c ranks (one destination)."""
    import ctypes as C

    import torch
    from lutgen_rs_b200 import _native, device
    device.init(0)
    pal = palettes["catppuccin_mocha"] if n == 26 else synthetic_palette(n, seed=n)
    side = level ** 3
    s = torch.cuda.current_stream().cuda_stream
    for r in (interpolation.GaussianRemapper(pal, 128.0, nearest, 1.0, False), interpolation.NearestNeighborRemapper(pal, 1.0, False)):
        whole = torch.from_numpy(r.generate_lut(level)).cuda().reshape(-1)
        for count in (2, 3, 8):
            lut = torch.full((3 * side * side,), 7, dtype=torch.uint8, device="cuda")
            ptrs = (C.c_void_p * 1)(lut.data_ptr())
            for idx in range(count):
                _native.check(_native.lib().lutgen_cuda_generate_lut_shard_multi_dev(r._h, level, ptrs, 1, idx, count, s))
            torch.cuda.synchronize()
            assert torch.equal(lut, whole), (count, int((lut != whole).sum()))


def test_multi_destination_rows_write_every_copy(palettes):
    """lutgen_cuda_generate_lut_rows_multi_dev with several destinations on ONE GPU (the peers'
    buffers of a multi-GPU job are ordinary device pointers to the kernel): every copy gets the rows,
    through the staged word stores (level 16, 8) and the entry-by-entry path (level 9)."""
    import ctypes as C

    import torch
    from lutgen_rs_b200 import _native, device
    device.init(0)
    pal = palettes["nord"]
    s = torch.cuda.current_stream().cuda_stream
    for level, rows in ((16, (1000, 2000)), (8, (0, 512)), (9, (3, 700))):
        side = level ** 3
        for r in (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), interpolation.NearestNeighborRemapper(pal, 1.0, False),
                  interpolation.ShepardRemapper(synthetic_palette(80, 3), 4.0, 16, 1.0, False)):
            ref = torch.zeros(3 * side * side, dtype=torch.uint8, device="cuda")
            device.generate_lut_rows_(r, level, ref, rows[0], rows[1])
            copies = [torch.zeros(3 * side * side, dtype=torch.uint8, device="cuda") for _ in range(4)]
            ptrs = (C.c_void_p * 4)(*[c.data_ptr() for c in copies])
            _native.check(_native.lib().lutgen_cuda_generate_lut_rows_multi_dev(r._h, level, ptrs, 4, rows[0], rows[1], s))
            torch.cuda.synchronize()
            for c in copies:
                assert torch.equal(c, ref)


@pytest.mark.parametrize("form", [None, "1", "0"], ids=["library-choice", "strips", "staged"])
def test_interleaved_shards_into_several_destinations(palettes, form, monkeypatch):
    """The exchange kernel's two forms on ONE GPU (the peers' CLUTs are ordinary device pointers to it): shard i of c
    into three destinations, strips sent whole by the warp that completes them (levels that are multiples of 4; from
    four shards on when the library chooses) or tiles staged and stored by their own warp. The c launches together must
    leave the one-GPU CLUT in every copy, and the per-strip counters must come back to zero (second round)."""
    import ctypes as C

    import torch
    from lutgen_rs_b200 import _native, device
    device.init(0)
    if form is None:
        monkeypatch.delenv("LUTGEN_CUDA_STRIPS", raising=False)
    else:
        monkeypatch.setenv("LUTGEN_CUDA_STRIPS", form)
    pal = palettes["catppuccin_mocha"]
    s = torch.cuda.current_stream().cuda_stream
    for level in (16, 8, 4, 9):
        side = level ** 3
        for r in (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), interpolation.NearestNeighborRemapper(pal, 1.0, False)):
            whole = torch.from_numpy(r.generate_lut(level)).cuda().reshape(-1)
            for count in (2, 4, 7):
                for _round in range(2):
                    copies = [torch.full((3 * side * side,), 7, dtype=torch.uint8, device="cuda") for _ in range(3)]
                    for idx in range(count):
                        # each shard's own CLUT first, as a rank passes them: the copies take turns being "own"
                        order = [copies[idx % 3]] + [c for j, c in enumerate(copies) if j != idx % 3]
                        ptrs = (C.c_void_p * 3)(*[c.data_ptr() for c in order])
                        _native.check(_native.lib().lutgen_cuda_generate_lut_shard_multi_dev(r._h, level, ptrs, 3, idx, count, s))
                    torch.cuda.synchronize()
                    for c in copies:
                        assert torch.equal(c, whole), (level, count, int((c != whole).sum()))
