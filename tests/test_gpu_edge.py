"""GPU parity tests, edge cases: parameter extremes, degenerate palettes, every dispatch branch
(fast tile kernel, exact kernel, flag list, colour-cube route) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import random_rgba, synthetic_palette

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import identity, interpolation  # noqa: E402


def _diff(a, b):
    return np.abs(a.astype(np.int16) - b.astype(np.int16))


def _check_rbf(cls, okind, pal, level, param, nearest, lum=1.0, preserve=False, tol=1):
    args = (pal, param, nearest, lum, preserve) if param is not None else (pal, nearest, lum, preserve)
    got = cls(*args).generate_lut(level)
    want = O.Remapper(okind, pal, param=param or 0.0, nearest=nearest, lum_factor=lum, preserve=preserve).generate_lut(level)
    d = _diff(got, want)
    assert d.max() <= tol, (int(d.max()), int((d > tol).sum()))
    return d


@pytest.mark.parametrize("level", [2, 3, 4, 7])
def test_small_and_odd_levels_all_kinds(palettes, level):
    """Bricks larger than the cube, non-power-of-two cube sides."""
    pal = palettes["catppuccin_mocha"]
    _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, pal, level, 128.0, 16)
    _check_rbf(interpolation.ShepardRemapper, O.SHEPARD, pal, level, 4.0, 8)
    _check_rbf(interpolation.LinearRemapper, O.LINEAR, pal, level, None, 4)
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).generate_lut(level)
    assert (interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(level) == want).all()


def test_nearest_larger_than_64_takes_the_multi_pass_selection():
    pal = synthetic_palette(200, seed=3)
    _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, pal, 5, 64.0, 100)
    _check_rbf(interpolation.ShepardRemapper, O.SHEPARD, pal, 5, 2.0, 150)


def test_huge_palette_falls_back_to_the_exact_kernel():
    pal = synthetic_palette(2500, seed=4)  # > 2048 colours: shared-memory budget of the tile kernel
    _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, pal, 4, 256.0, 16)
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).generate_lut(4)
    assert (interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(4) == want).all()


@pytest.mark.parametrize("lum", [-1.0, 1e-320, 0.0, float("inf"), float("nan"), 1e30, 3.5])
def test_unusual_lum_factors_match_the_oracle(palettes, lum):
    """lum <= 0 / non-normal values leave the fast path; the f64 kernel follows the reference's own
    arithmetic, NaNs and infinities included (f32_to_srgb8 maps NaN to 0)."""
    pal = palettes["nord"]
    _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, pal, 4, 128.0, 8, lum=lum)
    _check_rbf(interpolation.ShepardRemapper, O.SHEPARD, pal, 4, 4.0, 0, lum=lum, preserve=True)
    want = O.Remapper(O.NEAREST, pal, lum_factor=lum, preserve=True).generate_lut(4)
    assert (interpolation.NearestNeighborRemapper(pal, lum, True).generate_lut(4) == want).all()


def test_empty_and_single_colour_palettes():
    empty = np.zeros((0, 3), np.uint8)
    got = interpolation.GaussianRemapper(empty, 128.0, 16, 1.0, False).generate_lut(3)
    want = O.Remapper(O.GAUSSIAN, empty, param=128.0, nearest=16).generate_lut(3)
    assert (got == want).all() and (got == 0).all()  # 0/0 = NaN -> f32_to_srgb8 -> 0
    with pytest.raises(lg.LutgenPanic):  # kiddo nearest_one on an empty tree panics
        interpolation.NearestNeighborRemapper(empty, 1.0, False)
    one = np.array([[12, 200, 77]], np.uint8)
    for level in (2, 5):
        assert (interpolation.NearestNeighborRemapper(one, 1.0, False).generate_lut(level).reshape(-1, 3) == one[0]).all()
        _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, one, level, 128.0, 16)
        _check_rbf(interpolation.ShepardRemapper, O.SHEPARD, one, level, 4.0, 0, preserve=True)


@pytest.mark.parametrize("shape", [0.0, -8.0, 0.5, 2000.0, 1e6])
def test_gaussian_shape_extremes(palettes, shape):
    """Huge shapes underflow f32 (flagged, recomputed in f64; the reference itself degrades to 0/0 ->
    black only where f64 underflows too); negative shapes weight far colours more."""
    _check_rbf(interpolation.GaussianRemapper, O.GAUSSIAN, palettes["gruvbox_dark"], 6, shape, 16)


@pytest.mark.parametrize("power", [0.0, 0.5, 64.0, 300.0])
def test_shepard_power_extremes(palettes, power):
    _check_rbf(interpolation.ShepardRemapper, O.SHEPARD, palettes["catppuccin_mocha"], 6, power, 16)


def test_level16_contains_passthrough_every_palette_colour(palettes):
    """At level 16 every palette colour is a CLUT entry and must map to itself (rbf.rs:66-68)."""
    pal = palettes["catppuccin_mocha"]
    lut = interpolation.ShepardRemapper(pal, 4.0, 16, 1.0, False).generate_lut(16).reshape(-1, 3)
    idx = pal[:, 0].astype(np.int64) + 256 * pal[:, 1].astype(np.int64) + 65536 * pal[:, 2].astype(np.int64)
    assert (lut[idx] == pal).all()


def test_small_images_fast_kernel_coherent_and_random(palettes):
    pal = palettes["nord"]
    h, w = 300, 400
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([xx * 255 // w, yy * 255 // h, (xx + yy) * 255 // (w + h), np.full_like(xx, 200)], -1).astype(np.uint8)
    noisy = random_rgba(h * w, seed=8).reshape(h, w, 4)
    for img in (smooth, noisy):
        for make, okind, kw, tol in (
            (lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), O.GAUSSIAN, dict(param=128.0, nearest=16), 1),
            (lambda: interpolation.NearestNeighborRemapper(pal, 0.9, False), O.NEAREST, dict(lum_factor=0.9), 0),
            (lambda: interpolation.NearestNeighborRemapper(pal, 1.0, True), O.NEAREST, dict(preserve=True), 0),
        ):
            got = img.copy()
            make().remap_image(got)
            want = O.Remapper(okind, pal, **kw).remap_image(img)
            assert _diff(got, want).max() <= tol
            assert (got[..., 3] == img[..., 3]).all()


@pytest.mark.parametrize("iterations", [0, 3, 1025, 2048])
def test_sampling_iteration_counts(palettes, iterations):
    """K = 0 (black, gaussian_sample.rs:53 never runs), odd K (f64 replay), K beyond one offset chunk."""
    pal = palettes["carburetor"]
    g = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, iterations, 1.0, 9, False)
    o = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=iterations, lum_factor=1.0, seed=9)
    if iterations:
        o.set_noise(g.noise())
    level = 3 if iterations > 1000 else 4
    assert (g.generate_lut(level) == o.generate_lut(level)).all()


def test_sampling_large_sigma_and_mean_saturate(palettes):
    pal = palettes["nord"]
    for mean, std in ((0.0, 400.0), (300.0, 5.0), (-300.0, 5.0)):
        g = interpolation.GaussianSamplingRemapper(pal, mean, std, 64, 1.0, 3, False)
        o = O.Remapper(O.SAMPLING, pal, mean=mean, std_dev=std, iterations=64, lum_factor=1.0, seed=3)
        o.set_noise(g.noise())
        assert (g.generate_lut(4) == o.generate_lut(4)).all()
    with pytest.raises(lg.LutgenPanic):  # Normal::new(.., inf).unwrap()
        interpolation.GaussianSamplingRemapper(pal, 0.0, float("inf"), 8, 1.0, 1, False)


def test_sampling_set_noise_round_trip(palettes):
    pal = palettes["nord"]
    g = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, 16, 1.0, 1, False)
    o = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=16, lum_factor=1.0, seed=77)
    g.set_noise(o.get_noise())  # GPU consumes the oracle's stream
    assert (g.generate_lut(4) == o.generate_lut(4)).all()
    img = random_rgba(5000, seed=2).reshape(1, -1, 4)
    got = img.copy()
    g.remap_image(got)
    assert (got == o.remap_image(img)).all()


def test_apply_batch_with_empty_frames(palettes):
    lut = O.identity(3)[::-1].copy()
    frames = [np.zeros((0, 0, 4), np.uint8), random_rgba(17, 1).reshape(1, 17, 4), np.zeros((1, 0, 4), np.uint8)]
    want = O.correct_image(frames[1], lut, 3)
    identity.correct_images(frames, lut)
    assert (frames[1] == want).all()
    identity.correct_images([], lut)


def test_apply_wrong_clut_size_panics():
    img = random_rgba(10, 1).reshape(1, 10, 4)
    with pytest.raises(lg.LutgenPanic):
        identity.correct_image_with_level(img, np.zeros((27, 27, 3), np.uint8), 4)
    with pytest.raises(lg.LutgenPanic):
        identity.correct_image(img, np.zeros((10, 10, 3), np.uint8))


def test_baseline_config4_at_its_stated_size_sampled_against_the_oracle():
    """BASELINE.json config 4 as stated: GaussianSamplingRemapper, level 12, 512 iterations, 256 synthetic colours
    (gaussian_sample.rs:49-76). The whole CLUT is generated on the GPU; every 64th entry is checked against the
    oracle run with the SAME noise table (K is a power of two: bit exact)."""
    pal = synthetic_palette(256, 1)
    level, K = 12, 512
    r = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, K, 1.0, 42080085, False)
    got = r.par_generate_lut(level).reshape(-1, 3)
    ident = O.identity(level).reshape(-1, 3)
    idx = np.arange(7, ident.shape[0], 64)
    px = np.concatenate([ident[idx], np.full((idx.size, 1), 255, np.uint8)], axis=1)
    o = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=K, seed=42080085)
    o.set_noise(r.noise())
    want = o.remap_image(px)
    assert (want[:, 3] == 255).all()
    assert (got[idx] == want[:, :3]).all(), int((got[idx] != want[:, :3]).any(axis=1).sum())


@pytest.mark.parametrize("level,parts", [(16, 8), (16, 3), (8, 2), (5, 4), (3, 8), (12, 1)])
def test_parts_of_a_clut_tile_it_and_write_only_their_rows(palettes, level, parts):
    """generate_lut_part (one process per GPU, CLUT wanted in one shared host image): every part writes exactly the
    rows InterpolatedRemapper.part_rows names, nothing else, and all parts together give generate_lut's CLUT."""
    pal = palettes["catppuccin_mocha"]
    n = level ** 3
    for r in (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), interpolation.GaussianBlurRemapper(pal, 4.0, 1.0, False)):
        whole = r.generate_lut(level)
        img = np.full((n, n, 3), 0xA5, np.uint8)
        covered = np.zeros(n, bool)
        for p in range(parts):
            before = img.copy()
            assert r.generate_lut_part(level, img, p, parts) is img
            mine = np.zeros(n, bool)
            for r0, r1 in interpolation.InterpolatedRemapper.part_rows(level, p, parts):
                assert 0 <= r0 <= r1 <= n
                assert not mine[r0:r1].any() and not covered[r0:r1].any()
                mine[r0:r1] = True
            covered |= mine
            assert (img[mine] == whole[mine]).all()
            assert (img[~mine] == before[~mine]).all()
        assert covered.all() and (img == whole).all()
