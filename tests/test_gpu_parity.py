"""GPU parity tests: the CUDA path (through the C ABI, via the Python host mirror) against the
CPU oracle on the same seeded inputs, against the reference's golden LUT images, and through
size-independent properties at full size.

Bars (BASELINE.json north_star): bit-exact for identity, apply (correct_image), NearestNeighbor
and the sampling remapper given a shared noise table; +-1 LSB per channel for the RBF remappers.
"""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN, random_rgba, synthetic_palette

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import identity, interpolation  # noqa: E402


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))["rgb"]


def _diff(a, b):
    return np.abs(a.astype(np.int16) - b.astype(np.int16))


# ---------------------------------------------------------------- identity.rs
@pytest.mark.parametrize("level", list(range(2, 17)))
def test_identity_bit_exact(level):
    assert (identity.generate(level) == O.identity(level)).all()


def test_identity_rejects_levels_the_reference_panics_on():
    for level in (0, 1):
        with pytest.raises(lg.LutgenPanic):
            identity.generate(level)


def test_detect_level_and_panics():
    assert identity.detect_level(np.zeros((512, 512, 3), np.uint8)) == 8
    with pytest.raises(lg.LutgenPanic):
        identity.detect_level(np.zeros((100, 100, 3), np.uint8))
    with pytest.raises(lg.LutgenPanic):
        identity.detect_level(np.zeros((63, 64, 3), np.uint8))


@pytest.mark.parametrize("level", [2, 3, 5, 8, 12, 16])
@pytest.mark.parametrize("npix", [0, 1, 3, 4, 5, 1023, 65537])
def test_apply_bit_exact_ragged(level, npix):
    rng = np.random.default_rng(level * 1000 + npix)
    lut = rng.integers(0, 256, (level ** 3, level ** 3, 3), dtype=np.uint8)
    img = random_rgba(npix, seed=npix + level).reshape(1, npix, 4) if npix else np.zeros((1, 0, 4), np.uint8)
    want = O.correct_image(img, lut, level)
    got = img.copy()
    identity.correct_image_with_level(got, lut, level)
    assert (got == want).all()


def test_apply_4k_frame_level8_bit_exact(palettes):
    """BASELINE config 2: level-8 CLUT over a 3840x2160 RGBA8 frame, via correct_image."""
    lut = O.Remapper(O.GAUSSIAN, palettes["catppuccin_mocha"], param=128.0, nearest=16).generate_lut(8)
    img = random_rgba(3840 * 2160, seed=42080085).reshape(2160, 3840, 4)
    want = O.correct_image(img, lut, 8, threads=O.max_threads())
    got = img.copy()
    identity.correct_image(got, lut)
    assert (got == want).all()
    assert (got[..., 3] == img[..., 3]).all()


def test_apply_unaligned_views():
    """Frames that start off a 16-byte boundary take the scalar head path."""
    lut = O.identity(4)[:, ::-1].copy()  # any valid level-4 CLUT
    base = random_rgba(4099, seed=9)
    for off in (1, 2, 3):
        img = base[off:off + 4001].copy()
        want = O.correct_image(img, lut, 4)
        got = img.copy()
        identity.correct_image_with_level(got, lut, 4)
        assert (got == want).all()


def test_apply_level16_identity_is_noop():
    img = random_rgba(100003, seed=5)
    got = img.copy()
    identity.correct_image(got.reshape(1, -1, 4), identity.generate(16))
    assert (got == img).all()


def test_apply_chunked_pipeline(palettes, monkeypatch):
    """Host frames travel in chunks through an upload / kernel / download pipeline (api.cu: apply_batch_share). Tiny
    chunks force many of them, with ragged last chunks, across frame boundaries and over several streams."""
    lut = O.Remapper(O.NEAREST, palettes["nord"], lum_factor=1.0).generate_lut(6)
    sizes = (100003, 5, 4000, 999, 1000, 1001, 0, 12345)
    for chunk in ("1000", "4096", "0"):
        monkeypatch.setenv("LUTGEN_CUDA_APPLY_CHUNK", chunk)
        frames = [random_rgba(n, seed=n + 1).reshape(1, n, 4) if n else np.zeros((1, 0, 4), np.uint8) for n in sizes]
        want = [O.correct_image(f, lut, 6) if f.size else f.copy() for f in frames]
        identity.correct_images(frames, lut)
        for f, w in zip(frames, want):
            assert (f == w).all()
        one = random_rgba(77777, seed=3).reshape(1, -1, 4)
        w1 = O.correct_image(one, lut, 6)
        identity.correct_image(one, lut)
        assert (one == w1).all()


def test_apply_batch_matches_single(palettes):
    lut = O.Remapper(O.NEAREST, palettes["nord"], lum_factor=1.0).generate_lut(6)
    frames = [random_rgba(n, seed=n).reshape(1, n, 4) for n in (1000, 1, 77777, 4096, 31)]
    want = [O.correct_image(f, lut, 6) for f in frames]
    identity.correct_images(frames, lut)
    for f, w in zip(frames, want):
        assert (f == w).all()


def test_correct_pixel_cli_known_answer():
    """crates/cli/src/main.rs:1192-1226 through the GPU path end to end."""
    pal = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], np.uint8)
    lut = interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(2)
    assert identity.correct_pixel([255, 128, 0], lut, 2).tolist() == [255, 0, 0]


# --------------------------------------------------------------------- oklab
def test_srgb_to_oklab_bit_exact_whole_cube():
    """Every one of the 2^24 sRGB colours: f32 Oklab identical to the oracle's, bit for bit
    (this is what makes NearestNeighbor argmins and `palette.contains` byte-exact)."""
    import ctypes as C
    from lutgen_rs_b200 import _native
    L = _native.lib()
    step = 1 << 22
    for start in range(0, 1 << 24, step):
        i = np.arange(start, start + step, dtype=np.uint32)
        rgb = np.stack([i & 255, (i >> 8) & 255, i >> 16], 1).astype(np.uint8)
        got = np.empty((step, 3), np.float32)
        _native.check(L.lutgen_cuda_srgb_to_oklab(C.c_void_p(rgb.ctypes.data), step, C.c_void_p(got.ctypes.data)))
        want = O.srgb_to_oklab(rgb)
        assert (got.view(np.uint32) == want.view(np.uint32)).all()


def test_oklab_to_srgb_bit_exact():
    import ctypes as C
    from lutgen_rs_b200 import _native
    rng = np.random.default_rng(1)
    lab = np.concatenate([
        np.stack([rng.uniform(-0.1, 1.1, 400000), rng.uniform(-0.5, 0.5, 400000), rng.uniform(-0.5, 0.5, 400000)], 1),
        np.array([[np.nan, 0, 0], [np.inf, 0, 0], [-np.inf, 0, 0], [0, 0, 0], [1, 0, 0]]),
    ]).astype(np.float32)
    got = np.empty((lab.shape[0], 3), np.uint8)
    _native.check(_native.lib().lutgen_cuda_oklab_to_srgb(C.c_void_p(lab.ctypes.data), lab.shape[0], C.c_void_p(got.ctypes.data)))
    assert (got == O.oklab_to_srgb(lab)).all()


# ------------------------------------------------------- NearestNeighbor (a13/a14)
@pytest.mark.parametrize("pal_name,level,lum,preserve", [
    ("catppuccin_mocha", 8, 1.0, False),
    ("catppuccin_mocha", 8, 0.7, True),
    ("nord", 10, 1.0, False),
    ("carburetor", 12, 1.3, False),
    ("gruvbox_dark", 9, 1.0, True),
])
def test_nearest_neighbor_lut_bit_exact(palettes, pal_name, level, lum, preserve):
    pal = palettes[pal_name]
    want = O.Remapper(O.NEAREST, pal, lum_factor=lum, preserve=preserve).generate_lut(level)
    got = interpolation.NearestNeighborRemapper(pal, lum, preserve).par_generate_lut(level)
    assert (got == want).all()


def test_nearest_neighbor_256_colour_palette_bit_exact():
    pal = synthetic_palette(256, seed=1)
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).generate_lut(10)
    got = interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(10)
    assert (got == want).all()


def test_nearest_neighbor_image_and_pixel(palettes):
    pal = palettes["nord"]
    img = random_rgba(200003, seed=17).reshape(1, -1, 4)
    r = interpolation.NearestNeighborRemapper(pal, 1.0, False)
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).remap_image(img)
    got = img.copy()
    r.remap_image(got)
    assert (got == want).all()
    assert (got[..., 3] == img[..., 3]).all()
    px = np.array([63, 127, 255, 9], np.uint8)
    r.remap_pixel(px)
    assert (px == O.Remapper(O.NEAREST, pal, lum_factor=1.0).remap_image(np.array([[63, 127, 255, 9]], np.uint8))[0]).all()
    # idempotent: palette colours map to themselves
    again = got.copy()
    r.remap_image(again)
    assert (again == got).all()


# ----------------------------------------------------------------- RBF (a7-a12)
def _rbf_case(cls, okind, pal, level, param, nearest, lum, preserve):
    args = (pal, param, nearest, lum, preserve) if param is not None else (pal, nearest, lum, preserve)
    got = cls(*args).par_generate_lut(level)
    want = O.Remapper(okind, pal, param=param or 0.0, nearest=nearest, lum_factor=lum, preserve=preserve).generate_lut(level)
    d = _diff(got, want)
    return d


@pytest.mark.parametrize("pal_name,level,shape,nearest,lum,preserve", [
    ("catppuccin_mocha", 8, 128.0, 16, 1.0, False),   # BASELINE config 1
    ("carburetor", 8, 96.0, 16, 1.0, True),           # reference bench parameters
    ("gruvbox_dark", 8, 96.0, 0, 1.0, False),         # all-palette path
    ("nord", 6, 512.0, 4, 0.7, True),                 # Studio extremes
    ("catppuccin_mocha", 5, 8.0, 30, 1.2, False),     # nearest >= len -> all-palette path
])
def test_gaussian_rbf_lut_within_1lsb(palettes, pal_name, level, shape, nearest, lum, preserve):
    d = _rbf_case(interpolation.GaussianRemapper, O.GAUSSIAN, palettes[pal_name], level, shape, nearest, lum, preserve)
    assert d.max() <= 1, (d.max(), (d > 1).sum())
    assert (d > 0).mean() < 1e-3


@pytest.mark.parametrize("pal_name,level,power,nearest,lum,preserve", [
    ("catppuccin_mocha", 8, 4.0, 16, 1.0, False),
    ("carburetor", 8, 16.0, 16, 1.0, True),
    ("nord", 6, 64.0, 0, 1.0, False),
    ("gruvbox_dark", 6, 1.5, 3, 0.8, False),
])
def test_shepard_lut_within_1lsb(palettes, pal_name, level, power, nearest, lum, preserve):
    d = _rbf_case(interpolation.ShepardRemapper, O.SHEPARD, palettes[pal_name], level, power, nearest, lum, preserve)
    assert d.max() <= 1, (d.max(), (d > 1).sum())
    assert (d > 0).mean() < 1e-3


@pytest.mark.parametrize("nearest", [0, 2, 5])
def test_linear_lut_within_1lsb(palettes, nearest):
    d = _rbf_case(interpolation.LinearRemapper, O.LINEAR, palettes["nord"], 6, None, nearest, 1.0, False)
    assert d.max() <= 1, (d.max(), (d > 1).sum())


def test_shepard_256_colours_nearest16(palettes):
    """BASELINE config 3's remapper at a size the oracle finishes quickly."""
    pal = synthetic_palette(256, seed=1)
    d = _rbf_case(interpolation.ShepardRemapper, O.SHEPARD, pal, 8, 4.0, 16, 1.0, False)
    assert d.max() <= 1, (d.max(), (d > 1).sum())


def test_rbf_image_front_end_and_contains(palettes):
    pal = palettes["catppuccin_mocha"]
    img = random_rgba(100003, seed=23)
    img[: len(pal), :3] = pal  # exact palette colours: left untouched (rbf.rs:66-68)
    img = img.reshape(1, -1, 4)
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    got = img.copy()
    r.remap_image(got)
    want = O.Remapper(O.GAUSSIAN, pal, param=128.0, nearest=16).remap_image(img)
    assert _diff(got, want).max() <= 1
    assert (got[0, : len(pal)] == img[0, : len(pal)]).all()
    assert (got[..., 3] == img[..., 3]).all()


# the reference's own golden images, straight through the GPU path
def test_golden_gruvbox_level8_gpu(palettes):
    got = interpolation.GaussianRemapper(palettes["gruvbox_dark"], 96.0, 0, 1.0, False).par_generate_lut(8)
    d = _diff(got, _golden("gruvbox_dark_hald8"))
    assert d.max() <= 1 and (d > 0).sum() <= 64


@pytest.mark.parametrize("name,pal", [("catppuccin_mocha_hald10", "catppuccin_mocha"), ("nord_hald10", "nord")])
def test_golden_level10_gpu(palettes, name, pal):
    got = interpolation.GaussianRemapper(palettes[pal], 128.0, 16, 1.0, False).par_generate_lut(10)
    d = _diff(got, _golden(name))
    assert d.max() <= 1 and (d > 0).sum() <= 220


# --------------------------------------------------------- Gaussian sampling (a15/a16)
@pytest.mark.parametrize("iterations,level,preserve", [(64, 5, False), (512, 4, False), (100, 4, False), (7, 5, True), (1, 3, False)])
def test_sampling_shared_noise_bit_exact(palettes, iterations, level, preserve):
    pal = palettes["carburetor"]
    g = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, iterations, 1.0, 42080085, preserve)
    o = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=iterations, lum_factor=1.0, seed=1, preserve=preserve)
    o.set_noise(g.noise())
    assert (g.generate_lut(level) == o.generate_lut(level)).all()


def test_sampling_noise_is_standard_normal(palettes):
    g = interpolation.GaussianSamplingRemapper(palettes["nord"], 3.0, 20.0, 20000, 1.0, 7, False)
    nz = g.noise()
    assert abs(nz.mean() - 3.0) < 0.5 and abs(nz.std() - 20.0) < 0.5
    z = (nz - 3.0) / 20.0
    assert abs((np.abs(z) < 1).mean() - 0.6827) < 0.01 and abs((np.abs(z) < 2).mean() - 0.9545) < 0.01
    g2 = interpolation.GaussianSamplingRemapper(palettes["nord"], 3.0, 20.0, 20000, 1.0, 7, False)
    assert (g2.noise() == nz).all()  # counter-based: same seed, same table


def test_sampling_statistical_agreement(palettes):
    """Own RNG stream vs the oracle's own stream: the two differ no more than two oracle
    streams differ from each other (check ii)."""
    pal = palettes["carburetor"]
    level, K = 5, 512
    g = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, K, 1.0, 42080085, False).generate_lut(level)
    o1 = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=K, lum_factor=1.0, seed=1).generate_lut(level)
    o2 = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=K, lum_factor=1.0, seed=2).generate_lut(level)
    ref = _diff(o1, o2)
    for o in (o1, o2):
        d = _diff(g, o)
        assert d.mean() <= 1.5 * ref.mean() + 0.05
        assert np.percentile(d, 99.9) <= 1.5 * np.percentile(ref, 99.9) + 1


def test_sampling_image_front_end(palettes):
    pal = palettes["nord"]
    img = random_rgba(20011, seed=3).reshape(1, -1, 4)
    g = interpolation.GaussianSamplingRemapper(pal, 0.0, 20.0, 32, 1.0, 5, False)
    o = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=32, lum_factor=1.0, seed=5)
    o.set_noise(g.noise())
    got = img.copy()
    g.remap_image(got)
    assert (got == o.remap_image(img)).all()


# ------------------------------------------------------------- interrupts (lib.rs:149-170)
def test_interrupt_semantics(palettes):
    r = interpolation.GaussianRemapper(palettes["nord"], 128.0, 16, 1.0, False)
    flag = lg.AbortFlag()
    assert r.par_generate_lut_with_interrupt(6, flag) is not None
    flag.set()
    assert r.par_generate_lut_with_interrupt(6, flag) is None
    img = random_rgba(1000, seed=1).reshape(1, -1, 4)
    before = img.copy()
    r.par_remap_image_with_interrupt(img, flag)  # flag set: every pixel skipped (mod.rs:54-60)
    assert (img == before).all()


# ------------------------------------------------------------ fast path vs exact path
def _exact_only(monkeypatch, on):
    if on:
        monkeypatch.setenv("LUTGEN_CUDA_EXACT_ONLY", "1")
    else:
        monkeypatch.delenv("LUTGEN_CUDA_EXACT_ONLY", raising=False)


def test_fast_oklab_within_tile_padding():
    """The near-tie tolerance of the warp-tile kernels is built on a MEASURED bound: the fast f32 conversion stays within
    WT_OKLAB_ERR = 8e-7 (as a vector) of the reference's f32 Oklab over all 2^24 colours -- the whole domain, so the
    maximum is a bound (kernels_warp.cuh; measured 5.87e-7, tools/measure_fast_oklab.py). The tile boxes are padded by 2e-5."""
    import ctypes as C
    from lutgen_rs_b200 import _native
    L = _native.lib()
    worst = 0.0
    step = 1 << 22
    for start in range(0, 1 << 24, step):
        i = np.arange(start, start + step, dtype=np.uint32)
        rgb = np.stack([i & 255, (i >> 8) & 255, i >> 16], 1).astype(np.uint8)
        fast = np.empty((step, 3), np.float32)
        exact = np.empty((step, 3), np.float32)
        _native.check(L.lutgen_cuda_srgb_to_oklab_fast(C.c_void_p(rgb.ctypes.data), step, C.c_void_p(fast.ctypes.data)))
        _native.check(L.lutgen_cuda_srgb_to_oklab(C.c_void_p(rgb.ctypes.data), step, C.c_void_p(exact.ctypes.data)))
        d = fast.astype(np.float64) - exact.astype(np.float64)
        worst = max(worst, float(np.sqrt((d * d).sum(1)).max()))
    assert worst < 7e-7, worst   # WT_OKLAB_ERR less the 1e-7 allowed for the image front end's summation order


@pytest.mark.parametrize("lum,preserve", [(1.0, False), (0.7, True)])
def test_level16_nearest_neighbor_fast_equals_exact(palettes, monkeypatch, lum, preserve):
    """Full-size CLUT (16.7M entries): the tile-bounded kernel and the f64 brute-force kernel
    must agree on every byte (NearestNeighbor is the bit-exact row)."""
    pal = palettes["catppuccin_mocha"]
    _exact_only(monkeypatch, True)
    exact = interpolation.NearestNeighborRemapper(pal, lum, preserve).generate_lut(16)
    _exact_only(monkeypatch, False)
    fast = interpolation.NearestNeighborRemapper(pal, lum, preserve).generate_lut(16)
    assert (fast == exact).all()
    # idempotence at full size: every CLUT colour is a palette colour (preserve off)
    if not preserve:
        cols = np.unique(fast.reshape(-1, 3), axis=0)
        assert set(map(tuple, cols.tolist())) <= set(map(tuple, pal.tolist()))


@pytest.mark.parametrize("cls,param,nearest,pal_kind", [
    (interpolation.GaussianRemapper, 128.0, 16, "builtin"),
    (interpolation.ShepardRemapper, 4.0, 16, "synthetic256"),
    (interpolation.GaussianRemapper, 96.0, 0, "builtin"),
])
def test_level16_rbf_fast_vs_exact_within_1lsb(palettes, monkeypatch, cls, param, nearest, pal_kind):
    """BASELINE configs 1/3 at the full 4096x4096: fast f32 kernel vs the f64 kernel, +-1 LSB."""
    pal = palettes["catppuccin_mocha"] if pal_kind == "builtin" else synthetic_palette(256, seed=1)
    _exact_only(monkeypatch, True)
    exact = cls(pal, param, nearest, 1.0, False).generate_lut(16)
    _exact_only(monkeypatch, False)
    fast = cls(pal, param, nearest, 1.0, False).generate_lut(16)
    d = _diff(fast, exact)
    assert d.max() <= 1, (int(d.max()), int((d > 1).sum()))
    assert (d > 0).mean() < 2e-3, float((d > 0).mean())
    # and a strided 1/64 sample against the CPU oracle itself
    okind = O.GAUSSIAN if cls is interpolation.GaussianRemapper else O.SHEPARD
    orc = O.Remapper(okind, pal, param=param, nearest=nearest, lum_factor=1.0, preserve=False)
    rows = list(range(0, 4096, 64))
    for r0 in rows[::8]:
        want = orc.generate_lut(16, rows=(r0, r0 + 1))[r0]
        assert _diff(fast[r0], want).max() <= 1


def test_row_shards_tile_the_whole_lut(palettes):
    """Multi-GPU unit: any split of the rows reproduces the single-launch CLUT byte for byte."""
    import torch
    from lutgen_rs_b200 import device
    device.init(0)
    pal = palettes["nord"]
    for level, cuts in ((8, [0, 1, 100, 101, 300, 512]), (5, [0, 7, 64, 125]), (16, [0, 512, 1000, 4096])):
        r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
        whole = torch.from_numpy(r.generate_lut(level)).reshape(-1)
        lut = torch.zeros(3 * level ** 6, dtype=torch.uint8, device="cuda")
        for a, b in zip(cuts[:-1], cuts[1:]):
            device.generate_lut_rows_(r, level, lut, a, b)
        torch.cuda.synchronize()
        assert torch.equal(lut.cpu(), whole)


def test_large_image_goes_through_the_colour_cube(palettes, monkeypatch):
    """remap_image on >= 2^20 pixels tabulates the remapper over the sRGB cube and gathers; the
    result must equal the per-pixel kernels' (same bytes for NN, +-1 LSB for RBF)."""
    pal = palettes["catppuccin_mocha"]
    img = random_rgba((1 << 20) + 12345, seed=77).reshape(1, -1, 4)
    for make, exact_bytes in ((lambda: interpolation.NearestNeighborRemapper(pal, 1.0, False), True),
                              (lambda: interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), False)):
        monkeypatch.setenv("LUTGEN_CUDA_NO_TABLE", "1")
        direct = img.copy()
        make().remap_image(direct)
        monkeypatch.delenv("LUTGEN_CUDA_NO_TABLE")
        via = img.copy()
        make().remap_image(via)
        d = _diff(via, direct)
        assert d.max() <= (0 if exact_bytes else 1)
        assert (via[..., 3] == img[..., 3]).all()
    sample = img[:, :20000]
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).remap_image(sample)
    got = img.copy()
    interpolation.NearestNeighborRemapper(pal, 1.0, False).remap_image(got)
    assert (got[:, :20000] == want).all()


def test_duplicate_palette_colours(palettes):
    """A palette with repeated colours (4 built-ins have them): NN stays bit exact, RBF +-1 LSB."""
    pal = np.concatenate([palettes["nord"], palettes["nord"][:5], palettes["nord"][3:4]])
    want = O.Remapper(O.NEAREST, pal, lum_factor=1.0).generate_lut(8)
    assert (interpolation.NearestNeighborRemapper(pal, 1.0, False).generate_lut(8) == want).all()
    d = _rbf_case(interpolation.GaussianRemapper, O.GAUSSIAN, pal, 8, 128.0, 16, 1.0, False)
    assert d.max() <= 1
