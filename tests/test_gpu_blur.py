"""GaussianBlurRemapper (crates/lib/src/interpolation/gaussian_blur.rs; SURVEY 8(f) row 1, the CLI's
and Studio's default algorithm): the CUDA path must be BIT EXACT with the CPU restatement -- every
step is f32 arithmetic evaluated in the reference's order. Nothing in the reference pins this
remapper's output (no golden image, no test): parity is against the oracle's restatement only."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import synthetic_palette

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import identity, interpolation  # noqa: E402


def _case(pal, level, radius, lum=1.0, preserve=False):
    got = interpolation.GaussianBlurRemapper(pal, radius, lum, preserve).par_generate_lut(level)
    want = O.Remapper(O.BLUR, pal, param=radius, lum_factor=lum, preserve=preserve).generate_lut(level)
    return got, want


@pytest.mark.parametrize("pal_name,level,radius,lum,preserve", [
    ("catppuccin_mocha", 10, 8.0, 1.0, False),   # CLI defaults (crates/cli/src/main.rs:70, 139)
    ("carburetor", 8, 8.0, 1.0, False),          # reference bench parameters (benches/main.rs:136-138)
    ("nord", 8, 8.0, 0.7, True),                 # Studio defaults: lum 0.7, preserve (state.rs:277-288)
    ("gruvbox_dark", 12, 3.5, 1.0, False),       # Studio's native default level
    ("catppuccin_mocha", 6, 0.4, 1.3, True),     # 5 taps
    ("nord", 5, 20.0, 1.0, False),               # kernel wider than the cube
    ("nord", 2, 1.0, 1.0, False),                # smallest level
    ("carburetor", 7, 2.0, 1.0, True),           # odd level, cube side 49
])
def test_blur_lut_bit_exact(palettes, pal_name, level, radius, lum, preserve):
    got, want = _case(palettes[pal_name], level, radius, lum, preserve)
    assert got.shape == want.shape
    assert (got == want).all(), int((got != want).sum())


def test_blur_level16_bit_exact(palettes):
    got, want = _case(palettes["catppuccin_mocha"], 16, 8.0)
    assert (got == want).all(), int((got != want).sum())


@pytest.mark.parametrize("level,radius,preserve", [(8, 8.0, False), (16, 3.0, True), (4, 2.0, False)])
def test_blur_bulk_and_plain_staging_agree(palettes, monkeypatch, level, radius, preserve):
    """The R and G passes stage their tiles with cp.async.bulk rows + an mbarrier per buffer where every row is a 16-byte
    aligned run (kernels_blur.cuh), with 4-byte cp.async otherwise; LUTGEN_CUDA_BLUR_NO_BULK forces the latter."""
    pal = palettes["catppuccin_mocha"]
    monkeypatch.delenv("LUTGEN_CUDA_BLUR_NO_BULK", raising=False)
    bulk = interpolation.GaussianBlurRemapper(pal, radius, 1.0, preserve).par_generate_lut(level)
    monkeypatch.setenv("LUTGEN_CUDA_BLUR_NO_BULK", "1")
    plain = interpolation.GaussianBlurRemapper(pal, radius, 1.0, preserve).par_generate_lut(level)
    assert (bulk == plain).all()
    if level <= 8:
        want = O.Remapper(O.BLUR, pal, param=radius, preserve=preserve).generate_lut(level)
        assert (bulk == want).all()


def test_blur_256_colour_palette(palettes):
    got, want = _case(synthetic_palette(256, seed=1), 8, 4.0)
    assert (got == want).all()


@pytest.mark.parametrize("radius", [0.0, -1.0, float("nan"), 1e-3])
def test_blur_degenerate_radii(palettes, radius):
    """radius 0 -> one NaN tap -> black; negative -> no taps -> zeros; tiny -> identity kernel."""
    got, want = _case(palettes["nord"], 4, radius)
    assert (got == want).all()
    got, want = _case(palettes["nord"], 4, radius, preserve=True)
    assert (got == want).all()


def test_blur_sequential_hint_variant_agrees_here(palettes):
    """generate_lut_inner carries the NN hint across rows, par_generate_lut_inner restarts it; the two
    can only differ where a stale hint is within the early-exit threshold. Record that they agree on
    the CLI default configuration (the tests below use a palette where they do not)."""
    pal = palettes["catppuccin_mocha"]
    orc = O.Remapper(O.BLUR, pal, param=8.0, lum_factor=1.0, preserve=False)
    assert (orc.generate_lut(8) == orc.generate_lut(8, sequential_hints=True)).all()


def _near_duplicate_palette():
    """Colours a few sRGB steps apart: at coarse levels the early-exit threshold 0.5/size^2 reaches
    from one to the other, so the nearest-neighbour hint decides and the two variants differ."""
    base = np.array([[200, 60, 60], [60, 200, 60], [60, 60, 200], [128, 128, 128], [20, 20, 20], [240, 240, 240]], np.int16)
    near = base + np.array([[3, -2, 2], [-3, 2, 1], [2, 3, -2], [2, 2, 2], [3, 3, 3], [-3, -3, -3]], np.int16)
    near2 = base + np.array([[-2, 3, -3], [2, -3, -2], [-3, -2, 3], [-3, -3, -3], [5, 4, 6], [-6, -5, -4]], np.int16)
    return np.concatenate([base, near, near2]).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("level,radius,preserve", [(3, 1.0, False), (4, 2.0, False), (5, 1.5, True), (6, 3.0, False)])
def test_blur_sequential_variant_follows_generate_lut_inner(level, radius, preserve):
    """generate_lut (sequential: the hint carries on from row to row, gaussian_blur.rs:213) against
    the oracle's sequential restatement, on a palette where it differs from the parallel variant;
    par_generate_lut on the same palette against the parallel restatement."""
    pal = _near_duplicate_palette()
    r = interpolation.GaussianBlurRemapper(pal, radius, 1.0, preserve)
    orc = O.Remapper(O.BLUR, pal, param=radius, lum_factor=1.0, preserve=preserve)
    want_seq, want_par = orc.generate_lut(level, sequential_hints=True), orc.generate_lut(level)
    got_seq, got_par = r.generate_lut(level), r.par_generate_lut(level)
    assert (got_par == want_par).all(), int((got_par != want_par).sum())
    assert (got_seq == want_seq).all(), int((got_seq != want_seq).sum())
    # and alternating between them does not serve one from the other's cached volume
    assert (r.par_generate_lut(level) == want_par).all() and (r.generate_lut(level) == want_seq).all()


def test_blur_the_two_variants_do_differ_somewhere():
    pal = _near_duplicate_palette()
    orc = O.Remapper(O.BLUR, pal, param=1.0, lum_factor=1.0, preserve=False)
    assert any((orc.generate_lut(lv) != orc.generate_lut(lv, sequential_hints=True)).any() for lv in (3, 4, 5, 6))


def test_blur_generate_variants_and_interrupt(palettes):
    r = interpolation.GaussianBlurRemapper(palettes["nord"], 8.0, 1.0, False)
    a = r.generate_lut(6)
    assert (r.par_generate_lut(6) == a).all()
    flag = lg.AbortFlag()
    assert (r.par_generate_lut_with_interrupt(6, flag) == a).all()
    flag.set()
    assert r.generate_lut_with_interrupt(6, flag) is None
    with pytest.raises(AttributeError):
        r.remap_image(np.zeros((2, 2, 4), np.uint8))
    with pytest.raises(lg.LutgenPanic):
        interpolation.GaussianBlurRemapper(np.zeros((0, 3), np.uint8), 8.0, 1.0, False)


def test_blur_rows_and_apply(palettes):
    """Row shards of a blur CLUT tile the whole CLUT; the CLUT then drives correct_image."""
    import torch
    from lutgen_rs_b200 import device
    device.init(0)
    pal = palettes["catppuccin_mocha"]
    r = interpolation.GaussianBlurRemapper(pal, 8.0, 1.0, False)
    whole = r.generate_lut(8)
    lut = torch.zeros(3 * 8 ** 6, dtype=torch.uint8, device="cuda")
    for a, b in ((0, 100), (100, 101), (101, 512)):
        device.generate_lut_rows_(r, 8, lut, a, b)
    torch.cuda.synchronize()
    assert (lut.cpu().numpy().reshape(512, 512, 3) == whole).all()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (100, 200, 4), dtype=np.uint8)
    got = img.copy()
    identity.correct_image(got, whole)
    assert (got == O.correct_image(img, whole, 8)).all()
