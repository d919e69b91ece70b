"""CPU-only: pins the oracle (oracle/lutgen_oracle.c) to everything the reference's own
tests and committed assets hold for the hot path (SURVEY.md sections 4 and 8c)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN, random_rgba


def _golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))["rgb"]


# reference doctest crates/lib/src/lib.rs:51-58 wrote docs/assets/gruvbox-dark-hald-clut.png
def test_golden_gruvbox_dark_level8(palettes):
    g = _golden("gruvbox_dark_hald8")
    lut = O.Remapper(O.GAUSSIAN, palettes["gruvbox_dark"], param=96.0, nearest=0, lum_factor=1.0,
                     preserve=False).generate_lut(8)
    d = np.abs(lut.astype(np.int16) - g.astype(np.int16))
    assert d.max() <= 1
    assert (d > 0).sum() <= 64, (d > 0).sum()  # observed 34 of 786432 (cbrt/FMA flavour noise)


# CLI defaults of the time (shape 128, nearest 16, lum 1, preserve off), level 10
@pytest.mark.parametrize("name,pal,limit", [("catppuccin_mocha_hald10", "catppuccin_mocha", 200),
                                            ("nord_hald10", "nord", 220)])
def test_golden_level10_nearest16(palettes, name, pal, limit):
    g = _golden(name)
    lut = O.Remapper(O.GAUSSIAN, palettes[pal], param=128.0, nearest=16, lum_factor=1.0,
                     preserve=False).generate_lut(10)
    d = np.abs(lut.astype(np.int16) - g.astype(np.int16))
    assert d.max() <= 1
    assert (d > 0).sum() <= limit, (d > 0).sum()  # observed 118 / 134 of 3,000,000


# crates/cli/src/main.rs:1192-1226 `patch_hex_preserves_full_color`
def test_cli_known_answer_patch():
    pal = [[255, 0, 0], [0, 255, 0], [0, 0, 255]]
    lut = O.Remapper(O.NEAREST, pal, lum_factor=1.0, preserve=False).generate_lut(2)
    assert lut.shape == (8, 8, 3)
    assert O.correct_pixel([255, 128, 0], lut, 2).tolist() == [255, 0, 0]
    assert O.identity(2).reshape(-1, 3)[7].tolist() == [255, 85, 0]


def test_identity_matches_formula():
    for level in (2, 3, 5, 8, 16):
        cs = level * level
        ident = O.identity(level).reshape(-1, 3).astype(np.int64)
        i = np.arange(level ** 6)
        r, g, b = i % cs, (i // cs) % cs, i // (cs * cs)
        want = np.stack([r * 255 // (cs - 1), g * 255 // (cs - 1), b * 255 // (cs - 1)], 1)
        assert (ident == want).all()
    assert (O.identity(16).reshape(-1, 3)[:, 0] == np.arange(2 ** 24) % 256).all()


def test_detect_level():
    for level in range(2, 17):
        assert O.detect_level(level ** 3, level ** 3) == level
    assert O.detect_level(100, 100) == -1  # assert_eq!(width, level^3) panics
    assert O.detect_level(64, 63) == -1  # assert_eq!(width, height) panics


def test_apply_identity_is_quantisation():
    """correct_image with the identity CLUT returns q(c)*255/(cs-1) per channel; alpha kept."""
    img = random_rgba(10007, 7)
    for level in (2, 4, 8, 16):
        cs = level * level
        out = O.correct_image(img, O.identity(level), level)
        q = img[:, :3].astype(np.int64) * (cs - 1) // 255
        want = q * 255 // (cs - 1)
        assert (out[:, :3] == want).all()
        assert (out[:, 3] == img[:, 3]).all()
    assert (O.correct_image(img, O.identity(16), 16) == img).all()


def test_apply_threads_agree():
    img = random_rgba(50021, 3)
    lut = O.Remapper(O.NEAREST, [[0, 0, 0], [255, 255, 255], [200, 30, 30]], lum_factor=1.0).generate_lut(6)
    assert (O.correct_image(img, lut, 6, threads=1) == O.correct_image(img, lut, 6, threads=4)).all()


def test_cbrt_correctly_rounded():
    """cr_cbrtf vs an exact integer check on a sample of LMS-range floats."""
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(1e-5, 1.0, 20000), rng.uniform(0.0, 1e-3, 5000)]).astype(np.float32)
    from fractions import Fraction
    for x in xs[:3000].tolist():
        y = np.float32(O.lib().lutgen_oracle_cbrtf(x))
        lo = np.nextafter(y, np.float32(0))
        hi = np.nextafter(y, np.float32(2))
        fx = Fraction(float(x))
        m_lo = (Fraction(float(lo)) + Fraction(float(y))) / 2
        m_hi = (Fraction(float(hi)) + Fraction(float(y))) / 2
        assert m_lo ** 3 < fx < m_hi ** 3, x


def test_oklab_roundtrip_all_grey_and_primaries():
    rgb = np.array([[i, i, i] for i in range(256)] + [[255, 0, 0], [0, 255, 0], [0, 0, 255]], np.uint8)
    back = O.oklab_to_srgb(O.srgb_to_oklab(rgb))
    assert np.abs(back.astype(int) - rgb.astype(int)).max() <= 1


def test_rbf_contains_passthrough(palettes):
    """rbf.rs:66-68: a pixel equal to a palette colour is returned unchanged."""
    pal = palettes["catppuccin_mocha"]
    img = np.concatenate([pal, np.full((len(pal), 1), 77, np.uint8)], 1)
    for kind, param in ((O.GAUSSIAN, 128.0), (O.SHEPARD, 4.0), (O.LINEAR, 0.0)):
        out = O.Remapper(kind, pal, param=param, nearest=16, lum_factor=0.7, preserve=True).remap_image(img)
        assert (out == img).all()


def test_sampling_is_deterministic_stencil(palettes):
    """gaussian_sample.rs:52 re-seeds per pixel: equal colours map equally, alpha kept, and the
    result only depends on the shared noise table."""
    pal = palettes["carburetor"]
    r = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=64, lum_factor=1.0, seed=42080085)
    img = random_rgba(512, 11)
    img[256:] = img[:256]
    out = r.remap_image(img)
    assert (out[256:, :3] == out[:256, :3]).all()
    assert (out[:, 3] == img[:, 3]).all()
    r2 = O.Remapper(O.SAMPLING, pal, mean=0.0, std_dev=20.0, iterations=64, lum_factor=1.0, seed=1)
    r2.set_noise(r.get_noise())
    assert (r2.remap_image(img) == out).all()
