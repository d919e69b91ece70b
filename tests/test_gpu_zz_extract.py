"""Palette extraction on the GPU (csrc/kernels_extract.cuh; SURVEY 8(f) row 4, `lutgen extract`,
crates/cli/src/main.rs:772-815) through the C ABI: bit exact with the oracle's restatement of the specification
in csrc/extract_core.h. PARITY WITH THE REFERENCE IS UNPINNED (its `quantette` dependency is not vendored and
nothing in the reference pins its output); tests/test_extract_core.py holds the quality comparison with
scikit-learn's k-means and runs the same kernel bodies on the host."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import extract, interpolation  # noqa: E402


def _photo(h, w, seed, noise=6.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 19.0 + yy / 31.0), 128 + 90 * np.cos(xx / 13.0), 128 + 80 * np.sin(yy / 11.0)], -1)
    return np.clip(base + rng.normal(0, noise, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("k", [1, 2, 8, 31, 64])
@pytest.mark.parametrize("iters", [0, 1, 16])
def test_extract_palette_bit_exact(k, iters):
    img = _photo(48, 64, seed=k)
    got, want = extract.extract_palette(img, k, iters), O.extract_palette(img, k, iters)
    assert got.shape == want.shape and (got == want).all()


def test_extract_palette_larger_image_and_full_palette():
    img = _photo(300, 400, seed=11)
    for k, iters in ((16, 16), (255, 4), (256, 2)):
        got, want = extract.extract_palette(img, k, iters), O.extract_palette(img, k, iters)
        assert len(got) == k and (got == want).all(), k
    # twice the same answer (nothing depends on the order in which threads append or add)
    assert (extract.extract_palette(img, 16, 16) == extract.extract_palette(img, 16, 16)).all()


def test_extract_palette_degenerate_inputs():
    img = np.zeros((10, 10, 3), np.uint8)
    img[:5] = (200, 30, 30)
    img[5:, :3] = (10, 10, 250)
    for k in (1, 2, 3, 4, 200):
        got, want = extract.extract_palette(img, k, 16), O.extract_palette(img, k, 16)
        assert len(got) == min(k, 3) and (got == want).all()
    one = np.full((4, 4, 3), 77, np.uint8)
    assert extract.extract_palette(one, 8, 3).tolist() == [[77, 77, 77]]
    tie = np.array([[[9, 0, 0], [3, 0, 0], [3, 1, 0], [9, 0, 0], [3, 0, 0]]], np.uint8)
    assert extract.extract_palette(tie, 1, 0).tolist() == [[3, 0, 0]]
    rng = np.random.default_rng(3)
    heavy = np.concatenate([np.tile(np.array([[12, 200, 40]], np.uint8), (70000, 1)), rng.integers(0, 256, (5000, 3), dtype=np.uint8)])
    assert (extract.extract_palette(heavy, 40, 8) == O.extract_palette(heavy, 40, 8)).all()
    assert len(extract.extract_palette(np.zeros((0, 3), np.uint8), 8)) == 0
    with pytest.raises(lg.LutgenPanic):
        extract.extract_palette(img, 0)
    with pytest.raises(lg.LutgenPanic):
        extract.extract_palette(img, 257)


def test_extract_then_generate_like_the_cli():
    """`lutgen extract`: the extracted palette feeds a remapper (crates/cli/src/main.rs:805-811)."""
    img = _photo(120, 160, seed=4)
    pal = extract.extract_palette(img, 12)
    assert pal.shape == (12, 3)
    lut = interpolation.GaussianRemapper(pal, 128.0, 0, 1.0, False).par_generate_lut(4)
    want = O.Remapper(O.GAUSSIAN, O.extract_palette(img, 12), param=128.0, nearest=0).generate_lut(4)
    assert np.abs(lut.astype(np.int16) - want.astype(np.int16)).max() <= 1
