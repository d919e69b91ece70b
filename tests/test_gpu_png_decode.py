"""GPU tests of the PNG reader (lutgen_cuda_decode_png*, csrc/png_decode.cu): byte-equal pixels against PIL for every
colour type and filter, against this library's own writer (the parallel piece path), and the file-to-file
`lutgen apply` call. The reference's committed CLUT images are covered as fixtures: tests/golden/*.npz hold their
decoded pixels and tests/golden/gruvbox-dark-hald-clut.png is the file itself."""
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PIL = pytest.importorskip("PIL.Image")
lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import _native, identity, interpolation, png  # noqa: E402


def _png(arr, mode, **kw):
    bio = io.BytesIO()
    PIL.fromarray(arr, mode).save(bio, "PNG", **kw)
    return bio.getvalue()


def _photo(h, w, c, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 255 // max(w + h - 2, 1)), np.full_like(x, 200)], -1)[..., :c]
    return (base + rng.integers(-6, 7, base.shape)).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("mode,c", [("L", 1), ("LA", 2), ("RGB", 3), ("RGBA", 4)])
@pytest.mark.parametrize("shape", [(1, 1), (2, 77), (33, 64), (67, 5), (131, 213), (480, 641)])
def test_decode_equals_pil(mode, c, shape):
    h, w = shape
    img = _photo(h, w, 4, seed=h * 1000 + w)
    arr = np.ascontiguousarray(img[..., 0] if c == 1 else img[..., [0, 3]] if c == 2 else img[..., :c])
    data = _png(arr, mode, compress_level=6 if (h + w) % 2 else 1)
    want = np.array(PIL.open(io.BytesIO(data)))
    got = png.decode(data)
    assert got.shape == (h, w, c) and (got.reshape(want.shape) == want).all()
    assert png.info(data) == (w, h, c, c in (2, 4))
    assert (png.decode(data, 4) == np.array(PIL.open(io.BytesIO(data)).convert("RGBA"))).all()
    assert (png.decode(data, 3) == np.array(PIL.open(io.BytesIO(data)).convert("RGB"))).all()


def test_decode_palette_images():
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 200, (150, 211), dtype=np.uint8)
    im = PIL.fromarray(idx, "P")
    im.putpalette(rng.integers(0, 256, 200 * 3, dtype=np.uint8).tolist())
    bio = io.BytesIO()
    im.save(bio, "PNG")
    assert (png.decode(bio.getvalue()) == np.array(im.convert("RGB"))).all()
    bio = io.BytesIO()
    im.save(bio, "PNG", transparency=bytes(rng.integers(0, 256, 200, dtype=np.uint8).tolist()))
    data = bio.getvalue()
    got = png.decode(data)
    assert got.shape[2] == 4 and (got == np.array(PIL.open(io.BytesIO(data)).convert("RGBA"))).all()


def test_own_files_round_trip_through_the_piece_path_and_foreign_ones_through_the_serial_path():
    """encode -> decode on the device for a CLUT (3073 pieces at level 16 would be the headline; level 9 here keeps the
    test short), a frame with alpha and noise (stored pieces); the same pixels written by PIL come back the same."""
    pal = np.array([[30, 30, 46], [205, 214, 244], [243, 139, 168], [166, 227, 161], [137, 180, 250]], np.uint8)
    lut = interpolation.GaussianRemapper(pal, 128.0, 0, 1.0, False).par_generate_lut(9)
    for img in (lut, _photo(333, 517, 4, seed=2), np.random.default_rng(4).integers(0, 256, (100, 300, 3), dtype=np.uint8)):
        data = png.encode(img)
        assert (png.decode(data) == img).all()
        assert (png.decode(_png(img, "RGB" if img.shape[2] == 3 else "RGBA")) == img).all()


def test_level16_clut_file_round_trip():
    lut = interpolation.NearestNeighborRemapper(np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [20, 20, 20]], np.uint8), 1.0, False).generate_lut(16)
    data = png.encode(lut)
    assert (png.decode(data, 3) == lut).all()


def test_reference_clut_asset_decodes_to_the_golden_pixels():
    """tests/golden/gruvbox-dark-hald-clut.png is the reference's docs/assets file (a foreign, single-stream PNG: the
    serial path); gruvbox_dark_hald8.npz holds what PIL made of it (tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "gruvbox-dark-hald-clut.png")
    data = open(path, "rb").read()
    want = np.load(os.path.join(GOLDEN, "gruvbox_dark_hald8.npz"))
    want = want[want.files[0]]
    got = png.decode(data, 3)
    assert got.shape == want.shape and (got == want).all()
    assert identity.detect_level(got) == 8


def test_apply_file_to_file_equals_the_separate_steps():
    pal = np.array([[40, 40, 40], [235, 219, 178], [204, 36, 29], [152, 151, 26], [69, 133, 136]], np.uint8)
    lut = interpolation.GaussianRemapper(pal, 96.0, 0, 1.0, False).par_generate_lut(8)
    img = _photo(270, 480, 4, seed=11)
    out = png.correct_png_file(_png(img, "RGBA"), png.encode(lut))
    want = img.copy()
    identity.correct_image(want, lut)
    assert (np.array(PIL.open(io.BytesIO(out)).convert("RGBA")) == want).all()
    # an RGB image file gains alpha 255 (to_rgba8), a PIL-written CLUT file works as well as our own
    out = png.correct_png_file(_png(img[..., :3].copy(), "RGB"), _png(lut, "RGB"))
    want = np.concatenate([img[..., :3], np.full_like(img[..., :1], 255)], -1)
    identity.correct_image(want, lut)
    assert (png.decode(out) == want).all()


def test_damaged_and_unsupported_files_are_reported():
    img = _photo(60, 80, 3)
    data = bytearray(_png(img, "RGB"))
    with pytest.raises(_native.LutgenError):
        png.decode(b"GIF89a not a png")
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x10
    with pytest.raises(_native.LutgenError):
        png.decode(bytes(bad))
    with pytest.raises(_native.LutgenError):
        png.decode(_png((img[..., 0].astype(np.uint16) * 257), "I;16"))
    with pytest.raises(_native.LutgenError):   # not a Hald CLUT: 80 x 60
        png.correct_png_file(bytes(data), bytes(data))
