"""PNG writer on the GPU (csrc/png.cu; SURVEY 8(f) row 2: `lut.save(path)` / `image.save(path)`,
crates/cli/src/main.rs:767-871) through the C ABI. Parity property of the domain: lossless round trip
through independent decoders (PIL; a from-the-specification reader over zlib with every CRC and the
Adler-32 checked). On top of that the CUDA kernels and the host instantiation of the same phases
(tests/png_sim.cpp) must write the very same bytes -- the pipeline is deterministic."""
import io

import numpy as np
import pytest

import oracle_lib as O
import png_sim as P
from conftest import synthetic_palette

pytestmark = pytest.mark.gpu

lg = pytest.importorskip("lutgen_rs_b200")
from lutgen_rs_b200 import interpolation, png  # noqa: E402


def _pil(data):
    from PIL import Image
    a = np.array(Image.open(io.BytesIO(data)))
    return a[:, :, None] if a.ndim == 2 else a


def _check(img, own=True, sim=True):
    data = png.encode(img)
    assert (_pil(data) == img).all()
    if own:
        dec, _, _ = P.decode(data)
        assert (dec == img).all()
    if sim:
        assert data == P.encode(img)[0]
    assert len(data) <= png.bound(img.shape[1], img.shape[0], img.shape[2])
    return data


@pytest.mark.parametrize("c", [1, 2, 3, 4])
@pytest.mark.parametrize("h,w", [(1, 1), (1, 300), (300, 1), (37, 53), (128, 128), (3, 5461), (2, 10923)])
def test_png_round_trip_and_same_bytes_as_the_host_instantiation(h, w, c):
    rng = np.random.default_rng(h * 1000 + w * 10 + c)
    own = h * w * c < 60000
    _check(rng.integers(0, 256, (h, w, c), dtype=np.uint8), own)         # stored blocks
    _check(np.full((h, w, c), 200, np.uint8), own)                        # one long run
    _check(rng.integers(0, 4, (h, w, c), dtype=np.uint8), own)            # short codes, accidental runs
    g = (np.arange(w)[None, :, None] * 3 + np.arange(h)[:, None, None] + np.arange(c)[None, None, :] * 50) & 255
    _check(g.astype(np.uint8), own)


@pytest.mark.parametrize("run", [3, 4, 257, 258, 259, 260, 261, 517, 518, 16000, 40000])
def test_png_run_lengths_around_the_match_limits(run):
    rng = np.random.default_rng(run)
    row = rng.integers(0, 256, (1, (run + 2) // 3 + 40, 3), dtype=np.uint8)
    row.reshape(-1)[17:17 + run] = 77
    _check(row, own=False)
    _check(np.concatenate([row, row, row]), own=False)


def test_png_deep_huffman_code():
    fib = [1, 1]
    while len(fib) < 19:
        fib.append(fib[-1] + fib[-2])
    vals = np.concatenate([np.full(f, 3 * i + 1, np.uint8) for i, f in enumerate(fib)])
    np.random.default_rng(4).shuffle(vals)
    _check(vals.reshape(1, -1, 1))


def test_png_of_cluts(palettes):
    pal = palettes["catppuccin_mocha"]
    for level in (4, 8):
        lut = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False).par_generate_lut(level)
        data = _check(lut, own=level <= 4)
        assert len(data) < (0.5 if level >= 8 else 0.7) * lut.size
    nn = interpolation.NearestNeighborRemapper(pal, 1.0, False).par_generate_lut(8)
    assert len(_check(nn, own=False)) < 0.15 * nn.size


def test_generate_lut_png_is_generate_then_encode(palettes):
    pal = palettes["nord"]
    r = interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False)
    for level in (5, 8):
        lut = r.par_generate_lut(level)
        data = png.generate_lut_png(r, level)
        assert (_pil(data) == lut).all()
        assert data == png.encode(lut)
    flag = lg.AbortFlag(True)
    assert png.generate_lut_png(r, 4, flag) is None
    b = interpolation.GaussianBlurRemapper(pal, 4.0, 1.0, False)
    assert (_pil(png.generate_lut_png(b, 6)) == b.par_generate_lut(6)).all()


def test_png_level16_clut_and_4k_frame(palettes):
    """BASELINE sizes: the 4096 x 4096 CLUT of the headline configuration and a 4K RGBA frame."""
    r = interpolation.GaussianRemapper(palettes["catppuccin_mocha"], 128.0, 16, 1.0, False)
    lut = r.par_generate_lut(16)
    data = png.generate_lut_png(r, 16)
    assert (_pil(data) == lut).all()
    assert len(data) < 0.5 * lut.size
    rng = np.random.default_rng(9)
    h, w = 2160, 3840
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 90.0 + yy / 130.0), 128 + 90 * np.cos(xx / 70.0), 128 + 80 * np.sin(yy / 50.0)], -1)
    frame = np.clip(base + rng.normal(0, 2, (h, w, 3)), 0, 255).astype(np.uint8)
    frame = np.concatenate([frame, np.full((h, w, 1), 255, np.uint8)], -1)
    data = png.encode(frame)
    assert (_pil(data) == frame).all()
    assert len(data) < 0.7 * frame.size


def test_correct_image_png_is_correct_then_encode(palettes):
    """`lutgen apply` on a still image (crates/cli/src/main.rs:852-871): correct_image + save, fused."""
    from lutgen_rs_b200 import identity
    lut = interpolation.GaussianRemapper(palettes["catppuccin_mocha"], 128.0, 16, 1.0, False).par_generate_lut(8)
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (123, 457, 4), dtype=np.uint8)
    before = img.copy()
    data = png.correct_image_png(img, lut)
    assert (img == before).all()
    want = O.correct_image(before, lut, 8)
    assert (_pil(data) == want).all()
    corrected = before.copy()
    identity.correct_image(corrected, lut)
    assert data == png.encode(corrected)
    with pytest.raises(lg.LutgenPanic):
        png.correct_image_png(img, lut[:100])


def test_png_device_entry_point(palettes):
    import torch
    from lutgen_rs_b200 import device
    device.init(0)
    rng = np.random.default_rng(3)
    img = (rng.integers(0, 8, (200, 333, 3)) * 9).astype(np.uint8)
    t = torch.from_numpy(img).cuda()
    out = device.encode_png_(t)
    torch.cuda.synchronize()
    data = out.cpu().numpy().tobytes()
    assert data == png.encode(img)
    assert (_pil(data) == img).all()
    with pytest.raises(lg.LutgenPanic):
        device.encode_png_(t, out=torch.empty(100, dtype=torch.uint8, device="cuda"))


def test_png_argument_errors():
    with pytest.raises(lg.LutgenPanic):
        png.encode(np.zeros((0, 4, 3), np.uint8))
    with pytest.raises(TypeError):
        png.encode(np.zeros((4, 4, 5), np.uint8))
    with pytest.raises(TypeError):
        png.encode(np.zeros((4, 4, 3), np.float32))
    import ctypes as C
    from lutgen_rs_b200 import _native
    img = np.zeros((64, 64, 3), np.uint8)
    small = np.empty(16, np.uint8)
    n = C.c_size_t(0)
    rc = _native.lib().lutgen_cuda_encode_png(img.ctypes.data, 64, 64, 3, small.ctypes.data, small.size, C.byref(n))
    assert rc == _native.ERR_INVALID and n.value == len(png.encode(img))
