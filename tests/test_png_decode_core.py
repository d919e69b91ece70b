"""CPU tests of the PNG decoder core (lutgen_rs_b200/csrc/png_decode_core.h, host build tests/png_dec_sim.cpp): the
inflate, Adler-32, container parsing and per-pixel reconstruction the CUDA kernels run, against zlib and PIL."""
import io
import os
import zlib

import numpy as np
import pytest

import png_dec_sim as D
import png_sim

PIL = pytest.importorskip("PIL.Image")


def _png(arr, mode=None, **kw):
    bio = io.BytesIO()
    PIL.fromarray(arr, mode).save(bio, "PNG", **kw)
    return bio.getvalue()


def _photo(h, w, c, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 255 // max(w + h - 2, 1)), np.full_like(x, 200)], -1)[..., :c]
    return (base + rng.integers(-6, 7, base.shape)).clip(0, 255).astype(np.uint8)


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("kind", ["zeros", "text", "noise", "runs"])
def test_inflate_matches_zlib(level, kind):
    rng = np.random.default_rng(1)
    raw = {"zeros": bytes(70000), "text": b"the quick brown fox jumps over the lazy dog " * 3000,
           "noise": rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(),
           "runs": np.repeat(rng.integers(0, 256, 700, dtype=np.uint8), rng.integers(1, 400, 700)).tobytes()}[kind]
    rc, got = D.inflate(zlib.compress(raw, level), len(raw))
    assert rc == 0 and got == raw


def test_inflate_fixed_huffman_and_empty_stream():
    c = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)
    raw = b"abcabcabcabc" * 50
    rc, got = D.inflate(c.compress(raw) + c.flush(), len(raw))
    assert rc == 0 and got == raw
    rc, got = D.inflate(zlib.compress(b""), 0)
    assert rc == 0 and got == b""


def test_inflate_rejects_damage():
    raw = bytes(range(256)) * 40
    z = bytearray(zlib.compress(raw, 6))
    rc, _ = D.inflate(bytes(z[: len(z) // 2]), len(raw))
    assert rc != 0                       # truncated
    rc, _ = D.inflate(bytes(z), len(raw) - 1)
    assert rc != 0                       # more output than announced
    z[0] = 0x79
    rc, _ = D.inflate(bytes(z), len(raw))
    assert rc != 0                       # not a zlib header


@pytest.mark.parametrize("mode,c", [("L", 1), ("LA", 2), ("RGB", 3), ("RGBA", 4)])
@pytest.mark.parametrize("shape", [(1, 1), (3, 70), (67, 5), (120, 213)])
def test_decode_matches_pil_for_every_colour_type(mode, c, shape):
    h, w = shape
    img = _photo(h, w, max(c, 3), seed=h * 1000 + w)
    arr = img[..., 0] if c == 1 else (np.stack([img[..., 0], img[..., 1]], -1) if c == 2 else img[..., :3] if c == 3 else
                                      np.concatenate([img[..., :3], (255 - img[..., :1])], -1))
    data = _png(np.ascontiguousarray(arr), mode)
    rc, got, path = D.decode(data)
    assert rc == 0 and path == 2
    want = np.array(PIL.open(io.BytesIO(data)))
    assert (got.reshape(want.shape) == want).all()
    # to_rgba8 / to_rgb8 (crates/cli/src/main.rs:602-613)
    for oc, pm in ((4, "RGBA"), (3, "RGB")):
        rc, got, _ = D.decode(data, oc)
        assert rc == 0
        assert (got == np.array(PIL.open(io.BytesIO(data)).convert(pm))).all()


def test_decode_palette_images_with_and_without_transparency():
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 17, (40, 61), dtype=np.uint8)
    im = PIL.fromarray(idx, "P")
    pal = rng.integers(0, 256, 17 * 3, dtype=np.uint8).tolist()
    im.putpalette(pal)
    bio = io.BytesIO()
    im.save(bio, "PNG")
    rc, got, _ = D.decode(bio.getvalue())
    assert rc == 0 and got.shape == (40, 61, 3)
    assert (got == np.array(im.convert("RGB"))).all()
    bio = io.BytesIO()
    im.save(bio, "PNG", transparency=bytes(rng.integers(0, 256, 17, dtype=np.uint8).tolist()))
    data = bio.getvalue()
    rc, got, _ = D.decode(data)
    assert rc == 0 and got.shape == (40, 61, 4)
    assert (got == np.array(PIL.open(io.BytesIO(data)).convert("RGBA"))).all()


def test_every_filter_type_is_reconstructed():
    """PIL picks filters adaptively; a hand-made file forces each of the five on a noisy image."""
    img = _photo(37, 53, 3, seed=9)
    h, w, c = img.shape
    rb = w * c
    for ft in range(5):
        stream = bytearray()
        prev = np.zeros(rb, np.int32)
        for y in range(h):
            row = img[y].reshape(-1).astype(np.int32)
            left = np.concatenate([np.zeros(c, np.int32), row[:-c]])
            upleft = np.concatenate([np.zeros(c, np.int32), prev[:-c]])
            if ft == 0: pred = 0 * row
            elif ft == 1: pred = left
            elif ft == 2: pred = prev
            elif ft == 3: pred = (left + prev) // 2
            else:
                p = left + prev - upleft
                pa, pb, pc = abs(p - left), abs(p - prev), abs(p - upleft)
                pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            stream.append(ft)
            stream += ((row - pred) & 255).astype(np.uint8).tobytes()
            prev = row
        import struct
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
        data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(stream), 6)) + chunk(b"IEND", b"")
        rc, got, _ = D.decode(data)
        assert rc == 0 and (got == img).all(), ft


def test_files_of_this_librarys_writer_take_the_piece_path():
    """png_core.h writes one IDAT per 16 KB piece: the decoder recognises the pattern and inflates the pieces
    independently (what the GPU does with one thread per piece); a foreign file of the same image takes the serial path."""
    for shape in ((130, 257, 3), (64, 64, 4), (300, 301, 3), (5, 7, 3)):
        img = _photo(*shape, seed=shape[0])
        data, _, _ = png_sim.encode(img)
        rc, got, path = D.decode(data)
        assert rc == 0 and path == 1 and (got == img).all(), shape
        rc, got, path = D.decode(_png(img, "RGB" if shape[2] == 3 else "RGBA"))
        assert rc == 0 and path == 2 and (got == img).all()
    noise = np.random.default_rng(3).integers(0, 256, (90, 200, 3), dtype=np.uint8)   # pieces stored uncompressed
    data, _, stored = png_sim.encode(noise)
    rc, got, path = D.decode(data)
    assert stored > 0 and rc == 0 and path == 1 and (got == noise).all()


def test_container_errors():
    img = _photo(20, 30, 3)
    data = bytearray(_png(img, "RGB"))
    assert D.decode(b"not a png at all")[0] == 1
    bad = bytearray(data); bad[40] ^= 0x55
    assert D.decode(bytes(bad))[0] != 0                                   # a chunk CRC or the stream breaks
    assert D.decode(bytes(data[:-20]))[0] == 2                            # no IEND
    assert D.decode(_png(img.astype(np.uint16)[..., 0] * 257, "I;16"))[0] == 3      # 16-bit samples: unsupported, not wrong


@pytest.mark.skipif(not os.path.exists("/root/reference/docs/assets/gruvbox-dark-hald-clut.png"), reason="reference assets not on this machine")
def test_reference_hald_clut_assets_decode_like_pil():
    """The files `lutgen apply --hald-clut` reads: the reference's own committed CLUT images."""
    for name in ("gruvbox-dark-hald-clut.png", "catppuccin-mocha-hald-clut.png"):
        data = open(os.path.join("/root/reference/docs/assets", name), "rb").read()
        rc, got, _ = D.decode(data, 3)
        assert rc == 0
        assert (got == np.array(PIL.open(io.BytesIO(data)).convert("RGB"))).all()
