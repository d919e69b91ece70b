"""PNG encoder core (lutgen_rs_b200/csrc/png_core.h; SURVEY 8(f) row 2: the `lut.save` / `image.save`
calls of crates/cli/src/main.rs:767-871) run on the host through tests/png_sim.cpp -- the same per-thread
phases the CUDA kernels run, with a loop in place of the CTA. The domain's parity property is the
lossless round trip: what an independent decoder (a from-the-specification reader over zlib in
tests/png_sim.py, and PIL) reads back must be the input, bit for bit, with every chunk CRC and the
Adler-32 of the stream valid. The GPU tests (tests/test_gpu_png.py) repeat this through the C ABI."""
import ctypes as C
import heapq
import io

import numpy as np
import pytest

import png_sim as P

CHUNK = 16384


def _pil(data):
    from PIL import Image
    a = np.array(Image.open(io.BytesIO(data)))
    return a[:, :, None] if a.ndim == 2 else a


def _roundtrip(img, own=True):
    data, filters, stored = P.encode(img)
    assert (_pil(data) == img).all()
    if own:
        dec, f2, nidat = P.decode(data)
        assert dec.shape == img.shape and (dec == img).all()
        assert (f2 == filters).all()
        stream = img.shape[0] * (img.shape[1] * img.shape[2] + 1)
        assert nidat == (stream + CHUNK - 1) // CHUNK + 1  # one IDAT per piece + the Adler trailer
    assert len(data) <= P.lib().lpng_sim_bound(img.shape[1], img.shape[0], img.shape[2])
    return data, filters, stored


@pytest.mark.parametrize("c", [1, 2, 3, 4])
@pytest.mark.parametrize("h,w", [(1, 1), (1, 300), (300, 1), (2, 2), (37, 53), (128, 128)])
def test_noise_and_flat_images_round_trip(h, w, c):
    rng = np.random.default_rng(h * 1000 + w * 10 + c)
    _roundtrip(rng.integers(0, 256, (h, w, c), dtype=np.uint8))
    _roundtrip(np.full((h, w, c), 200, np.uint8))
    _roundtrip(rng.integers(0, 4, (h, w, c), dtype=np.uint8))  # few symbols: short codes, long accidental runs


def test_piece_boundaries():
    """Stream lengths around a multiple of the piece size; rows longer than a piece; a last piece of one byte."""
    rng = np.random.default_rng(5)
    # stream = h * (3 w + 1): w = 5461 -> 16384 exactly for one row
    for h, w in ((1, 5461), (2, 5461), (1, 5460), (1, 5462), (3, 5461), (1, 12000), (2, 10923)):
        img = (rng.integers(0, 3, (h, w, 3)) * 40).astype(np.uint8)
        _roundtrip(img, own=False)
        _roundtrip(np.zeros((h, w, 3), np.uint8), own=False)
    # 4 bytes per pixel: stream 16385 = piece + 1 byte
    _roundtrip(np.full((1, 4096, 4), 9, np.uint8), own=False)


@pytest.mark.parametrize("run", [3, 4, 5, 63, 64, 65, 257, 258, 259, 260, 261, 262, 515, 516, 517, 518, 519, 520, 1000, 16000, 40000])
@pytest.mark.parametrize("c", [1, 3, 4])
def test_run_lengths_around_the_match_limits(run, c):
    """A run of `run` identical filtered bytes between noise: tails of 0, 1 and 2 bytes after the 258-byte
    matches, runs crossing thread segments (64 bytes) and pieces (16384 bytes)."""
    rng = np.random.default_rng(run * 7 + c)
    npx = (run + c - 1) // c + 40
    row = rng.integers(0, 256, (1, npx, c), dtype=np.uint8)
    flat = row.reshape(-1)
    flat[17:17 + run] = 77  # with filter None this is a run of run - c matches; other filters shift it by a pixel
    data, filters, _ = _roundtrip(row, own=npx * c < 50000)
    # and the same content with the filter decision taken away (two identical rows -> Up gives all zeros)
    _roundtrip(np.concatenate([row, row, row]), own=False)


def test_gradients_and_clut_like_content_compress():
    g = (np.arange(512)[None, :, None] + np.arange(512)[:, None, None] + np.zeros((1, 1, 3))).astype(np.uint8)
    data, filters, stored = _roundtrip(g, own=False)
    assert stored == 0 and len(data) < g.size // 20
    import oracle_lib as O
    from conftest import synthetic_palette
    pal = synthetic_palette(26, seed=3)
    lut = O.Remapper(O.GAUSSIAN, pal, param=128.0, nearest=16).generate_lut(6)
    data, filters, stored = _roundtrip(lut)
    assert stored == 0 and len(data) < 0.5 * lut.size
    _roundtrip(O.identity(4))


def test_noise_falls_back_to_stored_blocks_within_the_bound():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    data, filters, stored = _roundtrip(img)
    assert stored > 0
    assert len(data) <= img.size + img.shape[0] + 17 * ((img.size + img.shape[0]) // CHUNK + 1) + 63


def _code(hist):
    hist = np.ascontiguousarray(hist, np.uint32)
    ln = np.zeros(286, np.uint8)
    code = np.zeros(286, np.uint16)
    P.lib().lpng_sim_code(hist.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p), code.ctypes.data_as(C.c_void_p))
    return ln, code


def _huffman_cost(freqs):
    h = [int(f) for f in freqs if f]
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, a + b)
    return cost


def _check_code(hist):
    ln, code = _code(hist)
    used = hist > 0
    assert (ln[used] > 0).all() and (ln[~used] == 0).all()
    assert ln.max() <= 15
    assert sum(2.0 ** -int(l) for l in ln[used]) == 1.0  # complete (zlib rejects anything else)
    # prefix free and canonical: codes of one length ascend with the symbol, all codes distinct
    seen = set()
    for s in np.nonzero(used)[0]:
        msb_first = int(format(int(code[s]), f"0{int(ln[s])}b")[::-1], 2)
        key = (int(ln[s]), msb_first)
        assert key not in seen
        seen.add(key)
    for l in range(1, 16):
        syms = [s for s in np.nonzero(used)[0] if ln[s] == l]
        vals = [int(format(int(code[s]), f"0{l}b")[::-1], 2) for s in syms]
        assert vals == sorted(vals)
    # more frequent symbols never get longer codes
    order = np.argsort(hist[used], kind="stable")
    ls = ln[used][order]
    assert (np.diff(ls.astype(int)) <= 0).all()
    return ln


def test_code_lengths_are_optimal_when_fifteen_bits_suffice():
    rng = np.random.default_rng(11)
    for trial in range(40):
        hist = np.zeros(286, np.uint32)
        n = int(rng.integers(2, 287))
        idx = rng.choice(286, n, replace=False)
        hist[idx] = rng.integers(1, 200, n)
        hist[256] = 1
        ln = _check_code(hist)
        cost = int((ln.astype(np.int64) * hist).sum())
        assert cost == _huffman_cost(hist), trial


def test_code_lengths_are_limited_to_fifteen_bits():
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    hist = np.zeros(286, np.uint32)
    hist[:24] = fib            # Huffman depth 23 without the limit
    hist[256] = 1
    ln = _check_code(hist)
    cost = int((ln.astype(np.int64) * hist).sum())
    assert cost < 1.02 * _huffman_cost(hist)
    hist = np.ones(286, np.uint32)   # 270 rare symbols under a Fibonacci head: the limit binds with many codes at 15 bits
    hist[:16] = [3 * f for f in fib[8:24]]
    ln = _check_code(hist)
    assert ln.max() == 15
    assert int((ln.astype(np.int64) * hist).sum()) < 1.02 * _huffman_cost(hist)
    hist = np.zeros(286, np.uint32)
    hist[0] = 16000            # two symbols
    hist[256] = 1
    ln = _check_code(hist)
    assert ln[0] == 1 and ln[256] == 1


def test_deep_code_reaches_the_stream():
    """A one-row grey image whose residual frequencies are Fibonacci numbers drives the length limiter
    through the whole pipeline (the decoder must accept the repaired code)."""
    fib = [1, 1]
    while len(fib) < 19:
        fib.append(fib[-1] + fib[-2])
    vals = np.concatenate([np.full(f, 3 * i + 1, np.uint8) for i, f in enumerate(fib)])
    rng = np.random.default_rng(4)
    rng.shuffle(vals)
    _roundtrip(vals.reshape(1, -1, 1))


def test_shared_state_size():
    assert P.lib().lpng_sim_shared_bytes() <= 45670  # four CTAs per SM today; 45,568 would allow five


def test_random_images_round_trip():
    """Seeded fuzz over geometry and content kinds (noise, few-valued, smooth, repeated rows, constant blocks)."""
    rng = np.random.default_rng(20261019)
    for trial in range(40):
        c = int(rng.integers(1, 5))
        h = int(rng.integers(1, 90))
        w = int(rng.integers(1, 700))
        kind = trial % 5
        if kind == 0:
            img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        elif kind == 1:
            img = (rng.integers(0, 3, (h, w, c)) * rng.integers(1, 80)).astype(np.uint8)
        elif kind == 2:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([(xx * (k + 1) + yy * (3 - k)) // (k + 2) for k in range(c)], -1).astype(np.uint8)
        elif kind == 3:
            row = rng.integers(0, 256, (1, w, c), dtype=np.uint8)
            img = np.repeat(row, h, axis=0)
            img[rng.integers(0, h)] ^= 1
        else:
            img = np.zeros((h, w, c), np.uint8)
            for _ in range(6):
                y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
                img[y0:y0 + int(rng.integers(1, 40)), x0:x0 + int(rng.integers(1, 300))] = rng.integers(0, 256, c)
        data, filters, stored = P.encode(img)
        assert (_pil(data) == img).all(), (trial, img.shape, kind)
        if img.size < 40000:
            dec, f2, _ = P.decode(data)
            assert (dec == img).all() and (f2 == filters).all(), (trial, img.shape, kind)


def test_python_mirror_checks_its_arguments_without_a_gpu():
    """lutgen_rs_b200.png: what does not need the device (bounds, argument errors)."""
    from lutgen_rs_b200 import png
    assert png.bound(4096, 4096, 3) == P.lib().lpng_sim_bound(4096, 4096, 3)
    assert png.bound(1, 1, 1) >= 8 + 25 + 12 + 2 + 5 + 2 + 16 + 12
    with pytest.raises(TypeError):
        png.encode(np.zeros((4, 4, 5), np.uint8))
    with pytest.raises(TypeError):
        png.encode(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(TypeError):
        png.encode(np.zeros((2, 2, 2, 3), np.uint8))
    import lutgen_rs_b200 as lg
    with pytest.raises(lg.LutgenPanic):
        png.bound(0, 4, 3)
    with pytest.raises(lg.LutgenPanic):
        png.bound(4, 4, 7)


def test_written_files_match_the_committed_digests():
    """tests/golden/png_files.json (made by tests/golden/make_png_golden.py) pins the bytes the encoder writes for
    four small images, and png_identity_level3.png is one of those files, readable by anything."""
    import hashlib
    import json
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import make_png_golden
    with open(os.path.join(golden, "png_files.json")) as f:
        want = json.load(f)
    for name, img in make_png_golden.images().items():
        w = want[name]
        if hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() != w["pixels_sha256"]:
            pytest.skip(f"{name}: this machine's oracle produces other pixels than the one that wrote the digests")
        data = P.encode(img)[0]
        assert len(data) == w["png_bytes"] and hashlib.sha256(data).hexdigest() == w["png_sha256"], name
    with open(os.path.join(golden, "png_identity_level3.png"), "rb") as f:
        data = f.read()
    import oracle_lib as O
    assert (_pil(data) == O.identity(3)).all()
    assert data == P.encode(O.identity(3))[0]


def test_no_phase_depends_on_the_order_of_its_threads():
    """The host executor runs the 256 threads of a phase one after the other; on the GPU they run in any order. A
    phase in which a thread reads what another one writes (a missing barrier), or two threads plainly write one place,
    gives an order-dependent result: demand the same file for ascending, descending and per-phase scrambled orders."""
    rng = np.random.default_rng(31)
    g = (np.arange(300)[None, :, None] * 2 + np.arange(70)[:, None, None] + np.zeros((1, 1, 3))).astype(np.uint8)
    imgs = [rng.integers(0, 256, (30, 200, 3), dtype=np.uint8), (rng.integers(0, 3, (60, 500, 4)) * 60).astype(np.uint8), g,
            np.full((20, 9000, 1), 3, np.uint8), np.repeat(rng.integers(0, 256, (1, 700, 3), dtype=np.uint8), 40, axis=0)]
    try:
        for img in imgs:
            files = []
            for mode in (0, 1, 2):
                P.set_order(mode)
                files.append(P.encode(img)[0])
            assert files[0] == files[1] == files[2]
            assert (_pil(files[2]) == img).all()
    finally:
        P.set_order(0)


def test_more_pieces_than_layout_threads():
    """More than 256 pieces: every thread of the layout pass (offsets, Adler-32 fold) then owns several of them."""
    rng = np.random.default_rng(8)
    img = (rng.integers(0, 4, (1300, 1200, 3)) * 30).astype(np.uint8)
    img[200:900, 100:1100] = 17
    data, _, stored = P.encode(img)
    assert (img.shape[0] * (img.shape[1] * 3 + 1) + CHUNK - 1) // CHUNK > 256
    assert (_pil(data) == img).all()
    # chunk CRCs, chunk count and the Adler-32 through the specification reader's front half
    import struct
    import zlib
    pos, idat = 8, []
    while pos < len(data):
        ln, typ = struct.unpack(">I4s", data[pos:pos + 8])
        assert zlib.crc32(data[pos + 4:pos + 8 + ln]) == struct.unpack(">I", data[pos + 8 + ln:pos + 12 + ln])[0]
        if typ == b"IDAT":
            idat.append(data[pos + 8:pos + 8 + ln])
        pos += 12 + ln
    assert len(idat) == (img.shape[0] * (img.shape[1] * 3 + 1) + CHUNK - 1) // CHUNK + 1
    assert len(zlib.decompress(b"".join(idat))) == img.shape[0] * (img.shape[1] * 3 + 1)
