"""Palette extraction (lutgen_rs_b200/csrc/extract_core.h; SURVEY 8(f) row 4, `lutgen extract`,
crates/cli/src/main.rs:772-815) on the host: the kernel bodies run in loops (tests/extract_sim.cpp, scrambled
order) against the oracle's independent restatement of the same specification -- bit exact -- and against
scikit-learn's k-means for palette quality. PARITY WITH THE REFERENCE IS UNPINNED: its `quantette` dependency is
not vendored and nothing in the reference pins its output; see extract_core.h."""
import numpy as np
import pytest

import extract_sim as S
import oracle_lib as O


def _photo(h, w, seed, noise=6.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 19.0 + yy / 31.0), 128 + 90 * np.cos(xx / 13.0), 128 + 80 * np.sin(yy / 11.0)], -1)
    return np.clip(base + rng.normal(0, noise, (h, w, 3)), 0, 255).astype(np.uint8)


def _blobs(n, centers, seed, spread=10):
    rng = np.random.default_rng(seed)
    c = rng.integers(20, 236, (centers, 3))
    idx = rng.integers(0, centers, n)
    return np.clip(c[idx] + rng.normal(0, spread, (n, 3)), 0, 255).astype(np.uint8).reshape(-1, 1, 3)


@pytest.mark.parametrize("k", [1, 2, 8, 31, 64])
@pytest.mark.parametrize("iters", [0, 1, 16])
def test_kernel_bodies_match_the_oracle_bit_for_bit(k, iters):
    img = _photo(48, 64, seed=k)
    want = O.extract_palette(img, k, iters)
    for seed in (0, 7):  # two different visiting orders
        got = S.extract_palette(img, k, iters, seed=seed)
        assert got.shape == want.shape and (got == want).all(), (k, iters, seed)


def test_few_distinct_colours_and_degenerate_inputs():
    img = np.zeros((10, 10, 3), np.uint8)
    img[:5] = (200, 30, 30)
    img[5:, :3] = (10, 10, 250)
    for k in (1, 2, 3, 4, 200):
        want = O.extract_palette(img, k, 16)
        got = S.extract_palette(img, k, 16)
        assert (got == want).all() and len(got) == min(k, 3)
    # with as many centroids as colours every colour is its own cluster and comes back exactly
    got = S.extract_palette(img, 3, 16)
    assert sorted(map(tuple, got.tolist())) == sorted([(0, 0, 0), (200, 30, 30), (10, 10, 250)])
    one = np.full((4, 4, 3), 77, np.uint8)
    assert S.extract_palette(one, 8, 3).tolist() == [[77, 77, 77]] == O.extract_palette(one, 8, 3).tolist()
    # the most frequent colour seeds the first centroid; ties go to the smallest key (r | g << 8 | b << 16)
    tie = np.array([[[9, 0, 0], [3, 0, 0], [3, 1, 0], [9, 0, 0], [3, 0, 0]]], np.uint8)
    assert S.extract_palette(tie, 1, 0).tolist() == [[3, 0, 0]] == O.extract_palette(tie, 1, 0).tolist()


def test_heavy_weights_and_many_colours():
    """Weights above the seeding cap (65535 pixels of one colour) and a few thousand distinct colours."""
    rng = np.random.default_rng(3)
    img = np.concatenate([np.tile(np.array([[12, 200, 40]], np.uint8), (70000, 1)), rng.integers(0, 256, (5000, 3), dtype=np.uint8)]).reshape(-1, 1, 3)
    for k in (5, 40):
        want = O.extract_palette(img, k, 8)
        assert (S.extract_palette(img, k, 8, seed=1) == want).all()
        assert [12, 200, 40] in want.tolist() or k < 40  # the dominant colour keeps a centroid of its own among 40


def _cost(img, pal):
    """Pixel-weighted mean squared Oklab distance to the nearest palette colour."""
    px = O.srgb_to_oklab(img.reshape(-1, 3)).astype(np.float64)
    pl = O.srgb_to_oklab(pal).astype(np.float64)
    d = ((px[:, None, :] - pl[None, :, :]) ** 2).sum(-1)
    return d.min(1).mean()


@pytest.mark.parametrize("maker,k", [(lambda: _photo(60, 80, 1), 12), (lambda: _blobs(6000, 7, 2), 7), (lambda: _photo(40, 40, 5, noise=25.0), 24)])
def test_palette_quality_is_that_of_kmeans(maker, k):
    """Not quantette's palette, but a k-means palette: its cost is within 15 % of scikit-learn's k-means++ (best of
    4 starts) on the same Oklab points, and each Lloyd pass can only lower it."""
    from sklearn.cluster import KMeans
    img = maker()
    pal = O.extract_palette(img, k, 32)
    assert len(pal) == k
    px = O.srgb_to_oklab(img.reshape(-1, 3)).astype(np.float64)
    km = KMeans(n_clusters=k, n_init=4, random_state=0).fit(px)
    ref_cost = km.inertia_ / len(px)
    assert _cost(img, pal) <= 1.15 * ref_cost + 1e-6, (_cost(img, pal), ref_cost)
    costs = [_cost(img, O.extract_palette(img, k, it)) for it in (0, 1, 2, 4, 32)]
    # sRGB8 rounding of the centroids adds a little noise on top of a monotone sequence
    assert all(b <= a * 1.02 + 1e-7 for a, b in zip(costs, costs[1:])), costs


def test_python_mirror_checks_its_arguments_without_a_gpu():
    from lutgen_rs_b200 import extract
    with pytest.raises(TypeError):
        extract.extract_palette(np.zeros((4, 4, 4), np.uint8), 8)
    with pytest.raises(TypeError):
        extract.extract_palette(np.zeros((4, 4, 3), np.float32), 8)
    with pytest.raises(OverflowError):
        extract.extract_palette(np.zeros((4, 4, 3), np.uint8), -1)


def test_tie_heavy_random_inputs():
    """Seeded fuzz on tiny images whose colours repeat, collide on coarse grids or sit a step apart: the tie rules
    (smallest key in the seeding, lowest index in the assignment) are what decides here."""
    rng = np.random.default_rng(123)
    for trial in range(150):
        ncol = int(rng.integers(1, 12))
        base = rng.integers(0, 256, (ncol, 3))
        if trial % 3 == 0:
            base = (base // 64) * 64
        if trial % 3 == 1:
            base = rng.integers(100, 104, (ncol, 3))
        idx = rng.integers(0, ncol, int(rng.integers(1, 60)))
        img = base[idx].astype(np.uint8).reshape(1, -1, 3)
        k, it = int(rng.integers(1, 10)), int(rng.integers(0, 6))
        want, got = O.extract_palette(img, k, it), S.extract_palette(img, k, it, seed=trial)
        assert want.shape == got.shape and (want == got).all(), (trial, k, it)


def _golden_images():
    rng = np.random.default_rng(2026)
    return {
        "random_40x50": rng.integers(0, 256, (40, 50, 3), dtype=np.uint8),
        "three_tones": np.concatenate([np.tile(np.array([[200, 30, 30]], np.uint8), (500, 1)), np.tile(np.array([[10, 10, 250]], np.uint8), (30, 1)),
                                       rng.integers(90, 110, (200, 3)).astype(np.uint8)]).reshape(-1, 1, 3),
    }


def test_palettes_match_the_committed_ones():
    """tests/golden/extract_palettes.json: palettes of two integer-generated images, written by the oracle when the
    specification was fixed. They pin the SPECIFICATION (seeding, tie rules, rounding) from round to round; they say
    nothing about the reference's quantette palettes (parity unpinned)."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extract_palettes.json")) as f:
        want = json.load(f)
    for name, img in _golden_images().items():
        w = want[name]
        assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == w["pixels_sha256"], "numpy's generator changed?"
        for k, it, key in ((8, 16, "k8_iters16"), (3, 0, "k3_iters0")):
            assert O.extract_palette(img, k, it).tolist() == w[key], (name, key)
            assert S.extract_palette(img, k, it, seed=5).tolist() == w[key], (name, key)
