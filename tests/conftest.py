import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def palettes():
    with open(os.path.join(GOLDEN, "palettes.json")) as f:
        return {k: np.array(v, np.uint8) for k, v in json.load(f).items()}


def splitmix64_stream(seed, n):
    """n uint64 draws of splitmix64 (SURVEY.md section 8d's synthetic-input generator)."""
    out = np.empty(n, np.uint64)
    s = np.uint64(seed)
    inc = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        idx = (np.arange(1, n + 1, dtype=np.uint64) * inc) + s
        z = idx
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        out = z ^ (z >> np.uint64(31))
    return out


def synthetic_palette(n, seed=1):
    """n distinct 24-bit colours from splitmix64(seed), duplicates rejected."""
    draws = splitmix64_stream(seed, 4 * n + 64) & np.uint64(0xFFFFFF)
    seen, out = set(), []
    for d in draws.tolist():
        if d not in seen:
            seen.add(d)
            out.append([(d >> 16) & 255, (d >> 8) & 255, d & 255])
            if len(out) == n:
                break
    return np.array(out, np.uint8)


def random_rgba(npix, seed):
    """Uniform random RGBA8 pixels from the splitmix64 stream."""
    words = splitmix64_stream(seed, (npix + 1) // 2)
    return words.view(np.uint8)[: npix * 4].reshape(npix, 4).copy()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
