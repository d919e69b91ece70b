"""fast_srgb8::f32_to_srgb8 as the kernels evaluate it (csrc/color.cuh: f32_to_srgb8_wide):
the per-bucket constant (bias << 9) - scale * (i << 8) folded into the table so that no mask or shift is needed.
CPU test of the identity over the clamped input range; the CUDA code itself is covered by the GPU parity tests."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINV, MAXV = 0x39000000, 0x3F7FFFFF


def table():
    src = open(os.path.join(ROOT, "lutgen_rs_b200", "csrc", "color.cuh")).read()
    body = re.search(r"k_srgb_encode_table\[104\] = \{(.*?)\};", src, re.S).group(1)
    t = np.array([int(x, 16) for x in re.findall(r"0x[0-9a-f]+", body)], dtype=np.uint64)
    assert len(t) == 104
    return t


def literal(fu, t):
    e = t[(fu - MINV) >> 20]
    return (((e >> 16) << 9) + (e & 0xFFFF) * ((fu >> 12) & 0xFF)) >> 16


def wide(fu, t):
    idx = np.arange(104, dtype=np.uint64)
    scale = t & 0xFFFF
    bias2 = (((t >> 16) << 9) - scale * ((idx + 0x390) << 8)) % (1 << 32)   # srgb_encode2_entry
    i = (fu >> 20) - 0x390
    x = (bias2[i] + scale[i] * (fu >> 12)) % (1 << 32)
    assert ((x >> 24) == 0).all(), "bits above the result byte must be zero (srgb8_pack3 relies on it)"
    return x >> 16


def test_wide_encoder_equals_the_literal_formula():
    t = table()
    # every 5th float of the range plus every bucket edge with its neighbours
    fu = np.arange(MINV, MAXV + 1, 5, dtype=np.uint64)
    edges = (np.arange(104, dtype=np.uint64) << 20) + MINV
    near = np.concatenate([edges + d for d in (0, 1, 2, 4095, 4096, 4097)] + [edges[1:] - d for d in (1, 2, 4096)])
    fu = np.concatenate([fu, near, np.array([MINV, MAXV], dtype=np.uint64)])
    fu = fu[(fu >= MINV) & (fu <= MAXV)]
    assert (literal(fu, t) == wide(fu, t)).all()
    out = literal(np.array([MINV, MAXV], dtype=np.uint64), t)
    assert out[0] == 0 and out[1] == 255
