#!/usr/bin/env python3
"""CPU simulation of the fast kernel's tile classification (kernels_fast.cuh steps 2-4) for
different brick shapes: how many palette colours end up IN / UNCERTAIN per tile, how many have to be
chosen per entry. Design aid only (numpy; uses the oracle's Oklab conversion).

    python tests/tools/sim_tiles.py [palette] [k] [shape] [level]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from conftest import synthetic_palette  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "catppuccin_mocha"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    shape = float(sys.argv[3]) if len(sys.argv) > 3 else 128.0
    level = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    kind = sys.argv[5] if len(sys.argv) > 5 else "gauss"
    NEG = -float(sys.argv[6]) if len(sys.argv) > 6 else -24.0
    if name.startswith("syn"):
        pal = synthetic_palette(int(name[3:]), 1)
    else:
        with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
            pal = np.array(json.load(f)[name], np.uint8)
    n = len(pal)
    cs = level * level
    chan = (np.arange(cs) * 255 // (cs - 1)).astype(np.uint8)
    P = O.srgb_to_oklab(pal).astype(np.float64)
    # whole cube in [b][g][r] order
    rr, gg, bb = np.meshgrid(chan, chan, chan, indexing="ij")  # r,g,b index order [r][g][b]
    rgb = np.stack([rr, gg, bb], -1).reshape(-1, 3)
    lab = O.srgb_to_oklab(rgb).reshape(cs, cs, cs, 3)  # [r][g][b]
    c1 = -shape * np.log2(np.e)
    for (tr, tg, tb) in [(8, 8, 16), (8, 4, 4), (8, 4, 8), (4, 4, 4)]:
        if cs % tr or cs % tg or cs % tb:
            continue
        v = lab.reshape(cs // tr, tr, cs // tg, tg, cs // tb, tb, 3)
        lo = v.min(axis=(1, 3, 5)).reshape(-1, 3) - 2e-5
        hi = v.max(axis=(1, 3, 5)).reshape(-1, 3) + 2e-5
        T = lo.shape[0]
        below = lo[:, None, :] - P[None]
        above = P[None] - hi[:, None, :]
        ex = np.maximum(0, np.maximum(below, above))
        fx = np.maximum(np.abs(P[None] - lo[:, None, :]), np.abs(P[None] - hi[:, None, :]))
        LB = (ex ** 2).sum(-1)
        UB = (fx ** 2).sum(-1)
        ubmin = UB.min(1)
        kk = k if 0 < k < n else 0
        if kind == "gauss":
            negl = c1 * (LB - ubmin[:, None]) < NEG
        else:  # shepard power = shape
            with np.errstate(divide="ignore"):
                negl = (-shape / 2) * (np.log2(np.maximum(LB, 1e-300)) - np.log2(ubmin[:, None])) < NEG
        if kk:
            uk = np.sort(UB, 1)[:, kk - 1]
            lk = np.sort(LB, 1)[:, kk - 1]
            if kind == "gauss":
                moot = c1 * (lk - ubmin) < NEG
            else:
                moot = (-shape / 2) * (np.log2(np.maximum(lk, 1e-300)) - np.log2(ubmin)) < NEG
            cand = LB <= uk[:, None]
            # rivals(t) = #{q != t: LB_q <= UB_t}
            LBs = np.sort(LB, 1)
            riv = np.empty_like(LB, dtype=np.int64)
            for t in range(n):
                riv[:, t] = (LBs <= UB[:, t:t + 1]).sum(1) - 1
            is_in = cand & (riv <= kk - 1)
            is_unc = cand & ~is_in
            is_in = np.where(moot[:, None], True, is_in)
            is_unc = np.where(moot[:, None], False, is_unc)
        else:
            moot = np.ones(T, bool)
            is_in = np.ones_like(LB, bool)
            is_unc = np.zeros_like(LB, bool)
        nin_all = is_in.sum(1)
        nunc_all = is_unc.sum(1)
        nin = (is_in & ~negl).sum(1)
        nunc = (is_unc & ~negl).sum(1)
        kp = kk - nin_all
        choose = (~moot) & (kp > 0) & (nunc_all > kp) & (nunc > kp)
        direct = np.where(moot | (kp <= 0) | choose, nin, nin + nunc)
        sel = np.where(choose, nunc, 0)
        hist = np.bincount(sel, minlength=1)
        print(f"brick {tr}x{tg}x{tb} ({tr*tg*tb:5d} entries, {T} tiles): direct {direct.mean():5.2f}  choose-tiles {choose.mean()*100:5.1f}%  "
              f"|UNC| when choosing {sel[choose].mean() if choose.any() else 0:5.2f}  mean sel {sel.mean():5.2f}  moot {moot.mean()*100:4.1f}%  "
              f"sel>4 {np.mean(sel>4)*100:4.1f}% sel>8 {np.mean(sel>8)*100:4.1f}% sel>16 {np.mean(sel>16)*100:4.1f}% max {sel.max()}  cand(in+unc all) {(nin_all+nunc_all).mean():.1f}")


if __name__ == "__main__":
    main()
