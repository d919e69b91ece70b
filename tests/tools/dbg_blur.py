import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import oracle_lib as O
from lutgen_rs_b200 import interpolation
pal = np.array(json.load(open("/root/repo/tests/golden/palettes.json"))["catppuccin_mocha"], np.uint8)
for level, radius, lum, pres in [(4, 1.0, 1.0, False), (8, 8.0, 1.0, False), (10, 8.0, 1.0, False), (10, 0.4, 1.0, False), (7, 2.0, 1.0, True), (12, 3.5, 1.0, False)]:
    got = interpolation.GaussianBlurRemapper(pal, radius, lum, pres).par_generate_lut(level)
    want = O.Remapper(O.BLUR, pal, param=radius, lum_factor=lum, preserve=pres).generate_lut(level)
    size = level * level
    g = got.reshape(-1, 3); w = want.reshape(-1, 3)
    bad = np.nonzero((g != w).any(1))[0]
    print(level, radius, pres, "bad entries", len(bad), "of", len(g))
    if len(bad):
        r = bad % size; gg = (bad // size) % size; b = bad // (size * size)
        print("  r range", r.min(), r.max(), "g", gg.min(), gg.max(), "b", b.min(), b.max())
        print("  r hist", np.bincount(r, minlength=size).nonzero()[0][:20], " g", np.bincount(gg, minlength=size).nonzero()[0][:20], " b", np.bincount(b, minlength=size).nonzero()[0][:20])
        print("  first", [(int(r[i]), int(gg[i]), int(b[i]), g[bad[i]].tolist(), w[bad[i]].tolist()) for i in range(min(5, len(bad)))])
