#!/usr/bin/env python3
"""Off-by-one bytes of the GPU CLUTs against the reference's golden CLUT images (tests/golden)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lutgen_rs_b200 import interpolation
import oracle_lib as O
pals = {k: np.array(v, np.uint8) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "palettes.json"))).items()}
for name, pal, level, shape, nearest in (("gruvbox_dark_hald8", "gruvbox_dark", 8, 96.0, 0), ("catppuccin_mocha_hald10", "catppuccin_mocha", 10, 128.0, 16), ("nord_hald10", "nord", 10, 128.0, 16)):
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    gold = gold[list(gold.keys())[0]]
    got = interpolation.GaussianRemapper(pals[pal], shape, nearest, 1.0, False).par_generate_lut(level)
    orc = O.Remapper(O.GAUSSIAN, pals[pal], param=shape, nearest=nearest).generate_lut(level)
    d = np.abs(got.astype(np.int16) - gold.reshape(got.shape).astype(np.int16))
    do = np.abs(orc.astype(np.int16) - gold.reshape(got.shape).astype(np.int16))
    dg = np.abs(got.astype(np.int16) - orc.astype(np.int16))
    print(f"{name}: GPU vs golden max {d.max()} off-by-one {int((d>0).sum())}; oracle vs golden {int((do>0).sum())}; GPU vs oracle max {dg.max()} differing {int((dg>0).sum())} of {d.size}")
