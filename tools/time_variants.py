#!/usr/bin/env python3
"""Times the headline CLUT (level-16 Gaussian RBF, catppuccin-mocha, nearest 16) and a few neighbours with every
tools/_variant_*.so (tools/build_variants.py), one subprocess per library, L2 flushed between repetitions as
bench.py does; prints the median per variant and an Adler-32 of each CLUT (all variants must agree).

    python tools/time_variants.py [reps]            # on the GPU box
"""
import glob
import json
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _synthetic_palette(n, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import synthetic_palette
    return synthetic_palette(n, seed)


def child(reps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from lutgen_rs_b200 import device, interpolation
    device.init(0)
    with open(os.path.join(ROOT, "tests", "golden", "palettes.json")) as f:
        pal = np.array(json.load(f)["catppuccin_mocha"], np.uint8)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    cases = {"gauss16": (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), 16),
             "nn16": (interpolation.NearestNeighborRemapper(pal, 1.0, False), 16),
             "shep16": (interpolation.ShepardRemapper(pal, 4.0, 16, 1.0, False), 16),
             "gauss12": (interpolation.GaussianRemapper(pal, 128.0, 16, 1.0, False), 12),
             "shep256": (interpolation.ShepardRemapper(_synthetic_palette(256, 1), 4.0, 16, 1.0, False), 16)}
    for name, (r, level) in cases.items():
        side = level ** 3
        lut = torch.empty(3 * side * side, dtype=torch.uint8, device="cuda")
        ts = []
        for i in range(reps + 3):
            flush.fill_(i & 255)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            device.generate_lut_rows_(r, level, lut, 0, side)
            ev1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(ev0.elapsed_time(ev1))
        ts.sort()
        host = lut.cpu().numpy()
        out[name] = {"ms_median": round(ts[len(ts) // 2], 4), "ms_min": round(ts[0], 4), "adler": zlib.adler32(host.tobytes())}
        ref_path = os.path.join("/tmp", "lutgen_variants", f"_ref_{name}.npy")
        if os.path.exists(ref_path):
            d = np.abs(host.astype(np.int16) - np.load(ref_path).astype(np.int16))
            out[name]["vs_first"] = [int(d.max()), int((d > 0).sum())]
        else:
            os.makedirs(os.path.dirname(ref_path), exist_ok=True)
            np.save(ref_path, host)
    print("RESULT " + json.dumps(out))


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    libs = sorted(glob.glob(os.path.join(ROOT, "tools", "_variant_*.so")))
    ref = None
    for lib in libs:
        env = dict(os.environ, LUTGEN_CUDA_LIB=lib)
        p = subprocess.run([sys.executable, __file__, "--child", str(reps)], env=env, capture_output=True, text=True, timeout=300)
        name = os.path.basename(lib)[len("_variant_"):-3]
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        if not res:
            print(name, "FAILED", p.stderr[-500:])
            continue
        r = json.loads(res[0][7:])
        if ref is None:
            ref = r
        same = all(r[k]["adler"] == ref[k]["adler"] for k in r)
        print(f"{name:12s}", " ".join(f"{k} {v['ms_median']:.4f}/{v['ms_min']:.4f}" for k, v in r.items()), "same_bytes" if same else
              "BYTES DIFFER (max, count vs the first variant): " + " ".join(f"{k} {v.get('vs_first')}" for k, v in r.items()))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
