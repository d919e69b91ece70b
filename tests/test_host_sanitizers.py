"""The PNG and palette-extraction cores (the code the GPU runs, instantiated on the host) under AddressSanitizer and
UndefinedBehaviorSanitizer (bounds-strict): compute-sanitizer is closed on the GPU pool, so out-of-range indices
into the shared-state arrays, bad shifts and signed overflow are looked for here."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_cores_are_clean_under_asan_and_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "sanitize_cores")
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fsanitize=bounds-strict", "-fno-sanitize-recover=all", "-o", exe,
           os.path.join(HERE, "cpp", "sanitize_cores.cpp"), os.path.join(HERE, "png_sim.cpp"), os.path.join(HERE, "extract_sim.cpp")]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and ("asan" in build.stderr or "ubsan" in build.stderr or "sanitize" in build.stderr):
        pytest.skip("this toolchain has no sanitizer runtimes")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and run.stdout.startswith("ok"), (run.stdout[-500:], run.stderr[-3000:])
