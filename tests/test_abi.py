"""CPU-only: the C-ABI library builds, loads, exports every symbol include/lutgen_cuda.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from lutgen_rs_b200 import _native


def _declared():
    text = open(os.path.join(ROOT, "include", "lutgen_cuda.h")).read()
    return sorted(set(re.findall(r"LUTGEN_API\s+[\w\s\*]+?\b(lutgen_cuda_\w+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _declared() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _native.load()
    for name in _declared():
        assert getattr(L, name) is not None
    assert b"sm_100a" in L.lutgen_cuda_version()


def test_detect_level_is_host_side():
    L = _native.load()
    lvl = C.c_uint8(0)
    for level in range(2, 17):
        assert L.lutgen_cuda_detect_level(level ** 3, level ** 3, C.byref(lvl)) == 0
        assert lvl.value == level
    assert L.lutgen_cuda_detect_level(100, 100, C.byref(lvl)) == _native.ERR_INVALID
    assert L.lutgen_cuda_detect_level(64, 63, C.byref(lvl)) == _native.ERR_INVALID
    assert b"identity.rs" in L.lutgen_cuda_last_error()


def _has_gpu():
    n = C.c_int(0)
    return _native.load().lutgen_cuda_device_count(C.byref(n)) == 0 and n.value > 0


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback_without_gpu():
    L = _native.load()
    assert L.lutgen_cuda_init(0) == _native.ERR_NO_DEVICE
    out = np.zeros(3 * 64, np.uint8)
    assert L.lutgen_cuda_identity(2, C.c_void_p(out.ctypes.data)) == _native.ERR_NO_DEVICE
    assert (out == 0).all()
    with pytest.raises(_native.LutgenError):
        from lutgen_rs_b200 import identity
        identity.generate(2)


def test_documented_environment_switches_exist_in_the_sources():
    """DESIGN.md's appendix lists the environment switches; each must be read somewhere in the sources it names
    (a guard against the table drifting from the code)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "DESIGN.md")).read()
    appendix = text[text.index("## Appendix: environment switches"):]
    names = set(re.findall(r"`(LUTGEN_[A-Z_]+)", appendix))
    assert len(names) >= 15
    blob = ""
    for d in ("lutgen_rs_b200", os.path.join("lutgen_rs_b200", "csrc")):
        for f in os.listdir(os.path.join(root, d)):
            if f.endswith((".cu", ".cuh", ".h", ".py")):
                blob += open(os.path.join(root, d, f), errors="ignore").read()
    missing = [n for n in names if f'"{n}"' not in blob]
    assert not missing, missing
