"""The C++ host mirror (include/lutgen.hpp): compiles and links everywhere, runs on the GPU box."""
import os
import shutil
import subprocess

import pytest

import oracle_lib as O
from conftest import ROOT
from lutgen_rs_b200 import _native

BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "host_mirror")
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")


def _compile():
    _native.load()
    O.build()
    os.makedirs(BUILD, exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "include", "lutgen.hpp"), os.path.join(ROOT, "include", "lutgen_cuda.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    libdir, odir = os.path.join(ROOT, "lutgen_rs_b200"), os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
                           "-L" + libdir, "-llutgen_cuda", "-L" + odir, "-llutgen_oracle",
                           "-Wl,-rpath,$ORIGIN/../../../lutgen_rs_b200", "-Wl,-rpath,$ORIGIN/../../../oracle/_build"])
    return EXE


def test_cpp_host_mirror_compiles_and_links():
    exe = _compile()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_host_mirror_runs():
    exe = _compile()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror OK" in out.stdout
