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


def _compile(src=SRC, exe=EXE):
    _native.load()
    O.build()
    os.makedirs(BUILD, exist_ok=True)
    deps = [src, os.path.join(ROOT, "include", "lutgen.hpp"), os.path.join(ROOT, "include", "lutgen_cuda.h")]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    libdir, odir = os.path.join(ROOT, "lutgen_rs_b200"), os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L" + libdir, "-llutgen_cuda", "-L" + odir, "-llutgen_oracle",
                           "-Wl,-rpath,$ORIGIN/../../../lutgen_rs_b200", "-Wl,-rpath,$ORIGIN/../../../oracle/_build"])
    return exe


MULTI_SRC = os.path.join(ROOT, "tests", "cpp", "host_multi.cpp")
MULTI_EXE = os.path.join(BUILD, "host_multi")


def test_cpp_host_mirror_compiles_and_links():
    exe = _compile()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_host_mirror_runs():
    exe = _compile()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror OK" in out.stdout


def test_cpp_multi_device_host_compiles_and_links():
    assert os.path.exists(_compile(MULTI_SRC, MULTI_EXE))


def _run_multi(devices):
    env = dict(os.environ)
    env.pop("LUTGEN_CUDA_DEVICES", None)
    if devices:
        env["LUTGEN_CUDA_DEVICES"] = ",".join(str(d) for d in devices)
    out = subprocess.run([_compile(MULTI_SRC, MULTI_EXE)], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    return [ln for ln in out.stdout.splitlines() if not ln.startswith("#")], out.stdout


@pytest.mark.gpu
def test_one_process_on_one_device_list_equals_the_default():
    """lutgen_cuda_init_devices with a single device is the plain single-GPU library."""
    plain, _ = _run_multi(None)
    listed, text = _run_multi([0])
    assert "devices in use: 1" in text
    assert plain == listed and len(plain) >= 10


@pytest.mark.gpu
def test_one_process_many_devices_is_byte_identical():
    """ONE process, several GPUs, no caller changes -- every CLUT and every corrected frame must be
    the bytes one GPU produces (needs at least two GPUs on the box)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    one, _ = _run_multi([0])
    many, text = _run_multi(list(range(min(n, 8))))
    assert f"devices in use: {min(n, 8)}" in text
    assert one == many, (one, many)
