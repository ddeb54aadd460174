"""The compiled (C++) host layer over the C ABI: include/crgpu.hpp + examples/host_cpp_hand_case.cpp.
Named so that it runs after the other GPU tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import cellranger_b200._lib as _lib

    _lib.load()  # makes sure libcrgpu.so exists (builds it if stale)
    exe = str(tmp_path / "host_cpp_hand_case")
    libdir = os.path.join(ROOT, "cellranger_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "host_cpp_hand_case.cpp"), "-L" + libdir, "-lcrgpu",
           "-Wl,-rpath," + libdir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cpp_host_layer_compiles_and_fails_loudly_without_a_device(tmp_path):
    """No GPU needed: the header-only wrapper and its example compile and link against libcrgpu.so; without a
    CUDA device the program reports the error of crgpu_ctx_create (there is no CPU fallback)."""
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test below")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "crgpu_ctx_create failed" in res.stdout and "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_cpp_host_layer_hand_case(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().splitlines()[-1] == "OK"
