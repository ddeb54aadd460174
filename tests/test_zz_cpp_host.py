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


def _build_c(tmp_path):
    import cellranger_b200._lib as _lib

    _lib.load()
    exe = str(tmp_path / "host_c_sharded")
    libdir = os.path.join(ROOT, "cellranger_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-pedantic", "-O1", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "host_c_sharded.c"), "-L" + libdir, "-lcrgpu", "-Wl,-rpath," + libdir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_sharded_example_compiles_and_fails_loudly_without_a_device(tmp_path):
    """The sharded run driven from plain C99 through include/crgpu.h (no Python, no torch in the step): compiles and
    links here; without a GPU it reports crgpu_ctx_create's error."""
    import torch

    exe = _build_c(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test below")
    res = subprocess.run([exe, "2", "20000"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "crgpu_ctx_create failed" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_c_sharded_example_matches_single_device(tmp_path, n_dev):
    """crgpu_group_run over n_dev devices (one host thread per device inside the library, in-library NCCL, peer
    stores) gives the matrix of one device, bit for bit. n_dev = 1 runs the whole sharded step (communicator of
    one rank, owner ranges, scatter into the own buffer, on-stream barrier) on any box."""
    import torch

    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    exe = _build_c(tmp_path)
    res = subprocess.run([exe, str(n_dev), "400000"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().splitlines()[-1] == "OK", res.stdout
