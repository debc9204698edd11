"""A C++ program hosting the C ABI (tests/cpp/abi_host.cpp): compiled with g++ against include/ppg_b200.h, linked to
libppg_b200.so, run here without a GPU (defaults, vocabulary reader, refusal to fall back) and on the B200 with one
frame through ppg_extract + ppg_extend_map_matches."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ppg_slam_b200")


def _build(tmp_path):
    assert os.path.exists(os.path.join(PKG, "libppg_b200.so")), "build the library first (python __graft_entry__.py)"
    exe = str(tmp_path / "abi_host")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "abi_host.cpp"), "-o", exe, "-L", PKG, "-lppg_b200",
           "-Wl,-rpath," + PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = _build(tmp_path)
    r = subprocess.run([exe, os.path.join(PKG, "weights")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_host ok (cpu)" in r.stdout


@pytest.mark.gpu
def test_cpp_host_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, os.path.join(PKG, "weights"), "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_host ok (gpu)" in r.stdout
