"""include/ppg_shim.hpp (the reference-signature drop-in classes) must at least compile: built here against
stand-in types because OpenCV/Eigen and the reference headers are not in this container."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_header_compiles_against_stub_types():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
           os.path.join(ROOT, "tests", "shim_stub"), os.path.join(ROOT, "tests", "shim_stub", "shim_check.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_header_is_plain_c():
    src = "#include \"ppg_b200.h\"\nint main(void){ppg_config c; ppg_default_config(&c); return c.max_batch;}\n"
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", "-I", os.path.join(ROOT, "include"), "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
