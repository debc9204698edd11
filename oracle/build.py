"""Build recipe for the C oracle (oracle/_ref/libppg_oracle.so).  TEST INFRASTRUCTURE ONLY.

The reference's own C++ for this path cannot be compiled here (needs OpenCV C++ headers, Eigen,
LibTorch-CUDA), so oracle/_ref holds only our restatement, built from
oracle/ppg_oracle.c.  -ffp-contract=off: the reference is built for baseline x86-64 without FMA.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
FLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC"]


def lib_path(variant=""):
    return os.path.join(OUT_DIR, "libppg_oracle%s.so" % variant)


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    src = os.path.join(HERE, "ppg_oracle.c")
    for variant, defs in (("", []), ("_libmf", ["-DPPGO_LIBM_FLOAT=1"])):
        out = lib_path(variant)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
            continue
        subprocess.check_call(["gcc"] + FLAGS + defs + [src, "-o", out, "-lm"])
    return lib_path()


if __name__ == "__main__":
    print(build(force=True))
