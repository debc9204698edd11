"""Build recipe of the reference-pinning harness (oracle/_ref/libppg_ref_*.so).  TEST INFRASTRUCTURE ONLY.

Compiles the reference's OWN C++ sources for the hot path (PPGExtractor.cpp, PPGGraph.cpp, GeometricCamera.cpp, Pinhole.cpp,
Matcher.cpp, MapPoint.cpp, Frame.cpp), from where they lie under /root/reference, against
  * the real LibTorch (CPU) of the Python environment, and
  * the stand-ins of oracle/ref_standins/ for OpenCV, Eigen and DBoW3 (none of which is installed here; the stand-ins
    say which of their arithmetic is an assumption),
plus oracle/ppg_oracle.c for the four OpenCV routines it restates (pinned against cv2 by tests/test_oracle_cv.py).
Outputs go to oracle/_ref/ only (git-ignored, travels to the GPU box).  Nothing of the reference is copied into the
repository: the translation units under oracle/ref_tu/ `#include` the reference files by absolute path.

`python -m oracle.ref_build` builds; build() returns None when /root/reference (or torch) is not available, which is the
normal case on the GPU box -- tests/test_ref_pin.py then runs from the committed fixtures only.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("PPG_REFERENCE_ROOT", "/root/reference")
REF_INCLUDES = ["feature/include", "sensors/include", "map/include", "matching/include"]


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "feature", "src", "PPGExtractor.cpp"))


def lib_path(name):
    return os.path.join(OUT_DIR, "libppg_ref_%s.so" % name)


def _torch_flags():
    import torch
    t = os.path.dirname(torch.__file__)
    inc = ["-I" + os.path.join(t, "include"), "-I" + os.path.join(t, "include", "torch", "csrc", "api", "include")]
    lib = ["-L" + os.path.join(t, "lib"), "-ltorch", "-ltorch_cpu", "-lc10", "-Wl,-rpath," + os.path.join(t, "lib")]
    abi = ["-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    return inc, lib, abi


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    if not available():
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    inc, lib, abi = _torch_flags()
    std = os.path.join(HERE, "ref_standins")
    deps = [os.path.join(dp, f) for dp, _, fs in os.walk(std) for f in fs] + [os.path.join(HERE, "ppg_oracle.c")] + \
           [os.path.join(HERE, "ref_tu", f) for f in os.listdir(os.path.join(HERE, "ref_tu")) if f.endswith(".hpp")]
    common = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
              "-ffunction-sections", "-fdata-sections", "-w", "-DNDEBUG_OFF",
              '-DREF_FILE(x)=<' + REF_ROOT + '/x>', "-I" + std] + \
             ["-I" + os.path.join(REF_ROOT, d) for d in REF_INCLUDES] + inc + abi
    oracle_o = os.path.join(OUT_DIR, "ppg_oracle_for_ref.o")
    if force or _stale(oracle_o, [os.path.join(HERE, "ppg_oracle.c")]):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-c",
                               os.path.join(HERE, "ppg_oracle.c"), "-o", oracle_o])
    built = {}
    repo = os.path.dirname(HERE)
    product = os.path.join(repo, "ppg_slam_b200", "libppg_b200.so")
    for name in ("extractor", "matcher", "shim"):
        tu = os.path.join(HERE, "ref_tu", name + "_tu.cpp")
        if not os.path.exists(tu):
            continue
        extra, extra_deps = [], []
        if name == "shim":  # include/ppg_shim.hpp executed on the reference's objects: links the product library
            if not os.path.exists(product):
                continue
            extra = ["-I" + os.path.join(repo, "include"), "-L" + os.path.dirname(product), "-lppg_b200",
                     "-Wl,-rpath," + os.path.dirname(product)]
            extra_deps = [os.path.join(repo, "include", "ppg_shim.hpp"), os.path.join(repo, "include", "ppg_b200.h")]
        out = lib_path(name)
        if force or _stale(out, [tu, oracle_o] + deps + extra_deps):
            cmd = ["g++"] + common + ["-shared", tu, oracle_o, "-o", out, "-Wl,--gc-sections", "-lm"] + lib + extra
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference harness %s failed to build:\n%s" % (name, r.stderr[-6000:]))
            if verbose:
                sys.stderr.write(r.stderr)
        built[name] = out
    return built


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
