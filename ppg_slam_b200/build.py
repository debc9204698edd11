"""In-tree build of libppg_b200.so (sm_100a only).  `python -m ppg_slam_b200.build [--force]`.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
post.cu, assoc.cu, extend.cu, bow.cu and triang.cu are built with -fmad=false: their results must be bit-identical to the reference's
SSE2 (no-FMA) float arithmetic (CMakeLists.txt:8-9 of the reference sets no -march).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libppg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr"]
SOURCES = {
    "api.cu": [],
    "conv_tc.cu": [],
    "conv_t64.cu": [],
    "conv_t128.cu": [],
    "net_direct.cu": [],
    "conv1a_tc.cu": [],
    "post.cu": ["-fmad=false"],
    "assoc.cu": ["-fmad=false"],
    "extend.cu": ["-fmad=false"],
    "bow.cu": ["-fmad=false"],
    "triang.cu": ["-fmad=false"],
    "voc_io.cu": [],
    "comm.cu": [],
}


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "ppg_b200.h")]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = _deps()
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + deps):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=5) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
