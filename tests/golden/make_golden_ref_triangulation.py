"""Golden vectors of Matcher::SearchForTriangulation and of both Matcher::SearchByBoW overloads made by the REFERENCE's own C++ (oracle/ref_build.py compiles
matching/src/Matcher.cpp and sensors/src/Pinhole.cpp from /root/reference): two pinhole key frames rebuilt from flat
arrays, the real function, its vMatchedPairs -- plus F12 and the epipole as the reference's classes compute them from the
poses.  Run in the build container: python tests/golden/make_golden_ref_triangulation.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402

CASES = [  # name, seed, generator arguments
    ("tri0", 11, dict(n1=110, n2=120, n_nodes=8, frac_mp=0.15, noise_px=0.5)),
    ("tri1", 12, dict(n1=96, n2=90, n_nodes=5, frac_mp=0.3, noise_px=0.8, forward=True)),
    ("tri2", 13, dict(n1=60, n2=130, n_nodes=30, frac_mp=0.0, noise_px=0.3)),
]


def main():
    cam = cameras.EUROC
    out = {}
    for name, seed, kw in CASES:
        x = synth.two_view_inputs(seed, cam, **kw)
        ref = R.search_for_triangulation(cam, x["R1"], x["t1"], x["R2"], x["t2"], x["pos1"], x["desc1"], x["node1"],
                                         x["has_mp1"], x["pos2"], x["desc2"], x["node2"], x["has_mp2"])
        for k, v in x.items():
            out[name + "/" + k] = v
        out[name + "/ref_match12"] = ref["match12"]
        out[name + "/ref_nmatches"] = np.array([ref["nmatches"]], np.int32)
        out[name + "/ref_F12"], out[name + "/ref_epipole"] = ref["F12"], ref["epipole"]
        print(name, kw, "nmatches", ref["nmatches"], "epipole", ref["epipole"])
    # Matcher::SearchByBoW, both overloads (Matcher.cpp:393-477, :663-754), by the reference's own functions
    for name, seed, kw, ratio in [("bow0", 21, dict(n1=90, n2=100, n_nodes=4), 0.8),
                                  ("bow1", 22, dict(n1=70, n2=60, n_nodes=1, frac_none=0.1, frac_bad=0.25), 0.6)]:
        x = synth.bow_pair_inputs(seed, **kw)
        a = R.search_by_bow_kf_f(cam, x["desc1"], x["node1"], x["state1"], x["desc2"], x["node2"], ratio)
        b = R.search_by_bow_kf_kf(cam, x["desc1"], x["node1"], x["state1"], x["desc2"], x["node2"], x["state2"], ratio)
        for k, v in x.items():
            out[name + "/" + k] = v
        out[name + "/ratio"] = np.array([ratio], np.float32)
        out[name + "/ref_f2kf"], out[name + "/ref_match12"] = a["f2kf"], b["match12"]
        out[name + "/ref_nmatches"] = np.array([a["nmatches"], b["nmatches"]], np.int32)
        print(name, kw, "KF-F", a["nmatches"], "KF-KF", b["nmatches"])
    path = os.path.join(ROOT, "tests", "golden", "ref_l2_triangulation.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
