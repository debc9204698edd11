#!/usr/bin/env python
"""Generates tests/golden/ref_l1.npz from the REFERENCE'S OWN C++ (oracle/ref_build.py: feature/src/PPGExtractor.cpp
compiled from /root/reference against LibTorch-CPU and the OpenCV / Eigen stand-ins).  Runs in the build container only
(the reference tree and its net/*.pt are not on the GPU box); the fixture travels.

Small cameras keep the fixture small: for each frame the dense maps the reference's stages consumed (prob, heat before
refine, raw descriptors; fp32, exactly as LibTorch produced them) and everything the reference produced from them
(keypoints, scores, undistorted positions, out flags, refined + remapped heat map, edges, line scores, mvConnected,
mvColine, normalised descriptors, image bounds).  tests/test_ref_pin.py feeds the maps to the oracle (CPU) and to the
CUDA post-processing (GPU) and compares bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_build, ref_harness  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402

# (name, camera, frame seed, synth arguments): a pinhole camera with EuRoC-like distortion (remap on), a fisheye one
# with the as-shipped parameter shift D = (0, k1, k2, k3) of the TUM-VI configuration (remap off)
CASES = [
    ("pinhole0", cameras.Camera("ref-pinhole", 256, 192, 156.0, 155.5, 125.2, 99.4,
                                (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), False), 3, dict(n_rect=10, n_line=8)),
    ("pinhole1", cameras.Camera("ref-pinhole", 256, 192, 156.0, 155.5, 125.2, 99.4,
                                (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05), False), 8, dict(n_rect=14, n_line=12)),
    ("fisheye0", cameras.Camera("ref-fisheye", 224, 224, 83.4, 83.3, 111.3, 112.6,
                                (0.0, 0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202), True), 5,
     dict(n_rect=12, n_line=10)),
]


def main():
    if not ref_build.build():
        raise SystemExit("the reference tree is not available here")
    out = {}
    for name, cam, seed, kw in CASES:
        g = synth.frame(seed, cam.width, cam.height, **kw)
        r = ref_harness.RefExtractor(cam, threads=1)
        rec, maps = r.run(g)
        b = r.image_bounds()
        r.close()
        print(name, "n_kp", rec["n_kp"], "edges", rec["n_edges"], "colines", len(rec["col_pairs"]),
              "nan lscore", int(np.isnan(rec["edge_score"]).sum()))
        out[name + "/cam"] = np.array([cam.width, cam.height, cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D) +
                                      [float(cam.fisheye)], np.float64)
        out[name + "/bounds"] = np.array([b["minX"], b["minY"], b["maxX"], b["maxY"], b["wInv"], b["hInv"]], np.float64)
        out[name + "/gray"] = g
        for k in ("prob", "heat_raw", "heat_final", "desc"):
            out[name + "/" + k] = maps[k]
        for k in ("pos", "xun", "yun", "score", "out", "edge_start", "edge_end", "edge_score", "conn_off", "conn_idx",
                  "col_off", "col_pairs", "desc"):
            out[name + "/rec_" + k] = rec[k]
    path = os.path.join(ROOT, "tests", "golden", "ref_l1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
