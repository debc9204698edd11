#!/usr/bin/env python
"""Host-to-host time of the single-call matchers (one frame / one key-frame pair per call, pageable host arrays in,
results out, the ctx synchronised inside the call): Matcher::SearchByProjection x 2, SearchForInitialization,
SearchForTriangulation (Pinhole / KannalaBrandt8), SearchByBoW.  Prints one JSON line (median of 30 calls after 5)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402


def med(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return round(1e3 * float(np.median(t)), 4)


def main():
    out = {}
    cam = cameras.EUROC
    e = capi.Extractor(cam, max_batch=1, max_map_points=2048)
    x = synth.projection_inputs(1, cam, n_src=400, n=350)
    valid = (x["state"] == 1) & (x["inside_numpy"] > 0)
    q = synth.projection_rows(x, 0, valid, x["uv_numpy"])
    r = {}

    def proj():
        r["p"] = e.search_by_projection(q["map_desc"], q["proj_uv"], q["observed"], x["kp_x"], x["kp_y"], x["desc"],
                                        q["kp_mp"], 15.0, 0.8)
    out["search_by_projection_ms"] = med(proj)
    out["search_by_projection"] = dict(rows=int(len(q["rows"])), keypoints=350, nmatches=int(r["p"]["nmatches"]))
    t = synth.two_view_inputs(3, cam, n1=350, n2=350, n_nodes=40)
    R1, t1, R2, t2 = (t[k].astype(np.float64) for k in ("R1", "t1", "R2", "t2"))
    R12, t12 = R1 @ R2.T, t1 - R1 @ R2.T @ t2
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(K.T) @ tx @ R12 @ np.linalg.inv(K)).astype(np.float32)
    C2 = R2 @ (-R1.T @ t1) + t2
    ep = np.array([fx * C2[0] / C2[2] + cx, fy * C2[1] / C2[2] + cy], np.float32)

    def tri():
        r["t"] = e.search_for_triangulation(t["desc1"], t["node1"], t["has_mp1"], t["pos1"], t["desc2"], t["node2"],
                                            t["has_mp2"], t["pos2"], F12, ep)
    out["search_for_triangulation_pinhole_ms"] = med(tri)
    out["search_for_triangulation_pinhole"] = dict(n1=350, n2=350, nmatches=int(r["t"]["nmatches"]))
    e.close()
    cam = cameras.TUMVI
    e = capi.Extractor(cam, max_batch=1, max_map_points=2048)
    t = synth.two_view_inputs(3, cam, n1=350, n2=350, n_nodes=40)
    R1, t1, R2, t2 = (t[k].astype(np.float64) for k in ("R1", "t1", "R2", "t2"))
    R12, t12 = (R1 @ R2.T).astype(np.float32), (t1 - R1 @ R2.T @ t2).astype(np.float32)
    cam8 = [cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D)
    ep = np.array([cam.K[2], cam.K[5]], np.float32) + 300  # far from the features

    def tri8():
        r["k"] = e.search_for_triangulation(t["desc1"], t["node1"], t["has_mp1"], t["pos1"], t["desc2"], t["node2"],
                                            t["has_mp2"], t["pos2"], np.zeros(9), ep, kb8=(cam8, R12, t12))
    out["search_for_triangulation_kb8_ms"] = med(tri8)
    out["search_for_triangulation_kb8"] = dict(n1=350, n2=350, nmatches=int(r["k"]["nmatches"]))
    e.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
