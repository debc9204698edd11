#!/usr/bin/env python
"""Generates tests/golden/ref_l2.npz from the REFERENCE'S OWN C++ (oracle/ref_build.py: matching/src/Matcher.cpp,
feature/src/MapPoint.cpp, map/src/Frame.cpp, feature/src/PPGGraph.cpp compiled from /root/reference against the Eigen /
OpenCV stand-ins).  Runs in the build container only; the fixture travels.

Every case: the flat inputs of Matcher::ExtendMapMatches (already renumbered so that table order = the order in which the
reference's unstable std::sort walked the candidates -- the one documented divergence, see oracle/ref_harness.py) and what
the reference made of them on its pointer graph: F.mvpMapPoints, F.mvpMapEdges, mnTrackedbyFrame, the returned count.
Plus window queries of Frame::GetFeaturesInArea in the reference's visiting order.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_build, ref_harness as R  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402
from tests.test_oracle_extend import random_frame_graph  # noqa: E402

CAMS = {"EuRoC": cameras.EUROC, "TUM-VI": cameras.TUMVI}
# (camera, seed, keypoints, key edges, map rows, th, ratio, clean, clustered)
CASES = [("EuRoC", 1, 40, 70, 120, 10.0, 0.8, False, False), ("EuRoC", 2, 44, 110, 139, 15.0, 0.95, False, True),
         ("EuRoC", 3, 30, 40, 90, 3.0, 0.6, True, False), ("TUM-VI", 4, 36, 80, 100, 10.0, 0.8, False, True),
         ("EuRoC", 5, 12, 0, 30, 10.0, 0.8, False, False), ("EuRoC", 6, 45, 120, 64, 10.0, 0.8, False, True)]


def init_pair(seed, cam, n1, n2, noise, jitter):
    rs = np.random.RandomState(seed)
    kx1 = rs.uniform(5, cam.width - 5, n1).astype(np.float32)
    ky1 = rs.uniform(5, cam.height - 5, n1).astype(np.float32)
    d1 = rs.normal(size=(n1, 256)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    k = min(n1, n2) * 2 // 3
    src = rs.choice(n1, k, replace=False)
    kx2 = np.concatenate([kx1[src] + rs.uniform(-jitter, jitter, k), rs.uniform(5, cam.width - 5, n2 - k)]).astype(np.float32)
    ky2 = np.concatenate([ky1[src] + rs.uniform(-jitter, jitter, k), rs.uniform(5, cam.height - 5, n2 - k)]).astype(np.float32)
    d2 = np.concatenate([d1[src] + rs.normal(0, noise, (k, 256)), rs.normal(size=(n2 - k, 256))]).astype(np.float32)
    q = min(k // 4, n2 - k)
    d2[k:k + q] = d1[src[:q]] + rs.normal(0, noise * 1.5, (q, 256))  # a second, slightly worse copy nearby
    kx2[k:k + q] = kx2[:q] + 3
    ky2[k:k + q] = ky2[:q] - 2
    d1[rs.choice(n1, n1 // 6, replace=False)] = d1[src[:n1 // 6]] + rs.normal(0, noise, (n1 // 6, 256))
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    p = rs.permutation(n2)
    return kx1, ky1, d1.astype(np.float32), np.stack([kx1, ky1], 1).copy(), kx2[p], ky2[p], d2[p].astype(np.float32)


def main():
    if not ref_build.build():
        raise SystemExit("the reference tree is not available here")
    out = {}
    for c, (cname, seed, n, ne, M, th, ratio, clean, clustered) in enumerate(CASES):
        cam = CAMS[cname]
        rs = np.random.RandomState(7000 + seed)
        kx, ky, fd, es, ee, coff, cidx, _ = random_frame_graph(rs, cam, n, ne)
        if clustered:  # everything inside a few search windows: many competing candidates per keypoint
            kx = (0.4 * cam.width + rs.uniform(0, 60, n)).astype(np.float32)
            ky = (0.4 * cam.height + rs.uniform(0, 40, n)).astype(np.float32)
        inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=th,
                                  planted_frac=0.7, clean=clean)
        ok = R.consistent_edge_ok(inp["bad"], inp["edge_off"], inp["edge_other"], inp["edge_ok"])
        perm = R.walk_order_permutation(inp["candidate"], inp["bad"], inp["edge_off"])
        t = R.permute_table(perm, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                            inp["edge_other"], ok, inp["proj_uv"], inp["view_cos"], inp["tracked"], inp["kp_mp"])
        # the reference runs on the ORIGINAL numbering; its result is translated into the renumbered table
        ref = R.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                                   inp["edge_other"], ok, inp["proj_uv"], inp["view_cos"], inp["tracked"], kx, ky, fd,
                                   inp["kp_mp"], es, ee, coff, cidx, th, ratio)
        km = np.where(ref["kp_mp"] >= 0, t["new_of_old"][np.maximum(ref["kp_mp"], 0)], ref["kp_mp"]).astype(np.int32)
        me = np.where(ref["kedge_me"] >= 0, t["pos_new_of_old"][np.maximum(ref["kedge_me"], 0)], -1).astype(np.int32)
        tr = ref["tracked"][t["old_of_new"]]
        pre = "case%d/" % c
        out[pre + "meta"] = np.array([list(CAMS).index(cname), th, ratio], np.float64)
        for k in ("map_desc", "candidate", "observed", "bad", "edge_off", "edge_other", "edge_ok", "proj_uv", "view_cos",
                  "tracked", "kp_mp"):
            out[pre + k] = t[k]
        for k, v in (("kp_x", kx), ("kp_y", ky), ("frame_desc", fd), ("edge_start", es), ("edge_end", ee),
                     ("conn_off", coff), ("conn_idx", cidx)):
            out[pre + k] = v
        out[pre + "ref_nmatches"] = np.array([ref["nmatches"]], np.int32)
        out[pre + "ref_kp_mp"], out[pre + "ref_kedge_me"], out[pre + "ref_tracked"] = km, me, tr
        print(pre, cname, "n", n, "M", M, "nmatches", ref["nmatches"], "matched kp", int((km >= 0).sum()),
              "map edges", int((me >= 0).sum()))
    # Frame::GetFeaturesInArea: keypoints + queries + the reference's answer (visiting order)
    for cname, cam in CAMS.items():
        rs = np.random.RandomState(11)
        n = 400
        kx = rs.uniform(-20, cam.width + 20, n).astype(np.float32)
        ky = rs.uniform(-20, cam.height + 20, n).astype(np.float32)
        q = np.stack([rs.uniform(-30, cam.width + 30, 60), rs.uniform(-30, cam.height + 30, 60),
                      rs.choice([7.5, 25.0, 40.0, 60.0], 60)], 1).astype(np.float32)
        ans, off = [], [0]
        for x, y, r in q:
            a = R.features_in_area(cam, kx, ky, float(x), float(y), float(r))
            ans.extend(a.tolist())
            off.append(len(ans))
        pre = "area_%s/" % cname
        out[pre + "kx"], out[pre + "ky"], out[pre + "queries"] = kx, ky, q
        out[pre + "ans"], out[pre + "off"] = np.array(ans, np.int32), np.array(off, np.int32)
        print(pre, "hits", len(ans))
    # Matcher::SearchForInitialization (Matcher.cpp:582-651): two frames, F2 = moved copies of part of F1 + clutter +
    # near-duplicates competing for the same feature
    for c, (cname, seed, n1, n2, noise, jitter, window, ratio) in enumerate(
            [("EuRoC", 1, 180, 200, 0.04, 30.0, 50, 0.9), ("TUM-VI", 2, 120, 90, 0.04, 40.0, 100, 0.85),
             ("EuRoC", 3, 60, 150, 0.02, 5.0, 20, 0.6)]):
        cam = CAMS[cname]
        a = init_pair(seed, cam, n1, n2, noise, jitter)
        ref = R.search_for_initialization(cam, a[0], a[1], a[2], a[3], a[4], a[5], a[6], window, ratio)
        pre = "init%d/" % c
        out[pre + "meta"] = np.array([list(CAMS).index(cname), window, ratio], np.float64)
        for k, v in zip(("kx1", "ky1", "desc1", "prev", "kx2", "ky2", "desc2"), a):
            out[pre + k] = v
        out[pre + "ref_nmatches"] = np.array([ref["nmatches"]], np.int32)
        out[pre + "ref_matches12"], out[pre + "ref_prev"] = ref["matches12"], ref["prev_matched"]
        print(pre, cname, "n1", n1, "n2", n2, "nmatches", ref["nmatches"])
    path = os.path.join(ROOT, "tests", "golden", "ref_l2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
