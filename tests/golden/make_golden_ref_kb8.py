"""Golden vectors made by the REFERENCE's own C++ for the two functions whose camera model matters (oracle/ref_build.py
compiles map/src/Frame.cpp, matching/src/Matcher.cpp, sensors/src/Pinhole.cpp and sensors/src/KannalaBrandt8.cpp from
/root/reference):
  * Frame::CheckInFrustum (Frame.cpp:223-260) on real MapPoint objects with the real Pinhole / KannalaBrandt8 camera;
  * Matcher::SearchForTriangulation (Matcher.cpp:767-885) with the real KannalaBrandt8 camera (epipolarConstrain =
    TriangulateMatches, KannalaBrandt8.cpp:167-236; its Eigen::JacobiSVD is the stand-in's, oracle/ref_standins), plus
    single TriangulateMatches calls with their return codes and triangulated points.
Run in the build container: python tests/golden/make_golden_ref_kb8.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402

TRI = [  # name, camera, seed, generator arguments
    ("kb8tri0", "TUM-VI", 31, dict(n1=110, n2=120, n_nodes=8, frac_mp=0.15, noise_px=0.5)),
    ("kb8tri1", "UMA-VI", 32, dict(n1=96, n2=90, n_nodes=5, frac_mp=0.3, noise_px=0.8, forward=True)),
    ("kb8tri2", "TUM-VI-1024", 33, dict(n1=60, n2=130, n_nodes=30, frac_mp=0.0, noise_px=0.3)),
]
FRUSTUM = [("fru0", "EuRoC", 41, 700), ("fru1", "TUM-VI", 42, 700), ("fru2", "UMA-VI", 43, 500)]


def main():
    out = {}
    for name, cname, seed, kw in TRI:
        cam = cameras.ALL[cname]
        x = synth.two_view_inputs(seed, cam, **kw)
        ref = R.search_for_triangulation(cam, x["R1"], x["t1"], x["R2"], x["t2"], x["pos1"], x["desc1"], x["node1"],
                                         x["has_mp1"], x["pos2"], x["desc2"], x["node2"], x["has_mp2"])
        for k, v in x.items():
            out[name + "/" + k] = v
        out[name + "/camera"] = np.array(cname)
        out[name + "/ref_match12"] = ref["match12"]
        out[name + "/ref_nmatches"] = np.array([ref["nmatches"]], np.int32)
        for k in ("epipole", "R12", "t12"):
            out[name + "/ref_" + k] = ref[k]
        # single pairs through KannalaBrandt8::TriangulateMatches: the matched ones and as many arbitrary ones
        rs = np.random.RandomState(seed)
        pairs = [(i, j) for i, j in enumerate(ref["match12"]) if j >= 0][:24]
        pairs += [(int(rs.randint(len(x["pos1"]))), int(rs.randint(len(x["pos2"])))) for _ in range(24)]
        val, x3d, r1 = [], [], []
        for i, j in pairs:
            t = R.kb8_triangulate(cam, x["pos1"][i], x["pos2"][j], ref["R12"], ref["t12"])
            val.append(t["value"])
            x3d.append(t["x3D"])
            r1.append(t["r1"])
        out[name + "/pairs"] = np.array(pairs, np.int32)
        out[name + "/ref_pair_value"] = np.array(val, np.float32)
        out[name + "/ref_pair_x3D"] = np.array(x3d, np.float32)
        out[name + "/ref_pair_r1"] = np.array(r1, np.float32)
        print(name, cname, kw, "nmatches", ref["nmatches"], "pair codes",
              sorted({int(v) if v < 0 else 1 for v in val}))
    for name, cname, seed, m in FRUSTUM:
        cam = cameras.ALL[cname]
        g = synth.frustum_inputs(seed, cam, m, n_frames=1)
        ref = R.check_in_frustum(cam, g["Rcw"][0], g["tcw"][0], g["Ow"][0], g["world_pos"], g["normal"], g["min_dist"],
                                 g["max_dist"], 0.5)
        for k, v in g.items():
            out[name + "/" + k] = v
        out[name + "/camera"] = np.array(cname)
        for k in ("in_view", "proj_uv", "depth", "view_cos", "visible"):
            out[name + "/ref_" + k] = ref[k]
        print(name, cname, "points", m, "in view", int(ref["in_view"].sum()))
    path = os.path.join(ROOT, "tests", "golden", "ref_l2_kb8.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
