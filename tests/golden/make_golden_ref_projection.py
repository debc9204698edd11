"""Golden vectors of Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (Matcher.cpp:31-87),
Matcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, descDist) (:1337-1411) and
Matcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:479-568) made by the REFERENCE's own C++
(oracle/ref_build.py compiles matching/src/Matcher.cpp, map/src/Frame.cpp, feature/src/MapPoint.cpp and the two camera
classes from /root/reference): Frame / MapPoint / KeyFrame objects rebuilt from flat arrays, the real function, the
resulting CurrentFrame.mvpMapPoints -- plus which source features the reference's projection tests let through and
where they project.  Run in the build container: python tests/golden/make_golden_ref_projection.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402

CASES = [  # name, camera, seed, mode, th, descDist / TH_HIGH
    ("proj0", "EuRoC", 51, 0, 15.0, 0.8),
    ("proj1", "TUM-VI", 52, 0, 7.0, 0.8),
    ("proj2", "EuRoC", 53, 1, 10.0, 0.5),
    ("proj3", "UMA-VI", 54, 1, 3.0, 64.0),
    # mode 2 = SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:479-568): the last number is
    # ratioHamming (accept <= TH_LOW * ratioHamming), the similarity's scale is stored as <case>/scale
    ("proj4", "EuRoC", 55, 2, 8.0, 1.5),
    ("proj5", "TUM-VI", 56, 2, 4.0, 1.0),
]
SCALE = {"proj4": 1.3, "proj5": 0.8}


def main():
    out = {}
    for name, cname, seed, mode, th, dd in CASES:
        cam = cameras.ALL[cname]
        x = synth.projection_inputs(seed, cam, n_src=96, n=110)
        x["scale"] = np.float32(SCALE.get(name, 1.0))
        ref = R.search_by_projection(cam, mode, x, th, dd)
        for k, v in x.items():
            out[name + "/" + k] = v
        out[name + "/camera"] = np.array(cname)
        out[name + "/mode_th_dd"] = np.array([mode, th, dd], np.float32)
        out[name + "/ref_kp_mp"] = ref["kp_mp"]
        out[name + "/ref_nmatches"] = np.array([ref["nmatches"]], np.int32)
        out[name + "/ref_row_valid"], out[name + "/ref_proj_uv"] = ref["row_valid"], ref["proj_uv"]
        print(name, cname, "mode", mode, "th", th, "rows", int(ref["row_valid"].sum()), "nmatches", ref["nmatches"])
    path = os.path.join(ROOT, "tests", "golden", "ref_l2_projection.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
