"""Golden vectors of Matcher::SearchBySim3 (matching/src/Matcher.cpp:1149-1335) made by the REFERENCE's own C++
(oracle/ref_build.py): two raw key frames whose features carry real MapPoint objects, the real function, the resulting
vpMatches12 -- plus which features the loop heads let search and where they project into the other key frame.
Run in the build container: python tests/golden/make_golden_ref_sim3.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as R  # noqa: E402
from ppg_slam_b200 import cameras, synth  # noqa: E402

CASES = [("sim0", "EuRoC", 61, 7.5, 1.0), ("sim1", "TUM-VI", 62, 15.0, 1.03)]


def main():
    out = {}
    for name, cname, seed, th, scale in CASES:
        cam = cameras.ALL[cname]
        x = synth.sim3_inputs(seed, cam, n1=90, n2=100, scale=scale)
        ref = R.search_by_sim3(cam, x, th)
        for k, v in x.items():
            out[name + "/" + k] = v
        out[name + "/camera"] = np.array(cname)
        out[name + "/th"] = np.float32(th)
        for k, v in ref.items():
            out[name + "/ref_" + k] = np.asarray(v)
        print(name, cname, "th", th, "searching", int(ref["valid1"].sum()), int(ref["valid2"].sum()), "found", ref["nfound"])
    path = os.path.join(ROOT, "tests", "golden", "ref_l2_sim3.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
