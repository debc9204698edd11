"""GPU sweep of ppg_extend_map_matches over small random configurations against the oracle (the CPU twin of this sweep
is tests/test_oracle_extend.py::test_extend_map_matches_random_small_graphs).  Written after the round-1 GPU budget was
spent: run it first thing in the next round (python tools/extend_sweep_check.py [n_cases]) and move it into
tests/test_gpu_extend.py once it has passed on a B200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import post_ref as O  # noqa: E402  (checker)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402
from tests.test_oracle_extend import random_frame_graph  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    cam = cameras.EUROC
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    bad = 0
    for seed in range(cases):
        rs = np.random.RandomState(1000 + seed)
        n = int(rs.randint(3, 45))
        M = int(rs.randint(5, 140))
        ne = int(rs.randint(0, min(3 * n, n * (n - 1) // 2) + 1))
        kx, ky, fd, es, ee, coff, cidx, _ = random_frame_graph(rs, cam, n, ne)
        if seed % 3 == 0:
            kx = (300 + rs.uniform(0, 60, n)).astype(np.float32)
            ky = (200 + rs.uniform(0, 40, n)).astype(np.float32)
        th = float(rs.choice([3.0, 10.0, 15.0]))
        inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=th,
                                  planted_frac=float(rs.uniform(0.2, 0.9)), clean=bool(seed % 4 == 1))
        if seed % 5 == 2 and len(inp["edge_other"]) > 4:
            inp["edge_other"][1::7] = inp["edge_other"][0::7][:len(inp["edge_other"][1::7])]
        ratio = float(rs.choice([0.6, 0.8, 0.95]))
        ref = O.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                                   inp["edge_other"], inp["edge_ok"], inp["proj_uv"], inp["view_cos"], inp["tracked"],
                                   kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, th=th, ratio=ratio)
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        got = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                   inp["tracked"], th, ratio)
        ok = (got["nmatches"] == ref["nmatches"] and np.array_equal(got["kp_mp"], ref["kp_mp"]) and
              np.array_equal(got["kedge_me"], ref["kedge_me"]) and np.array_equal(got["tracked"], ref["tracked"]))
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, "n", n, "M", M, "edges", ne, "th", th, "ratio", ratio)
    e.close()
    print("extend sweep: %d cases, %d mismatches" % (cases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
