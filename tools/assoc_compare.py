#!/usr/bin/env python
"""BASELINE config 4 (UMA-VI scale: ~1000 keypoints against 50 000 map descriptors) through the two association paths:
  gemm  : brute-force bf16 tcgen05 GEMM + fused window test / top-4 (assoc_gemm_kernel) + exact re-score
  lists : windowed exact-distance lists (extend_lists_kernel), the state-independent half of ExtendMapMatches
and the same comparison at the benchmark's scale (357 x 8192, batch 32).  Both produce the same best / second best
(tests/test_gpu_assoc.py, tests/test_gpu_extend.py); this tool only times them.  Writes gpurun_out/assoc_compare.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402


def med(ts, k):
    return float(np.median([dict(t)[k] for t in ts if k in dict(t)]))


def one(cam, N, M, th=10.0, ratio=0.8, reps=7):
    rs = np.random.RandomState(3)
    kx = rs.uniform(8, cam.width - 8, N).astype(np.float32)
    ky = rs.uniform(8, cam.height - 8, N).astype(np.float32)
    fd = rs.normal(size=(N, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    es = np.zeros(0, np.int32)
    coff = np.zeros(N + 1, np.int32)
    inp = synth.extend_inputs(4, fd, np.stack([kx, ky], 1), es, es, M, cam.width, cam.height, th=th, clean=True)
    x = capi.Extractor(cam, max_batch=1, max_map_points=M, junction_max_num=1024)
    out = {"keypoints": N, "rows": M}
    try:
        x.upload_map(inp["map_desc"])
        x.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"], inp["edge_ok"])
        x.assoc_stage(kx, ky, fd, np.ones(N, np.uint8), inp["proj_uv"], inp["view_cos"], th, ratio)
        for _ in range(3):
            x.assoc_run()
        x.sync()
        x.set_profiling(True)
        ts = []
        for _ in range(reps):
            x.assoc_run()
            x.sync()
            ts.append(x.stage_times())
        out["gemm_ms"] = med(ts, "assoc.gemm(top4)")
        out["rescore_ms"] = med(ts, "assoc.rescore")
        out["gemm_tflops"] = 2.0 * M * N * 256 / (out["gemm_ms"] * 1e-3) / 1e12
        ts = []
        for _ in range(reps):
            x.extend_map_matches(kx, ky, fd, np.full(N, -1, np.int32), es, es, coff, np.zeros(0, np.int32),
                                 inp["proj_uv"], inp["view_cos"], np.zeros(M, np.uint8), th, ratio)
            ts.append(x.stage_times())
        x.set_profiling(False)
        out["lists_ms"] = med(ts, "extend.lists")
        out["walk_ms"] = med(ts, "extend.walk")
        out["fallback_rows"] = x.assoc_fallback_rows()
    finally:
        x.close()
    return out


def main():
    res = {"uma_1000x50000": one(cameras.UMA, 1000, 50000), "euroc_357x8192": one(cameras.EUROC, 357, 8192),
           "euroc_500x50000": one(cameras.EUROC, 500, 50000)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "assoc_compare.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
