#!/usr/bin/env python
"""Short workload for ncu captures of the streaming post-processing kernels and the association GEMM: two batch-32 EuRoC
steps (scan, refine, remap, descriptor sampling, ...) and two runs of the 1000 x 50 000 association (UMA-VI scale)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402

B = 32
cam, frames = bench.make_workload(B)
e = capi.Extractor(cam, max_batch=B)
e.run(frames)
for _ in range(2):
    e.run_device(B)
e.sync()
e.close()
cam = cameras.UMA
N, M = 1000, 50000
rs = np.random.RandomState(3)
kx = rs.uniform(8, cam.width - 8, N).astype(np.float32)
ky = rs.uniform(8, cam.height - 8, N).astype(np.float32)
fd = rs.normal(size=(N, 256)).astype(np.float32)
fd /= np.linalg.norm(fd, axis=1, keepdims=True)
inp = synth.association_inputs(4, fd, np.stack([kx, ky], 1), M, cam.width, cam.height, th=10.0)
x = capi.Extractor(cam, max_batch=1, max_map_points=M, junction_max_num=1024)
x.upload_map(inp["map_desc"])
x.assoc_stage(kx, ky, fd, np.ones(N, np.uint8), inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
for _ in range(3):
    x.assoc_run()
x.sync()
x.close()
print("done")
