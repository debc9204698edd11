#!/usr/bin/env python
"""Short deterministic workload for ncu: two batch-32 extract steps + association of 2 frames per step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402

B = int(os.environ.get("PPG_NCU_BATCH", "32"))
cam = cameras.EUROC
frames = [synth.frame(s, cam.width, cam.height) for s in range(B)]
e = capi.Extractor(cam, max_batch=B, max_map_points=8192)
recs = e.run(frames)
r0 = recs[0]
inp = synth.association_inputs(17, r0["desc"], np.stack([r0["kp_x"], r0["kp_y"]], 1), 8192, cam.width, cam.height)
e.upload_map(inp["map_desc"])
e.assoc_stage(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 256), np.float32), np.zeros(0, np.uint8),
              inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
for step in range(2):
    e.run_device(B)
    for f in range(2):
        e.assoc_run_frame(f)
e.sync()
print("launches", e.launch_count())
e.close()
