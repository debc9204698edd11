#!/usr/bin/env python
"""Short deterministic workload for ncu: the bench step (batch-32 extract + Matcher::ExtendMapMatches vs 8192 map rows),
one warm-up step through the host path and two device-resident steps (PPG_NCU_ASSOC=core: search core only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import capi  # noqa: E402

B = int(os.environ.get("PPG_NCU_BATCH", "32"))
cam, frames = bench.make_workload(B)
e = capi.Extractor(cam, max_batch=B, max_map_points=bench.MAP_ROWS)
recs = e.run(frames)
base = bench.make_assoc_inputs(cam, recs, bench.MAP_ROWS)
bench.upload_map(e, base)
e.assoc_stage_batch(base["proj_all"], base["vcos_all"], bench.TH, bench.RATIO)
for step in range(2):
    e.run_device(B)
    if os.environ.get("PPG_NCU_ASSOC", "extend") == "core":
        e.assoc_run_batch(B)
    else:
        e.extend_run_batch(B)
e.sync()
print("launches", e.launch_count())
e.close()
