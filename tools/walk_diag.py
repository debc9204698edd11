"""Where the sequential ExtendMapMatches kernel spends its cycles on the benchmark workload (every frame with its own
local map): mean of the per-frame counters the kernel keeps (SM cycles / 16).  python tools/walk_diag.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import capi  # noqa: E402

B = 32
cam, frames = bench.make_workload(B)
e = capi.Extractor(cam, max_batch=B, max_map_points=bench.MAP_ROWS)
recs = e.run(frames)
base = bench.make_assoc_inputs(cam, recs, bench.MAP_ROWS)
bench.upload_map(e, base)
e.assoc_stage_batch(base["proj_all"], base["vcos_all"], bench.TH, bench.RATIO)
for _ in range(3):
    e.extend_run_batch(B)
e.sync()
e.timer_start()
for _ in range(10):
    e.extend_run_batch(B)
ms = e.timer_stop() / 10
got = e.extend_fetch_batch(B)
d = np.array([g["diag"] for g in got], dtype=np.float64)
names = ["rounds", "setup", "chunk", "eval", "event", "seed", "seeds", "weights", "seeds_skipped"]
out = {"lists+walk_ms_per_batch": ms, "accepted": float(np.mean([g["n_accepted"] for g in got])),
       "grown": float(np.mean([g["n_grown"] for g in got]))}
for i, n in enumerate(names[:d.shape[1]]):
    out[n] = {"mean": float(d[:, i].mean()), "max": float(d[:, i].max())}
for n in ("setup", "chunk", "eval", "event", "seed", "weights"):
    if n in out:
        out[n]["mean_us_at_1.9GHz"] = out[n]["mean"] * 16 / 1900.0
print(json.dumps(out, indent=1))
e.close()
