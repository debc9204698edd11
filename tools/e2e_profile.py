#!/usr/bin/env python
"""Host-side time of each public call of one e2e step (single context), to see what pipelining has to hide."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import capi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cam, frames = bench.make_workload(B)
e = capi.Extractor(cam, max_batch=B, max_map_points=bench.MAP_ROWS)
recs = e.run(frames)
e.run(frames)
map_desc, per_frame = bench.make_assoc_inputs(cam, recs, bench.MAP_ROWS)
e.upload_map(map_desc)
proj_all = np.stack([uv for uv, _ in per_frame])
vcos_all = np.stack([vc for _, vc in per_frame])
keep, fptrs, fstrides, _ = e._frame_ptrs(frames)
acc = {}


def t(name, fn):
    t0 = time.perf_counter()
    r = fn()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
    return r


for it in range(13):
    if it == 3:
        acc.clear()
    t("upload", lambda: e.lib.ppg_upload_frames(e.h, fptrs, fstrides, B))
    t("run(enqueue)", lambda: e.lib.ppg_run(e.h, B))
    t("sync", lambda: e.sync())
    t("download", lambda: e.lib.ppg_download(e.h, B, e._outs))
    t("assoc_stage_batch", lambda: e.assoc_stage_batch(proj_all, vcos_all, bench.TH, bench.RATIO))
    t("assoc_run_batch", lambda: e.assoc_run_batch(B))
    t("assoc_fetch_batch", lambda: e.assoc_fetch_batch(B))
for k, v in acc.items():
    print("%-20s %.3f ms/step" % (k, v / 10 * 1e3))
print("total %.3f ms/step" % (sum(acc.values()) / 10 * 1e3))
e.close()
