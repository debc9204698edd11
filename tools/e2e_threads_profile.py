#!/usr/bin/env python
"""Per-call wall time of the e2e step when T contexts run on T host threads (what inflates under concurrency?)."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import capi  # noqa: E402

B, T, STEPS = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 16
if len(sys.argv) > 2:
    sys.setswitchinterval(float(sys.argv[2]))
cam, frames = bench.make_workload(B)
ctxs = [capi.Extractor(cam, max_batch=B, max_map_points=bench.MAP_ROWS) for _ in range(T)]
recs = ctxs[0].run(frames)
map_desc, per_frame = bench.make_assoc_inputs(cam, recs, bench.MAP_ROWS)
proj_all = np.stack([uv for uv, _ in per_frame])
vcos_all = np.stack([vc for _, vc in per_frame])
for x in ctxs:
    x.upload_map(map_desc)
keep, fptrs, fstrides, _ = ctxs[0]._frame_ptrs(frames)
acc = [dict() for _ in range(T)]


def step(i, x, record):
    def t(name, fn):
        t0 = time.perf_counter()
        fn()
        if record:
            acc[i][name] = acc[i].get(name, 0.0) + time.perf_counter() - t0
    t("upload", lambda: x.lib.ppg_upload_frames(x.h, fptrs, fstrides, B))
    t("run", lambda: x.lib.ppg_run(x.h, B))
    t("download", lambda: x.lib.ppg_download(x.h, B, x._outs))
    t("stage", lambda: x.assoc_stage_batch(proj_all, vcos_all, bench.TH, bench.RATIO))
    t("assoc_run", lambda: x.assoc_run_batch(B))
    t("fetch", lambda: x.assoc_fetch_batch(B))


def worker(i):
    for k in range(STEPS + 2):
        step(i, ctxs[i], k >= 2)


th = [threading.Thread(target=worker, args=(i,)) for i in range(T)]
t0 = time.perf_counter()
for t_ in th:
    t_.start()
for t_ in th:
    t_.join()
dt = time.perf_counter() - t0
print("contexts %d: %.0f frames/s (incl. 2 warm-up steps per context)" % (T, T * (STEPS + 2) * B / dt))
tot = {}
for a in acc:
    for k, v in a.items():
        tot[k] = tot.get(k, 0.0) + v / (T * STEPS) * 1e3
print({k: round(v, 3) for k, v in tot.items()}, "sum %.3f ms per step per context" % sum(tot.values()))
for x in ctxs:
    x.close()
