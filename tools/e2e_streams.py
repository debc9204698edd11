#!/usr/bin/env python
"""Where does the multi-context e2e arm lose time?  (a) device-only steps from T threads / contexts (kernel
interference between streams), (b) full host-in / host-out steps from T threads."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppg_slam_b200 import capi  # noqa: E402

B = 32
STEPS = 12
cam, frames = bench.make_workload(B)


def make():
    e = capi.Extractor(cam, max_batch=B, max_map_points=bench.MAP_ROWS)
    recs = e.run(frames)
    return e, recs


e0, recs = make()
map_desc, per_frame = bench.make_assoc_inputs(cam, recs, bench.MAP_ROWS)
proj_all = np.stack([uv for uv, _ in per_frame])
vcos_all = np.stack([vc for _, vc in per_frame])
ctxs = [e0] + [make()[0] for _ in range(int(sys.argv[1]) - 1 if len(sys.argv) > 1 else 3)]
for x in ctxs:
    x.upload_map(map_desc)
    x.upload(frames)
    x.assoc_stage_batch(proj_all, vcos_all, bench.TH, bench.RATIO)
keep, fptrs, fstrides, _ = e0._frame_ptrs(frames)


def dev_step(x):
    x.run_device(B)
    x.assoc_run_batch(B)
    x.sync()


def e2e_step(x):
    x.lib.ppg_extract(x.h, fptrs, fstrides, B, x._outs)
    x.assoc_stage_batch(proj_all, vcos_all, bench.TH, bench.RATIO)
    x.assoc_run_batch(B)
    x.assoc_fetch_batch(B)


def run(fn, T):
    for x in ctxs[:T]:
        fn(x)
        fn(x)

    def worker(x):
        for _ in range(STEPS):
            fn(x)
    th = [threading.Thread(target=worker, args=(x,)) for x in ctxs[:T]]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return T * STEPS * B / dt


for T in range(1, len(ctxs) + 1):
    print("contexts %d: device-only %.0f frames/s, e2e %.0f frames/s" % (T, run(dev_step, T), run(e2e_step, T)),
          flush=True)
for x in ctxs:
    x.close()
