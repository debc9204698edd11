#!/usr/bin/env python
"""A/B of the convolution kernel choice (PPG_CONV_KERNEL = 1 generic, 2 halo, 3 transposed for the 64 -> 64 layers):
self-test errors against the CUDA-core reference convolution + per-layer timings at batch 32."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402

cam = cameras.EUROC
B = 32
frames = [synth.frame(s, cam.width, cam.height) for s in range(B)]
out = {}
for v in sys.argv[1:] or ["2", "3"]:
    os.environ["PPG_CONV_KERNEL"] = v
    try:
        e = capi.Extractor(cam, max_batch=B)
        e.upload(frames)
        e.run_device(B)
        e.sync()
        st = e.selftest_conv()
        e.set_profiling(True)
        e.run_device(B)
        times = dict(e.stage_times())
        e.set_profiling(False)
        e.timer_start()
        for _ in range(5):
            e.run_device(B)
        ms = e.timer_stop() / 5
        out[v] = dict(selftest={n: d for n, d, r in st}, ms=ms, fps=B / ms * 1e3,
                      conv={k: round(t, 4) for k, t in times.items() if k.startswith(("conv", "edge"))})
        e.close()
    except Exception as ex:  # noqa: BLE001
        out[v] = dict(error=str(ex))
    print(v, json.dumps(out[v]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_variants.json"), "w"), indent=1)
