#!/usr/bin/env python
"""Timing experiments on the halo convolution kernel: one subprocess per PPG_CONV_FLAGS value (see conv_tc2_kernel),
prints the per-stage device times of one batch-32 EuRoC step.  Flags >= 0x100 produce wrong results on purpose."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json
sys.path.insert(0, %r)
from ppg_slam_b200 import cameras, capi, synth
cam = cameras.EUROC
B = 32
e = capi.Extractor(cam, max_batch=B)
e.upload([synth.frame(s, cam.width, cam.height) for s in range(B)])
for _ in range(3):
    e.run_device(B)
e.sync()
e.set_profiling(True)
acc = {}
for _ in range(5):
    e.run_device(B)
    for k, v in e.stage_times():
        acc[k] = acc.get(k, 0.0) + v / 5
print("RES " + json.dumps(acc))
e.close()
""" % ROOT


def main():
    flags = sys.argv[1:] or ["0", "1"]
    out = {}
    for f in flags:
        env = dict(os.environ, PPG_CONV_FLAGS=f)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RES ")]
        if not line:
            print(f, "FAILED", r.stderr[-500:])
            continue
        d = json.loads(line[0][4:])
        out[f] = d
        print(f, " ".join("%s=%.3f" % (k, d[k]) for k in ("conv1a+conv1b" if "conv1a+conv1b" in d else "conv1b", "conv2a", "conv2b", "conv3a", "edge1")), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_exp.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
