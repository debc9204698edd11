#!/usr/bin/env python
"""Two contexts driven from two host threads at the same time: does a context's dense output depend on what the
other one is doing?  Prints, per layer output that can be fetched (feature map, prob, heat), how many values differ from
the context's own single-threaded run and where (bounding box), for PPG_CONV_KERNEL settings given on the command line."""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from ppg_slam_b200 import cameras, capi, synth
    cam = cameras.EUROC
    frames = [synth.frame(s, cam.width, cam.height) for s in (10, 11, 12, 13)]
    a = capi.Extractor(cam, max_batch=4)
    b = capi.Extractor(cam, max_batch=4)
    rep = {"kernel": os.environ.get("PPG_CONV_KERNEL", "default")}
    try:
        a.run(frames)
        base = [a.get_maps(f, feature=True) for f in range(4)]
        base_rec = a.run(frames)
        stop = []

        def other():
            while not stop:
                b.run(frames[::-1])

        th = threading.Thread(target=other)
        th.start()
        bad = []
        for rep_i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
            recs = a.run(frames)
            for f in range(4):
                m = a.get_maps(f, feature=True)
                for k in ("feature", "prob", "heat"):
                    d = np.argwhere(m[k] != base[f][k])
                    if len(d):
                        bad.append(dict(rep=rep_i, frame=f, map=k, n=int(len(d)), lo=d.min(0).tolist(),
                                        hi=d.max(0).tolist(),
                                        maxabs=float(np.abs(m[k] - base[f][k]).max())))
                if recs[f]["n_kp"] != base_rec[f]["n_kp"]:
                    bad.append(dict(rep=rep_i, frame=f, n_kp=[int(recs[f]["n_kp"]), int(base_rec[f]["n_kp"])]))
        stop.append(1)
        th.join()
        rep["mismatches"] = bad[:40]
        rep["n_mismatches"] = len(bad)
    finally:
        a.close()
        b.close()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
