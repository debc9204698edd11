"""Times the association step of one extraction batch on the GPU: search core only (ppg_assoc_run_batch) against the
whole ExtendMapMatches (ppg_extend_run_batch: lists + sequential walk with seed growing), CUDA events on the ctx
stream.  python tools/extend_time.py [batch] [map_rows]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cam = cameras.EUROC
e = capi.Extractor(cam, max_batch=B, max_map_points=max(M, 1024))
recs = e.run([synth.frame(s, cam.width, cam.height) for s in range(B)])
r0 = recs[0]
inp = synth.extend_inputs(17, r0["desc"], np.stack([r0["kp_x"], r0["kp_y"]], 1), r0["edge_start"], r0["edge_end"], M,
                          cam.width, cam.height, th=10.0, clean=True)
e.upload_map(inp["map_desc"])
e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"], inp["edge_ok"])
uv = np.stack([inp["proj_uv"] + np.random.RandomState(100 + f).uniform(-2, 2, inp["proj_uv"].shape).astype(np.float32)
               for f in range(B)])
vc = np.stack([inp["view_cos"]] * B)
e.assoc_stage_batch(uv, vc, 10.0, 0.8)
out = {"batch": B, "map_rows": M}
for name, fn in (("search_core_ms", e.assoc_run_batch), ("extend_map_matches_ms", e.extend_run_batch)):
    for _ in range(3):
        fn(B)
    e.sync()
    e.timer_start()
    for _ in range(10):
        fn(B)
    out[name] = e.timer_stop() / 10
import os as _os
from ppg_slam_b200 import vocabulary as _voc
e.upload_vocabulary(_voc.load_blob(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                                  "ppg_slam_b200", "weights", "voc_euroc_9x3.bin")))
for _ in range(3):
    e.bow_run_batch(B, 4)
e.sync()
e.timer_start()
for _ in range(10):
    e.bow_run_batch(B, 4)
out["bow_transform_ms"] = e.timer_stop() / 10
got = e.extend_fetch_batch(B)
out["accepted_per_frame"] = float(np.mean([g["n_accepted"] for g in got]))
out["grown_per_frame"] = float(np.mean([g["n_grown"] for g in got]))
out["rescans_per_frame"] = float(np.mean([g["n_rescans"] for g in got]))
out["matched_keypoints_per_frame"] = float(np.mean([(g["kp_mp"] >= 0).sum() for g in got]))
out["n_kp_frame0"] = int(r0["n_kp"])
out["walk_diag_mean(rounds,setup,chunk,eval,event,seed cycles/16,seeds,weights cycles/16,seeds skipped)"] = np.mean([g["diag"] for g in got], 0).round(0).tolist()
out["walk_diag_frame0"] = got[0]["diag"]
print(json.dumps(out))
