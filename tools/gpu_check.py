#!/usr/bin/env python
"""First-contact diagnostics on a B200: prints what every stage produces next to the oracle (no asserts),
so that one gpurun call localises a problem.  Writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppg_slam_b200 import cameras, capi, synth  # noqa: E402


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rep = {}
    from oracle.net_ref import NetRef
    from tests.parity_util import diff_records, oracle_post
    net = NetRef()
    cam = cameras.EUROC
    g = synth.frame(0, cam.width, cam.height)
    t = time.time()
    ref = net.forward_u8(g)
    rep["oracle_net_s"] = time.time() - t
    e = capi.Extractor(cam, max_batch=2)
    try:
        try:
            got = e.run_from_maps(ref["prob"][None], ref["heat"][None], ref["desc"][None], allow_capacity=True)[0]
            t = time.time()
            orc = oracle_post(cam, ref["prob"], ref["heat"], ref["desc"])
            rep["oracle_post_s"] = time.time() - t
            rep["post_from_maps"] = dict(diff=diff_records(got, orc),
                                         got={k: int(got[k]) for k in ("n_kp", "n_edges", "n_colines", "status",
                                                                       "n_cand", "n_pairs_ok", "n_candidate_lines",
                                                                       "nms_rounds")},
                                         want=dict(n_kp=int(orc["n_kp"]), n_edges=int(orc["n_edges"]),
                                                   n_col=len(orc["col_pairs"]), n_cand=int(orc["n_cand"]),
                                                   stats=orc["stats"].tolist()))
            hf = e.get_maps(0)["heat_final"]
            rep["post_from_maps"]["heat_final_mismatch_px"] = int((hf.view(np.uint32) !=
                                                                   orc["heat_final"].view(np.uint32)).sum())
        except Exception:
            rep["post_from_maps_error"] = traceback.format_exc()
        print(json.dumps(rep, indent=1), flush=True)
        try:
            e.set_profiling(True)
            rec = e.run([g], allow_capacity=True)[0]
            rep["full"] = {k: int(rec[k]) for k in ("n_kp", "n_edges", "n_colines", "status", "n_cand")}
            rep["stage_ms_b1"] = e.stage_times()
            rep["conv_selftest"] = e.selftest_conv()
            m = e.get_maps(0, feature=True)
            rep["dense"] = dict(
                feat_err=float(np.abs(m["feature"] - ref["feature"]).max()), feat_max=float(np.abs(ref["feature"]).max()),
                prob_err=float(np.abs(m["prob"] - ref["prob"]).max()),
                heat_err=float(np.abs(m["heat"] - ref["heat"]).max()),
                desc_err=float(np.abs(m["desc"] - ref["desc"]).max()), desc_max=float(np.abs(ref["desc"]).max()))
            a, b = m["desc"].reshape(256, -1), ref["desc"].reshape(256, -1)
            cos = (a * b).sum(0) / (np.linalg.norm(a, axis=0) * np.linalg.norm(b, axis=0) + 1e-12)
            rep["dense"]["desc_cos_min"] = float(cos.min())
        except Exception:
            rep["full_error"] = traceback.format_exc()
    finally:
        e.close()
    print(json.dumps(rep, indent=1), flush=True)
    try:
        B = 32
        e = capi.Extractor(cam, max_batch=B)
        frames = [synth.frame(s, cam.width, cam.height) for s in range(B)]
        e.upload(frames)
        for _ in range(2):
            e.run_device(B)
        e.sync()
        e.set_profiling(True)
        e.run_device(B)
        rep["stage_ms_b32"] = e.stage_times()
        e.set_profiling(False)
        e.timer_start()
        for _ in range(5):
            e.run_device(B)
        ms = e.timer_stop() / 5
        rep["b32_ms"] = ms
        rep["b32_fps"] = B / ms * 1e3
        recs = e.download(B, allow_capacity=True)
        rep["b32_records"] = [[r["n_kp"], r["n_edges"], r["n_colines"], r["status"], r["nms_rounds"],
                               r["n_pairs_ok"], r["n_candidate_lines"]] for r in recs]
        rep["b32_lines_phase_cycles_x16_mean"] = np.mean([r["diag"] for r in recs], axis=0).tolist()
        rep["b32_lines_phase_cycles_x16_max"] = np.max([r["diag"] for r in recs], axis=0).tolist()
        e.close()
    except Exception:
        rep["b32_error"] = traceback.format_exc()
    print(json.dumps(rep, indent=1), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
