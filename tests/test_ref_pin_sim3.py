"""Matcher::SearchBySim3 (matching/src/Matcher.cpp:1149-1335; loop closing) pinned to the reference's own C++.

The function is two FROZEN-state passes -- every map point of KF1 looks for its best feature of KF2 within th and
TH_HIGH, and the other way round -- followed by a mutual check (:1310-1327): nothing a pass accepts changes what a later
map point may take.  Each pass therefore IS the best-only window core the product offers as `ppg_associate` with
`PPG_SEARCH_WINDOW` (bit-identical to the oracle's `ppgo_search_all(mode = 1)` on the GPU:
tests/test_gpu_assoc.py::test_window_search_modes_match_oracle), and this file shows that the core + the mutual check
reproduce the reference's real function on two key frames with real MapPoint objects (tests/golden/ref_l2_sim3.npz from
tests/golden/make_golden_ref_sim3.py, and live where the harness is present)."""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TH_HIGH = 0.8


def _by_core(cam, x, ref, th):
    """two passes of the frozen best-only window core + the mutual check"""
    from oracle import post_ref as O

    def one_pass(valid, uv, mp_desc, pos_o, desc_o):
        rows = np.nonzero(valid)[0]
        best = np.full(len(valid), -1, np.int32)
        if len(rows) and len(pos_o):
            r = O.search_all(cam, pos_o[:, 0], pos_o[:, 1], desc_o, np.ones(len(pos_o), np.uint8), mp_desc[rows], uv[rows],
                             np.zeros(len(rows), np.float32), th, 1.0, mode=1, max_dist=TH_HIGH)
            best[rows] = np.where(r["accept"] > 0, r["best_idx"], -1)
        return best
    m1 = one_pass(ref["valid1"], ref["uv1"], x["mp_desc1"], x["pos2"], x["desc2"])
    m2 = one_pass(ref["valid2"], ref["uv2"], x["mp_desc2"], x["pos1"], x["desc1"])
    out, nf = np.array(x["matches12"], np.int32).copy(), 0
    for i1 in range(len(m1)):
        if m1[i1] >= 0 and m2[m1[i1]] == i1:
            out[i1] = m1[i1]
            nf += 1
    return nf, out


def _names():
    return sorted({k.split("/")[0] for k in np.load(os.path.join(GOLD, "ref_l2_sim3.npz")).files})


@pytest.mark.parametrize("name", _names())
def test_window_core_and_mutual_check_reproduce_the_reference_search_by_sim3(name):
    z = np.load(os.path.join(GOLD, "ref_l2_sim3.npz"))
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    cam = cameras.ALL[str(d["camera"])]
    ref = {k[4:]: d[k] for k in d if k.startswith("ref_")}
    nf, out = _by_core(cam, d, ref, float(d["th"]))
    assert nf == int(ref["nfound"]) and nf >= 5
    np.testing.assert_array_equal(out, ref["matches12"])
    # already matched features stay as they were and do not search (:1166-1176)
    pre = d["matches12"] >= 0
    assert pre.any() and not ref["valid1"][pre].any()
    np.testing.assert_array_equal(ref["matches12"][pre], d["matches12"][pre])


def test_window_core_equals_reference_search_by_sim3_live():
    """72 random key-frame pairs on three calibrations (both camera models), similarity scales 0.97 - 1.03, 1 - 500
    features, bad points, features without a point, pre-filled matches, two radii."""
    from oracle import ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    total = 0
    for cam in (cameras.EUROC, cameras.TUMVI, cameras.UMA):
        for seed in range(12):
            x = synth.sim3_inputs(700 + seed, cam, n1=[300, 60, 500, 1][seed % 4], n2=[320, 400, 50, 7][(seed // 2) % 4],
                                  scale=[1.0, 1.03, 0.97][seed % 3])
            for th in (7.5, 15.0):
                ref = R.search_by_sim3(cam, x, th)
                nf, out = _by_core(cam, x, ref, th)
                assert nf == ref["nfound"], (cam.name, seed, th)
                np.testing.assert_array_equal(out, ref["matches12"], err_msg="%s seed %d th %g" % (cam.name, seed, th))
                total += nf
    assert total > 400
