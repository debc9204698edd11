"""Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (matching/src/Matcher.cpp:31-87: every frame tracked with
the motion model), Matcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, descDist) (:1337-1411:
relocalisation) and Matcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:479-568: loop
closing), WHOLE -- both are sequential, an accepted match occupies its keypoint for the later map points -- pinned
to the reference's own C++ (oracle/ref_build.py): tests/golden/ref_l2_projection.npz holds what the real functions did to
Frame / MapPoint / KeyFrame objects (tests/golden/make_golden_ref_projection.py); the oracle (CPU) and
ppg_search_by_projection (GPU) must reproduce CurrentFrame.mvpMapPoints and the count exactly, and the drop-in class
ppg_shim::Matcher is executed next to ::Matcher on identical objects."""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _z():
    return np.load(os.path.join(GOLD, "ref_l2_projection.npz"))


def _names():
    return sorted({k.split("/")[0] for k in _z().files})


def _case(name):
    z = _z()
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    d["cam"] = cameras.ALL[str(d["camera"])]
    d["mode"], d["th"], d["dd"] = int(d["mode_th_dd"][0]), float(d["mode_th_dd"][1]), float(d["mode_th_dd"][2])
    d["accept_dist"] = _max_dist(d["mode"], d["dd"])
    return d


def _max_dist(mode, dd):
    """what the functions accept: TH_HIGH (mode 0, stored), descDist (mode 1), TH_LOW * ratioHamming (mode 2, :559)"""
    return float(np.float32(0.7) * np.float32(dd)) if mode == 2 else dd


def _run(fn, d, valid, uv):
    """rows / recoding as a caller of the C ABI does it (synth.projection_rows) -> the reference's coding of the result"""
    q = synth.projection_rows(d, d["mode"], valid, uv)
    got = fn(q["map_desc"], q["proj_uv"], q["observed"], d["kp_x"], d["kp_y"], d["desc"], q["kp_mp"], d["th"],
             d["accept_dist"])
    return got["nmatches"], synth.projection_result(d, q["rows"], got["kp_mp"]), q


@pytest.mark.parametrize("name", _names())
def test_oracle_reproduces_the_reference_search_by_projection(name):
    from oracle import post_ref as O
    d = _case(name)
    nm, res, q = _run(lambda *a: O.search_by_projection(d["cam"], *a), d, d["ref_row_valid"], d["ref_proj_uv"])
    assert nm == int(d["ref_nmatches"][0]) and nm >= 10
    np.testing.assert_array_equal(res, d["ref_kp_mp"])


def test_fixture_needs_the_live_state():
    """The walk is not a frozen-state search: matching every row against the INITIAL CurrentFrame.mvpMapPoints gives other
    answers (several map points want the same keypoint)."""
    from oracle import post_ref as O
    differ = 0
    for name in _names():
        d = _case(name)
        q = synth.projection_rows(d, d["mode"], d["ref_row_valid"], d["ref_proj_uv"])
        _, live, _ = _run(lambda *a: O.search_by_projection(d["cam"], *a), d, d["ref_row_valid"], d["ref_proj_uv"])
        # the initial occupancy alone, with no row in the table: a keypoint held by an observed point is taken (-2)
        if q["observed"] is None:
            frozen = np.minimum(q["kp_mp"], -1)
        else:
            held = q["observed"][np.maximum(q["kp_mp"], 0)] > 0
            frozen = np.where(q["kp_mp"] >= 0, np.where(held, -2, -1), q["kp_mp"]).astype(np.int32)
        taken = {}
        for r in range(len(q["rows"])):
            one = O.search_by_projection(d["cam"], q["map_desc"][r:r + 1], q["proj_uv"][r:r + 1],
                                         None if q["observed"] is None else q["observed"][r:r + 1], d["kp_x"], d["kp_y"],
                                         d["desc"], frozen, d["th"], d["accept_dist"])
            k = np.nonzero(one["kp_mp"] == 0)[0]
            if len(k):
                taken.setdefault(int(k[0]), []).append(int(q["rows"][r]))
        differ += sum(1 for v in taken.values() if len(v) > 1)
    assert differ >= 5


def _live_cases():
    for ci, cam in enumerate((cameras.EUROC, cameras.TUMVI, cameras.UMA, cameras.TUMVI1024)):
        for seed in range(6):
            for mode, th, dd in ((0, 15.0, 0.8), (0, 7.0, 0.8), (0, 30.0, 0.8), (1, 10.0, 0.5), (1, 3.0, 64.0),
                                 (2, 8.0, 1.5), (2, 4.0, 1.0)):
                yield cam, 400 + 10 * ci + seed, mode, th, dd, dict(n_src=[300, 40, 500, 1][seed % 4],
                                                                      n=[340, 500, 60, 5][(seed // 2) % 4])


def test_oracle_equals_reference_search_by_projection_live():
    """168 random configurations on four calibrations (both camera models), the three functions, window radii 3 - 30, 1 - 500
    source features against 5 - 500 keypoints: the reference's own functions (here) and the oracle."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    total = 0
    for cam, seed, mode, th, dd, kw in _live_cases():
        x = synth.projection_inputs(seed, cam, **kw)
        x["scale"] = [1.0, 1.7, 0.6][seed % 3]
        ref = R.search_by_projection(cam, mode, x, th, dd)
        d = dict(x, mode=mode, th=th, dd=dd, accept_dist=_max_dist(mode, dd))
        nm, res, _ = _run(lambda *a: O.search_by_projection(cam, *a), d, ref["row_valid"], ref["proj_uv"])
        assert nm == ref["nmatches"], (cam.name, seed, mode, th)
        np.testing.assert_array_equal(res, ref["kp_mp"], err_msg="%s seed %d mode %d th %g" % (cam.name, seed, mode, th))
        total += nm
    assert total > 4000


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names())
def test_cuda_search_by_projection_reproduces_the_reference(name):
    from ppg_slam_b200 import capi
    d = _case(name)
    e = capi.Extractor(d["cam"], max_batch=1, max_map_points=1024)
    try:
        nm, res, _ = _run(e.search_by_projection, d, d["ref_row_valid"], d["ref_proj_uv"])
    finally:
        e.close()
    assert nm == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(res, d["ref_kp_mp"])


@pytest.mark.gpu
def test_cuda_search_by_projection_equals_oracle_sweep():
    """ppg_search_by_projection against the oracle on 40 configurations up to 2000 rows x 1000 keypoints, window radii up
    to 60 px (windows of more than 32 keypoints: the stored lists are cut and the best-only rule reads their first free
    entry, or rescans the window when none is left), pre-assigned and unobserved points, empty sides."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    total = rescans = 0
    for cam in (cameras.EUROC, cameras.TUMVI):
        e = capi.Extractor(cam, max_batch=1, max_map_points=2048)
        try:
            for seed in range(20):
                n_src, n = [300, 2000, 40, 1, 900][seed % 5], [340, 1000, 60, 5][(seed // 2) % 4]
                th = [15.0, 7.0, 60.0, 30.0][seed % 4]
                mode = seed % 2
                x = synth.projection_inputs(500 + seed, cam, n_src=n_src, n=n, frac_dup=[0.25, 0.6][seed % 2])
                valid = (x["state"] == 1) & (x["inside_numpy"] > 0)  # the split is the caller's: numpy's projection here
                uv = x["uv_numpy"]
                d = dict(x, mode=mode, th=th, dd=[0.8, 0.5][mode], accept_dist=[0.8, 0.5][mode])
                nm0, want, _ = _run(lambda *a: O.search_by_projection(cam, *a), d, valid, uv)
                q = synth.projection_rows(d, mode, valid, uv)
                got = e.search_by_projection(q["map_desc"], q["proj_uv"], q["observed"], d["kp_x"], d["kp_y"], d["desc"],
                                             q["kp_mp"], d["th"], d["accept_dist"])
                assert got["nmatches"] == nm0, (cam.name, seed)
                np.testing.assert_array_equal(synth.projection_result(d, q["rows"], got["kp_mp"]), want,
                                              err_msg="%s seed %d" % (cam.name, seed))
                total += nm0
                rescans += got["n_rescans"]
        finally:
            e.close()
    assert total > 1500


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names())
def test_shim_search_by_projection_equals_the_reference_on_real_objects(name):
    """include/ppg_shim.hpp compiled against the reference's real Frame.h / KeyFrame.h / MapPoint.h and executed: the shim
    runs the projection tests with the reference's own classes, flattens, calls the GPU and writes
    CurrentFrame.mvpMapPoints back; the same objects go through ::Matcher."""
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    d = _case(name)
    ref, shim = R.shim_projection_both(d["cam"], d["mode"], d, d["th"], d["dd"])
    assert shim["nmatches"] == ref["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(ref["kp_mp"], d["ref_kp_mp"])
    np.testing.assert_array_equal(shim["kp_mp"], ref["kp_mp"])
