"""Parity pinned to the REFERENCE'S OWN C++: feature/src/PPGExtractor.cpp (detectKeyPoint, detectLines, refineHeatMap,
heatMapInlierRate, heatMapLineScore, bilinearInterpolation, genPointDescriptor), feature/src/PPGGraph.cpp and
sensors/src/GeometricCamera.cpp compiled unmodified from /root/reference against LibTorch-CPU and the OpenCV / Eigen
stand-ins (oracle/ref_build.py), and matching/src/Matcher.cpp (ExtendMapMatches) the same way.

 * tests/golden/ref_l1.npz / ref_l2.npz hold what that code produced (tests/golden/make_golden_ref*.py, run in the build
   container); the oracle (CPU) and the CUDA path (GPU) must reproduce every discrete output bit for bit from the dense
   maps / inputs stored beside them.  These tests need neither the reference tree nor the harness.
 * where the harness is present (this container), the oracle is also compared LIVE with the reference on the full-size
   EuRoC / TUM-VI / UMA-VI frames.
"""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(GOLD, "ref_l1.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    out = []
    for n in names:
        c = z[n + "/cam"]
        cam = cameras.Camera("ref-" + n, int(c[0]), int(c[1]), float(c[2]), float(c[3]), float(c[4]), float(c[5]),
                             tuple(float(v) for v in c[6:10]), bool(c[10]))
        out.append((n, cam, {k[len(n) + 1:]: z[k] for k in z.files if k.startswith(n + "/")}))
    return out


def _same_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def _check_against_reference(got, d, cam, heat_final=None, desc_tol=1e-6):
    """got: record of the oracle / the CUDA path; d: what the reference's C++ produced."""
    n = len(d["rec_score"])
    assert int(got["n_kp"]) == n
    np.testing.assert_array_equal(got["px"], d["rec_pos"][:, 0].astype(np.int32))  # mPos: integer-valued floats
    np.testing.assert_array_equal(got["py"], d["rec_pos"][:, 1].astype(np.int32))
    assert _same_f32(got["score"], d["rec_score"])
    assert _same_f32(got["xun"], d["rec_xun"]) and _same_f32(got["yun"], d["rec_yun"])  # cv::undistortPoints
    np.testing.assert_array_equal(np.asarray(got["out"]).astype(np.uint8), d["rec_out"])
    assert int(got["n_edges"]) == len(d["rec_edge_start"])
    np.testing.assert_array_equal(got["edge_start"], d["rec_edge_start"])
    np.testing.assert_array_equal(got["edge_end"], d["rec_edge_end"])
    assert _same_f32(got["edge_score"], d["rec_edge_score"])  # lscore, NaN for the 5 <= dist < 6 lines
    np.testing.assert_array_equal(got["conn_off"], d["rec_conn_off"])
    np.testing.assert_array_equal(got["conn_idx"], d["rec_conn_idx"])
    np.testing.assert_array_equal(got["col_off"], d["rec_col_off"])
    np.testing.assert_array_equal(np.asarray(got["col_pairs"]).reshape(-1, 2), d["rec_col_pairs"].reshape(-1, 2))
    if heat_final is not None:
        assert _same_f32(heat_final, d["heat_final"])  # refineHeatMap + cv::remap
    # torch::grid_sampler + F::normalize: same bilinear weights, other summation order
    assert np.abs(np.asarray(got["desc"]) - d["rec_desc"]).max() <= desc_tol


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_oracle_reproduces_the_reference_extractor(case):
    from oracle import post_ref as O
    name, cam, d = case
    got = O.extract_post(cam, d["prob"], d["heat_raw"], d["desc"])
    _check_against_reference(got, d, cam, heat_final=got["heat_final"])
    b = O.image_bounds(cam)  # GeometricCamera::InitializeImageBounds
    assert [b.minX, b.minY, b.maxX, b.maxY] == [int(v) for v in d["bounds"][:4]]
    assert np.float32(b.wInv) == np.float32(d["bounds"][4]) and np.float32(b.hInv) == np.float32(d["bounds"][5])


def test_fixture_exercises_the_quirks():
    """The fixture is only worth something if the reference took its odd paths in it."""
    nan_lines = outs = cols = 0
    for _, _, d in _cases():
        nan_lines += int(np.isnan(d["rec_edge_score"]).sum())  # NaN < 0.8 is false: accepted (PPGExtractor.cpp:376)
        outs += int(d["rec_out"].sum())                        # undistorted out of the image: excluded from the graph
        cols += len(d["rec_col_pairs"])
    assert nan_lines >= 3 and outs >= 5 and cols >= 20


@pytest.mark.parametrize("cam,seed", [(cameras.EUROC, 0), (cameras.EUROC, 5), (cameras.TUMVI, 1), (cameras.UMA, 2),
                                      (cameras.TUMVI1024, 3)],
                         ids=["euroc-0", "euroc-5", "tumvi-1", "uma-2", "tumvi1024-3"])
def test_oracle_equals_reference_live(cam, seed):
    """Full-size frames through the reference's own C++ (here) and through the oracle fed with the dense maps the
    reference computed: every discrete output identical."""
    from oracle import post_ref as O, ref_harness as R
    if not R.available():
        pytest.skip("reference harness not built (no /root/reference on this host)")
    r = R.RefExtractor(cam, threads=8)
    try:
        rec, maps = r.run(synth.frame(seed, cam.width, cam.height))
    finally:
        r.close()
    assert rec["n_kp"] > 100
    d = {"rec_" + k: v for k, v in rec.items() if not isinstance(v, int)}
    d["heat_final"] = maps["heat_final"]
    got = O.extract_post(cam, maps["prob"], maps["heat_raw"], maps["desc"])
    _check_against_reference(got, d, cam, heat_final=got["heat_final"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_cuda_post_processing_reproduces_the_reference_extractor(case):
    """The CUDA post-processing fed with the dense maps the reference's stages saw, against what the reference's C++
    made of them -- no oracle in between."""
    from ppg_slam_b200 import capi
    name, cam, d = case
    e = capi.Extractor(cam, max_batch=1)
    try:
        got = e.run_from_maps(d["prob"][None], d["heat_raw"][None], d["desc"][None])[0]
        hf = e.get_maps(0)["heat_final"]
        _check_against_reference(got, d, cam, heat_final=hf)
    finally:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_full_cuda_path_stays_close_to_the_reference(case):
    """End to end (fp16 tensor-core networks + post-processing) equality with the fp32 reference is statistical, not bit
    exact: a probability within 5e-3 of the 1/128 threshold or of a neighbour's score can flip a keypoint.  Bound the
    difference of the keypoint sets on the fixture frames."""
    from ppg_slam_b200 import capi
    name, cam, d = case
    e = capi.Extractor(cam, max_batch=1)
    try:
        got = e.run([d["gray"]])[0]
    finally:
        e.close()
    ref = {(int(x), int(y)) for x, y in d["rec_pos"]}
    mine = {(int(x), int(y)) for x, y in zip(got["px"], got["py"])}
    common = len(ref & mine)
    assert common >= 0.85 * len(ref) and len(mine) <= 1.15 * len(ref) + 2, (common, len(ref), len(mine))


# ------------------------------------------------------------------------------------------------ Matcher (L2)
def _l2():
    return np.load(os.path.join(GOLD, "ref_l2.npz"))


def _l2_cases():
    z = _l2()
    return sorted({k.split("/")[0] for k in z.files if k.startswith("case")})


_L2_CAMS = [cameras.EUROC, cameras.TUMVI]


def _l2_case(name):
    z = _l2()
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    cam = _L2_CAMS[int(d["meta"][0])]
    return cam, float(d["meta"][1]), float(d["meta"][2]), d


def _same_as_reference(got, d):
    assert int(got["nmatches"]) == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(got["kp_mp"], d["ref_kp_mp"])       # F.mvpMapPoints
    np.testing.assert_array_equal(got["kedge_me"], d["ref_kedge_me"])  # F.mvpMapEdges
    np.testing.assert_array_equal(got["tracked"], d["ref_tracked"])    # mnTrackedbyFrame == F.mnId


@pytest.mark.parametrize("name", _l2_cases())
def test_oracle_reproduces_the_reference_extend_map_matches(name):
    """Matcher::ExtendMapMatches as the reference's own C++ ran it on MapPoint / MapEdge / Frame objects (non-candidate,
    bad, unobserved and already-tracked rows, invalid and dangling map edges, keypoints that hold a map point): the
    oracle reproduces F.mvpMapPoints, F.mvpMapEdges, the tracked flags and the count."""
    from oracle import post_ref as O
    cam, th, ratio, d = _l2_case(name)
    got = O.extend_map_matches(cam, d["map_desc"], d["candidate"], d["observed"], d["bad"], d["edge_off"], d["edge_other"],
                               d["edge_ok"], d["proj_uv"], d["view_cos"], d["tracked"], d["kp_x"], d["kp_y"],
                               d["frame_desc"], d["kp_mp"], d["edge_start"], d["edge_end"], d["conn_off"], d["conn_idx"],
                               th=th, ratio=ratio)
    _same_as_reference(got, d)


@pytest.mark.parametrize("cname,cam", [("EuRoC", cameras.EUROC), ("TUM-VI", cameras.TUMVI)])
def test_oracle_reproduces_the_reference_get_features_in_area(cname, cam):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea of the reference (keypoints outside the image bounds
    included): same indices in the same visiting order."""
    from oracle import post_ref as O
    z = _l2()
    kx, ky, q = z["area_%s/kx" % cname], z["area_%s/ky" % cname], z["area_%s/queries" % cname]
    ans, off = z["area_%s/ans" % cname], z["area_%s/off" % cname]
    for i, (x, y, r) in enumerate(q):
        got = O.features_in_area(cam, kx, ky, float(x), float(y), float(r))
        np.testing.assert_array_equal(got, ans[off[i]:off[i + 1]])
    assert off[-1] > 100


def test_oracle_equals_reference_extend_live():
    """60 random configurations through the reference's own Matcher::ExtendMapMatches (here) and the oracle."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this host)")
    from tests.test_oracle_extend import random_frame_graph
    cam = cameras.EUROC
    for seed in range(60):
        rs = np.random.RandomState(1000 + seed)
        n, M = int(rs.randint(3, 45)), int(rs.randint(5, 140))
        ne = int(rs.randint(0, min(3 * n, n * (n - 1) // 2) + 1))
        kx, ky, fd, es, ee, coff, cidx, _ = random_frame_graph(rs, cam, n, ne)
        if seed % 3 == 0:
            kx = (300 + rs.uniform(0, 60, n)).astype(np.float32)
            ky = (200 + rs.uniform(0, 40, n)).astype(np.float32)
        th = float(rs.choice([3.0, 10.0, 15.0]))
        inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=th,
                                  planted_frac=float(rs.uniform(0.2, 0.9)), clean=bool(seed % 4 == 1))
        ok = R.consistent_edge_ok(inp["bad"], inp["edge_off"], inp["edge_other"], inp["edge_ok"])
        ratio = float(rs.choice([0.6, 0.8, 0.95]))
        ref = R.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                                   inp["edge_other"], ok, inp["proj_uv"], inp["view_cos"], inp["tracked"], kx, ky, fd,
                                   inp["kp_mp"], es, ee, coff, cidx, th, ratio)
        perm = R.walk_order_permutation(inp["candidate"], inp["bad"], inp["edge_off"])
        t = R.permute_table(perm, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                            inp["edge_other"], ok, inp["proj_uv"], inp["view_cos"], inp["tracked"], inp["kp_mp"])
        got = O.extend_map_matches(cam, t["map_desc"], t["candidate"], t["observed"], t["bad"], t["edge_off"],
                                   t["edge_other"], t["edge_ok"], t["proj_uv"], t["view_cos"], t["tracked"], kx, ky, fd,
                                   t["kp_mp"], es, ee, coff, cidx, th=th, ratio=ratio)
        km = np.where(ref["kp_mp"] >= 0, t["new_of_old"][np.maximum(ref["kp_mp"], 0)], ref["kp_mp"])
        me = np.where(ref["kedge_me"] >= 0, t["pos_new_of_old"][np.maximum(ref["kedge_me"], 0)], -1)
        assert got["nmatches"] == ref["nmatches"], seed
        np.testing.assert_array_equal(got["kp_mp"], km)
        np.testing.assert_array_equal(got["kedge_me"], me)
        np.testing.assert_array_equal(got["tracked"], ref["tracked"][t["old_of_new"]])


def _init_cases():
    z = _l2()
    return sorted({k.split("/")[0] for k in z.files if k.startswith("init")})


def _init_case(name):
    z = _l2()
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    return _L2_CAMS[int(d["meta"][0])], int(d["meta"][1]), float(d["meta"][2]), d


@pytest.mark.parametrize("name", _init_cases())
def test_oracle_reproduces_the_reference_search_for_initialization(name):
    """Matcher::SearchForInitialization as the reference's own C++ ran it on two Frames (vector<int> distance quirk
    included): vnMatches12, the updated vbPrevMatched and the count."""
    from oracle import post_ref as O
    cam, window, ratio, d = _init_case(name)
    got = O.search_for_initialization(cam, d["desc1"], d["prev"], d["kx2"], d["ky2"], d["desc2"], window, ratio)
    assert got["nmatches"] == int(d["ref_nmatches"][0]) and got["nmatches"] > 10
    np.testing.assert_array_equal(got["matches12"], d["ref_matches12"])
    np.testing.assert_array_equal(got["prev_matched"], d["ref_prev"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", _init_cases())
def test_cuda_search_for_initialization_reproduces_the_reference(name):
    from ppg_slam_b200 import capi
    cam, window, ratio, d = _init_case(name)
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    try:
        got = e.search_for_initialization(d["desc1"], d["prev"], d["kx2"], d["ky2"], d["desc2"], window, ratio)
    finally:
        e.close()
    assert got["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(got["matches12"], d["ref_matches12"])
    np.testing.assert_array_equal(got["prev_matched"], d["ref_prev"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", _l2_cases())
def test_cuda_extend_map_matches_reproduces_the_reference(name):
    """ppg_extend_map_matches against what the reference's C++ produced -- no oracle in between."""
    from ppg_slam_b200 import capi
    cam, th, ratio, d = _l2_case(name)
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    try:
        e.upload_map(d["map_desc"])
        e.upload_map_graph(d["candidate"], d["observed"], d["bad"], d["edge_off"], d["edge_other"], d["edge_ok"])
        got = e.extend_map_matches(d["kp_x"], d["kp_y"], d["frame_desc"], d["kp_mp"], d["edge_start"], d["edge_end"],
                                   d["conn_off"], d["conn_idx"], d["proj_uv"], d["view_cos"], d["tracked"], th, ratio)
    finally:
        e.close()
    _same_as_reference(got, d)


# ------------------------------------------------------------------------------------------------ the drop-in classes
@pytest.mark.gpu
@pytest.mark.parametrize("name", _l2_cases())
def test_shim_matcher_class_equals_the_reference_on_real_objects(name):
    """include/ppg_shim.hpp compiled against the reference's real Frame.h / MapPoint.h / PPGGraph.h / Matcher.h and
    EXECUTED: one pointer graph of MapPoint / MapEdge / Frame objects goes through the reference's own
    Matcher::ExtendMapMatches, an identical one through ppg_shim::Matcher::ExtendMapMatches (row ordering, row_of map,
    C ABI, GPU, write-back into F.mvpMapPoints / F.mvpMapEdges / mnTrackedbyFrame); both must leave the same state."""
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    cam, th, ratio, d = _l2_case(name)
    ref, shim = R.shim_extend_both(cam, d["map_desc"], d["candidate"], d["observed"], d["bad"], d["edge_off"],
                                   d["edge_other"], d["edge_ok"], d["proj_uv"], d["view_cos"], d["tracked"], d["kp_x"],
                                   d["kp_y"], d["frame_desc"], d["kp_mp"], d["edge_start"], d["edge_end"], d["conn_off"],
                                   d["conn_idx"], th, ratio)
    assert shim["nmatches"] == ref["nmatches"]
    np.testing.assert_array_equal(shim["kp_mp"], ref["kp_mp"])
    np.testing.assert_array_equal(shim["kedge_me"], ref["kedge_me"])
    np.testing.assert_array_equal(shim["tracked"], ref["tracked"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", _init_cases())
def test_shim_search_for_initialization_equals_the_reference_on_real_frames(name):
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    cam, window, ratio, d = _init_case(name)
    ref, shim = R.shim_init_both(cam, d["kx1"], d["ky1"], d["desc1"], d["prev"], d["kx2"], d["ky2"], d["desc2"], window,
                                 ratio)
    assert shim["nmatches"] == ref["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(shim["matches12"], ref["matches12"])
    np.testing.assert_array_equal(shim["prev_matched"], ref["prev_matched"])


# ---------------------------------------------------------------- Matcher::SearchForTriangulation (pinhole)
def _tri():
    return np.load(os.path.join(GOLD, "ref_l2_triangulation.npz"))


def _tri_cases():
    return sorted({k.split("/")[0] for k in _tri().files if k.startswith("tri")})


def _bow_cases():
    return sorted({k.split("/")[0] for k in _tri().files if k.startswith("bow")})


def _bow_from_oracle(O, x, ratio, kf_kf):
    """Flat arrays -> the rows / candidates the C ABI and the oracle take -> per-feature result in the reference's terms."""
    from ppg_slam_b200 import synth
    rows, rd, rn = synth.bow_rows(x["desc1"], x["node1"], x["state1"])
    n1, n2 = len(x["desc1"]), len(x["desc2"])
    kp_node = np.where(x["state2"] == 1, x["node2"], -1).astype(np.int32) if kf_kf else x["node2"]
    if len(rows) == 0:
        return 0, np.full(n1 if kf_kf else n2, -1, np.int32), (rows, rd, rn, kp_node)
    got = O(x["desc2"], kp_node, rd, rn, ratio, 0.7, bool(kf_kf))
    return got["nmatches"], _bow_result(got["kp_row"], rows, n1, kf_kf), (rows, rd, rn, kp_node)


def _bow_result(kp_row, rows, n1, kf_kf):
    if not kf_kf:  # f2kf: frame feature -> key-frame feature
        return np.where(kp_row >= 0, rows[np.maximum(kp_row, 0)], -1).astype(np.int32)
    m12 = np.full(n1, -1, np.int32)  # KF1 feature -> KF2 feature
    for i2, r in enumerate(kp_row):
        if r >= 0:
            m12[rows[r]] = i2
    return m12


def _tri_case(name):
    z = _tri()
    return {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}


@pytest.mark.parametrize("name", _tri_cases())
def test_oracle_reproduces_the_reference_search_for_triangulation(name):
    """Matcher::SearchForTriangulation (Matcher.cpp:767-885) as the reference's own C++ ran it, with its own Pinhole
    camera (epipolarConstrain, Pinhole.cpp:98-114), on two key frames: vMatches12 and the count, given the F12 / epipole
    the reference's classes computed from the poses."""
    from oracle import post_ref as O
    d = _tri_case(name)
    got = O.search_for_triangulation(d["desc1"], d["node1"], d["has_mp1"], d["pos1"], d["desc2"], d["node2"], d["has_mp2"],
                                     d["pos2"], d["ref_F12"], d["ref_epipole"])
    assert got["nmatches"] == int(d["ref_nmatches"][0]) and got["nmatches"] >= 10
    np.testing.assert_array_equal(got["match12"], d["ref_match12"])


def test_oracle_equals_reference_search_for_triangulation_live():
    """40 random key-frame pairs (sideways and forward motion -- the epipole inside the image --, 6 to 40 vocabulary nodes,
    0 to 50 % of the features already tracked, empty and one-feature frames) through the reference's own function (here)
    and the oracle; the fixture generator's poses go through the reference's SE3 / Pinhole classes."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    from ppg_slam_b200 import synth
    cam = cameras.EUROC
    total = 0
    for seed in range(40):
        kw = dict(n_nodes=[6, 12, 40][seed % 3], noise_px=[0.3, 0.6, 1.2][(seed // 3) % 3], forward=(seed % 4 == 3),
                  frac_mp=[0.0, 0.2, 0.5][(seed // 2) % 3], n1=[0, 1, 57, 300, 500][seed % 5] if seed < 10 else 300)
        x = synth.two_view_inputs(seed, cam, **kw)
        ref = R.search_for_triangulation(cam, x["R1"], x["t1"], x["R2"], x["t2"], x["pos1"], x["desc1"], x["node1"],
                                         x["has_mp1"], x["pos2"], x["desc2"], x["node2"], x["has_mp2"])
        got = O.search_for_triangulation(x["desc1"], x["node1"], x["has_mp1"], x["pos1"], x["desc2"], x["node2"],
                                         x["has_mp2"], x["pos2"], ref["F12"], ref["epipole"])
        assert got["nmatches"] == ref["nmatches"], seed
        np.testing.assert_array_equal(got["match12"], ref["match12"], err_msg="seed %d" % seed)
        total += ref["nmatches"]
    assert total > 800


@pytest.mark.gpu
@pytest.mark.parametrize("name", _tri_cases())
def test_cuda_search_for_triangulation_reproduces_the_reference(name):
    from ppg_slam_b200 import capi
    d = _tri_case(name)
    e = capi.Extractor(cameras.EUROC, max_batch=1)
    try:
        got = e.search_for_triangulation(d["desc1"], d["node1"], d["has_mp1"], d["pos1"], d["desc2"], d["node2"],
                                         d["has_mp2"], d["pos2"], d["ref_F12"], d["ref_epipole"])
    finally:
        e.close()
    assert got["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(got["match12"], d["ref_match12"])


@pytest.mark.gpu
def test_cuda_search_for_triangulation_equals_oracle_sweep():
    """24 random key-frame pairs up to 500 x 520 features: ppg_search_for_triangulation against the oracle, F12 / epipole
    from the fixture generator's poses (plain numpy here: /root/reference is not on the GPU box)."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi, synth
    cam = cameras.EUROC
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    e = capi.Extractor(cam, max_batch=1)
    total = 0
    try:
        for seed in range(24):
            kw = dict(n_nodes=[6, 12, 40][seed % 3], noise_px=[0.3, 0.6, 1.2][(seed // 3) % 3], forward=(seed % 4 == 3),
                      frac_mp=[0.0, 0.2, 0.5][(seed // 2) % 3], n1=[0, 1, 57, 300, 500][seed % 5],
                      n2=[320, 2, 33, 520][seed % 4])
            x = synth.two_view_inputs(100 + seed, cam, **kw)
            R1, t1, R2, t2 = (x[k].astype(np.float64) for k in ("R1", "t1", "R2", "t2"))
            R12, t12 = R1 @ R2.T, t1 - R1 @ R2.T @ t2  # T12 = T1w * Tw2
            tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
            F12 = (np.linalg.inv(K.T) @ tx @ R12 @ np.linalg.inv(K)).astype(np.float32)
            C2 = R2 @ (-R1.T @ t1) + t2
            ep = np.array([fx * C2[0] / C2[2] + cx, fy * C2[1] / C2[2] + cy], np.float32)
            args = (x["desc1"], x["node1"], x["has_mp1"], x["pos1"], x["desc2"], x["node2"], x["has_mp2"], x["pos2"], F12, ep)
            want = O.search_for_triangulation(*args)
            got = e.search_for_triangulation(*args)
            assert got["nmatches"] == want["nmatches"], seed
            np.testing.assert_array_equal(got["match12"], want["match12"], err_msg="seed %d" % seed)
            total += want["nmatches"]
    finally:
        e.close()
    assert total > 200


@pytest.mark.gpu
@pytest.mark.parametrize("name", _tri_cases())
def test_shim_search_for_triangulation_equals_the_reference_on_real_keyframes(name):
    """include/ppg_shim.hpp compiled against the reference's real KeyFrame.h / Pinhole.h / SE3.h and executed: the shim
    derives F12 and the epipole from the key frames' poses with the reference's own classes, flattens the FeatureVectors,
    calls the GPU and must hand back the vMatchedPairs of the reference's host function."""
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    d = _tri_case(name)
    ref, shim = R.shim_triangulation_both(cameras.EUROC, d["R1"], d["t1"], d["R2"], d["t2"], d["pos1"], d["desc1"],
                                          d["node1"], d["has_mp1"], d["pos2"], d["desc2"], d["node2"], d["has_mp2"])
    assert shim["nmatches"] == ref["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(ref["match12"], d["ref_match12"])
    np.testing.assert_array_equal(shim["match12"], ref["match12"])


# ---------------------------------------------------------------- Matcher::SearchByBoW, both overloads
@pytest.mark.parametrize("name", _bow_cases())
@pytest.mark.parametrize("kf_kf", [False, True], ids=["kf-frame", "kf-kf"])
def test_oracle_reproduces_the_reference_search_by_bow(name, kf_kf):
    """Matcher::SearchByBoW(KF, F) (Matcher.cpp:393-477) and (KF, KF) (:663-754) as the reference's own C++ ran them on a
    raw key frame / Frame with real MapPoint objects (good, bad, none): which feature got which map point, and the count."""
    from oracle import post_ref as O
    d = _tri_case(name)
    nm, res, _ = _bow_from_oracle(O.search_by_bow, d, float(d["ratio"][0]), kf_kf)
    assert nm == int(d["ref_nmatches"][int(kf_kf)]) and nm >= 15
    np.testing.assert_array_equal(res, d["ref_match12"] if kf_kf else d["ref_f2kf"])


def test_oracle_equals_reference_search_by_bow_live():
    """30 random feature-set pairs x 2 ratios x both overloads (1 to 40 nodes, up to 60 % of the features without and 30 %
    with a bad map point, empty and one-feature key frames) through the reference's own functions and the oracle."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    from ppg_slam_b200 import synth
    cam = cameras.EUROC
    total = 0
    for seed in range(30):
        kw = dict(n_nodes=[1, 4, 12, 40][seed % 4], frac_none=[0.0, 0.25, 0.6][seed % 3],
                  frac_bad=[0.0, 0.1, 0.3][(seed // 3) % 3], n1=[0, 1, 40, 300, 500][seed % 5] if seed < 10 else 300,
                  n2=[280, 3, 500][seed % 3])
        x = synth.bow_pair_inputs(seed, **kw)
        for ratio in (0.6, 0.9):
            a = R.search_by_bow_kf_f(cam, x["desc1"], x["node1"], x["state1"], x["desc2"], x["node2"], ratio)
            nm, res, _ = _bow_from_oracle(O.search_by_bow, x, ratio, False)
            assert nm == a["nmatches"], (seed, ratio)
            np.testing.assert_array_equal(res, a["f2kf"], err_msg="KF-F seed %d" % seed)
            b = R.search_by_bow_kf_kf(cam, x["desc1"], x["node1"], x["state1"], x["desc2"], x["node2"], x["state2"], ratio)
            nm, res, _ = _bow_from_oracle(O.search_by_bow, x, ratio, True)
            assert nm == b["nmatches"], (seed, ratio)
            np.testing.assert_array_equal(res, b["match12"], err_msg="KF-KF seed %d" % seed)
            total += a["nmatches"] + b["nmatches"]
    assert total > 3000


@pytest.mark.gpu
@pytest.mark.parametrize("name", _bow_cases())
@pytest.mark.parametrize("kf_kf", [False, True], ids=["kf-frame", "kf-kf"])
def test_cuda_search_by_bow_reproduces_the_reference(name, kf_kf):
    from ppg_slam_b200 import capi, synth
    d = _tri_case(name)
    rows, rd, rn = synth.bow_rows(d["desc1"], d["node1"], d["state1"])
    kp_node = np.where(d["state2"] == 1, d["node2"], -1).astype(np.int32) if kf_kf else d["node2"]
    e = capi.Extractor(cameras.EUROC, max_batch=1, max_map_points=1024)
    try:
        e.upload_map(rd)
        got = e.search_by_bow(rn, d["desc2"], kp_node, float(d["ratio"][0]), 0.7, kf_kf)
    finally:
        e.close()
    assert got["nmatches"] == int(d["ref_nmatches"][int(kf_kf)])
    np.testing.assert_array_equal(_bow_result(got["kp_row"], rows, len(d["desc1"]), kf_kf),
                                  d["ref_match12"] if kf_kf else d["ref_f2kf"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", _bow_cases())
@pytest.mark.parametrize("kf_kf", [False, True], ids=["kf-frame", "kf-kf"])
def test_shim_search_by_bow_equals_the_reference_on_real_objects(name, kf_kf):
    """ppg_shim::Matcher::SearchByBoW (both overloads) compiled against the reference's real headers and executed on raw
    key frames / a Frame with real MapPoint objects, next to ::Matcher."""
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    d = _tri_case(name)
    ref, shim = R.shim_bow_both(cameras.EUROC, kf_kf, d["desc1"], d["node1"], d["state1"], d["desc2"], d["node2"],
                                d["state2"], float(d["ratio"][0]))
    assert shim["nmatches"] == ref["nmatches"] == int(d["ref_nmatches"][int(kf_kf)])
    np.testing.assert_array_equal(ref["out"], d["ref_match12"] if kf_kf else d["ref_f2kf"])
    np.testing.assert_array_equal(shim["out"], ref["out"])
