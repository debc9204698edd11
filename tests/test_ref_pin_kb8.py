"""Parity pinned to the REFERENCE'S OWN C++ for the two functions whose camera model matters: Frame::CheckInFrustum
(map/src/Frame.cpp:223-260) with the real Pinhole / KannalaBrandt8 classes, and Matcher::SearchForTriangulation
(matching/src/Matcher.cpp:767-885) with the real KannalaBrandt8 camera -- epipolarConstrain = TriangulateMatches
(sensors/src/KannalaBrandt8.cpp:167-236), compiled unmodified (oracle/ref_build.py).

tests/golden/ref_l2_kb8.npz holds what that code produced (tests/golden/make_golden_ref_kb8.py); the oracle (CPU) and the
CUDA path (GPU) are compared with it, and -- where the harness is present -- the oracle is compared LIVE with the reference.

What is bit exact and what is not.  Everything but the float transcendentals: the oracle's literal-libm build (`_libmf`:
atan2f / tanf as the reference calls them) is bit identical to the reference on every number; the default build and the
CUDA kernels evaluate atan2f / tanf as the correctly rounded float of the double routine (divergence 2),
which moves projections by an ulp -- decisions (mbTrackInView, vMatches12, TriangulateMatches' return code) are equal.
The one step that is not the reference's arithmetic is Eigen::JacobiSVD inside KannalaBrandt8::Triangulate: the harness,
the oracle and the kernel all use the null vector of ppgo_null_vector4 (checked against numpy's SVD below).
"""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _z():
    return np.load(os.path.join(GOLD, "ref_l2_kb8.npz"))


def _names(prefix):
    return sorted({k.split("/")[0] for k in _z().files if k.startswith(prefix)})


def _case(name):
    z = _z()
    d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    d["cam"] = cameras.ALL[str(d["camera"])]
    return d


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---------------------------------------------------------------- the null vector that stands in for Eigen::JacobiSVD
def test_null_vector_equals_numpy_svd():
    """ppgo_null_vector4 = the right singular vector of the smallest singular value, to float accuracy, on random and on
    nearly rank-deficient matrices (the shape of a DLT matrix of a true correspondence)."""
    from oracle import post_ref as O
    rs = np.random.RandomState(0)
    for i in range(600):
        A = rs.randn(4, 4)
        if i % 2:
            x = rs.randn(4)
            x /= np.linalg.norm(x)
            A = A - np.outer(A @ x, x) + 1e-4 * rs.randn(4, 4)
        A = A.astype(np.float32)
        v = O.null_vector4(A).astype(np.float64)
        _, s, vt = np.linalg.svd(A.astype(np.float64))
        w = vt[3] if np.dot(v, vt[3]) >= 0 else -vt[3]
        gap = min((s[2] - s[3]) / s[0], 1.0)
        assert np.abs(v - w).max() * gap < 2e-7, i
    # exact cases: a diagonal matrix, and a matrix with a zero column
    np.testing.assert_array_equal(np.abs(O.null_vector4(np.diag([3.0, 2.0, 0.5, 4.0]))), [0, 0, 1, 0])
    A = rs.randn(4, 4).astype(np.float32)
    A[:, 1] = 0
    np.testing.assert_array_equal(np.abs(O.null_vector4(A)), [0, 1, 0, 0])


def test_kb8_unproject_inverts_project():
    from oracle import post_ref as O
    rs = np.random.RandomState(1)
    for cam in (cameras.TUMVI, cameras.UMA, cameras.TUMVI1024):
        for _ in range(200):
            p = np.array([rs.uniform(0, cam.width), rs.uniform(0, cam.height)], np.float32)
            if np.hypot((p[0] - cam.K[2]) / cam.K[0], (p[1] - cam.K[5]) / cam.K[4]) > 0.8:
                continue  # far field: the UMA-VI polynomial stops being monotonic near 1.15 rad and (x, y, 1) * tan(theta) ends at 90 degrees
            r = O.kb8_unproject(cam, p)
            assert r[2] == 1.0
            q = O.kb8_project(cam, r * np.float32(rs.uniform(0.5, 8)))
            assert np.abs(q - p).max() < 2e-3, (cam.name, p, q)
    c = cameras.TUMVI  # the principal point: theta_d == 0 -> scale stays 1 (KannalaBrandt8.cpp:70)
    np.testing.assert_array_equal(O.kb8_unproject(c, [c.K[2], c.K[5]]), [0, 0, 1])


# ---------------------------------------------------------------- KannalaBrandt8::TriangulateMatches, single pairs
@pytest.mark.parametrize("name", _names("kb8tri"))
def test_oracle_reproduces_the_reference_triangulate_matches(name):
    """Return code / depth, triangulated point and unprojected ray of the reference's own KannalaBrandt8 class on 48
    pairs per case: bit for bit with the literal-libm oracle; the default oracle takes the same decision on every pair
    and lands within 1e-4 relative of the point."""
    from oracle import post_ref as O
    d = _case(name)
    n_pos = 0
    for k, (i, j) in enumerate(d["pairs"]):
        want, wx = d["ref_pair_value"][k], d["ref_pair_x3D"][k]
        z, x = O.kb8_triangulate_matches(d["cam"], d["pos1"][i], d["pos2"][j], d["ref_R12"], d["ref_t12"], variant="_libmf")
        assert _bits(z) == _bits(want), (k, z, want)
        np.testing.assert_array_equal(_bits(O.kb8_unproject(d["cam"], d["pos1"][i], variant="_libmf")),
                                      _bits(d["ref_pair_r1"][k]))
        z2, x2 = O.kb8_triangulate_matches(d["cam"], d["pos1"][i], d["pos2"][j], d["ref_R12"], d["ref_t12"])
        if want > 0:
            n_pos += 1
            np.testing.assert_array_equal(_bits(x), _bits(wx))
            assert z2 > 0 and np.abs(x2 - wx).max() <= 1e-4 * np.abs(wx).max()
        else:
            assert z2 == want
    assert n_pos >= 6


# ---------------------------------------------------------------- Matcher::SearchForTriangulation, KannalaBrandt8
def _tri_args(d):
    return (d["cam"], d["desc1"], d["node1"], d["has_mp1"], d["pos1"], d["desc2"], d["node2"], d["has_mp2"], d["pos2"],
            d["ref_R12"], d["ref_t12"], d["ref_epipole"])


@pytest.mark.parametrize("variant", ["", "_libmf"], ids=["default", "literal-libm"])
@pytest.mark.parametrize("name", _names("kb8tri"))
def test_oracle_reproduces_the_reference_search_for_triangulation_kb8(name, variant):
    from oracle import post_ref as O
    d = _case(name)
    got = O.search_for_triangulation_kb8(*_tri_args(d), variant=variant)
    assert got["nmatches"] == int(d["ref_nmatches"][0]) and got["nmatches"] >= 5
    np.testing.assert_array_equal(got["match12"], d["ref_match12"])


def _kw(seed):
    return dict(n_nodes=[6, 12, 40][seed % 3], noise_px=[0.3, 0.6, 1.2][(seed // 3) % 3], forward=(seed % 4 == 3),
                frac_mp=[0.0, 0.2, 0.5][(seed // 2) % 3], n1=[0, 1, 57, 300, 500][seed % 5] if seed < 10 else 260)


def test_oracle_equals_reference_search_for_triangulation_kb8_live():
    """36 random key-frame pairs on the three fisheye calibrations (sideways and forward motion, 6 to 40 vocabulary
    nodes, empty and one-feature frames) through the reference's own function with its own KannalaBrandt8 camera (here)
    and through the oracle, both libm variants."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    total = 0
    for seed in range(36):
        cam = (cameras.TUMVI, cameras.UMA, cameras.TUMVI1024)[seed % 3]
        x = synth.two_view_inputs(200 + seed, cam, **_kw(seed))
        ref = R.search_for_triangulation(cam, x["R1"], x["t1"], x["R2"], x["t2"], x["pos1"], x["desc1"], x["node1"],
                                         x["has_mp1"], x["pos2"], x["desc2"], x["node2"], x["has_mp2"])
        for variant in ("", "_libmf"):
            got = O.search_for_triangulation_kb8(cam, x["desc1"], x["node1"], x["has_mp1"], x["pos1"], x["desc2"],
                                                 x["node2"], x["has_mp2"], x["pos2"], ref["R12"], ref["t12"],
                                                 ref["epipole"], variant=variant)
            assert got["nmatches"] == ref["nmatches"], (seed, variant)
            np.testing.assert_array_equal(got["match12"], ref["match12"], err_msg="seed %d %s" % (seed, variant))
        total += ref["nmatches"]
    assert total > 600


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("kb8tri"))
def test_cuda_search_for_triangulation_kb8_reproduces_the_reference(name):
    from ppg_slam_b200 import capi
    d = _case(name)
    cam = d["cam"]
    cam8 = [cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D)
    e = capi.Extractor(cameras.TUMVI, max_batch=1)
    try:
        got = e.search_for_triangulation(d["desc1"], d["node1"], d["has_mp1"], d["pos1"], d["desc2"], d["node2"],
                                         d["has_mp2"], d["pos2"], np.zeros(9), d["ref_epipole"],
                                         kb8=(cam8, d["ref_R12"], d["ref_t12"]))
    finally:
        e.close()
    assert got["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(got["match12"], d["ref_match12"])


def _relative_pose(x):
    """T12 = T1w * Tw2 and the second camera's view of the first camera's centre, in plain numpy (the GPU box has no
    reference tree)."""
    R1, t1, R2, t2 = (x[k].astype(np.float64) for k in ("R1", "t1", "R2", "t2"))
    R12, t12 = R1 @ R2.T, t1 - R1 @ R2.T @ t2
    C2 = R2 @ (-R1.T @ t1) + t2
    return R12.astype(np.float32), t12.astype(np.float32), C2.astype(np.float32)


@pytest.mark.gpu
def test_cuda_search_for_triangulation_kb8_equals_oracle_sweep():
    """24 random key-frame pairs up to 500 x 520 features on the three fisheye calibrations: ppg_search_for_triangulation
    (camera_model 1) against the oracle -- every match index."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    e = capi.Extractor(cameras.TUMVI, max_batch=1)
    total = 0
    try:
        for seed in range(24):
            cam = (cameras.TUMVI, cameras.UMA, cameras.TUMVI1024)[seed % 3]
            kw = _kw(seed)
            kw.update(n1=[0, 1, 57, 300, 500][seed % 5], n2=[320, 2, 33, 520][seed % 4])
            x = synth.two_view_inputs(300 + seed, cam, **kw)
            R12, t12, C2 = _relative_pose(x)
            ep = O.kb8_project(cam, C2)
            cam8 = [cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D)
            want = O.search_for_triangulation_kb8(cam, x["desc1"], x["node1"], x["has_mp1"], x["pos1"], x["desc2"],
                                                  x["node2"], x["has_mp2"], x["pos2"], R12, t12, ep)
            got = e.search_for_triangulation(x["desc1"], x["node1"], x["has_mp1"], x["pos1"], x["desc2"], x["node2"],
                                             x["has_mp2"], x["pos2"], np.zeros(9), ep, kb8=(cam8, R12, t12))
            assert got["nmatches"] == want["nmatches"], seed
            np.testing.assert_array_equal(got["match12"], want["match12"], err_msg="seed %d" % seed)
            total += want["nmatches"]
    finally:
        e.close()
    assert total > 150


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("kb8tri"))
def test_shim_search_for_triangulation_kb8_equals_the_reference_on_real_keyframes(name):
    """include/ppg_shim.hpp compiled against the reference's real KeyFrame.h / KannalaBrandt8.h / SE3.h and executed: the
    shim derives R12 / t12 and the epipole from the key frames' poses with the reference's own classes, flattens the
    FeatureVectors, calls the GPU and must hand back the vMatchedPairs of the reference's host function."""
    from oracle import ref_harness as R
    if not R.shim_available():
        pytest.skip("shim harness not built (built in the build container: oracle/ref_build.py)")
    d = _case(name)
    ref, shim = R.shim_triangulation_both(d["cam"], d["R1"], d["t1"], d["R2"], d["t2"], d["pos1"], d["desc1"], d["node1"],
                                          d["has_mp1"], d["pos2"], d["desc2"], d["node2"], d["has_mp2"])
    assert shim["nmatches"] == ref["nmatches"] == int(d["ref_nmatches"][0])
    np.testing.assert_array_equal(ref["match12"], d["ref_match12"])
    np.testing.assert_array_equal(shim["match12"], ref["match12"])


# ---------------------------------------------------------------- Frame::CheckInFrustum
def _check_frustum(got, d, exact_uv):
    np.testing.assert_array_equal(got["in_view"], d["ref_in_view"])
    iv = d["ref_in_view"].astype(bool)
    assert 0.2 * len(iv) < iv.sum() < 0.9 * len(iv)
    np.testing.assert_array_equal(d["ref_visible"], d["ref_in_view"])  # IncreaseVisible exactly for the rows in view (:259)
    for k in ("depth", "view_cos"):
        np.testing.assert_array_equal(_bits(got[k][iv]), _bits(d["ref_" + k][iv]), err_msg=k)
    if exact_uv:
        np.testing.assert_array_equal(_bits(got["proj_uv"][iv]), _bits(d["ref_proj_uv"][iv]))
    else:  # atan2f rounded from the double routine: an ulp of theta / psi
        assert np.abs(got["proj_uv"][iv] - d["ref_proj_uv"][iv]).max() < 2e-4
    # rows out of view keep the reset values of :225-228
    assert (got["proj_uv"][~iv] == -1).all() and (got["depth"][~iv] == -1).all()


@pytest.mark.parametrize("name", _names("fru"))
def test_oracle_reproduces_the_reference_check_in_frustum(name):
    """mbTrackInView, mTrackProjX / Y, mTrackDepth, mTrackViewCos as the reference's own Frame::CheckInFrustum left them on
    real MapPoint objects, with the reference's own Pinhole::project / KannalaBrandt8::project and IsInImage."""
    from oracle import post_ref as O
    d = _case(name)
    args = (d["cam"], d["Rcw"][0], d["tcw"][0], d["Ow"][0], d["world_pos"], d["normal"], d["min_dist"], d["max_dist"], 0.5)
    _check_frustum(O.check_in_frustum(*args, variant="_libmf"), d, True)
    _check_frustum(O.check_in_frustum(*args), d, not d["cam"].fisheye)


def test_oracle_equals_reference_check_in_frustum_live():
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    for cam in (cameras.EUROC, cameras.TUMVI, cameras.UMA, cameras.TUMVI1024):
        for seed in range(4):
            g = synth.frustum_inputs(50 + seed, cam, 3000, n_frames=2)
            for f in range(2):
                args = (cam, g["Rcw"][f], g["tcw"][f], g["Ow"][f], g["world_pos"], g["normal"], g["min_dist"],
                        g["max_dist"], [0.5, 0.3][f])
                ref = R.check_in_frustum(*args)
                d = {"ref_" + k: v for k, v in ref.items()}
                _check_frustum(O.check_in_frustum(*args, variant="_libmf"), d, True)
                _check_frustum(O.check_in_frustum(*args), d, not cam.fisheye)


@pytest.mark.gpu
@pytest.mark.parametrize("name", _names("fru"))
def test_cuda_check_in_frustum_reproduces_the_reference(name):
    from ppg_slam_b200 import capi
    d = _case(name)
    cam, M = d["cam"], len(d["min_dist"])
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    try:
        e.upload_map(np.zeros((M, 256), np.float32))
        e.upload_map_geometry(d["world_pos"], d["normal"], d["min_dist"], d["max_dist"])
        e.assoc_stage_poses(d["Rcw"], d["tcw"], d["Ow"], M, 0.5, 10.0, 0.8)
        got = e.frustum_fetch(1)
    finally:
        e.close()
    _check_frustum({k: got[k][0] for k in ("in_view", "proj_uv", "depth", "view_cos")}, d, not cam.fisheye)
