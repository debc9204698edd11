"""Pins the OpenCV routines restated in oracle/ppg_oracle.c bit-for-bit against cv2 4.13 outputs
(fixtures from tests/golden/make_golden.py; call sites PPGExtractor.cpp:66,69,221,223,262)."""
import os

import numpy as np
import pytest

from oracle import post_ref as O
from ppg_slam_b200 import cameras


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "cv_kat.npz"))


@pytest.mark.parametrize("cam", [cameras.EUROC, cameras.TUMVI, cameras.UMA], ids=lambda c: c.name)
def test_undistort_points_bit_exact(kat, cam):
    pts, und = kat[cam.name + "_pts"], kat[cam.name + "_und"]
    got = O.undistort_points(cam, pts)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got.view(np.uint32), und.view(np.uint32))
    if cam.fisheye:  # the sentinel for non-converged points must be hit by the fixture
        assert (und[:, 0] == -1000000.0).any()


def test_init_undistort_map_pinhole(kat):
    mx, my = O.init_undistort_map(cameras.EUROC)
    rows = kat["euroc_map_rows"]
    np.testing.assert_array_equal(mx[rows].view(np.uint32), kat["euroc_mx_rows"].view(np.uint32))
    np.testing.assert_array_equal(my[rows].view(np.uint32), kat["euroc_my_rows"].view(np.uint32))
    s = kat["euroc_map_sum"]
    assert mx.astype(np.float64).sum() == s[0] and my.astype(np.float64).sum() == s[1]


def test_init_undistort_map_fisheye(kat):
    mx, my = O.init_undistort_map(cameras.TUMVI)
    rows = kat["euroc_map_rows"]
    # cv::fisheye::initUndistortRectifyMap goes through float intermediates for CV_32F maps; it is unused
    # by every shipped config (D[0]==0 => no remap), so a tolerance pin is enough here.
    assert np.abs(mx[rows] - kat["tumvi_mx_rows"]).max() < 2e-3
    assert np.abs(my[rows] - kat["tumvi_my_rows"]).max() < 2e-3


def test_remap_bit_exact(kat):
    cam = cameras.EUROC
    mx, my = O.init_undistort_map(cam)
    vv, uu = np.mgrid[0:cam.height, 0:cam.width]
    src = (((uu * 7 + vv * 13) % 251).astype(np.float32) / np.float32(250))
    dst = O.remap_linear(src, mx, my)
    rows = kat["euroc_map_rows"]
    np.testing.assert_array_equal(dst[rows].view(np.uint32), kat["remap_dst_rows"].view(np.uint32))
    assert dst.astype(np.float64).sum() == kat["remap_dst_sum"][0]


def test_live_cv2_if_present():
    cv2 = pytest.importorskip("cv2")
    cam = cameras.EUROC
    K = np.array(cam.K, np.float32).reshape(3, 3)
    D = np.array(cam.D, np.float32).reshape(4, 1)
    vv, uu = np.mgrid[0:cam.height, 0:cam.width]
    pts = np.stack([uu.ravel(), vv.ravel()], 1).astype(np.float32)[::7]
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, None, K).reshape(-1, 2)
    np.testing.assert_array_equal(O.undistort_points(cam, pts).view(np.uint32), ref.view(np.uint32))
    mx, my = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (cam.width, cam.height), cv2.CV_32F)
    omx, omy = O.init_undistort_map(cam)
    np.testing.assert_array_equal(omx.view(np.uint32), mx.view(np.uint32))
    np.testing.assert_array_equal(omy.view(np.uint32), my.view(np.uint32))
    rs = np.random.RandomState(5)
    src = rs.rand(cam.height, cam.width).astype(np.float32)
    np.testing.assert_array_equal(O.remap_linear(src, mx, my).view(np.uint32),
                                  cv2.remap(src, mx, my, cv2.INTER_LINEAR).view(np.uint32))
