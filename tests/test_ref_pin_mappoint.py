"""MapPoint::ComputeDistinctiveDescriptors (feature/src/MapPoint.cpp:234-302) pinned to the reference's own C++: a real
MapPoint observed by raw key frames, the real function (oracle/ref_build.py), the descriptor it chose -- committed in
tests/golden/ref_l2_mappoint.npz (tests/golden/make_golden_ref_mappoint.py).  The oracle (CPU) and
ppg_distinctive_descriptors (GPU) return the index of the same observation."""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(GOLD, "ref_l2_mappoint.npz"))
    names = sorted({k.split("/")[0] for k in z.files}, key=lambda s: int(s[2:]))
    return [(n, z[n + "/obs_desc"], z[n + "/ref_descriptor"]) for n in names]


def _packed():
    cs = _cases()
    desc = np.concatenate([c[1] for c in cs])
    off = np.concatenate([[0], np.cumsum([len(c[1]) for c in cs])]).astype(np.int32)
    return cs, desc, off


def _check(best, cs):
    for (name, d, ref), b in zip(cs, best):
        np.testing.assert_array_equal(d[int(b)].view(np.uint32), ref.view(np.uint32), err_msg=name)


def test_oracle_reproduces_the_reference_distinctive_descriptor():
    from oracle import post_ref as O
    cs, desc, off = _packed()
    assert len(cs) >= 8
    _check(O.distinctive_all(desc, off), cs)


def test_oracle_equals_reference_distinctive_descriptor_live():
    """300 map points with 1 - 128 observations (tight and loose clusters, duplicate observations with equal medians) and
    the early returns of :241-262 (bad point, bad key frames, index -1) through the reference's own function."""
    from oracle import post_ref as O, ref_harness as R
    if not R.matcher_available():
        pytest.skip("reference harness not built (no /root/reference on this machine)")
    rs = np.random.RandomState(0)
    for trial in range(300):
        n = [1, 2, 3, 4, 5, 8, 13, 40, 128][trial % 9]
        d = rs.randn(256) + rs.randn(n, 256) * rs.choice([0.05, 0.3, 1.0])
        if trial % 7 == 0 and n > 2:
            d[1] = d[0]
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        ref = R.distinctive_descriptor(d)
        best = int(O.distinctive_all(d, np.array([0, n], np.int32))[0])
        np.testing.assert_array_equal(d[best].view(np.uint32), ref.view(np.uint32), err_msg="trial %d" % trial)
    d = rs.randn(4, 256).astype(np.float32)
    assert (R.distinctive_descriptor(d, point_bad=True) == -1).all()            # :242-243
    assert (R.distinctive_descriptor(d, state=[1, 1, 1, 1]) == -1).all()        # every key frame bad: :262
    # observations that are skipped (:253-258) are not candidates: the caller of the C ABI packs the rest
    ref = R.distinctive_descriptor(d, state=[2, 0, 1, 0])
    keep = d[[1, 3]]
    best = int(O.distinctive_all(keep, np.array([0, 2], np.int32))[0])
    np.testing.assert_array_equal(keep[best].view(np.uint32), ref.view(np.uint32))


@pytest.mark.gpu
def test_cuda_distinctive_descriptors_reproduce_the_reference():
    from ppg_slam_b200 import capi
    cs, desc, off = _packed()
    e = capi.Extractor(cameras.EUROC, max_batch=1, max_map_points=1024)
    try:
        best = e.distinctive_descriptors(desc, off)
    finally:
        e.close()
    _check(best, cs)
