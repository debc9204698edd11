"""GPU parity of the image<->map association (search core of Matcher::ExtendMapMatches,
matching/src/Matcher.cpp:224-281) against the L2 oracle: best / second-best indices, distances (bit-exact:
both sides use the same fixed summation order) and the accept flag of the ratio test."""
import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

pytestmark = pytest.mark.gpu


def _unit(a):
    return (a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-12)).astype(np.float32)


def _synthetic_keypoints(seed, cam, n):
    rs = np.random.RandomState(seed)
    gx, gy = np.meshgrid(np.arange(8, cam.width - 8, 6), np.arange(8, cam.height - 8, 6))
    sel = rs.choice(gx.size, n, replace=False)
    kx = gx.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    ky = gy.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    return kx, ky, _unit(rs.normal(size=(n, 256)))


def _check(e, cam, kx, ky, fdesc, free, inp, th, ratio):
    from oracle import post_ref as O
    got = e.associate(kx, ky, fdesc, free, inp["proj_uv"], inp["view_cos"], th, ratio)
    ref = O.search_all(cam, kx, ky, fdesc, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], th, ratio)
    np.testing.assert_array_equal(got["best_idx"], ref["best_idx"])
    np.testing.assert_array_equal(got["second_idx"], ref["second_idx"])
    np.testing.assert_array_equal(got["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
    np.testing.assert_array_equal(got["second_d"].view(np.uint32), ref["second_d"].view(np.uint32))
    np.testing.assert_array_equal(got["accept"], ref["accept"])
    return got, ref


@pytest.mark.parametrize("cam,n,m,th", [(cameras.UMA, 1000, 50000, 10.0), (cameras.EUROC, 357, 8192, 10.0),
                                        (cameras.TUMVI, 500, 3000, 15.0), (cameras.EUROC, 37, 300, 3.0)],
                         ids=["uma-1000x50k", "euroc-357x8k", "tumvi-500x3k", "euroc-37x300"])
def test_assoc_matches_oracle(cam, n, m, th):
    from ppg_slam_b200 import capi
    kx, ky, fdesc = _synthetic_keypoints(11, cam, n)
    inp = synth.association_inputs(5, fdesc, np.stack([kx, ky], 1), m, cam.width, cam.height, th=th)
    rs = np.random.RandomState(3)
    free = (rs.rand(n) > 0.1).astype(np.uint8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=max(m, 1024))
    try:
        e.upload_map(inp["map_desc"])
        got, ref = _check(e, cam, kx, ky, fdesc, free, inp, th, 0.8)
        assert ref["accept"].sum() > 0.2 * m / 2  # planted matches are found
        fb = e.assoc_fallback_rows()
        assert fb <= 0.25 * m, "too many rows (%d of %d) fell back to the exact window scan" % (fb, m)
    finally:
        e.close()


def test_assoc_edge_cases():
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    kx, ky, fdesc = _synthetic_keypoints(2, cam, 200)
    # exact duplicates: identical descriptors on neighbouring keypoints -> tie broken by grid order
    fdesc[10:20] = fdesc[10]
    kx[10:20] = 300 + np.arange(10, dtype=np.float32) * 1.5
    ky[10:20] = 200 + (np.arange(10, dtype=np.float32) % 3) * 2.0
    # keypoints outside the undistorted image bounds are not indexable (Frame.cpp:317-327)
    kx[0], ky[0] = -80.0, 100.0
    kx[1], ky[1] = 900.0, 600.0
    m = 600
    inp = synth.association_inputs(9, fdesc, np.stack([kx, ky], 1), m, cam.width, cam.height, th=6.0)
    inp["map_desc"][:40] = fdesc[10]          # rows that tie exactly on the duplicated keypoints
    inp["proj_uv"][:40] = [306.0, 202.0]
    inp["proj_uv"][40:60] = [-500.0, -500.0]  # windows entirely outside the grid (early returns :270-292)
    inp["proj_uv"][60:70] = [5000.0, 100.0]
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    try:
        e.upload_map(inp["map_desc"])
        for free in (np.ones(200, np.uint8), np.zeros(200, np.uint8), (np.arange(200) % 2).astype(np.uint8)):
            for ratio in (0.6, 0.9):
                got, ref = _check(e, cam, kx, ky, fdesc, free, inp, 6.0, ratio)
        assert (got["second_idx"] == -1).any()
    finally:
        e.close()


@pytest.mark.parametrize("name,th,max_dist,e2_max", [("by_projection", 15.0, 0.8, 0.0), ("by_projection_desc", 10.0, 0.75, 0.0),
                                                      ("fuse", 3.0, 0.7, 5.99)], ids=lambda v: str(v))
def test_window_search_modes_match_oracle(name, th, max_dist, e2_max):
    """PPG_SEARCH_WINDOW: the best-only cores of SearchByProjection (Matcher.cpp:31-87, :1337-1411) and Fuse
    (:897-1036, incl. the e2 > 5.99 skip) -- same kernels, r = th, accept = best <= threshold."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    n, m = 420, 6000
    kx, ky, fdesc = _synthetic_keypoints(21, cam, n)
    inp = synth.association_inputs(8, fdesc, np.stack([kx, ky], 1), m, cam.width, cam.height, th=th)
    rs = np.random.RandomState(4)
    k = min(n, 300)
    inp["proj_uv"][:k] = np.stack([kx[:k], ky[:k]], 1) + rs.uniform(-2.0, 2.0, (k, 2)).astype(np.float32)
    free = (rs.rand(n) > 0.15).astype(np.uint8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=8192)
    try:
        e.upload_map(inp["map_desc"])
        got = e.associate(kx, ky, fdesc, free, inp["proj_uv"], inp["view_cos"], th, 0.8, mode=capi.SEARCH_WINDOW,
                          max_dist=max_dist, e2_max=e2_max)
        ref = O.search_all(cam, kx, ky, fdesc, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], th, 0.8,
                           mode=1, max_dist=max_dist, e2_max=e2_max)
        np.testing.assert_array_equal(got["best_idx"], ref["best_idx"])
        np.testing.assert_array_equal(got["second_idx"], ref["second_idx"])
        np.testing.assert_array_equal(got["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
        np.testing.assert_array_equal(got["accept"], ref["accept"])
        assert ref["accept"].sum() > 20 and (ref["best_idx"] < 0).sum() > 0
        # the ExtendMapMatches rule is back after a mode-0 call on the same ctx
        _check(e, cam, kx, ky, fdesc, free, inp, th, 0.8)
    finally:
        e.close()


def test_assoc_stress_huge_windows_and_global_ties():
    """Windows that cover the whole image (every keypoint is a candidate of every row: the hit queues drain on
    every chunk and most rows need the exact window scan), and a frame whose descriptors are all identical (every
    distance ties: the reference keeps the first keypoint in GetFeaturesInArea order)."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    n, m = 357, 1500
    kx, ky, fdesc = _synthetic_keypoints(17, cam, n)
    inp = synth.association_inputs(13, fdesc, np.stack([kx, ky], 1), m, cam.width, cam.height, th=10.0)
    free = np.ones(n, np.uint8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=2048)
    try:
        e.upload_map(inp["map_desc"])
        _check(e, cam, kx, ky, fdesc, free, inp, 250.0, 0.8)       # r = 625 / 1000 px
        same = np.repeat(fdesc[:1], n, 0)
        got, ref = _check(e, cam, kx, ky, same, free, inp, 10.0, 0.8)
        assert (got["best_idx"] >= 0).sum() > 100
        _check(e, cam, kx, ky, same, free, inp, 250.0, 0.9)
    finally:
        e.close()


def test_assoc_on_extracted_frame():
    """extract -> associate with the frame's own keypoints/descriptors left on the device."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    e = capi.Extractor(cam, max_batch=2, max_map_points=8192)
    try:
        recs = e.run([synth.frame(0, cam.width, cam.height), synth.frame(1, cam.width, cam.height)])
        r = recs[1]
        kp = np.stack([r["kp_x"], r["kp_y"]], 1)
        inp = synth.association_inputs(4, r["desc"], kp, 8192, cam.width, cam.height, th=10.0)
        e.upload_map(inp["map_desc"])
        n = r["n_kp"]
        e.assoc_stage(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 256), np.float32),
                      np.zeros(0, np.uint8), inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        e.assoc_run_frame(1)
        got = e.assoc_fetch()
        ref = O.search_all(cam, r["kp_x"], r["kp_y"], r["desc"], np.ones(n, np.uint8), inp["map_desc"],
                           inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        np.testing.assert_array_equal(got["best_idx"], ref["best_idx"])
        np.testing.assert_array_equal(got["second_idx"], ref["second_idx"])
        np.testing.assert_array_equal(got["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
        np.testing.assert_array_equal(got["accept"], ref["accept"])
    finally:
        e.close()


def test_sharded_association_nccl_two_gpus():
    """Row-sharded table over 2 GPUs + NCCL all-gather of the top-2 records == un-sharded oracle."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "sharded_assoc_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"parity_vs_oracle": true' in r.stdout


def test_allgather_of_records_single_rank_nccl():
    """The library's own exchange step (csrc/comm.cu: pack kernel + ncclAllGather on the ctx stream) on ONE GPU: a
    world of one rank, the records gathered equal the per-row results fetched the ordinary way, padding rows are empty."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    rs = np.random.RandomState(21)
    n, m = 300, 3000
    kx = rs.uniform(8, cam.width - 8, n).astype(np.float32)
    ky = rs.uniform(8, cam.height - 8, n).astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    inp = synth.association_inputs(6, fd, np.stack([kx, ky], 1), m, cam.width, cam.height)
    e = capi.Extractor(cam, max_batch=1, max_map_points=4096)
    try:
        e.comm_init(capi.comm_unique_id(), 0, 1)
        e.upload_map(inp["map_desc"])
        e.assoc_stage(kx, ky, fd, np.ones(n, np.uint8), inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        e.assoc_run()
        e.assoc_allgather(m, m + 7)  # common send count larger than the shard
        rec, us = e.assoc_allgather_fetch()
        want = e.assoc_fetch()
        assert rec.shape == (1, m + 7, 5) and us >= 0
        np.testing.assert_array_equal(rec[0, :m, 0], want["best_idx"])
        np.testing.assert_array_equal(rec[0, :m, 1], want["second_idx"])
        np.testing.assert_array_equal(rec[0, :m, 2], want["best_d"].view(np.int32))
        np.testing.assert_array_equal(rec[0, :m, 3], want["second_d"].view(np.int32))
        np.testing.assert_array_equal(rec[0, :m, 4].astype(np.uint8), want["accept"])
        assert (rec[0, m:, 0] == -1).all() and (rec[0, m:, 4] == 0).all()
        assert want["accept"].sum() > 50
        e.comm_destroy()
    finally:
        e.close()


def test_assoc_batch_equals_oracle_per_frame():
    """Throughput form: all frames of an extraction batch against the resident table in one set of launches."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    Bn, M = 4, 4096
    e = capi.Extractor(cam, max_batch=Bn, max_map_points=M)
    try:
        recs = e.run([synth.frame(s, cam.width, cam.height) for s in range(Bn)])
        kp0 = np.stack([recs[0]["kp_x"], recs[0]["kp_y"]], 1)
        base = synth.association_inputs(4, recs[0]["desc"], kp0, M, cam.width, cam.height, th=10.0)
        e.upload_map(base["map_desc"])
        rs = np.random.RandomState(0)
        uv = np.stack([base["proj_uv"] + rs.uniform(-3, 3, base["proj_uv"].shape).astype(np.float32) for _ in range(Bn)])
        vc = np.stack([base["view_cos"]] * Bn)
        e.assoc_stage_batch(uv, vc, 10.0, 0.8)
        e.assoc_run_batch(Bn)
        got = e.assoc_fetch_batch(Bn)
        for f in range(Bn):
            r = recs[f]
            ref = O.search_all(cam, r["kp_x"], r["kp_y"], r["desc"], np.ones(r["n_kp"], np.uint8), base["map_desc"],
                               uv[f], vc[f], 10.0, 0.8)
            np.testing.assert_array_equal(got[f]["best_idx"], ref["best_idx"])
            np.testing.assert_array_equal(got[f]["second_idx"], ref["second_idx"])
            np.testing.assert_array_equal(got[f]["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
            np.testing.assert_array_equal(got[f]["accept"], ref["accept"])
        assert got[0]["accept"].sum() > 100
    finally:
        e.close()


def test_distinctive_descriptors_match_oracle():
    """ppg_distinctive_descriptors / ppg_upload_map_distinctive == MapPoint::ComputeDistinctiveDescriptors
    (MapPoint.cpp:234-302) restated by the oracle, and the table built on the device gives the same association
    as uploading the chosen rows from the host."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    rs = np.random.RandomState(31)
    P = 700
    sizes = rs.randint(1, 40, P)
    sizes[:6] = [1, 2, 3, 128, 64, 127]
    centre = _unit(rs.normal(size=(P, 256)))
    desc = np.concatenate([_unit(centre[k] + rs.normal(size=(n, 256)) * rs.choice([0.05, 0.2, 2.0]))
                           for k, n in enumerate(sizes)])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    want = O.distinctive_all(desc, off)
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    try:
        got = e.distinctive_descriptors(desc, off)
        np.testing.assert_array_equal(got, want)
        assert len(set(got.tolist())) > 5
        # device-built table == host upload of the same rows
        n = 300
        kx, ky, fdesc = _synthetic_keypoints(3, cam, n)
        table = desc[off[:-1] + want]
        inp = synth.association_inputs(2, fdesc, np.stack([kx, ky], 1), P, cam.width, cam.height, th=10.0)
        free = np.ones(n, np.uint8)
        e.upload_map(table)
        a = e.associate(kx, ky, fdesc, free, inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        got2 = e.distinctive_descriptors(desc, off, to_table=True)
        np.testing.assert_array_equal(got2, want)
        b = e.associate(kx, ky, fdesc, free, inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        for k in ("best_idx", "second_idx", "accept"):
            np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_array_equal(a["best_d"].view(np.uint32), b["best_d"].view(np.uint32))
        # more than PPG_MAX_OBSERVATIONS observations -> capacity error, no crash
        big = np.concatenate([[0], [129]]).astype(np.int32)
        with pytest.raises(capi.PpgError):
            e.distinctive_descriptors(_unit(rs.normal(size=(129, 256))), big)
    finally:
        e.close()
