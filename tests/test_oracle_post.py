"""Known-answer tests of the L1/L2 oracle (oracle/ppg_oracle.c) against small independent restatements of the
reference statements (pure Python / numpy / torch), on the edge cases SURVEY.md s.8c lists."""
import numpy as np
import pytest

from oracle import post_ref as O
from ppg_slam_b200 import cameras, synth

CAM = cameras.Camera("kat", 96, 64, 60.0, 60.0, 47.5, 31.5, (-0.05, 0.01, 0.0005, -0.0003), False)


def py_detect(prob, W, H, R=4, thr=1 / 128, cap=500):
    """feature/src/PPGExtractor.cpp:168-206 restated literally in Python (stable tie order = raster)."""
    cand = [(x, y, prob[y, x]) for y in range(H) for x in range(W) if not (prob[y, x] < thr)]
    cand.sort(key=lambda t: -t[2])  # Python's sort is stable -> ties stay in raster order
    flag = np.zeros((H + 1, W + 1), np.uint8)
    kps = []
    for x, y, s in cand:
        if x < R or x > W - R - 1 or y < R or y > H - R - 1 or flag[y, x] != 0:
            continue
        flag[y, x] = 1
        kps.append((x, y, s))
        if len(kps) + 1 > cap:
            break
        for i in range(y - R, y + R + 1):
            for j in range(x - R, x + R + 1):
                if i < 0 or i > H or j < 0 or j > W:
                    continue
                flag[i, j] = 255
    return kps


@pytest.mark.parametrize("case", ["random", "ties", "border", "cap"])
def test_detect_keypoints_kat(case):
    W, H = CAM.width, CAM.height
    rs = np.random.RandomState(3)
    if case == "random":
        prob = (rs.rand(H, W) ** 6).astype(np.float32)
        cap = 500
    elif case == "ties":
        prob = np.zeros((H, W), np.float32)
        prob[8:56:2, 8:88:2] = 0.5
        cap = 500
    elif case == "border":
        prob = np.zeros((H, W), np.float32)
        prob[3, 10] = 0.9   # rejected (y < R) and must NOT suppress its neighbour
        prob[5, 11] = 0.5
        prob[30, W - 4] = 0.9  # x > W-R-1 rejected
        prob[30, W - 5] = 0.4  # accepted
        prob[H - 5, 40] = 0.3
        cap = 500
    else:
        prob = (rs.rand(H, W) ** 2).astype(np.float32)
        cap = 7
    cfg = O.make_cfg(CAM, junction_max_num=cap)
    got = O.detect_keypoints(cfg, prob)
    want = py_detect(prob, W, H, cap=cap)
    assert got["n"] == len(want)
    assert list(zip(got["x"].tolist(), got["y"].tolist())) == [(x, y) for x, y, _ in want]
    np.testing.assert_array_equal(got["score"], np.array([s for _, _, s in want], np.float32))
    und = O.undistort_points(CAM, np.stack([got["x"], got["y"]], 1).astype(np.float32))
    np.testing.assert_array_equal(got["xun"], und[:, 0])
    inb = (und[:, 0] >= 1) & (und[:, 0] < W - 1) & (und[:, 1] >= 1) & (und[:, 1] < H - 1)
    np.testing.assert_array_equal(got["out"], (~inb).astype(np.uint8))
    if case == "border":
        assert (11, 5) in [(x, y) for x, y, _ in want] and (10, 3) not in [(x, y) for x, y, _ in want]


def py_refine_tile(t, thr=0.01, ratio=0.3):
    """feature/src/PPGExtractor.cpp:540-578 on one tile (numpy)."""
    t = t.copy()
    v = t[t > np.float32(thr)]  # raster order
    n = v.size
    val_count = int(np.float32(ratio) * np.float32(n))
    if val_count < 1:
        return t
    if n >= t.size * 0.9 and float(v[int(n * 0.9)]) > 0.1:
        return np.zeros_like(t)
    top = np.sort(v)[::-1][:val_count]
    ave = np.float32(top.astype(np.float64).sum() / np.float64(np.float32(val_count)))
    out = np.zeros_like(t)
    m = t > np.float32(thr)
    ns = (t[m] / ave).astype(np.float32)
    out[m] = np.where(ns.astype(np.float64) > 1.0, np.float32(1.0), ns)
    return out


def test_refine_heat_kat():
    W, H = CAM.width, CAM.height
    rs = np.random.RandomState(5)
    heat = np.zeros((H, W), np.float32)
    heat[0:16, 0:16] = (rs.rand(16, 16) * 0.9).astype(np.float32)          # ordinary tile
    heat[0:16, 16:32] = 0.005                                               # nothing valid
    heat[0:16, 16:18] = 0.5                                                 # n = 32 -> valCount 9
    heat[0, 32:35] = 0.7                                                    # n = 3 -> valCount 0: tile UNCHANGED
    heat[1, 40] = 0.004                                                     # ... including values <= thr
    heat[16:32, 0:16] = (0.2 + 0.7 * rs.rand(16, 16)).astype(np.float32)    # n = 256 >= 231 and v[230] > 0.1 -> zero
    heat[16:32, 16:32] = (0.02 + 0.05 * rs.rand(16, 16)).astype(np.float32)  # n = 256 but v[230] <= 0.1 -> rescale
    heat[32:48, 0:16] = (rs.rand(16, 16) > 0.5).astype(np.float32)          # saturating values
    got = O.refine_heat(O.make_cfg(CAM), heat)
    want = heat.copy()
    for ty in range(H // 16):
        for tx in range(W // 16):
            want[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16] = py_refine_tile(heat[ty * 16:(ty + 1) * 16,
                                                                                tx * 16:(tx + 1) * 16])
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got[0:16, 32:48], heat[0:16, 32:48])  # unchanged tile
    assert not got[16:32, 0:16].any()                            # zero-filled tile
    assert got[16:32, 16:32].max() == 1.0


def test_nan_score_lines_are_accepted():
    """5 <= dist < ~6 => segNum = 1 => 0/0 = NaN inlier rate; NaN < 0.8 is false => accepted (:376)."""
    cam = cameras.Camera("flat", 96, 64, 60.0, 60.0, 47.5, 31.5, (0.0, 0.0, 0.0, 0.0), False)
    cfg = O.make_cfg(cam)
    heat = np.ones((64, 96), np.float32)
    kp = dict(n=3, xun=np.array([20.0, 25.5, 60.0], np.float32), yun=np.array([20.0, 20.0, 40.0], np.float32),
              out=np.zeros(3, np.uint8))
    r = O.detect_lines(cfg, heat, kp)
    edges = list(zip(r["edge_start"].tolist(), r["edge_end"].tolist()))
    assert (0, 1) in edges
    assert np.isnan(r["edge_score"][edges.index((0, 1))])
    assert (0, 2) in edges and np.isfinite(r["edge_score"][edges.index((0, 2))])
    # adjacency is consistent with the edge list
    for p in range(3):
        for e in r["conn_idx"][r["conn_off"][p]:r["conn_off"][p + 1]]:
            assert p in edges[e]


def test_out_points_make_no_lines():
    cam = cameras.Camera("flat", 96, 64, 60.0, 60.0, 47.5, 31.5, (0.0, 0.0, 0.0, 0.0), False)
    heat = np.ones((64, 96), np.float32)
    kp = dict(n=3, xun=np.array([20.0, 40.0, 60.0], np.float32), yun=np.array([20.0, 20.0, 40.0], np.float32),
              out=np.array([0, 1, 0], np.uint8))
    r = O.detect_lines(O.make_cfg(cam), heat, kp)
    assert list(zip(r["edge_start"].tolist(), r["edge_end"].tolist())) == [(0, 2)]


def test_overlap_filter_keeps_shorter_of_two_nearly_parallel_lines():
    """Two candidates from the same point within 0.2*pi and < 2 px apart: the longer one is dropped (:331-334)."""
    cam = cameras.Camera("flat", 96, 64, 60.0, 60.0, 47.5, 31.5, (0.0, 0.0, 0.0, 0.0), False)
    heat = np.ones((64, 96), np.float32)
    kp = dict(n=3, xun=np.array([10.0, 40.0, 70.0], np.float32), yun=np.array([30.0, 30.0, 30.5], np.float32),
              out=np.zeros(3, np.uint8))
    r = O.detect_lines(O.make_cfg(cam), heat, kp)
    edges = list(zip(r["edge_start"].tolist(), r["edge_end"].tolist()))
    assert (0, 1) in edges and (0, 2) not in edges and (1, 2) in edges
    # the three points are colinear: point 1 gets the coline pair (0, 2) or (2, 0)
    pairs = r["col_pairs"][r["col_off"][1]:r["col_off"][2]].tolist()
    assert sorted(pairs[0]) == [0, 2]


def test_libm_variant_does_not_change_the_graph():
    """Divergence 2 of the oracle header: float vs correctly-rounded-double transcendentals."""
    from oracle.net_ref import NetRef
    cam = cameras.EUROC
    m = NetRef().forward_u8(synth.frame(0, cam.width, cam.height))
    a = O.extract_post(cam, m["prob"], m["heat"], m["desc"])
    b = O.extract_post(cam, m["prob"], m["heat"], m["desc"], variant="_libmf")
    for k in ("edge_start", "edge_end", "conn_idx", "col_pairs"):
        np.testing.assert_array_equal(a[k], b[k])
    assert a["n_edges"] > 100


def test_descriptor_sampling_matches_torch():
    """genPointDescriptor (:515-538) = grid_sampler(bilinear, zeros, align_corners=False) + F.normalize."""
    import torch
    import torch.nn.functional as F
    cam = cameras.EUROC
    rs = np.random.RandomState(1)
    Hc, Wc = cam.height // 8, cam.width // 8
    desc = rs.normal(size=(256, Hc, Wc)).astype(np.float32) * 50
    kx = rs.randint(4, cam.width - 4, 40).astype(np.int32)
    ky = rs.randint(4, cam.height - 4, 40).astype(np.int32)
    got = O.sample_descriptors(O.make_cfg(cam), desc, kx, ky)
    gx = (kx.astype(np.float32) / np.float32(cam.width)).astype(np.float64) * 2. - 1.
    gy = (ky.astype(np.float32) / np.float32(cam.height)).astype(np.float64) * 2. - 1.
    grid = torch.from_numpy(np.stack([gx, gy], 1).astype(np.float32))[None, None]
    s = F.grid_sample(torch.from_numpy(desc)[None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    want = F.normalize(s[0, :, 0].T, dim=1).numpy()
    assert np.abs(got - want).max() < 2e-6
    # fewer than 10 keypoints -> zero descriptors (:520-524)
    assert not O.sample_descriptors(O.make_cfg(cam), desc, kx[:9], ky[:9]).any()


@pytest.mark.parametrize("cam", [cameras.EUROC, cameras.TUMVI], ids=lambda c: c.name)
def test_grid_query_equals_brute_force_mask(cam):
    """Frame::GetFeaturesInArea (Frame.cpp:262-315) restated literally == indexable & |dx|<r & |dy|<r."""
    rs = np.random.RandomState(8)
    n = 400
    kx = rs.uniform(-250, cam.width + 250, n).astype(np.float32)
    ky = rs.uniform(-250, cam.height + 250, n).astype(np.float32)
    idxable = O.indexable(cam, kx, ky)
    assert 0 < idxable.sum() < n
    for _ in range(300):
        x, y = np.float32(rs.uniform(-300, cam.width + 300)), np.float32(rs.uniform(-300, cam.height + 300))
        r = np.float32(rs.choice([7.5, 12.0, 25.0, 40.0, 60.0]))
        got = O.features_in_area(cam, kx, ky, x, y, r)
        mask = (idxable != 0) & (np.abs(kx - x) < r) & (np.abs(ky - y) < r)
        assert sorted(got.tolist()) == np.nonzero(mask)[0].tolist()


def test_search_core_kat():
    """best / second-best over the window with the strict-< update and the OR-accept rule (Matcher.cpp:251-276)."""
    cam = cameras.EUROC
    rs = np.random.RandomState(2)
    n, m = 120, 200
    kx = rs.uniform(20, cam.width - 20, n).astype(np.float32)
    ky = rs.uniform(20, cam.height - 20, n).astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    inp = synth.association_inputs(3, fd, np.stack([kx, ky], 1), m, cam.width, cam.height, th=10.0)
    free = (rs.rand(n) > 0.2).astype(np.uint8)
    ref = O.search_all(cam, kx, ky, fd, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
    idxable = O.indexable(cam, kx, ky)
    for j in range(m):
        r = np.float32(10.0 * (2.5 if inp["view_cos"][j] > 0.998 else 4.0))
        x, y = inp["proj_uv"][j]
        order = O.features_in_area(cam, kx, ky, x, y, r)
        best, best2, bi, bi2 = 1e6, 1e6, -1, -1
        for i in order:
            if not free[i]:
                continue
            d = O.descriptor_distance(inp["map_desc"][j], fd[i])
            if d < best:
                best2, bi2, best, bi = best, bi, d, i
            elif d < best2:
                best2, bi2 = d, i
        assert ref["best_idx"][j] == bi and ref["second_idx"][j] == bi2
        acc = int(bi >= 0 and not (best > 0.8 and best > 0.8 * best2))
        assert ref["accept"][j] == acc
        if bi >= 0:
            assert abs(ref["best_d"][j] - np.linalg.norm(inp["map_desc"][j] - fd[bi])) < 1e-5
    assert ref["accept"].sum() > 20 and (ref["best_idx"] < 0).sum() > 0


@pytest.mark.parametrize("name,th,max_dist,e2_max", [("SearchByProjection(Cur,Last)", 15.0, 0.8, 0.0),
                                                      ("SearchByProjection(F,KF,descDist)", 10.0, 0.75, 0.0),
                                                      ("Fuse", 3.0, 0.7, 5.99)], ids=lambda v: str(v))
def test_window_search_core_kat(name, th, max_dist, e2_max):
    """The best-only projection cores (Matcher.cpp:31-87, :1337-1411, :897-1036): r = th, strict-< best over the
    window in GetFeaturesInArea order, Fuse's e2 > 5.99 skip, accept = best <= threshold."""
    cam = cameras.EUROC
    rs = np.random.RandomState(5)
    n, m = 150, 250
    kx = rs.uniform(20, cam.width - 20, n).astype(np.float32)
    ky = rs.uniform(20, cam.height - 20, n).astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    inp = synth.association_inputs(6, fd, np.stack([kx, ky], 1), m, cam.width, cam.height, th=th)
    # planted rows project close to their keypoint so that Fuse's 2.45 px circle keeps some of them
    inp["proj_uv"][:60] = np.stack([kx[:60], ky[:60]], 1) + rs.uniform(-1.5, 1.5, (60, 2)).astype(np.float32)
    free = (rs.rand(n) > 0.2).astype(np.uint8)
    ref = O.search_all(cam, kx, ky, fd, free, inp["map_desc"], inp["proj_uv"], inp["view_cos"], th, 0.8, mode=1,
                       max_dist=max_dist, e2_max=e2_max)
    for j in range(m):
        x, y = inp["proj_uv"][j]
        order = O.features_in_area(cam, kx, ky, x, y, np.float32(th))
        best, bi = np.float32(1e6), -1
        for i in order:
            if not free[i]:
                continue
            if e2_max > 0:
                ex, ey = np.float32(x) - kx[i], np.float32(y) - ky[i]
                e2 = np.float32(np.float32(ex * ex) + np.float32(ey * ey))
                if float(e2) > e2_max:
                    continue
            d = O.descriptor_distance(inp["map_desc"][j], fd[i])
            if d < best:
                best, bi = d, i
        assert ref["best_idx"][j] == bi, (name, j)
        assert ref["accept"][j] == int(bi >= 0 and best <= np.float32(max_dist))
    assert ref["accept"].sum() > 5 and (ref["best_idx"] < 0).sum() > 0


def test_distinctive_descriptor_kat():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cpp:234-302): least median distance, median =
    sorted[(int)(0.5 (N-1))] (the self distance 0 is part of every row), BestMedian starts at 1.0, first minimum."""
    rs = np.random.RandomState(12)
    sizes = [1, 2, 3, 4, 7, 8, 15, 33, 64, 128, 5, 2]
    centre = rs.normal(size=(len(sizes), 256)).astype(np.float32)
    chunks = []
    for k, n in enumerate(sizes):
        d = centre[k] + rs.normal(size=(n, 256)).astype(np.float32) * (0.15 if k != 4 else 3.0)  # k == 4: all far apart
        chunks.append((d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32))
    desc = np.concatenate(chunks)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    got = O.distinctive_all(desc, off)
    for k, n in enumerate(sizes):
        d = chunks[k]
        D = np.zeros((n, n), np.float32)
        for i in range(n):
            for j in range(i + 1, n):
                D[i, j] = D[j, i] = O.descriptor_distance(d[i], d[j])
        best, best_med = 0, np.float32(1.0)
        for i in range(n):
            med = np.sort(D[i])[int(0.5 * (n - 1))]
            if med < best_med:
                best, best_med = i, med
        assert got[k] == best, (k, n)
    assert got[1] == 0 and got[0] == 0  # N <= 2: the median is the self distance, index 0 wins
