"""GPU parity of the whole Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381: window search with the live
frame state, assignment, seed growing over the point-pair graph) through ppg_extend_map_matches /
ppg_extend_run_batch against the oracle (oracle/ppg_oracle.c::ppgo_extend_map_matches, itself pinned by
tests/test_oracle_extend.py).  Everything compared is an index or a count: bit-exact."""
import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

pytestmark = pytest.mark.gpu


def _frame_graph(rs, cam, n, n_edges, spacing=6):
    gx, gy = np.meshgrid(np.arange(8, cam.width - 8, spacing), np.arange(8, cam.height - 8, spacing))
    sel = rs.choice(gx.size, n, replace=False)
    kx = gx.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    ky = gy.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    pairs = set()
    while len(pairs) < n_edges:
        a, b = rs.randint(0, n, 2)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    pairs = sorted(pairs)
    es = np.array([p[0] for p in pairs], np.int32)
    ee = np.array([p[1] for p in pairs], np.int32)
    conn = [[] for _ in range(n)]
    for e, (a, b) in enumerate(pairs):
        conn[a].append(e)
        conn[b].append(e)
    off = np.zeros(n + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in conn])
    idx = np.array([e for c in conn for e in c], np.int32)
    return kx, ky, fd.astype(np.float32), es, ee, off, idx


def _oracle(cam, inp, kx, ky, fd, es, ee, coff, cidx, th, ratio, kp_mp=None, tracked=None):
    from oracle import post_ref as O
    return O.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                                inp["edge_other"], inp["edge_ok"], inp["proj_uv"], inp["view_cos"],
                                inp["tracked"] if tracked is None else tracked, kx, ky, fd,
                                inp["kp_mp"] if kp_mp is None else kp_mp, es, ee, coff, cidx, th=th, ratio=ratio)


def _same(got, ref):
    assert got["nmatches"] == ref["nmatches"]
    np.testing.assert_array_equal(got["kp_mp"], ref["kp_mp"])
    np.testing.assert_array_equal(got["kedge_me"], ref["kedge_me"])
    np.testing.assert_array_equal(got["tracked"], ref["tracked"])


@pytest.mark.parametrize("cam,n,n_edges,m,th,clean,seed",
                         [(cameras.EUROC, 357, 700, 8192, 10.0, True, 0), (cameras.EUROC, 357, 700, 8192, 10.0, False, 1),
                          (cameras.UMA, 1000, 2500, 30000, 10.0, False, 2), (cameras.TUMVI, 500, 900, 3000, 15.0, False, 3),
                          (cameras.EUROC, 40, 30, 300, 3.0, False, 4)],
                         ids=["euroc-clean", "euroc-state", "uma-1000x30k", "tumvi-th15", "euroc-small"])
def test_extend_map_matches_equals_oracle(cam, n, n_edges, m, th, clean, seed):
    from ppg_slam_b200 import capi
    rs = np.random.RandomState(40 + seed)
    kx, ky, fd, es, ee, coff, cidx = _frame_graph(rs, cam, n, n_edges)
    inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, m, cam.width, cam.height, th=th,
                              planted_frac=0.4, clean=clean)
    ref = _oracle(cam, inp, kx, ky, fd, es, ee, coff, cidx, th, 0.8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=max(m, 1024), junction_max_num=max(500, n))
    try:
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        got = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                   inp["tracked"], th, 0.8)
        _same(got, ref)
        assert got["status"] == 0
        assert got["n_accepted"] >= min(20, n // 4) and got["n_grown"] >= 5
        assert got["nmatches"] == 2 * got["n_accepted"]
        # a second call on the same ctx (state buffers are reused) gives the same answer
        got2 = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                    inp["tracked"], th, 0.8)
        _same(got2, ref)
    finally:
        e.close()


def test_extend_dense_windows_force_rescans():
    """Keypoints every 3 px and th = 15: windows hold far more than the 16 stored candidates and most of the first 16
    get taken, so rows run out of stored candidates and are rescanned over the whole window by the CTA."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    rs = np.random.RandomState(9)
    n = 1000
    gx, gy = np.meshgrid(np.arange(200, 200 + 3 * 40, 3), np.arange(100, 100 + 3 * 25, 3))
    kx = gx.ravel()[:n].astype(np.float32)
    ky = gy.ravel()[:n].astype(np.float32)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    es, ee = np.arange(0, n - 1, dtype=np.int32), np.arange(1, n, dtype=np.int32)
    conn = [[] for _ in range(n)]
    for k in range(n - 1):
        conn[k].append(k)
        conn[k + 1].append(k)
    coff = np.zeros(n + 1, np.int32)
    coff[1:] = np.cumsum([len(c) for c in conn])
    cidx = np.array([x for c in conn for x in c], np.int32)
    m = 6000
    inp = synth.extend_inputs(5, fd, np.stack([kx, ky], 1), es, ee, m, cam.width, cam.height, th=15.0,
                              planted_frac=0.9, clean=False)
    # almost every keypoint already holds an observed map point -> the lists' heads are occupied
    kp_mp = inp["kp_mp"].copy()
    kp_mp[rs.rand(n) < 0.85] = -2
    ref = _oracle(cam, inp, kx, ky, fd, es, ee, coff, cidx, 15.0, 0.8, kp_mp=kp_mp)
    e = capi.Extractor(cam, max_batch=1, max_map_points=8192, junction_max_num=1000)
    try:
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        got = e.extend_map_matches(kx, ky, fd, kp_mp, es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                   inp["tracked"], 15.0, 0.8)
        _same(got, ref)
        assert got["n_rescans"] > 10, got["n_rescans"]
    finally:
        e.close()


def test_extend_edge_cases():
    """No keypoints; no map edges at all; no candidates; a map point with more edges than the capacity."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    rs = np.random.RandomState(1)
    kx, ky, fd, es, ee, coff, cidx = _frame_graph(rs, cam, 120, 200)
    m = 1500
    inp = synth.extend_inputs(7, fd, np.stack([kx, ky], 1), es, ee, m, cam.width, cam.height, th=10.0, clean=True)
    e = capi.Extractor(cam, max_batch=1, max_map_points=2048)
    try:
        e.upload_map(inp["map_desc"])
        g = [inp[k] for k in ("candidate", "observed", "bad", "edge_off", "edge_other", "edge_ok")]
        e.upload_map_graph(*g)
        z = np.zeros(0, np.int32)
        got = e.extend_map_matches(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 256), np.float32),
                                   None, z, z, np.zeros(1, np.int32), z, inp["proj_uv"], inp["view_cos"], None, 10.0, 0.8)
        assert got["nmatches"] == 0 and len(got["kp_mp"]) == 0 and not got["tracked"].any()
        # map without edges: pure window search with live occupancy
        noedge = dict(inp, edge_off=np.zeros(m + 1, np.int32), edge_other=z, edge_ok=np.zeros(0, np.uint8))
        e.upload_map_graph(noedge["candidate"], noedge["observed"], noedge["bad"], noedge["edge_off"], z,
                           np.zeros(0, np.uint8))
        got = e.extend_map_matches(kx, ky, fd, None, es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"], None, 10.0, 0.8)
        _same(got, _oracle(cam, noedge, kx, ky, fd, es, ee, coff, cidx, 10.0, 0.8))
        assert got["n_grown"] == 0 and got["n_accepted"] > 20
        # nothing trackable
        none = dict(inp, candidate=np.zeros(m, np.uint8))
        e.upload_map_graph(none["candidate"], *g[1:])
        got = e.extend_map_matches(kx, ky, fd, None, es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"], None, 10.0, 0.8)
        assert got["nmatches"] == 0 and (got["kp_mp"] == -1).all()
        # frame without key edges: seed growing has nothing to grow along
        e.upload_map_graph(*g)
        got = e.extend_map_matches(kx, ky, fd, None, z, z, np.zeros(len(kx) + 1, np.int32), z, inp["proj_uv"],
                                   inp["view_cos"], None, 10.0, 0.8)
        _same(got, _oracle(cam, inp, kx, ky, fd, z, z, np.zeros(len(kx) + 1, np.int32), z, 10.0, 0.8))
        # capacity: one accepted map point with 200 edges -> PPG_ERR_CAPACITY, never a wrong answer
        row = int(inp["planted_rows"][0])
        deg = np.diff(inp["edge_off"]).copy()
        adj = [list(zip(inp["edge_other"][inp["edge_off"][p]:inp["edge_off"][p + 1]].tolist(),
                        inp["edge_ok"][inp["edge_off"][p]:inp["edge_off"][p + 1]].tolist())) for p in range(m)]
        adj[row] = [((row + 1 + k) % m, 1) for k in range(200)]
        off = np.zeros(m + 1, np.int32)
        off[1:] = np.cumsum([len(a) for a in adj])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], off,
                           np.array([o for a in adj for o, _ in a], np.int32),
                           np.array([k for a in adj for _, k in a], np.uint8))
        with pytest.raises(capi.PpgError) as ei:
            e.extend_map_matches(kx, ky, fd, None, es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"], None, 10.0, 0.8)
        assert ei.value.code == capi.PPG_ERR_CAPACITY
        with pytest.raises(capi.PpgError):  # ratio >= 1: the reference itself is undefined there
            e.extend_map_matches(kx, ky, fd, None, es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"], None, 10.0, 1.0)
    finally:
        e.close()


def test_extend_batch_on_extracted_frames():
    """extract -> ExtendMapMatches of every frame of the batch with keypoints, descriptors and the point-pair graph
    left on the device; the map graph mirrors frame 0's point-pair graph."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    Bn, M = 4, 4096
    e = capi.Extractor(cam, max_batch=Bn, max_map_points=M)
    try:
        recs = e.run([synth.frame(s, cam.width, cam.height) for s in range(Bn)])
        r0 = recs[0]
        inp = synth.extend_inputs(4, r0["desc"], np.stack([r0["kp_x"], r0["kp_y"]], 1), r0["edge_start"], r0["edge_end"],
                                  M, cam.width, cam.height, th=10.0, clean=True)
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        rs = np.random.RandomState(0)
        uv = np.stack([inp["proj_uv"] + rs.uniform(-3, 3, inp["proj_uv"].shape).astype(np.float32) for _ in range(Bn)])
        vc = np.stack([inp["view_cos"]] * Bn)
        e.assoc_stage_batch(uv, vc, 10.0, 0.8)
        e.extend_run_batch(Bn)
        got = e.extend_fetch_batch(Bn)
        for f in range(Bn):
            r = recs[f]
            fi = dict(inp, proj_uv=uv[f], view_cos=vc[f], kp_mp=np.full(r["n_kp"], -1, np.int32))
            ref = _oracle(cam, fi, r["kp_x"], r["kp_y"], r["desc"], r["edge_start"], r["edge_end"], r["conn_off"],
                          r["conn_idx"], 10.0, 0.8)
            _same(got[f], ref)
        assert got[0]["n_accepted"] > 50 and got[0]["n_grown"] > 50
    finally:
        e.close()


@pytest.mark.parametrize("cam", [cameras.EUROC, cameras.TUMVI, cameras.UMA], ids=lambda c: c.name)
def test_check_in_frustum_equals_oracle(cam):
    """Frame::CheckInFrustum on the device (ppg_assoc_stage_poses) == the CPU restatement, bit for bit: flags,
    projections, depth and viewing cosine of every map point under three poses (pinhole and KannalaBrandt8)."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    M, F = 5000, 3
    g = synth.frustum_inputs(8, cam, M, n_frames=F)
    rs = np.random.RandomState(0)
    desc = rs.normal(size=(M, 256)).astype(np.float32)
    e = capi.Extractor(cam, max_batch=F, max_map_points=8192)
    try:
        e.upload_map(desc)
        e.upload_map_geometry(g["world_pos"], g["normal"], g["min_dist"], g["max_dist"])
        e.assoc_stage_poses(g["Rcw"], g["tcw"], g["Ow"], M, 0.5, 10.0, 0.8)
        got = e.frustum_fetch(F)
        for f in range(F):
            ref = O.check_in_frustum(cam, g["Rcw"][f], g["tcw"][f], g["Ow"][f], g["world_pos"], g["normal"],
                                     g["min_dist"], g["max_dist"], 0.5)
            np.testing.assert_array_equal(got["in_view"][f], ref["in_view"])
            for k in ("proj_uv", "depth", "view_cos"):
                np.testing.assert_array_equal(got[k][f].view(np.uint32), ref[k].view(np.uint32), err_msg=k)
            assert 0.1 * M < ref["in_view"].sum() < 0.9 * M
    finally:
        e.close()


def test_extend_batch_with_device_projection():
    """The tracking step end to end on the device: extract -> CheckInFrustum of the resident map under each frame's
    pose -> ExtendMapMatches.  Oracle: CheckInFrustum restatement, then ExtendMapMatches with mbTrackInView folded
    into the candidate flags (Matcher.cpp:212)."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    Bn, M = 3, 4096
    e = capi.Extractor(cam, max_batch=Bn, max_map_points=M)
    try:
        recs = e.run([synth.frame(s, cam.width, cam.height) for s in range(Bn)])
        r0 = recs[0]
        inp = synth.extend_inputs(6, r0["desc"], np.stack([r0["kp_x"], r0["kp_y"]], 1), r0["edge_start"], r0["edge_end"],
                                  M, cam.width, cam.height, th=10.0, clean=False)
        # world points that project (under the identity pose) where extend_inputs put the projections
        rs = np.random.RandomState(2)
        z = rs.uniform(2.0, 9.0, M).astype(np.float32)
        fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
        P = np.stack([(inp["proj_uv"][:, 0] - cx) / fx * z, (inp["proj_uv"][:, 1] - cy) / fy * z, z], 1).astype(np.float32)
        nrm = (P / np.linalg.norm(P, axis=1, keepdims=True) + rs.normal(0, 0.3, P.shape)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        d = np.linalg.norm(P, axis=1)
        dmin, dmax = (d * rs.uniform(0.4, 1.02, M)).astype(np.float32), (d * rs.uniform(0.98, 2.5, M)).astype(np.float32)
        g = synth.frustum_inputs(3, cam, 8, n_frames=Bn)  # poses only
        Rcw, tcw, Ow = g["Rcw"], g["tcw"] * 0.1, None
        Ow = np.stack([-(Rcw[f].T @ tcw[f]) for f in range(Bn)]).astype(np.float32)
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        e.upload_map_geometry(P, nrm, dmin, dmax)
        e.assoc_stage_poses(Rcw, tcw, Ow, M, 0.5, 10.0, 0.8)
        e.extend_run_batch(Bn)
        got = e.extend_fetch_batch(Bn)
        n_acc = 0
        for f in range(Bn):
            r = recs[f]
            fr = O.check_in_frustum(cam, Rcw[f], tcw[f], Ow[f], P, nrm, dmin, dmax, 0.5)
            fi = dict(inp, proj_uv=fr["proj_uv"], view_cos=fr["view_cos"], candidate=inp["candidate"] & fr["in_view"],
                      kp_mp=np.full(r["n_kp"], -1, np.int32), tracked=np.zeros(M, np.uint8))
            ref = _oracle(cam, fi, r["kp_x"], r["kp_y"], r["desc"], r["edge_start"], r["edge_end"], r["conn_off"],
                          r["conn_idx"], 10.0, 0.8)
            _same(got[f], ref)
            n_acc += got[f]["n_accepted"]
            assert 0.2 * M < fr["in_view"].sum() < M
        assert n_acc > 100
    finally:
        e.close()


def test_extend_is_deterministic_over_repeats():
    """The walk synchronises 256 threads through shared memory; a missing barrier would show up as run-to-run
    differences.  40 repeats of a seed-rich case (and of the batch form on three contexts' worth of frames) must
    give the same record every time."""
    from ppg_slam_b200 import capi
    cam = cameras.UMA
    rs = np.random.RandomState(77)
    kx, ky, fd, es, ee, coff, cidx = _frame_graph(rs, cam, 900, 2600)
    m = 12000
    inp = synth.extend_inputs(9, fd, np.stack([kx, ky], 1), es, ee, m, cam.width, cam.height, th=10.0,
                              planted_frac=0.5, clean=False)
    ref = _oracle(cam, inp, kx, ky, fd, es, ee, coff, cidx, 10.0, 0.8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=16384, junction_max_num=1000)
    try:
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        for _ in range(40):
            got = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                       inp["tracked"], 10.0, 0.8)
            _same(got, ref)
        assert got["n_grown"] > 100
    finally:
        e.close()


def test_extend_single_frame_with_staged_poses():
    """ppg_extend_map_matches with proj_uv = view_cos = NULL consumes the projections Frame::CheckInFrustum left on the
    device (ppg_assoc_stage_poses) -- the SearchLocalPoints sequence of the shim."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    rs = np.random.RandomState(21)
    kx, ky, fd, es, ee, coff, cidx = _frame_graph(rs, cam, 300, 600)
    M = 5000
    inp = synth.extend_inputs(8, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=10.0, clean=False)
    z = rs.uniform(2.0, 9.0, M).astype(np.float32)
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    P = np.stack([(inp["proj_uv"][:, 0] - cx) / fx * z, (inp["proj_uv"][:, 1] - cy) / fy * z, z], 1).astype(np.float32)
    nrm = (P / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32)
    d = np.linalg.norm(P, axis=1)
    dmin, dmax = (0.5 * d).astype(np.float32), (2.0 * d).astype(np.float32)
    g = synth.frustum_inputs(5, cam, 8, n_frames=1)
    Rcw, tcw = g["Rcw"], (g["tcw"] * 0.1).astype(np.float32)
    Ow = np.stack([-(Rcw[0].T @ tcw[0])]).astype(np.float32)
    fr = O.check_in_frustum(cam, Rcw[0], tcw[0], Ow[0], P, nrm, dmin, dmax, 0.5)
    fi = dict(inp, proj_uv=fr["proj_uv"], view_cos=fr["view_cos"], candidate=inp["candidate"] & fr["in_view"])
    ref = _oracle(cam, fi, kx, ky, fd, es, ee, coff, cidx, 10.0, 0.8)
    e = capi.Extractor(cam, max_batch=1, max_map_points=8192)
    try:
        e.upload_map(inp["map_desc"])
        e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                           inp["edge_ok"])
        e.upload_map_geometry(P, nrm, dmin, dmax)
        with pytest.raises(capi.PpgError):  # nothing staged yet
            e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, None, None, inp["tracked"], 10.0, 0.8)
        e.assoc_stage_poses(Rcw, tcw, Ow, M, 0.5, 10.0, 0.8)
        got = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, None, None, inp["tracked"], 10.0, 0.8)
        _same(got, ref)
        assert got["n_accepted"] > 50
    finally:
        e.close()


def test_extend_map_matches_random_small_graphs_sweep():
    """60 small random configurations (3..44 keypoints, 5..139 map rows, clustered keypoints, aliased map edges,
    th 3 / 10 / 15, ratio 0.6 / 0.8 / 0.95): the GPU twin of
    tests/test_oracle_extend.py::test_extend_map_matches_random_small_graphs (same generator, same oracle)."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    from tests.test_oracle_extend import random_frame_graph
    cam = cameras.EUROC
    e = capi.Extractor(cam, max_batch=1, max_map_points=1024)
    bad = []
    try:
        for seed in range(60):
            rs = np.random.RandomState(1000 + seed)
            n = int(rs.randint(3, 45))
            M = int(rs.randint(5, 140))
            ne = int(rs.randint(0, min(3 * n, n * (n - 1) // 2) + 1))
            kx, ky, fd, es, ee, coff, cidx, _ = random_frame_graph(rs, cam, n, ne)
            if seed % 3 == 0:  # everything inside a few search windows
                kx = (300 + rs.uniform(0, 60, n)).astype(np.float32)
                ky = (200 + rs.uniform(0, 40, n)).astype(np.float32)
            th = float(rs.choice([3.0, 10.0, 15.0]))
            inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=th,
                                      planted_frac=float(rs.uniform(0.2, 0.9)), clean=bool(seed % 4 == 1))
            if seed % 5 == 2 and len(inp["edge_other"]) > 4:  # several map edges of a point lead to the same row
                inp["edge_other"][1::7] = inp["edge_other"][0::7][:len(inp["edge_other"][1::7])]
            ratio = float(rs.choice([0.6, 0.8, 0.95]))
            ref = O.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"],
                                       inp["edge_off"], inp["edge_other"], inp["edge_ok"], inp["proj_uv"],
                                       inp["view_cos"], inp["tracked"], kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx,
                                       th=th, ratio=ratio)
            e.upload_map(inp["map_desc"])
            e.upload_map_graph(inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"], inp["edge_other"],
                               inp["edge_ok"])
            got = e.extend_map_matches(kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, inp["proj_uv"], inp["view_cos"],
                                       inp["tracked"], th, ratio)
            ok = (got["nmatches"] == ref["nmatches"] and np.array_equal(got["kp_mp"], ref["kp_mp"]) and
                  np.array_equal(got["kedge_me"], ref["kedge_me"]) and np.array_equal(got["tracked"], ref["tracked"]))
            if not ok:
                bad.append((seed, n, M, ne, th, ratio))
    finally:
        e.close()
    assert not bad, "mismatching (seed, n, M, edges, th, ratio): %s" % bad


def test_pipelined_extend_equals_synchronous_fetch():
    """ppg_assoc_stage_batch_async (pinned projections) + ppg_extend_run_batch + ppg_extend_fetch_batch_async +
    ppg_extend_collect == ppg_assoc_stage_batch + ppg_extend_run_batch + ppg_extend_fetch_batch, on a batch in which every
    frame has its local map (synth.extend_inputs_multi: the benchmark workload)."""
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    Bn, M = 4, 2048
    e = capi.Extractor(cam, max_batch=Bn, max_map_points=M)
    try:
        frames = [synth.frame(s, cam.width, cam.height) for s in range(Bn)]
        recs = e.run(frames)
        base = synth.extend_inputs_multi(17, recs, M, cam.width, cam.height, th=10.0)
        e.upload_map(base["map_desc"])
        e.upload_map_graph(base["candidate"], base["observed"], base["bad"], base["edge_off"], base["edge_other"],
                           base["edge_ok"])
        e.assoc_stage_batch(base["proj_all"], base["vcos_all"], 10.0, 0.8)
        e.extend_run_batch(Bn)
        want = e.extend_fetch_batch(Bn)
        assert min(w["n_accepted"] for w in want) > 100  # every frame matches its own slice of the table
        for f in range(Bn):  # and equals the oracle
            r = recs[f]
            ref = _oracle(cam, dict(base, proj_uv=base["proj_all"][f], view_cos=base["vcos_all"][f],
                                    kp_mp=np.full(r["n_kp"], -1, np.int32)),
                          r["kp_x"], r["kp_y"], r["desc"], r["edge_start"], r["edge_end"], r["conn_off"], r["conn_idx"],
                          10.0, 0.8)
            _same(want[f], ref)
        pu = capi.pinned_array(base["proj_all"].shape, np.float32)
        pv = capi.pinned_array(base["vcos_all"].shape, np.float32)
        pu[...] = base["proj_all"]
        pv[...] = base["vcos_all"]
        pf = capi.pinned_array((Bn, cam.height, cam.width), np.uint8)
        for i, g in enumerate(frames):
            pf[i] = g
        e.extract_async([pf[i] for i in range(Bn)])
        e.assoc_stage_batch_async(pu, pv, 10.0, 0.8)
        e.extend_run_batch(Bn)
        e.extend_fetch_batch_async(Bn)
        e.extract_wait(Bn)
        got = e.extend_collect(Bn)
        for f in range(Bn):
            _same(got[f], want[f])
        with pytest.raises(capi.PpgError):  # pageable projections are refused by the asynchronous call
            e.assoc_stage_batch_async(base["proj_all"], base["vcos_all"], 10.0, 0.8)
    finally:
        e.close()
        capi.drop_pinned()
