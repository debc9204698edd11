"""GPU parity tests of the extraction path, driven through the C ABI (ppg_slam_b200/capi.py).

  * tcgen05 convolutions vs a plain CUDA-core convolution over the same operands (self-test);
  * dense network outputs vs the L0 oracle (torch fp32 CPU with the reference weights) -- tolerance:
        junction prob map  max-abs error <= 5e-3
        heat score map     max-abs error <= 1e-2
        sampled descriptors cosine >= 0.999          (north-star tolerances);
  * keypoints / point-pair graph / colines given the REFERENCE maps: bit-exact vs the L1 oracle.
"""
import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth

pytestmark = pytest.mark.gpu

PROB_TOL, HEAT_TOL, COS_MIN = 5e-3, 1e-2, 0.999


@pytest.fixture(scope="module")
def net():
    from oracle.net_ref import NetRef
    return NetRef()


@pytest.fixture(scope="module")
def ex_euroc():
    from ppg_slam_b200 import capi
    e = capi.Extractor(cameras.EUROC, max_batch=4)
    yield e
    e.close()


def test_conv_selftest(ex_euroc):
    ex_euroc.run([synth.frame(0, 752, 480)])
    for name, d, r in ex_euroc.selftest_conv():
        assert d <= 2e-2 * max(1.0, r), "layer %s: tcgen05 vs CUDA-core conv differ by %g (ref max %g)" % (name, d, r)


@pytest.mark.parametrize("kernel", ["1", "2", "3", "4", "5", "6", "8"],
                         ids=["generic", "halo", "transposed-unfused", "transposed-everywhere", "generic-for-cin128",
                              "halo-for-conv2a", "t128-for-every-layer-it-can-run"])
def test_conv_selftest_other_kernel_choices(kernel, monkeypatch):
    """PPG_CONV_KERNEL routes the Cin = 64 layers to the other tcgen05 kernels (A/B switch of conv_tc_plan): every choice
    must give the same layer outputs -- in particular the non-pooled epilogue of the transposed kernel (conv2a under "4"),
    which the default configuration does not use."""
    from ppg_slam_b200 import capi
    monkeypatch.setenv("PPG_CONV_KERNEL", kernel)
    e = capi.Extractor(cameras.EUROC, max_batch=2)
    try:
        e.run([synth.frame(1, 752, 480), synth.frame(2, 752, 480)])
        for name, d, r in e.selftest_conv():
            assert d <= 2e-2 * max(1.0, r), "kernel %s layer %s: differ by %g (ref max %g)" % (kernel, name, d, r)
    finally:
        e.close()


def test_fused_conv1a_is_bit_identical_to_the_separate_kernel(monkeypatch):
    """Default: conv1a runs inside conv1b's producer warps (same tcgen05 instruction on the same operands as the
    stand-alone conv1a kernel, zero rows for conv1b's padding).  PPG_CONV_KERNEL=3 runs the two kernels with the
    full-resolution map in HBM between them.  Everything downstream must be bit-identical, ragged tile edges (752 = 19 * 38
    + 30) and a batch that leaves SMs with different tile counts included."""
    from ppg_slam_b200 import capi
    frames = [synth.frame(s, 752, 480) for s in (4, 5, 6)]
    out = {}
    for mode in ("5", "3"):  # both run conv2a on the halo kernel and the Cin = 128 layers on the generic one
        monkeypatch.setenv("PPG_CONV_KERNEL", mode)
        e = capi.Extractor(cameras.EUROC, max_batch=3)
        try:
            recs = e.run(frames, allow_capacity=True)
            out[mode] = (recs, [e.get_maps(i) for i in range(3)])
        finally:
            e.close()
    for i in range(3):
        for k in ("prob", "heat", "desc"):
            assert np.array_equal(out["5"][1][i][k], out["3"][1][i][k]), (i, k)
        a, b = out["5"][0][i], out["3"][0][i]
        assert a["n_kp"] == b["n_kp"] and a["n_kp"] > 50
        for k in a:
            if isinstance(a[k], np.ndarray):
                assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (i, k)


@pytest.mark.parametrize("seed", [0, 3])
def test_dense_maps_within_tolerance(ex_euroc, net, seed):
    g = synth.frame(seed, 752, 480)
    rec = ex_euroc.run([g], allow_capacity=True)[0]
    m = ex_euroc.get_maps(0)
    ref = net.forward_u8(g)
    assert np.abs(m["prob"] - ref["prob"]).max() <= PROB_TOL
    assert np.abs(m["heat"] - ref["heat"]).max() <= HEAT_TOL
    a, b = m["desc"].reshape(256, -1), ref["desc"].reshape(256, -1)
    cos = (a * b).sum(0) / (np.linalg.norm(a, axis=0) * np.linalg.norm(b, axis=0) + 1e-12)
    assert cos.min() >= COS_MIN
    assert rec["n_kp"] > 50


@pytest.mark.parametrize("cam", [cameras.TUMVI, cameras.UMA, cameras.TUMVI1024], ids=lambda c: c.name)
def test_dense_maps_other_shapes(net, cam):
    """The BASELINE.json shapes besides EuRoC (TUM-VI 512x512, UMA-VI 1024x768, TUM-VI-1024): the tensor-core
    networks stay within the same tolerances, and the full path finds a plausible number of keypoints."""
    from ppg_slam_b200 import capi
    e = capi.Extractor(cam, max_batch=1)
    try:
        g = synth.frame(7, cam.width, cam.height)
        rec = e.run([g], allow_capacity=True)[0]
        m = e.get_maps(0)
        ref = net.forward_u8(g)
        assert np.abs(m["prob"] - ref["prob"]).max() <= PROB_TOL
        assert np.abs(m["heat"] - ref["heat"]).max() <= HEAT_TOL
        a, b = m["desc"].reshape(256, -1), ref["desc"].reshape(256, -1)
        cos = (a * b).sum(0) / (np.linalg.norm(a, axis=0) * np.linalg.norm(b, axis=0) + 1e-12)
        assert cos.min() >= COS_MIN
        assert rec["n_kp"] > 50
        for name, d, r in e.selftest_conv():
            assert d <= 2e-2 * max(1.0, r), "%s %s: tcgen05 vs CUDA-core conv differ by %g" % (cam.name, name, d)
    finally:
        e.close()


@pytest.mark.parametrize("cam,seeds", [(cameras.EUROC, [0, 1, 2, 5]), (cameras.TUMVI, [1, 4]),
                                       (cameras.UMA, [2]), (cameras.TUMVI1024, [0])], ids=lambda v: getattr(v, "name", str(v)))
def test_post_bit_exact_from_reference_maps(net, cam, seeds):
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    e = capi.Extractor(cam, max_batch=len(seeds))
    try:
        maps = [net.forward_u8(synth.frame(s, cam.width, cam.height)) for s in seeds]
        got = e.run_from_maps(np.stack([m["prob"] for m in maps]), np.stack([m["heat"] for m in maps]),
                              np.stack([m["desc"] for m in maps]))
        for f, m in enumerate(maps):
            ref = oracle_post(cam, m["prob"], m["heat"], m["desc"])
            assert got[f]["status"] == 0
            bad = diff_records(got[f], ref)
            assert not bad, "%s seed %d: %s" % (cam.name, seeds[f], "; ".join(bad))
            hf = e.get_maps(f)["heat_final"]
            np.testing.assert_array_equal(hf.view(np.uint32), ref["heat_final"].view(np.uint32))
            assert got[f]["n_cand"] == ref["n_cand"]
    finally:
        e.close()


def test_edge_cases_from_maps():
    """Empty map (N == 0), fewer than 10 keypoints (zero descriptors, :520-524), keypoints on the NMS
    border, the 500 cap, exact score ties."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    cam = cameras.EUROC
    H, W = cam.height, cam.width
    e = capi.Extractor(cam, max_batch=1)
    try:
        rs = np.random.RandomState(0)
        desc = rs.normal(size=(256, H // 8, W // 8)).astype(np.float32)
        heat = (rs.rand(H, W) * 0.5).astype(np.float32)
        cases = {}
        cases["empty"] = np.zeros((H, W), np.float32)
        few = np.zeros((H, W), np.float32)
        for k, (x, y) in enumerate([(100, 100), (3, 50), (200, 3), (W - 5, 60), (W - 4, 90), (300, H - 5), (400, 200)]):
            few[y, x] = 0.5 + 0.01 * k
        cases["few_and_border"] = few
        ties = np.zeros((H, W), np.float32)
        ties[40:440:3, 40:700:3] = 0.25  # dense lattice of exactly tied scores -> raster order decides, cap 500
        cases["ties_and_cap"] = ties
        ramp = np.zeros((H, W), np.float32)
        ramp[100, 50:700] = np.linspace(0.1, 0.9, 650).astype(np.float32)  # long suppression chain
        cases["ramp_chain"] = ramp
        # degenerate maps: every pixel is a candidate and more survivors (14 400 / ~9 000) than ranking slots exist --
        # the 500 best must still be exactly the reference's (hundreds of NMS rounds on the uniform map)
        cases["uniform_all_candidates"] = np.full((H, W), 0.5, np.float32)
        cases["dense_noise"] = (rs.rand(H, W) * 0.9 + 0.05).astype(np.float32)
        for name, prob in cases.items():
            got = e.run_from_maps(prob[None], heat[None], desc[None])[0]
            ref = oracle_post(cam, prob, heat, desc)
            bad = diff_records(got, ref)
            assert not bad, "%s: %s" % (name, "; ".join(bad))
        assert e.run_from_maps(cases["empty"][None], heat[None], desc[None])[0]["n_kp"] == 0
    finally:
        e.close()


def test_widest_nms_radius_bit_exact(net):
    """junction_nms_radius = 8 (17-pixel window rows: handled by the global-memory NMS variant) and = 7 (the widest
    the shared-memory bitmap variant takes)."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    cam = cameras.EUROC
    m = net.forward_u8(synth.frame(2, cam.width, cam.height))
    for R in (7, 8):
        e = capi.Extractor(cam, max_batch=1, junction_nms_radius=R)
        try:
            got = e.run_from_maps(m["prob"][None], m["heat"][None], m["desc"][None])[0]
            ref = oracle_post(cam, m["prob"], m["heat"], m["desc"], junction_nms_radius=R)
            bad = diff_records(got, ref)
            assert not bad, "R = %d: %s" % (R, "; ".join(bad))
        finally:
            e.close()


def test_non_default_tunables_bit_exact(net):
    """The ten static tunables of PPGExtractor (PPGExtractor.cpp:44-53) travel through ppg_config: a run with
    JUNCTION_MAX_NUM = 1000 (BASELINE config 4 raises it), a lower junction threshold, NMS radius 3 and other line
    thresholds stays bit-exact against the oracle given the same overrides."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    cam = cameras.UMA
    over = dict(junction_max_num=1000, junction_thresh=1.0 / 256.0, junction_nms_radius=3, line_valid_ratio=0.25,
                line_dist_thresh=1.5, line_heatmap_thresh=0.25, line_inlier_rate=0.7)
    m = net.forward_u8(synth.frame(11, cam.width, cam.height))
    e = capi.Extractor(cam, max_batch=1, max_edges=8192, max_colines=8192, **over)
    try:
        got = e.run_from_maps(m["prob"][None], m["heat"][None], m["desc"][None])[0]
        ref = oracle_post(cam, m["prob"], m["heat"], m["desc"], **over)
        assert got["status"] == 0
        bad = diff_records(got, ref)
        assert not bad, "; ".join(bad)
        assert got["n_kp"] > 500, "the raised cap must be exercised (n_kp = %d)" % got["n_kp"]
    finally:
        e.close()


def test_dense_graph_stress_and_capacity():
    """A lattice of keypoints under a mostly-hot heat map: thousands of candidate pairs, long interaction lists (the
    spill pool of the overlap filter), many colinear triples -- still bit-exact.  With the heat map fully hot every
    pair passes the 3-point test and the candidate table overflows: that must come back as PPG_ERR_CAPACITY with
    the overflow bit set, not as a crash or a silently truncated graph."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    cam = cameras.EUROC
    H, W = cam.height, cam.width
    rs = np.random.RandomState(5)
    desc = rs.normal(size=(256, H // 8, W // 8)).astype(np.float32)
    prob = np.zeros((H, W), np.float32)
    ys, xs = np.meshgrid(np.arange(60, 420, 24), np.arange(80, 680, 40), indexing="ij")
    prob[ys, xs] = (0.3 + 0.5 * rs.rand(*ys.shape)).astype(np.float32)      # 15 x 15 = 225 keypoints
    # three quarters of the pixels hot: a fully hot 16 x 16 tile would be zeroed by refineHeatMap (:557-560)
    yy, xx = np.mgrid[0:H, 0:W]
    pattern = np.where((yy % 2 == 0) & (xx % 2 == 0), 0.0, 0.9).astype(np.float32)
    cold = (rs.rand(H // 16, W // 16) < 0.25).repeat(16, 0).repeat(16, 1)   # a quarter of the tiles stay cold
    heat = np.where(cold, 0.0, pattern).astype(np.float32)
    e = capi.Extractor(cam, max_batch=1, max_edges=16384, max_colines=16384)
    try:
        got = e.run_from_maps(prob[None], heat[None], desc[None], allow_capacity=True)[0]
        ref = oracle_post(cam, prob, heat, desc)
        assert got["status"] == 0, "capacity bits %d on the stress case" % got["status"]
        bad = diff_records(got, ref)
        assert not bad, "; ".join(bad)
        assert got["n_pairs_ok"] > 1500 and got["n_edges"] > 50, (got["n_pairs_ok"], got["n_edges"])
        hot = pattern
        dense = np.zeros((H, W), np.float32)
        dense[20:460:10, 20:740:10] = 0.5     # 44 x 72 candidates -> 500 keypoints, ~125 k passing pairs
        with pytest.raises(capi.PpgError):
            e.run_from_maps(dense[None], hot[None], desc[None])
        rec = e.run_from_maps(dense[None], hot[None], desc[None], allow_capacity=True)[0]
        assert rec["status"] & 2, "ST_OVF_PAIRS expected, status = %d" % rec["status"]
        # the ctx stays usable after an overflow
        again = e.run_from_maps(prob[None], heat[None], desc[None])[0]
        assert not diff_records(again, ref)
    finally:
        e.close()
    # small edge / coline capacities: flagged, no crash, keypoints still exact
    e = capi.Extractor(cam, max_batch=1, max_edges=64, max_colines=16)
    try:
        rec = e.run_from_maps(prob[None], heat[None], desc[None], allow_capacity=True)[0]
        assert rec["status"] & 8 and rec["status"] & 16, "ST_OVF_EDGES | ST_OVF_COLINE expected, status = %d" % rec["status"]
        np.testing.assert_array_equal(rec["px"], ref["px"])
        np.testing.assert_array_equal(rec["py"], ref["py"])
    finally:
        e.close()


def test_nms_global_memory_variant(net, monkeypatch):
    """Frames whose 2-bit state map does not fit in shared memory (1024x1024) take nms_global_kernel; force it on an
    EuRoC frame and on the tie / chain cases and require the same bit-exact records."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    monkeypatch.setenv("PPG_NMS_GLOBAL", "1")
    cam = cameras.EUROC
    H, W = cam.height, cam.width
    e = capi.Extractor(cam, max_batch=1)
    try:
        m = net.forward_u8(synth.frame(1, W, H))
        ties = np.zeros((H, W), np.float32)
        ties[40:440:3, 40:700:3] = 0.25
        ramp = np.zeros((H, W), np.float32)
        ramp[100, 50:700] = np.linspace(0.1, 0.9, 650).astype(np.float32)
        for name, prob in (("frame", m["prob"]), ("ties", ties), ("ramp", ramp)):
            got = e.run_from_maps(prob[None], m["heat"][None], m["desc"][None])[0]
            ref = oracle_post(cam, prob, m["heat"], m["desc"])
            bad = diff_records(got, ref)
            assert not bad, "%s: %s" % (name, "; ".join(bad))
    finally:
        e.close()


def test_batch_equals_single(ex_euroc):
    """Frames are independent: a batch of 4 gives the same records as four single-frame calls."""
    frames = [synth.frame(s, 752, 480) for s in (0, 1, 2, 3)]
    batch = ex_euroc.run(frames)
    for f in range(4):
        single = ex_euroc.run([frames[f]])[0]
        for k in ("px", "py", "edge_start", "edge_end", "col_pairs"):
            np.testing.assert_array_equal(batch[f][k], single[k])
        np.testing.assert_array_equal(batch[f]["desc"], single["desc"])


def test_padded_rows_empty_frames_and_mixed_batch(ex_euroc):
    """cv::Mat rows may be padded (run() takes data + step); a batch may mix ordinary frames with frames that yield no
    keypoints at all (detectLines / genPointDescriptor return early, :210, :239-240); the batched association must
    cope with a frame of zero keypoints."""
    from oracle import post_ref as O
    cam = cameras.EUROC
    g0, g1 = synth.frame(4, 752, 480), synth.frame(6, 752, 480)
    padded = np.zeros((480, 800), np.uint8)
    padded[:, :752] = g0
    view = padded[:, :752]
    assert view.strides[0] == 800
    black = np.zeros((480, 752), np.uint8)
    single0, single1 = ex_euroc.run([g0])[0], ex_euroc.run([g1])[0]
    batch = ex_euroc.run([view, black, g1])
    assert batch[1]["n_kp"] == 0 and batch[1]["n_edges"] == 0 and batch[1]["status"] == 0
    for got, want in ((batch[0], single0), (batch[2], single1)):
        for k in ("px", "py", "edge_start", "edge_end", "conn_idx", "col_pairs"):
            np.testing.assert_array_equal(got[k], want[k])
        np.testing.assert_array_equal(got["desc"], want["desc"])
    # association of the three frames in one batch: the empty frame has no candidates anywhere
    M = 2048
    kp0 = np.stack([single0["kp_x"], single0["kp_y"]], 1)
    inp = synth.association_inputs(3, single0["desc"], kp0, M, cam.width, cam.height, th=10.0)
    ex_euroc.upload_map(inp["map_desc"])
    uv = np.stack([inp["proj_uv"]] * 3)
    vc = np.stack([inp["view_cos"]] * 3)
    ex_euroc.assoc_stage_batch(uv, vc, 10.0, 0.8)
    ex_euroc.assoc_run_batch(3)
    got = ex_euroc.assoc_fetch_batch(3)
    assert (got[1]["best_idx"] == -1).all() and not got[1]["accept"].any()
    for f, r in ((0, single0), (2, single1)):
        ref = O.search_all(cam, r["kp_x"], r["kp_y"], r["desc"], np.ones(r["n_kp"], np.uint8), inp["map_desc"],
                           inp["proj_uv"], inp["view_cos"], 10.0, 0.8)
        np.testing.assert_array_equal(got[f]["best_idx"], ref["best_idx"])
        np.testing.assert_array_equal(got[f]["accept"], ref["accept"])


def test_full_path_agrees_with_oracle_graph(ex_euroc, net):
    """End to end (fp16 tensor-core networks): keypoint sets must overlap the fp32 oracle's almost entirely."""
    from tests.parity_util import oracle_post
    g = synth.frame(0, 752, 480)
    got = ex_euroc.run([g])[0]
    m = net.forward_u8(g)
    ref = oracle_post(cameras.EUROC, m["prob"], m["heat"], m["desc"])
    a = set(zip(got["px"].tolist(), got["py"].tolist()))
    b = set(zip(ref["px"].tolist(), ref["py"].tolist()))
    assert len(a & b) >= 0.9 * len(b)
    # descriptors of the common keypoints: cosine >= 0.999
    ia = {p: i for i, p in enumerate(zip(got["px"].tolist(), got["py"].tolist()))}
    ib = {p: i for i, p in enumerate(zip(ref["px"].tolist(), ref["py"].tolist()))}
    common = sorted(a & b)
    da = got["desc"][[ia[p] for p in common]]
    db = ref["desc"][[ib[p] for p in common]]
    assert (da * db).sum(1).min() >= COS_MIN


def test_contexts_of_different_shapes_coexist(net):
    """Two contexts with different frame shapes (and so different shared-memory footprints of the per-frame kernels)
    alive in one process, used alternately."""
    from ppg_slam_b200 import capi
    from tests.parity_util import diff_records, oracle_post
    big, small = cameras.UMA, cameras.TUMVI
    mb = net.forward_u8(synth.frame(3, big.width, big.height))
    ms = net.forward_u8(synth.frame(3, small.width, small.height))
    eb = capi.Extractor(big, max_batch=1)
    es = capi.Extractor(small, max_batch=1)   # created second: must not shrink what the first one needs
    try:
        for _ in range(2):
            gb = eb.run_from_maps(mb["prob"][None], mb["heat"][None], mb["desc"][None])[0]
            gs = es.run_from_maps(ms["prob"][None], ms["heat"][None], ms["desc"][None])[0]
        assert not diff_records(gb, oracle_post(big, mb["prob"], mb["heat"], mb["desc"]))
        assert not diff_records(gs, oracle_post(small, ms["prob"], ms["heat"], ms["desc"]))
    finally:
        eb.close()
        es.close()


def test_pipelined_calls_equal_synchronous_calls(ex_euroc):
    """ppg_extract_async / ppg_extract_wait with frames in pinned memory (and with pageable frames, which go through the
    staging buffer) return the records ppg_extract returns; two contexts driven as a ring from one thread."""
    from ppg_slam_b200 import capi
    frames = [synth.frame(s, 752, 480) for s in (20, 21, 22)]
    want = ex_euroc.run(frames)
    pf = capi.pinned_array((3, 480, 752), np.uint8)
    for i, g in enumerate(frames):
        pf[i] = g
    other = capi.Extractor(cameras.EUROC, max_batch=4)
    try:
        ring = [ex_euroc, other]
        got = {}
        for step in range(6):
            x = ring[step % 2]
            if step >= 2:
                got[step - 2] = x.extract_wait(3)
            if step < 4:
                x.extract_async([pf[i] for i in range(3)] if step != 1 else frames)  # step 1: pageable frames
        for k in range(4):
            for f in range(3):
                for key in ("px", "py", "edge_start", "edge_end", "col_pairs", "conn_idx"):
                    np.testing.assert_array_equal(got[k][f][key], want[f][key])
                np.testing.assert_array_equal(got[k][f]["desc"], want[f]["desc"])
    finally:
        other.close()
        capi.drop_pinned()


def test_two_contexts_on_two_threads(ex_euroc):
    """Different contexts are independent: two host threads driving one ctx each (the pipelined e2e arm of bench.py)
    get the records a single-threaded run gets."""
    import threading
    from ppg_slam_b200 import capi
    frames = [synth.frame(s, 752, 480) for s in (10, 11, 12, 13)]
    want = ex_euroc.run(frames)
    other = capi.Extractor(cameras.EUROC, max_batch=4)
    res, errs = {}, []

    def work(tag, e, fr):
        try:
            for _ in range(4):
                res[tag] = e.run(fr)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    try:
        th = [threading.Thread(target=work, args=("a", ex_euroc, frames)),
              threading.Thread(target=work, args=("b", other, frames[::-1]))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        for f in range(4):
            for got in (res["a"][f], res["b"][3 - f]):
                for k in ("px", "py", "edge_start", "edge_end", "col_pairs"):
                    np.testing.assert_array_equal(got[k], want[f][k])
                np.testing.assert_array_equal(got["desc"], want[f]["desc"])
    finally:
        other.close()


@pytest.mark.parametrize("variant", ["fused-scan", "scan-kernel", "fused-scan-global-nms"])
def test_full_path_equals_post_processing_of_its_own_maps(variant, monkeypatch):
    """The threshold scan fused into the junction head's epilogue (candidates, NMS state map, counters written by
    convPb) must give exactly the record the stand-alone post-processing gives on the SAME dense maps -- the path that
    is bit-exact against the oracle (ppg_extract_from_maps).  The fused variant is an option (PPG_FUSE_SCAN=1; it
    measured no faster than the separate scan kernel, which is the default)."""
    from ppg_slam_b200 import capi
    monkeypatch.setenv("PPG_FUSE_SCAN", "0" if variant == "scan-kernel" else "1")
    if variant == "fused-scan-global-nms":
        monkeypatch.setenv("PPG_NMS_GLOBAL", "1")
    cam = cameras.EUROC
    frames = [synth.frame(s, cam.width, cam.height) for s in (0, 5, 9)]
    frames.append(np.full((cam.height, cam.width), 90, np.uint8))  # flat frame: no keypoints
    e = capi.Extractor(cam, max_batch=4)
    try:
        full = e.run(frames)
        maps = [e.get_maps(f) for f in range(4)]
        again = e.run_from_maps(np.stack([m["prob"] for m in maps]), np.stack([m["heat"] for m in maps]),
                                np.stack([m["desc"] for m in maps]))
        for f in range(4):
            a, b = full[f], again[f]
            assert a["n_kp"] == b["n_kp"] and a["n_edges"] == b["n_edges"] and a["n_cand"] == b["n_cand"]
            for k in ("px", "py", "out", "edge_start", "edge_end", "conn_off", "conn_idx", "col_off", "col_pairs"):
                np.testing.assert_array_equal(a[k], b[k], err_msg="%s frame %d" % (k, f))
            np.testing.assert_array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32))
            np.testing.assert_array_equal(a["desc"].view(np.uint32), b["desc"].view(np.uint32))
        assert full[0]["n_kp"] > 200 and full[3]["n_kp"] == 0
    finally:
        e.close()
