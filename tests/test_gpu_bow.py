"""GPU parity of the bag-of-words transform (DBoW3::Vocabulary::transform as Frame::ComputeBoW calls it,
map/src/Frame.cpp:331-340) against the CPU restatement: word / weight / FeatureVector node of every feature and the
BowVector (keys and double values) bit for bit -- on the reference's own EuRoC vocabulary and on synthetic trees."""
import os

import numpy as np
import pytest

from ppg_slam_b200 import cameras, synth, vocabulary

pytestmark = pytest.mark.gpu
WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ppg_slam_b200", "weights")


def _same(got, ref):
    np.testing.assert_array_equal(got["word"], ref["word"])
    np.testing.assert_array_equal(got["weight"].view(np.uint64), ref["weight"].view(np.uint64))
    np.testing.assert_array_equal(got["node"], ref["node"])
    np.testing.assert_array_equal(got["bow_word"], ref["bow_word"])
    np.testing.assert_array_equal(got["bow_value"].view(np.uint64), ref["bow_value"].view(np.uint64))


@pytest.mark.parametrize("name,levelsup,n", [("euroc", 4, 500), ("tum", 4, 357), ("euroc", 2, 500), ("k4L2-l1", 1, 200),
                                              ("k3L4-dot", 2, 1000), ("k32L1", 0, 64)])
def test_bow_transform_equals_oracle(name, levelsup, n):
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    if name in ("euroc", "tum"):
        voc = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_%s_9x3.bin" % name))
    elif name == "k4L2-l1":
        voc = vocabulary.random_vocabulary(1, 4, 2, scoring=0, zero_weight_frac=0.3)
    elif name == "k3L4-dot":
        voc = vocabulary.random_vocabulary(2, 3, 4, scoring=5, zero_weight_frac=0.3)
    else:
        voc = vocabulary.random_vocabulary(3, 32, 1, scoring=1, zero_weight_frac=0.2)
    rs = np.random.RandomState(7)
    # descriptors near the vocabulary's own nodes (so that many words are hit) plus pure noise
    leaves = np.nonzero(voc.word_id >= 0)[0]
    feats = voc.desc[leaves[rs.randint(0, len(leaves), n)]] + rs.normal(0, 0.03, (n, 256)).astype(np.float32)
    feats[::7] = rs.normal(size=(len(feats[::7]), 256))
    feats = (feats / np.linalg.norm(feats, axis=1, keepdims=True)).astype(np.float32)
    feats[3] = feats[2]
    ref = O.bow_transform(voc, feats, levelsup)
    e = capi.Extractor(cameras.EUROC, max_batch=1, junction_max_num=1000)
    try:
        e.upload_vocabulary(voc)
        got = e.bow_transform(feats, levelsup)
        _same(got, ref)
        assert len(ref["bow_word"]) > 3 and len(ref["bow_word"]) < n  # words are shared, weights accumulate
        got = e.bow_transform(feats[:0], levelsup)
        assert len(got["word"]) == 0 and len(got["bow_word"]) == 0
    finally:
        e.close()


def test_bow_batch_on_extracted_frames():
    """Frame::ComputeBoW for every frame of an extraction batch with the descriptors left on the device."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    cam = cameras.EUROC
    voc = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))
    Bn = 3
    e = capi.Extractor(cam, max_batch=Bn)
    try:
        recs = e.run([synth.frame(s, cam.width, cam.height) for s in range(Bn)])
        e.load_vocabulary(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))  # the library's own reader (ppg_load_vocabulary)
        e.bow_run_batch(Bn, 4)
        got = e.bow_fetch_batch(Bn)
        for f in range(Bn):
            ref = O.bow_transform(voc, recs[f]["desc"], 4)
            _same(got[f], ref)
            assert len(got[f]["word"]) == recs[f]["n_kp"] and len(ref["bow_word"]) > 20
    finally:
        e.close()


def test_search_by_bow_core_equals_oracle():
    """PPG_SEARCH_NODE: the frozen-state core of Matcher::SearchByBoW (Matcher.cpp:393-477) with the FeatureVector
    nodes the GPU transform produced for both the keyframe's and the frame's descriptors."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    voc = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))
    rs = np.random.RandomState(11)
    n, m = 480, 400
    leaves = np.nonzero(voc.word_id >= 0)[0]
    fd = voc.desc[leaves[rs.randint(0, len(leaves), n)]] + rs.normal(0, 0.05, (n, 256)).astype(np.float32)
    fd = (fd / np.linalg.norm(fd, axis=1, keepdims=True)).astype(np.float32)
    kd = fd[rs.randint(0, n, m)] + rs.normal(0, 0.02, (m, 256)).astype(np.float32)
    kd = (kd / np.linalg.norm(kd, axis=1, keepdims=True)).astype(np.float32)
    free = (rs.rand(n) > 0.1).astype(np.uint8)
    e = capi.Extractor(cameras.EUROC, max_batch=1, max_map_points=1024)
    try:
        e.upload_vocabulary(voc)
        for levelsup in (4, 2):  # 4: everything under the root (the reference's setting), 2: level-1 nodes
            kp_node = e.bow_transform(fd, levelsup)["node"]
            row_node = e.bow_transform(kd, levelsup)["node"]
            e.upload_map(kd)
            z = np.zeros(n, np.float32)
            got = e.associate(z, z, fd, free, np.zeros((m, 2), np.float32), np.zeros(m, np.float32), 0.0, 0.8,
                              mode=capi.SEARCH_NODE, max_dist=0.7, row_node=row_node, kp_node=kp_node)
            ref = O.search_node_all(fd, free, kp_node, kd, row_node, 0.8, 0.7)
            np.testing.assert_array_equal(got["best_idx"], ref["best_idx"])
            np.testing.assert_array_equal(got["second_idx"], ref["second_idx"])
            np.testing.assert_array_equal(got["best_d"].view(np.uint32), ref["best_d"].view(np.uint32))
            np.testing.assert_array_equal(got["second_d"].view(np.uint32), ref["second_d"].view(np.uint32))
            np.testing.assert_array_equal(got["accept"], ref["accept"])
            assert ref["accept"].sum() > 50
    finally:
        e.close()


@pytest.mark.parametrize("strict,n,m,levelsup", [(False, 480, 400, 4), (True, 480, 400, 4), (False, 1000, 900, 2),
                                                  (False, 60, 0, 4)])
def test_search_by_bow_whole_equals_oracle(strict, n, m, levelsup):
    """ppg_search_by_bow: the whole Matcher::SearchByBoW (live vpMapPointMatches) on the GPU == the oracle."""
    from oracle import post_ref as O
    from ppg_slam_b200 import capi
    voc = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))
    rs = np.random.RandomState(13)
    leaves = np.nonzero(voc.word_id >= 0)[0]
    fd = voc.desc[leaves[rs.randint(0, len(leaves), n)]] + rs.normal(0, 0.05, (n, 256)).astype(np.float32)
    fd = (fd / np.linalg.norm(fd, axis=1, keepdims=True)).astype(np.float32)
    mm = max(m, 1)
    # three keyframe features per frame feature on average: the later ones find their best taken
    kd = fd[rs.randint(0, max(n // 3, 1), mm)] + rs.normal(0, 0.03, (mm, 256)).astype(np.float32)
    kd = (kd / np.linalg.norm(kd, axis=1, keepdims=True)).astype(np.float32)
    e = capi.Extractor(cameras.EUROC, max_batch=1, max_map_points=1024, junction_max_num=1000)
    try:
        e.upload_vocabulary(voc)
        kp_node = e.bow_transform(fd, levelsup)["node"]
        row_node = e.bow_transform(kd, levelsup)["node"]
        order = np.argsort(np.where(row_node < 0, 1 << 30, row_node), kind="stable")  # visiting order: node, then index
        kd, row_node = kd[order], row_node[order]
        if m == 0:
            row_node[:] = -1  # no keyframe feature holds a map point
        e.upload_map(kd)
        got = e.search_by_bow(row_node, fd, kp_node, 0.8, 0.7, strict)
        ref = O.search_by_bow(fd, kp_node, kd, row_node, 0.8, 0.7, strict)
        assert got["nmatches"] == ref["nmatches"]
        np.testing.assert_array_equal(got["kp_row"], ref["kp_row"])
        if m:
            assert ref["nmatches"] > 50
    finally:
        e.close()
