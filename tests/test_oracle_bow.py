"""Known-answer tests of the bag-of-words path: the vocabulary reader against the reference's own files (when the
reference tree is present) and against the committed blobs, and the transform oracle (oracle/ppg_oracle.c::
ppgo_bow_transform) against a restatement of DBoW3::Vocabulary::transform written the way DBoW3 is (std::map ->
dict kept sorted, per-feature descent, BowVector::normalize)."""
import math
import os

import numpy as np
import pytest

from oracle import post_ref as O
from ppg_slam_b200 import vocabulary

WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ppg_slam_b200", "weights")
REF_VOC = "/root/reference/Vocabulary"


def py_transform(voc, feats, levelsup):
    ch = voc.child_table()
    nid_level = voc.L - levelsup
    bow, fv = {}, {}
    words, weights = [], []
    for i, f in enumerate(feats):
        final_id, level, nid = 0, 0, 0
        while True:
            level += 1
            best = float("inf")
            nodes = [c for c in ch[final_id] if c >= 0]
            for c in nodes:
                diff = (f - voc.desc[c]).astype(np.float32)
                sq = (diff * diff).astype(np.float32)
                d = 0.0
                for x in sq:  # double accumulation in index order
                    d += float(x)
                if d < best:
                    best, final_id = d, c
            if level == nid_level:
                nid = final_id
            if ch[final_id][0] < 0:
                break
        w = float(voc.weight[final_id])
        words.append(int(voc.word_id[final_id]))
        weights.append(w)
        if w > 0:
            bow[words[-1]] = bow.get(words[-1], 0.0) + w
            fv.setdefault(nid, []).append(i)
    keys = sorted(bow)
    vals = [bow[k] for k in keys]
    if keys:
        if voc.scoring == 5:
            vals = [v / float(len(keys)) for v in vals]
        else:
            norm = 0.0
            if voc.scoring == 1:
                for v in vals:
                    norm += v * v
                norm = math.sqrt(norm)
            else:
                for v in vals:
                    norm += abs(v)
            if norm > 0:
                vals = [v / norm for v in vals]
    return words, weights, keys, vals, fv


@pytest.mark.parametrize("k,L,scoring,levelsup", [(9, 3, 1, 4), (4, 2, 0, 1), (3, 4, 5, 2), (9, 3, 1, 2)])
def test_bow_transform_kat(k, L, scoring, levelsup):
    voc = vocabulary.random_vocabulary(k * 10 + L, k, L, scoring=scoring)
    rs = np.random.RandomState(1)
    feats = rs.normal(size=(70, 256)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    feats[5] = feats[4]  # two features in the same word: weights accumulate
    got = O.bow_transform(voc, feats, levelsup)
    words, weights, keys, vals, fv = py_transform(voc, feats, levelsup)
    assert got["word"].tolist() == words
    assert got["weight"].tolist() == weights
    assert got["bow_word"].tolist() == keys
    assert got["bow_value"].tolist() == vals  # bit-exact doubles
    want_node = np.full(len(feats), -1)
    for nid, idx in fv.items():
        want_node[idx] = nid
    assert got["node"].tolist() == want_node.tolist()
    assert len(keys) >= 2 and len(set(words)) < len(words)  # shared words accumulate


def test_committed_vocabulary_blobs_are_the_reference_files():
    for name in ("voc_euroc_9x3", "voc_tum_9x3"):
        b = vocabulary.load_blob(os.path.join(WEIGHTS, name + ".bin"))
        assert (b.k, b.L, b.scoring, b.weighting, b.n_nodes, b.n_words) == (9, 3, 1, 0, 820, 729)
        nch = (b.children >= 0).sum(1)
        assert sorted(set(nch.tolist())) == [0, 9] and (nch == 9).sum() == 91
        assert sorted(b.word_id[b.word_id >= 0].tolist()) == list(range(729))
        assert ((b.word_id >= 0) == (nch == 0)).all()  # words are exactly the leaves
        src = os.path.join(REF_VOC, name + ".gz")
        if os.path.exists(src):  # in the build container: the blob is a faithful export of the reference file
            v = vocabulary.Vocabulary(src)
            np.testing.assert_array_equal(v.child_table(), b.children)
            np.testing.assert_array_equal(v.word_id, b.word_id)
            np.testing.assert_array_equal(v.weight, b.weight)
            np.testing.assert_array_equal(v.desc, b.desc)


def test_real_vocabulary_transform_is_consistent():
    """On the reference's EuRoC vocabulary: every feature lands on a leaf, levelsup = 4 puts all features with a
    positive weight under the root (L = 3), the BowVector has unit L2 norm."""
    voc = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))
    rs = np.random.RandomState(3)
    feats = voc.desc[rs.randint(91, 820, 200)] + rs.normal(0, 0.02, (200, 256)).astype(np.float32)
    got = O.bow_transform(voc, feats.astype(np.float32), 4)
    assert (got["word"] >= 0).all() and (got["word"] < 729).all()
    assert set(got["node"].tolist()) <= {0, -1}
    assert abs(np.sqrt((got["bow_value"] ** 2).sum()) - 1.0) < 1e-12
    assert (np.diff(got["bow_word"]) > 0).all()


def test_search_by_bow_core_kat():
    """Inner loop of Matcher::SearchByBoW(KF, F) (Matcher.cpp:421-461) restated with the FeatureVector as a dict."""
    rs = np.random.RandomState(5)
    n, m = 150, 120
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    kd = fd[rs.randint(0, n, m)] + rs.normal(0, 0.03, (m, 256)).astype(np.float32)
    kd = (kd / np.linalg.norm(kd, axis=1, keepdims=True)).astype(np.float32)
    kp_node = rs.choice([-1, 0, 0, 0, 4, 7], n).astype(np.int32)
    row_node = rs.choice([-1, 0, 0, 4, 7, 9], m).astype(np.int32)
    free = (rs.rand(n) > 0.15).astype(np.uint8)
    got = O.search_node_all(fd, free, kp_node, kd, row_node, 0.8, 0.7)
    fv = {}
    for i, nd in enumerate(kp_node.tolist()):
        if nd >= 0:
            fv.setdefault(nd, []).append(i)
    n_acc = 0
    for j in range(m):
        best1, best2, bi = 1e6, 1e6, -1
        for idx in fv.get(int(row_node[j]), []) if row_node[j] >= 0 else []:
            if not free[idx]:
                continue
            d = O.descriptor_distance(kd[j], fd[idx])
            if d < best1:
                best2, best1, bi = best1, d, idx
            elif d < best2:
                best2 = d
        acc = int(bi >= 0 and best1 <= np.float32(0.7) and best1 < float(np.float32(np.float32(0.8) * np.float32(best2))))
        assert got["best_idx"][j] == bi and got["accept"][j] == acc
        if bi >= 0:
            assert got["best_d"][j] == np.float32(best1) and got["second_d"][j] == np.float32(best2)
        n_acc += acc
    assert n_acc > 20 and (got["best_idx"] < 0).sum() > 5


def _bow_match_case(seed, n=160, m=130, nodes=(0,)):
    rs = np.random.RandomState(seed)
    fd = rs.normal(size=(n, 256)).astype(np.float32)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    # several keyframe features near the same frame feature: later rows must take their second choice or fail
    kd = fd[rs.randint(0, n // 3, m)] + rs.normal(0, 0.04, (m, 256)).astype(np.float32)
    kd = (kd / np.linalg.norm(kd, axis=1, keepdims=True)).astype(np.float32)
    kp_node = rs.choice(list(nodes) + [-1], n).astype(np.int32)
    row_node = np.sort(rs.choice(list(nodes), m)).astype(np.int32)
    row_node[rs.rand(m) < 0.05] = -1
    return fd, kp_node, kd, row_node


@pytest.mark.parametrize("strict,nodes", [(False, (0,)), (True, (0,)), (False, (3, 5, 9))])
def test_search_by_bow_whole_kat(strict, nodes):
    """Matcher::SearchByBoW (Matcher.cpp:393-477 / :663-754) with the FeatureVectors as ordered dicts and the live
    vpMapPointMatches vector, statement by statement."""
    fd, kp_node, kd, row_node = _bow_match_case(4, nodes=nodes)
    got = O.search_by_bow(fd, kp_node, kd, row_node, 0.8, 0.7, strict)
    fvF, fvK = {}, {}
    for i, nd in enumerate(kp_node.tolist()):
        if nd >= 0:
            fvF.setdefault(nd, []).append(i)
    for j, nd in enumerate(row_node.tolist()):
        if nd >= 0:
            fvK.setdefault(nd, []).append(j)
    matches = [None] * len(kp_node)
    nm = 0
    for node in sorted(set(fvF) & set(fvK)):
        for j in fvK[node]:
            best1, best2, bi = 1e6, 1e6, -1
            for idx in fvF[node]:
                if matches[idx] is not None:
                    continue
                d = O.descriptor_distance(kd[j], fd[idx])
                if d < best1:
                    best2, best1, bi = best1, d, idx
                elif d < best2:
                    best2 = d
            ok = best1 < np.float32(0.7) if strict else best1 <= np.float32(0.7)
            if ok and best1 < float(np.float32(np.float32(0.8) * np.float32(best2))):
                matches[bi] = j
                nm += 1
    assert got["nmatches"] == nm
    assert got["kp_row"].tolist() == [(-1 if x is None else x) for x in matches]
    assert nm > 15


def test_library_vocabulary_reader_matches_python_reader(tmp_path):
    """ppg_vocabulary_open (C++ host code of libppg_b200.so, no GPU): the committed blob, an uncompressed DBoW3 stream
    written here, and -- in the build container -- the reference's own QuickLZ-compressed files."""
    import struct
    from ppg_slam_b200 import capi
    blob = vocabulary.load_blob(os.path.join(WEIGHTS, "voc_euroc_9x3.bin"))

    def same(a, b):
        assert (a.k, a.L, a.scoring, a.weighting, a.n_nodes) == (b.k, b.L, b.scoring, b.weighting, b.n_nodes)
        np.testing.assert_array_equal(a.child_table(), b.child_table())
        np.testing.assert_array_equal(a.word_id, b.word_id)
        np.testing.assert_array_equal(a.weight, b.weight)
        np.testing.assert_array_equal(a.desc[1:], b.desc[1:])  # the root has no descriptor in the file

    same(capi.read_vocabulary(os.path.join(WEIGHTS, "voc_euroc_9x3.bin")), blob)
    # DBoW3 stream, uncompressed: depth-first node records in the order Vocabulary::toStream writes them
    v = vocabulary.random_vocabulary(5, 3, 2)
    out = struct.pack("<QBI", vocabulary.SIGNATURE, 0, v.n_nodes) + struct.pack("<iiii", v.k, v.L, v.scoring, v.weighting)
    parents = [0]
    while parents:
        pid = parents.pop()
        for c in v.children[pid]:
            if c < 0:
                continue
            out += struct.pack("<IId", c, pid, v.weight[c]) + struct.pack("<iii", 256, 1, 5) + v.desc[c].tobytes()
            if v.children[c][0] >= 0:
                parents.append(c)
    words = np.nonzero(v.word_id >= 0)[0]
    out += struct.pack("<I", len(words)) + b"".join(struct.pack("<II", int(v.word_id[n]), int(n)) for n in words)
    path = tmp_path / "voc_raw.bin"
    path.write_bytes(out)
    got = capi.read_vocabulary(str(path))
    same(got, v)
    same(vocabulary.Vocabulary(str(path)), v)  # the Python reader agrees
    bad = tmp_path / "bad.bin"
    bad.write_bytes(out[:-3])
    with pytest.raises(capi.PpgError):
        capi.read_vocabulary(str(bad))
    for name in ("voc_euroc_9x3", "voc_tum_9x3"):
        src = os.path.join(REF_VOC, name + ".gz")
        if os.path.exists(src):
            same(capi.read_vocabulary(src), vocabulary.load_blob(os.path.join(WEIGHTS, name + ".bin")))
