"""Association-level golden vectors (tests/golden/l2_association.npz, made by tests/golden/make_golden_l2.py):
the oracle must still give them (CPU), and the CUDA path must give them without any oracle at run time (GPU)."""
import os

import numpy as np
import pytest

from tests.golden_l2 import load

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "l2_association.npz")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64) if a.dtype == np.float64 else (a.view(np.uint32) if a.dtype == np.float32 else a)


def test_oracle_reproduces_the_l2_golden_vectors():
    from oracle import post_ref as O
    x, g = load(GOLD)
    e = x["ext"]
    r = O.extend_map_matches(x["cam"], e["map_desc"], e["candidate"], e["observed"], e["bad"], e["edge_off"],
                             e["edge_other"], e["edge_ok"], e["proj_uv"], e["view_cos"], e["tracked"], x["kx"], x["ky"],
                             x["fd"], e["kp_mp"], x["es"], x["ee"], x["coff"], x["cidx"], th=10.0, ratio=0.8)
    assert r["nmatches"] == int(g["ext_nmatches"])
    np.testing.assert_array_equal(r["kp_mp"], g["ext_kp_mp"])
    np.testing.assert_array_equal(r["kedge_me"], g["ext_kedge_me"])
    np.testing.assert_array_equal(r["tracked"], g["ext_tracked"])
    q = x["geo"]
    f = O.check_in_frustum(x["cam"], q["Rcw"][0], q["tcw"][0], q["Ow"][0], q["world_pos"], q["normal"], q["min_dist"],
                           q["max_dist"], 0.5)
    np.testing.assert_array_equal(f["in_view"], g["fr_in_view"])
    for k, gk in (("proj_uv", "fr_proj"), ("depth", "fr_depth"), ("view_cos", "fr_cos")):
        np.testing.assert_array_equal(_bits(f[k]), _bits(g[gk]))
    b = O.bow_transform(x["voc"], x["fd"], 4)
    np.testing.assert_array_equal(b["word"], g["bow_word"])
    np.testing.assert_array_equal(_bits(b["bow_value"]), _bits(g["bow_vec_value"]))
    kb = O.bow_transform(x["voc"], x["kd"], 4)
    m = O.search_by_bow(x["fd"], b["node"], x["kd"], kb["node"], 0.8, 0.7, False)
    assert m["nmatches"] == int(g["sbb_nmatches"])
    np.testing.assert_array_equal(m["kp_row"], g["sbb_kp_row"])


@pytest.mark.gpu
def test_cuda_path_reproduces_the_l2_golden_vectors():
    from ppg_slam_b200 import capi
    x, g = load(GOLD)
    e = x["ext"]
    ex = capi.Extractor(x["cam"], max_batch=1, max_map_points=1024)
    try:
        ex.upload_map(e["map_desc"])
        ex.upload_map_graph(e["candidate"], e["observed"], e["bad"], e["edge_off"], e["edge_other"], e["edge_ok"])
        r = ex.extend_map_matches(x["kx"], x["ky"], x["fd"], e["kp_mp"], x["es"], x["ee"], x["coff"], x["cidx"],
                                  e["proj_uv"], e["view_cos"], e["tracked"], 10.0, 0.8)
        assert r["nmatches"] == int(g["ext_nmatches"])
        np.testing.assert_array_equal(r["kp_mp"], g["ext_kp_mp"])
        np.testing.assert_array_equal(r["kedge_me"], g["ext_kedge_me"])
        np.testing.assert_array_equal(r["tracked"], g["ext_tracked"])
        q = x["geo"]
        ex.upload_map_geometry(q["world_pos"], q["normal"], q["min_dist"], q["max_dist"])
        ex.assoc_stage_poses(q["Rcw"], q["tcw"], q["Ow"], len(q["min_dist"]), 0.5, 10.0, 0.8)
        f = ex.frustum_fetch(1)
        np.testing.assert_array_equal(f["in_view"][0], g["fr_in_view"])
        for k, gk in (("proj_uv", "fr_proj"), ("depth", "fr_depth"), ("view_cos", "fr_cos")):
            np.testing.assert_array_equal(_bits(f[k][0]), _bits(g[gk]))
        ex.upload_vocabulary(x["voc"])
        b = ex.bow_transform(x["fd"], 4)
        np.testing.assert_array_equal(b["word"], g["bow_word"])
        np.testing.assert_array_equal(_bits(b["weight"]), _bits(g["bow_weight"]))
        np.testing.assert_array_equal(b["node"], g["bow_node"])
        np.testing.assert_array_equal(b["bow_word"], g["bow_vec_word"])
        np.testing.assert_array_equal(_bits(b["bow_value"]), _bits(g["bow_vec_value"]))
        kb = ex.bow_transform(x["kd"], 4)
        ex.upload_map(x["kd"])
        m = ex.search_by_bow(kb["node"], x["fd"], b["node"], 0.8, 0.7)
        assert m["nmatches"] == int(g["sbb_nmatches"])
        np.testing.assert_array_equal(m["kp_row"], g["sbb_kp_row"])
    finally:
        ex.close()
