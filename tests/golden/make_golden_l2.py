"""Golden vectors of the association level ("L2"): outputs of the CPU oracle for the whole
Matcher::ExtendMapMatches, Frame::CheckInFrustum and the bag-of-words transform / SearchByBoW on small seeded inputs
(inputs from tests/golden_l2.py::inputs, stored in the fixture together with the outputs).  The oracle
functions are themselves pinned by the statement-by-statement restatements in tests/test_oracle_extend.py and
tests/test_oracle_bow.py; this file freezes their answers so that a later change of the oracle (or of the synthetic
generators) cannot go unnoticed, and gives the GPU tests a target that needs no oracle at run time.
python tests/golden/make_golden_l2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import post_ref as O  # noqa: E402
from tests.golden_l2 import inputs, pack  # noqa: E402


def main():
    x = inputs()
    out = pack(x)
    r = O.extend_map_matches(x["cam"], x["ext"]["map_desc"], x["ext"]["candidate"], x["ext"]["observed"],
                             x["ext"]["bad"], x["ext"]["edge_off"], x["ext"]["edge_other"], x["ext"]["edge_ok"],
                             x["ext"]["proj_uv"], x["ext"]["view_cos"], x["ext"]["tracked"], x["kx"], x["ky"], x["fd"],
                             x["ext"]["kp_mp"], x["es"], x["ee"], x["coff"], x["cidx"], th=10.0, ratio=0.8)
    out.update(ext_nmatches=r["nmatches"], ext_kp_mp=r["kp_mp"], ext_kedge_me=r["kedge_me"], ext_tracked=r["tracked"])
    g = x["geo"]
    f = O.check_in_frustum(x["cam"], g["Rcw"][0], g["tcw"][0], g["Ow"][0], g["world_pos"], g["normal"], g["min_dist"],
                           g["max_dist"], 0.5)
    out.update(fr_in_view=f["in_view"], fr_proj=f["proj_uv"], fr_depth=f["depth"], fr_cos=f["view_cos"])
    b = O.bow_transform(x["voc"], x["fd"], 4)
    out.update(bow_word=b["word"], bow_weight=b["weight"], bow_node=b["node"], bow_vec_word=b["bow_word"],
               bow_vec_value=b["bow_value"])
    kb = O.bow_transform(x["voc"], x["kd"], 4)
    m = O.search_by_bow(x["fd"], b["node"], x["kd"], kb["node"], 0.8, 0.7, False)
    out.update(sbb_kp_row=m["kp_row"], sbb_nmatches=m["nmatches"])
    path = os.path.join(ROOT, "tests", "golden", "l2_association.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", r["nmatches"], int(f["in_view"].sum()), len(b["bow_word"]),
          m["nmatches"])


if __name__ == "__main__":
    main()
