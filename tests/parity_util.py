"""Shared helpers of the GPU parity tests: run the oracle and compare records field by field."""
import numpy as np

from oracle import post_ref as O

INT_FIELDS = ["px", "py", "out", "edge_start", "edge_end", "conn_off", "conn_idx", "col_off", "col_pairs"]
F32_EXACT_FIELDS = ["score", "xun", "yun", "kp_x", "kp_y", "edge_score"]


def oracle_post(cam, prob, heat, desc, **over):
    return O.extract_post(cam, prob, heat, desc, **over)


def diff_records(got, ref, desc_tol=1e-6):
    """-> list of human-readable mismatches (empty = parity).  Integer/index fields and the float fields
    that are pure copies or fixed-order arithmetic must be bit-identical (NaN == NaN by bit pattern)."""
    bad = []
    for k in ("n_kp", "n_edges"):
        if int(got[k]) != int(ref[k]):
            bad.append("%s: got %d want %d" % (k, got[k], ref[k]))
    if bad:
        return bad
    if int(got["n_colines"]) != len(ref["col_pairs"]):
        bad.append("n_colines: got %d want %d" % (got["n_colines"], len(ref["col_pairs"])))
    for k in INT_FIELDS:
        a, b = np.asarray(got[k]).astype(np.int64), np.asarray(ref[k]).astype(np.int64)
        if a.shape != b.shape or not np.array_equal(a, b):
            bad.append("%s differs (%s vs %s)" % (k, a.shape, b.shape))
    for k in F32_EXACT_FIELDS:
        a, b = np.asarray(got[k], np.float32), np.asarray(ref[k], np.float32)
        # bit-identical, except that any NaN matches any NaN (x86 0/0 gives the negative quiet NaN, the GPU the
        # positive one; lscore of the 5 <= dist < 6 lines is NaN in the reference too, PPGExtractor.cpp:376)
        same = a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))
        if not same:
            n = int(((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))).sum()) if a.shape == b.shape else -1
            bad.append("%s differs in %d entries" % (k, n))
    if got["n_kp"] > 0:
        d = np.abs(np.asarray(got["desc"]) - np.asarray(ref["desc"])).max()
        if not d <= desc_tol:
            bad.append("desc max abs diff %g > %g" % (d, desc_tol))
    return bad
