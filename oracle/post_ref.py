"""ctypes front of oracle/ppg_oracle.c: L1 (post-processing) and L2 (association) oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows feature/src/PPGExtractor.cpp:118-147
(run) for the orchestration: detectKeyPoint -> detectLines -> genPointDescriptor, then the
pinhole-only `mPos = mPosUn` copy (:141-145).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build


class Cfg(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("K", C.c_float * 9), ("D", C.c_float * 4),
                ("fisheye", C.c_int), ("junction_thresh", C.c_float), ("junction_nms_radius", C.c_int),
                ("junction_max_num", C.c_int), ("line_valid_thresh", C.c_float),
                ("line_valid_ratio", C.c_float), ("line_dist_thresh", C.c_float),
                ("heatmap_refine_sz", C.c_int), ("line_heatmap_thresh", C.c_float),
                ("line_inlier_rate", C.c_float)]


class Bounds(C.Structure):
    _fields_ = [("minX", C.c_int), ("minY", C.c_int), ("maxX", C.c_int), ("maxY", C.c_int),
                ("wInv", C.c_float), ("hInv", C.c_float)]


_libs = {}


def lib(variant=""):
    if variant not in _libs:
        _build.build()
        _libs[variant] = C.CDLL(_build.lib_path(variant))
        _libs[variant].ppgo_descriptor_distance.restype = C.c_float
    return _libs[variant]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_cfg(cam, **over):
    c = Cfg()
    lib().ppgo_default_cfg(C.byref(c))
    c.width, c.height, c.fisheye = cam.width, cam.height, int(cam.fisheye)
    c.K[:] = cam.K
    c.D[:] = cam.D
    for k, v in over.items():
        setattr(c, k, v)
    return c


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def undistort_points(cam, xy):
    xy = f32(xy).reshape(-1, 2)
    out = np.empty_like(xy)
    K, D = f32(cam.K), f32(cam.D)
    fn = lib().ppgo_undistort_points_fisheye if cam.fisheye else lib().ppgo_undistort_points_pinhole
    fn(_p(K, C.c_float), _p(D, C.c_float), _p(xy, C.c_float), len(xy), _p(out, C.c_float))
    return out


def init_undistort_map(cam):
    mx = np.empty((cam.height, cam.width), np.float32)
    my = np.empty_like(mx)
    K, D = f32(cam.K), f32(cam.D)
    fn = lib().ppgo_init_undistort_map_fisheye if cam.fisheye else lib().ppgo_init_undistort_map_pinhole
    fn(_p(K, C.c_float), _p(D, C.c_float), cam.width, cam.height, _p(mx, C.c_float), _p(my, C.c_float))
    return mx, my


def remap_linear(src, mx, my):
    src = f32(src)
    H, W = src.shape
    dst = np.empty_like(src)
    lib().ppgo_remap_linear(_p(src, C.c_float), W, H, _p(f32(mx), C.c_float), _p(f32(my), C.c_float),
                            _p(dst, C.c_float))
    return dst


def detect_keypoints(cfg, prob):
    prob = f32(prob)
    cap = max(cfg.junction_max_num, 1)
    kx, ky = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    sc, xu, yu = np.zeros(cap, np.float32), np.zeros(cap, np.float32), np.zeros(cap, np.float32)
    out = np.zeros(cap, np.uint8)
    nc = C.c_int(0)
    n = lib().ppgo_detect_keypoints(C.byref(cfg), _p(prob, C.c_float), _p(kx, C.c_int), _p(ky, C.c_int),
                                    _p(sc, C.c_float), _p(xu, C.c_float), _p(yu, C.c_float),
                                    _p(out, C.c_uint8), C.byref(nc))
    return dict(n=n, x=kx[:n].copy(), y=ky[:n].copy(), score=sc[:n].copy(), xun=xu[:n].copy(),
                yun=yu[:n].copy(), out=out[:n].copy(), n_cand=nc.value)


def refine_heat(cfg, heat):
    h = f32(heat).copy()
    lib().ppgo_refine_heat(C.byref(cfg), _p(h, C.c_float))
    return h


def detect_lines(cfg, heat_final, kp, max_edges=131072, max_col=131072, variant=""):
    n = kp["n"]
    heat_final = f32(heat_final)
    es, ee = np.zeros(max_edges, np.int32), np.zeros(max_edges, np.int32)
    esc = np.zeros(max_edges, np.float32)
    coff, cidx = np.zeros(n + 1, np.int32), np.zeros(2 * max_edges, np.int32)
    loff, lpairs = np.zeros(n + 1, np.int32), np.zeros(2 * max_col, np.int32)
    ncol = C.c_int(0)
    stats = np.zeros(4, np.int32)
    xun, yun, kout = f32(kp["xun"]), f32(kp["yun"]), np.ascontiguousarray(kp["out"], np.uint8)
    E = lib(variant).ppgo_detect_lines(C.byref(cfg), _p(heat_final, C.c_float), n, _p(xun, C.c_float),
                                       _p(yun, C.c_float), _p(kout, C.c_uint8), max_edges, _p(es, C.c_int),
                                       _p(ee, C.c_int), _p(esc, C.c_float), _p(coff, C.c_int),
                                       _p(cidx, C.c_int), max_col, _p(loff, C.c_int), _p(lpairs, C.c_int),
                                       C.byref(ncol), _p(stats, C.c_int))
    assert E >= 0 and ncol.value >= 0, "oracle capacity exceeded"
    return dict(n_edges=E, edge_start=es[:E].copy(), edge_end=ee[:E].copy(), edge_score=esc[:E].copy(),
                conn_off=coff, conn_idx=cidx[:coff[n]].copy(), col_off=loff,
                col_pairs=lpairs[:2 * ncol.value].reshape(-1, 2).copy(), stats=stats)


def sample_descriptors(cfg, desc_chw, kx, ky):
    desc_chw = f32(desc_chw)
    Cc, Hc, Wc = desc_chw.shape
    n = len(kx)
    out = np.zeros((n, Cc), np.float32)
    kx, ky = np.ascontiguousarray(kx, np.int32), np.ascontiguousarray(ky, np.int32)
    lib().ppgo_sample_descriptors(C.byref(cfg), _p(desc_chw, C.c_float), Cc, Hc, Wc, n, _p(kx, C.c_int),
                                  _p(ky, C.c_int), _p(out, C.c_float))
    return out


_prev_edges = {}


def extract_post(cam, prob, heat_raw, desc_chw, maps=None, variant="", **over):
    """PPGExtractor::run minus the networks (PPGExtractor.cpp:126-146).

    prob (H,W) junction map, heat_raw (H,W) softmax[:,1] BEFORE refine, desc_chw (256,Hc,Wc).
    -> dict with the output record of run(): kp_x/kp_y = output mPos (pinhole: = mPosUn, :141-145),
    px/py = detection pixel, xun/yun, out, score, edges, CSR adjacency, colines, desc, plus the
    intermediate refined/remapped heat map (heat_final).
    """
    cfg = make_cfg(cam, **over)
    kp = detect_keypoints(cfg, prob)
    res = dict(n_kp=kp["n"], px=kp["x"], py=kp["y"], score=kp["score"], xun=kp["xun"], yun=kp["yun"],
               out=kp["out"], n_cand=kp["n_cand"])
    n = kp["n"]
    if n == 0:
        # detectLines returns before touching mvKeyEdges (:239-240): the reference leaks the previous
        # frame's edges here; the oracle (and the product) report zero edges.  DESIGN.md "divergences".
        res.update(n_edges=0, edge_start=np.zeros(0, np.int32), edge_end=np.zeros(0, np.int32),
                   edge_score=np.zeros(0, np.float32), conn_off=np.zeros(1, np.int32),
                   conn_idx=np.zeros(0, np.int32), col_off=np.zeros(1, np.int32),
                   col_pairs=np.zeros((0, 2), np.int32), heat_final=None,
                   desc=np.zeros((0, desc_chw.shape[0]), np.float32),
                   kp_x=np.zeros(0, np.float32), kp_y=np.zeros(0, np.float32))
        return res
    heat = refine_heat(cfg, heat_raw)
    if cam.D[0] != 0.0:  # :261
        mx, my = maps if maps is not None else init_undistort_map(cam)
        heat = remap_linear(heat, mx, my)
    res["heat_final"] = heat
    res.update(detect_lines(cfg, heat, kp, variant=variant))
    res["desc"] = sample_descriptors(cfg, desc_chw, kp["x"], kp["y"])
    if cam.fisheye:
        res["kp_x"], res["kp_y"] = kp["x"].astype(np.float32), kp["y"].astype(np.float32)
    else:
        res["kp_x"], res["kp_y"] = kp["xun"].copy(), kp["yun"].copy()
    return res


# ---------------------------------------------------------------- association (L2)
def image_bounds(cam):
    b = Bounds()
    cfg = make_cfg(cam)
    lib().ppgo_image_bounds(C.byref(cfg), C.byref(b))
    return b


def descriptor_distance(a, b):
    a, b = f32(a), f32(b)
    return float(lib().ppgo_descriptor_distance(_p(a, C.c_float), _p(b, C.c_float), a.size))


def features_in_area(cam, kx, ky, x, y, r):
    b = image_bounds(cam)
    kx, ky = f32(kx), f32(ky)
    n = len(kx)
    goff, gidx = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(n, 1), np.int32)
    lib().ppgo_grid_build(C.byref(b), n, _p(kx, C.c_float), _p(ky, C.c_float), _p(goff, C.c_int), _p(gidx, C.c_int))
    out = np.zeros(max(n, 1), np.int32)
    cnt = lib().ppgo_features_in_area(C.byref(b), _p(goff, C.c_int), _p(gidx, C.c_int), _p(kx, C.c_float),
                                      _p(ky, C.c_float), C.c_float(x), C.c_float(y), C.c_float(r),
                                      _p(out, C.c_int))
    return out[:cnt].copy()


def indexable(cam, kx, ky):
    b = image_bounds(cam)
    px, py = C.c_int(0), C.c_int(0)
    return np.array([lib().ppgo_pos_in_grid(C.byref(b), C.c_float(float(x)), C.c_float(float(y)),
                                            C.byref(px), C.byref(py)) for x, y in zip(kx, ky)], np.uint8)


def search_all(cam, kx, ky, frame_desc, free_mask, map_desc, proj_uv, view_cos, th, ratio, th_high=0.8, mode=0,
               max_dist=0.8, e2_max=0.0):
    """Search core of Matcher::ExtendMapMatches (Matcher.cpp:224-281, mode 0) or of the best-only projection
    matchers (SearchByProjection :31-87 / :1337-1411, Fuse :897-1036; mode 1: r = th, accept = best <= max_dist,
    optional e2_max) for every map point, frame state frozen.
    -> dict(best_idx, second_idx, best_d, second_d, accept)."""
    cfg = make_cfg(cam)
    kx, ky, frame_desc = f32(kx), f32(ky), f32(frame_desc)
    map_desc, proj_uv, view_cos = f32(map_desc), f32(proj_uv), f32(view_cos)
    free_mask = np.ascontiguousarray(free_mask, np.uint8)
    n, m = len(kx), len(map_desc)
    bi, si = np.zeros(m, np.int32), np.zeros(m, np.int32)
    bd, sd = np.zeros(m, np.float32), np.zeros(m, np.float32)
    acc = np.zeros(m, np.uint8)
    lib().ppgo_search_all(C.byref(cfg), n, _p(kx, C.c_float), _p(ky, C.c_float), _p(frame_desc, C.c_float),
                          _p(free_mask, C.c_uint8), m, _p(map_desc, C.c_float), _p(proj_uv, C.c_float),
                          _p(view_cos, C.c_float), C.c_float(th), C.c_float(ratio), C.c_float(th_high),
                          C.c_int(mode), C.c_float(max_dist), C.c_double(e2_max), _p(bi, C.c_int), _p(si, C.c_int), _p(bd, C.c_float), _p(sd, C.c_float),
                          _p(acc, C.c_uint8))
    return dict(best_idx=bi, second_idx=si, best_d=bd, second_d=sd, accept=acc)


def distinctive_all(desc, offsets):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cpp:234-302) for a batch of map points: desc (total, 256)
    packed observation descriptors, offsets (P + 1).  -> BestIdx per map point (index into its own list)."""
    desc = f32(desc)
    offsets = np.ascontiguousarray(offsets, np.int32)
    p = len(offsets) - 1
    out = np.zeros(p, np.int32)
    lib().ppgo_distinctive_all(_p(desc, C.c_float), _p(offsets, C.c_int), p, _p(out, C.c_int))
    return out


def extend_map_matches(cam, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos,
                       tracked, kx, ky, frame_desc, kp_mp, kedge_start, kedge_end, conn_off, conn_idx, kedge_me=None,
                       th=10.0, ratio=0.8, th_high=0.8):
    """Matcher::ExtendMapMatches (Matcher.cpp:203-381) in POD form, whole function (sequential walk + seed
    growing); argument meaning in oracle/ppg_oracle.c::ppgo_extend_map_matches.
    -> dict(nmatches, kp_mp, kedge_me, tracked) (inputs are not modified)."""
    cfg = make_cfg(cam)
    map_desc, proj_uv, view_cos = f32(map_desc), f32(proj_uv), f32(view_cos)
    kx, ky, frame_desc = f32(kx), f32(ky), f32(frame_desc)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    candidate, observed, bad, edge_ok = u8(candidate), u8(observed), u8(bad), u8(edge_ok)
    edge_off, edge_other = i32(edge_off), i32(edge_other)
    kedge_start, kedge_end, conn_off, conn_idx = i32(kedge_start), i32(kedge_end), i32(conn_off), i32(conn_idx)
    tracked = u8(tracked).copy()
    kp_mp = i32(kp_mp).copy()
    ne = len(kedge_start)
    kedge_me = np.full(max(ne, 1), -1, np.int32) if kedge_me is None else i32(kedge_me).copy()
    pad = lambda a, t: a if a.size else np.zeros(1, t)
    nm = lib().ppgo_extend_map_matches(
        C.byref(cfg), len(map_desc), _p(map_desc, C.c_float), _p(candidate, C.c_uint8), _p(observed, C.c_uint8),
        _p(bad, C.c_uint8), _p(edge_off, C.c_int), _p(pad(edge_other, np.int32), C.c_int),
        _p(pad(edge_ok, np.uint8), C.c_uint8), _p(proj_uv, C.c_float), _p(view_cos, C.c_float),
        _p(tracked, C.c_uint8), len(kx), _p(pad(kx, np.float32), C.c_float), _p(pad(ky, np.float32), C.c_float),
        _p(pad(frame_desc, np.float32), C.c_float), _p(pad(kp_mp, np.int32), C.c_int),
        _p(pad(kedge_start, np.int32), C.c_int), _p(pad(kedge_end, np.int32), C.c_int), _p(conn_off, C.c_int),
        _p(pad(conn_idx, np.int32), C.c_int), _p(kedge_me, C.c_int), C.c_float(th), C.c_float(ratio),
        C.c_float(th_high))
    return dict(nmatches=int(nm), kp_mp=kp_mp, kedge_me=kedge_me[:ne], tracked=tracked)


def check_in_frustum(cam, Rcw, tcw, Ow, world_pos, normal, min_dist, max_dist, cos_limit=0.5, variant=""):
    """Frame::CheckInFrustum (map/src/Frame.cpp:223-260) for every map point under one pose.
    -> dict(in_view u8, proj_uv (M,2), depth, view_cos)."""
    cfg = make_cfg(cam)
    Rcw, tcw, Ow = f32(Rcw).reshape(9), f32(tcw).reshape(3), f32(Ow).reshape(3)
    wp, nr, mn, mx = f32(world_pos), f32(normal), f32(min_dist), f32(max_dist)
    m = len(mn)
    iv, uv = np.zeros(m, np.uint8), np.zeros((m, 2), np.float32)
    dp, vc = np.zeros(m, np.float32), np.zeros(m, np.float32)
    lib(variant).ppgo_check_in_frustum_all(C.byref(cfg), _p(Rcw, C.c_float), _p(tcw, C.c_float), _p(Ow, C.c_float), m,
                                           _p(wp, C.c_float), _p(nr, C.c_float), _p(mn, C.c_float), _p(mx, C.c_float),
                                           C.c_float(cos_limit), _p(iv, C.c_uint8), _p(uv, C.c_float),
                                           _p(dp, C.c_float), _p(vc, C.c_float))
    return dict(in_view=iv, proj_uv=uv, depth=dp, view_cos=vc)


def bow_transform(voc, feat, levelsup=4):
    """DBoW3::Vocabulary::transform as the reference calls it (map/src/Frame.cpp:331-340).  voc: object with k, L,
    scoring, child_table(), weight, word_id, desc (ppg_slam_b200.vocabulary.Vocabulary or a synthetic stand-in).
    -> dict(word, weight, node per feature; bow_word, bow_value sorted by word)."""
    feat = f32(feat)
    n = feat.shape[0]
    ch = np.ascontiguousarray(voc.child_table(), np.int32)
    w = np.ascontiguousarray(voc.weight, np.float64)
    wid = np.ascontiguousarray(voc.word_id, np.int32)
    nd = f32(voc.desc)
    fw, fwt, fn = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
    bw, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    nb = lib().ppgo_bow_transform(voc.k, voc.L, voc.scoring, _p(ch, C.c_int), _p(w, C.c_double), _p(wid, C.c_int),
                                  _p(nd, C.c_float), nd.shape[1], n, _p(feat if n else np.zeros(1, np.float32), C.c_float),
                                  levelsup, _p(fw, C.c_int), _p(fwt, C.c_double), _p(fn, C.c_int), _p(bw, C.c_int),
                                  _p(bv, C.c_double))
    return dict(word=fw[:n], weight=fwt[:n], node=fn[:n], bow_word=bw[:nb].copy(), bow_value=bv[:nb].copy())


def search_node_all(frame_desc, free_mask, kp_node, row_desc, row_node, ratio, max_dist):
    """Frozen-state search core of Matcher::SearchByBoW (Matcher.cpp:421-461 / :700-740): candidates = frame
    features of the row's FeatureVector node.  -> dict(best_idx, second_idx, best_d, second_d, accept)."""
    frame_desc, row_desc = f32(frame_desc), f32(row_desc)
    free_mask = np.ascontiguousarray(free_mask, np.uint8)
    kp_node, row_node = np.ascontiguousarray(kp_node, np.int32), np.ascontiguousarray(row_node, np.int32)
    n, m = len(kp_node), len(row_node)
    bi, si = np.zeros(m, np.int32), np.zeros(m, np.int32)
    bd, sd = np.zeros(m, np.float32), np.zeros(m, np.float32)
    acc = np.zeros(m, np.uint8)
    lib().ppgo_search_node_all(n, _p(frame_desc, C.c_float), _p(free_mask, C.c_uint8), _p(kp_node, C.c_int), m,
                               _p(row_desc, C.c_float), _p(row_node, C.c_int), C.c_float(ratio), C.c_float(max_dist),
                               _p(bi, C.c_int), _p(si, C.c_int), _p(bd, C.c_float), _p(sd, C.c_float), _p(acc, C.c_uint8))
    return dict(best_idx=bi, second_idx=si, best_d=bd, second_d=sd, accept=acc)


def search_by_bow(frame_desc, kp_node, row_desc, row_node, ratio, max_dist, strict=False):
    """Matcher::SearchByBoW whole (Matcher.cpp:393-477, :663-754).  -> dict(kp_row, nmatches)."""
    frame_desc, row_desc = f32(frame_desc), f32(row_desc)
    kp_node, row_node = np.ascontiguousarray(kp_node, np.int32), np.ascontiguousarray(row_node, np.int32)
    kr = np.zeros(max(len(kp_node), 1), np.int32)
    nm = lib().ppgo_search_by_bow(len(kp_node), _p(frame_desc, C.c_float), _p(kp_node, C.c_int), len(row_node),
                                  _p(row_desc, C.c_float), _p(row_node, C.c_int), C.c_float(ratio),
                                  C.c_float(max_dist), int(strict), _p(kr, C.c_int))
    return dict(kp_row=kr[:len(kp_node)].copy(), nmatches=int(nm))


def search_for_initialization(cam, desc1, prev_matched, kx2, ky2, desc2, window=50, ratio=0.9, th_low=0.7):
    """Matcher::SearchForInitialization (Matcher.cpp:582-651).  -> dict(nmatches, matches12, prev_matched)."""
    L = lib()
    cfg = make_cfg(cam)
    d1, d2 = f32(desc1), f32(desc2)
    prev = f32(prev_matched).copy()
    kx2, ky2 = f32(kx2), f32(ky2)
    n1 = len(d1)
    m12 = np.full(max(n1, 1), -1, np.int32)
    L.ppgo_search_for_initialization.restype = C.c_int
    nm = L.ppgo_search_for_initialization(C.byref(cfg), n1, _p(d1, C.c_float), _p(prev, C.c_float), len(kx2),
                                          _p(kx2, C.c_float), _p(ky2, C.c_float), _p(d2, C.c_float), int(window),
                                          C.c_float(ratio), C.c_float(th_low), _p(m12, C.c_int))
    return dict(nmatches=int(nm), matches12=m12[:n1], prev_matched=prev)


def search_for_triangulation(desc1, node1, has_mp1, pos1, desc2, node2, has_mp2, pos2, F12, epipole, th_low=0.7):
    """Matcher::SearchForTriangulation (Matcher.cpp:767-885) with the pinhole epipolar test (Pinhole.cpp:98-114).
    F12 row-major 3x3, epipole (2,).  -> dict(nmatches, match12)."""
    L = lib()
    d1, d2 = f32(desc1), f32(desc2)
    n1, n2 = len(d1), len(d2)
    nd1, nd2 = np.ascontiguousarray(node1, np.int32), np.ascontiguousarray(node2, np.int32)
    m1, m2 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(has_mp2, np.uint8)
    p1, p2 = f32(pos1), f32(pos2)
    F, ep = f32(F12).reshape(9), f32(epipole).reshape(2)
    m12 = np.full(max(n1, 1), -1, np.int32)
    L.ppgo_search_for_triangulation.restype = C.c_int
    nm = L.ppgo_search_for_triangulation(n1, _p(d1, C.c_float), _p(nd1, C.c_int), _p(m1, C.c_ubyte), _p(p1, C.c_float),
                                         n2, _p(d2, C.c_float), _p(nd2, C.c_int), _p(m2, C.c_ubyte), _p(p2, C.c_float),
                                         _p(F, C.c_float), _p(ep, C.c_float), C.c_float(th_low), _p(m12, C.c_int))
    return dict(nmatches=int(nm), match12=m12[:n1])


def search_by_projection(cam, map_desc, proj_uv, observed, kp_x, kp_y, frame_desc, kp_mp, th, max_dist):
    """The sequential part of Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (Matcher.cpp:31-87) and of
    SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, descDist) (:1337-1411).  Rows = projected map points in loop
    order; kp_mp = CurrentFrame.mvpMapPoints as rows (-1 none / unobserved, -2 occupied by a point outside the table).
    -> dict(nmatches, kp_mp)"""
    cfg = make_cfg(cam)
    md, uv, kx, ky, fd = f32(map_desc), f32(proj_uv), f32(kp_x), f32(kp_y), f32(frame_desc)
    m, n = len(md), len(kx)
    ob = np.ones(max(m, 1), np.uint8) if observed is None else np.ascontiguousarray(observed, np.uint8)
    km = np.full(max(n, 1), -1, np.int32) if kp_mp is None else np.ascontiguousarray(kp_mp, np.int32).copy()
    L = lib()
    L.ppgo_search_by_projection.restype = C.c_int
    nm = L.ppgo_search_by_projection(C.byref(cfg), m, _p(md, C.c_float), _p(uv, C.c_float), _p(ob, C.c_ubyte), n,
                                     _p(kx, C.c_float), _p(ky, C.c_float), _p(fd, C.c_float), _p(km, C.c_int),
                                     C.c_float(th), C.c_float(max_dist))
    return dict(nmatches=int(nm), kp_mp=km[:n])


def _cam8(cam):
    return f32([cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D))


def search_for_triangulation_kb8(cam, desc1, node1, has_mp1, pos1, desc2, node2, has_mp2, pos2, R12, t12, epipole,
                                 th_low=0.7, variant=""):
    """Matcher::SearchForTriangulation (Matcher.cpp:767-885) with KannalaBrandt8::epipolarConstrain
    (KannalaBrandt8.cpp:167-236).  R12 row-major 3x3, t12 (3,), epipole (2,).  -> dict(nmatches, match12)."""
    L = lib(variant)
    d1, d2 = f32(desc1), f32(desc2)
    n1, n2 = len(d1), len(d2)
    nd1, nd2 = np.ascontiguousarray(node1, np.int32), np.ascontiguousarray(node2, np.int32)
    m1, m2 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(has_mp2, np.uint8)
    p1, p2 = f32(pos1), f32(pos2)
    R, t, ep, c8 = f32(R12).reshape(9), f32(t12).reshape(3), f32(epipole).reshape(2), _cam8(cam)
    m12 = np.full(max(n1, 1), -1, np.int32)
    L.ppgo_search_for_triangulation_kb8.restype = C.c_int
    nm = L.ppgo_search_for_triangulation_kb8(n1, _p(d1, C.c_float), _p(nd1, C.c_int), _p(m1, C.c_ubyte),
                                             _p(p1, C.c_float), n2, _p(d2, C.c_float), _p(nd2, C.c_int),
                                             _p(m2, C.c_ubyte), _p(p2, C.c_float), _p(c8, C.c_float), _p(R, C.c_float),
                                             _p(t, C.c_float), _p(ep, C.c_float), C.c_float(th_low), _p(m12, C.c_int))
    return dict(nmatches=int(nm), match12=m12[:n1])


def null_vector4(A):
    """Right singular vector of the smallest singular value of a 4 x 4 float matrix (the stand-in for Eigen::JacobiSVD in
    KannalaBrandt8::Triangulate, KannalaBrandt8.cpp:233-234)."""
    a, v = f32(A).reshape(16), np.zeros(4, np.float32)
    lib().ppgo_null_vector4(_p(a, C.c_float), _p(v, C.c_float))
    return v


def kb8_unproject(cam, p2d, variant=""):
    c8, p, out = _cam8(cam), f32(p2d).reshape(2), np.zeros(3, np.float32)
    lib(variant).ppgo_kb8_unproject(_p(c8, C.c_float), _p(p, C.c_float), _p(out, C.c_float))
    return out


def kb8_project(cam, p3d, variant=""):
    c8, p, out = _cam8(cam), f32(p3d).reshape(3), np.zeros(2, np.float32)
    lib(variant).ppgo_kb8_project(_p(c8, C.c_float), _p(p, C.c_float), _p(out, C.c_float))
    return out


def kb8_triangulate_matches(cam, pos1, pos2, R12, t12, variant=""):
    """KannalaBrandt8::TriangulateMatches (KannalaBrandt8.cpp:175-222) -> (value, x3D)."""
    L = lib(variant)
    c8, p1, p2 = _cam8(cam), f32(pos1).reshape(2), f32(pos2).reshape(2)
    R, t = f32(R12).reshape(9), f32(t12).reshape(3)
    r1, x = np.zeros(3, np.float32), np.zeros(3, np.float32)
    L.ppgo_kb8_unproject(_p(c8, C.c_float), _p(p1, C.c_float), _p(r1, C.c_float))
    L.ppgo_kb8_triangulate_matches.restype = C.c_float
    z = L.ppgo_kb8_triangulate_matches(_p(c8, C.c_float), _p(r1, C.c_float), _p(p1, C.c_float), _p(p2, C.c_float),
                                       _p(R, C.c_float), _p(t, C.c_float), _p(x, C.c_float))
    return float(z), x
