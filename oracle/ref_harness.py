"""Python side of the reference-pinning harness (oracle/ref_build.py, oracle/ref_tu/*.cpp).  TEST INFRASTRUCTURE ONLY.

RefExtractor drives the reference's own PPGExtractor (compiled from /root/reference, LibTorch CPU) and returns the same
record layout oracle.post_ref.extract_post returns, plus the dense maps the reference's stages consumed, so that the
oracle can be fed exactly what the reference saw and every discrete output compared bit for bit.
"""
import ctypes as C
import os

import numpy as np

from . import ref_build

MODEL_DIR = os.path.join(ref_build.REF_ROOT, "net")


def available():
    return ref_build.available() and os.path.exists(ref_build.lib_path("extractor"))


_libs = {}


def _lib(name):
    if name not in _libs:
        import torch  # noqa: F401  (libtorch must be in the process before the harness resolves its symbols)
        lib = C.CDLL(ref_build.lib_path(name))
        if name == "extractor":
            lib.ref_extractor_create.restype = C.c_void_p
            lib.ref_extractor_destroy.argtypes = [C.c_void_p]
        _libs[name] = lib
    return _libs[name]


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


class RefExtractor:
    def __init__(self, cam, threads=0):
        self.lib = _lib("extractor")
        self.cam = cam
        params = np.array([cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D), np.float32)
        self.h = self.lib.ref_extractor_create(_p(params), cam.width, cam.height, int(cam.fisheye),
                                               MODEL_DIR.encode(), threads)
        if not self.h:
            raise RuntimeError("the reference PPGExtractor could not be constructed")
        self.h = C.c_void_p(self.h)

    def close(self):
        if self.h:
            self.lib.ref_extractor_destroy(self.h)
            self.h = None

    def image_bounds(self):
        mm = np.zeros(4, np.int32)
        inv = np.zeros(2, np.float32)
        self.lib.ref_image_bounds(self.h, _p(mm, C.c_int), _p(inv))
        return dict(minX=int(mm[0]), minY=int(mm[1]), maxX=int(mm[2]), maxY=int(mm[3]), wInv=inv[0], hInv=inv[1])

    def run(self, gray):
        """gray (H, W) uint8 -> (record, maps): record as oracle.post_ref.extract_post, maps = dict(prob, heat_raw,
        heat_ref, heat_final, desc) exactly as the reference's stages saw / produced them."""
        H, W = self.cam.height, self.cam.width
        g = np.ascontiguousarray(gray, np.uint8)
        assert g.shape == (H, W)
        prob = np.zeros((H, W), np.float32)
        heat_raw = np.zeros((H, W), np.float32)
        heat_ref = np.zeros((H, W), np.float32)
        heat_final = np.zeros((H, W), np.float32)
        desc = np.zeros((256, H // 8, W // 8), np.float32)
        n = self.lib.ref_extract(self.h, _p(g, C.c_uint8), _p(prob), _p(heat_raw), _p(heat_ref), _p(heat_final),
                                 _p(desc))
        if n < 0:
            raise RuntimeError("reference extraction failed")
        ne, nc, nl = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_counts(self.h, C.byref(ne), C.byref(nc), C.byref(nl))
        ne, nc, nl = ne.value, nc.value, nl.value
        pos = np.zeros((max(n, 1), 2), np.float32)
        posun = np.zeros((max(n, 1), 2), np.float32)
        score = np.zeros(max(n, 1), np.float32)
        out = np.zeros(max(n, 1), np.uint8)
        edge_se = np.zeros((max(ne, 1), 2), np.int32)
        lscore = np.zeros(max(ne, 1), np.float32)
        conn_off = np.zeros(n + 1, np.int32)
        conn_idx = np.zeros(max(nc, 1), np.int32)
        col_off = np.zeros(n + 1, np.int32)
        col_pairs = np.zeros((max(nl, 1), 2), np.int32)
        nd = np.zeros((max(n, 1), 256), np.float32)
        self.lib.ref_fetch(self.h, _p(pos), _p(posun), _p(score), _p(out, C.c_uint8), _p(edge_se, C.c_int), _p(lscore),
                           _p(conn_off, C.c_int), _p(conn_idx, C.c_int), _p(col_off, C.c_int), _p(col_pairs, C.c_int),
                           _p(nd))
        rec = dict(n_kp=n, pos=pos[:n], xun=posun[:n, 0].copy(), yun=posun[:n, 1].copy(), score=score[:n],
                   out=out[:n], n_edges=ne, edge_start=edge_se[:ne, 0].copy(), edge_end=edge_se[:ne, 1].copy(),
                   edge_score=lscore[:ne], conn_off=conn_off, conn_idx=conn_idx[:nc], col_off=col_off,
                   col_pairs=col_pairs[:nl], desc=nd[:n])
        maps = dict(prob=prob, heat_raw=heat_raw, heat_ref=heat_ref, heat_final=heat_final, desc=desc)
        return rec, maps


# ---------------------------------------------------------------- matcher (matching/src/Matcher.cpp)
def matcher_available():
    return ref_build.available() and os.path.exists(ref_build.lib_path("matcher"))


def _cam_params(cam):
    return np.array([cam.K[0], cam.K[4], cam.K[2], cam.K[5]] + list(cam.D), np.float32)


def consistent_edge_ok(bad, edge_off, edge_other, edge_ok):
    """What `!pME->isBad() && pME->mbValid` gives for every CSR entry: MapEdge::isBad (PPGGraph.cpp:90-94) is also true when
    an end point is bad, so a flattening of real objects never has edge_ok = 1 next to a bad end point."""
    ok = np.array(edge_ok, np.uint8).copy()
    for p in range(len(bad)):
        for k in range(edge_off[p], edge_off[p + 1]):
            q = edge_other[k]
            if bad[p] or (q >= 0 and bad[q]):
                ok[k] = 0
    return ok


def extend_map_matches(cam, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos, tracked,
                       kp_x, kp_y, frame_desc, kp_mp, edge_start, edge_end, conn_off, conn_idx, th, ratio, kedge_me=None):
    """The reference's own Matcher::ExtendMapMatches on a pointer graph rebuilt from the flat arrays; same arguments and
    result as oracle.post_ref.extend_map_matches."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    md, cd, ob, bd = f32(map_desc), u8(candidate), u8(observed), u8(bad)
    eo, et, ek = i32(edge_off), i32(edge_other), u8(edge_ok)
    uv, vc = f32(proj_uv), f32(view_cos)
    tr = u8(tracked).copy()
    kx, ky, fd = f32(kp_x), f32(kp_y), f32(frame_desc)
    km = i32(kp_mp).copy()
    es, ee, co, ci = i32(edge_start), i32(edge_end), i32(conn_off), i32(conn_idx)
    ne = len(es)
    me = np.full(max(ne, 1), -1, np.int32) if kedge_me is None else i32(kedge_me).copy()
    if len(et) == 0:
        et, ek = np.zeros(1, np.int32), np.zeros(1, np.uint8)
    if len(ci) == 0:
        ci = np.zeros(1, np.int32)
    if ne == 0:
        es, ee = np.zeros(1, np.int32), np.zeros(1, np.int32)
    params = _cam_params(cam)
    u8p, i32p = C.c_uint8, C.c_int
    nm = lib.ref_extend_map_matches(_p(params), cam.width, cam.height, int(cam.fisheye), len(cd), _p(md), _p(cd, u8p),
                                    _p(ob, u8p), _p(bd, u8p), _p(eo, i32p), _p(et, i32p), _p(ek, u8p), _p(uv), _p(vc),
                                    _p(tr, u8p), len(kx), _p(kx), _p(ky), _p(fd), _p(km, i32p), ne, _p(es, i32p),
                                    _p(ee, i32p), _p(co, i32p), _p(ci, i32p), _p(me, i32p), C.c_float(th),
                                    C.c_float(ratio))
    if nm < 0:
        raise RuntimeError("reference ExtendMapMatches failed")
    return dict(nmatches=nm, kp_mp=km, kedge_me=me[:ne], tracked=tr)


def features_in_area(cam, kx, ky, x, y, r):
    lib = _lib("matcher")
    kx, ky = np.ascontiguousarray(kx, np.float32), np.ascontiguousarray(ky, np.float32)
    out = np.zeros(max(len(kx), 1), np.int32)
    n = lib.ref_features_in_area(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), len(kx), _p(kx), _p(ky),
                                 C.c_float(x), C.c_float(y), C.c_float(r), _p(out, C.c_int))
    return out[:n].copy()


def descriptor_distance(a, b):
    lib = _lib("matcher")
    lib.ref_descriptor_distance.restype = C.c_float
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return float(lib.ref_descriptor_distance(_p(a), _p(b)))


def candidate_order(candidate, bad, edge_off):
    """Rows in the order the reference's ExtendMapMatches walks them (its unstable std::sort included)."""
    lib = _lib("matcher")
    cd, bd, eo = np.ascontiguousarray(candidate, np.uint8), np.ascontiguousarray(bad, np.uint8), np.ascontiguousarray(edge_off, np.int32)
    out = np.zeros(max(len(cd), 1), np.int32)
    n = lib.ref_candidate_order(len(cd), _p(cd, C.c_uint8), _p(bd, C.c_uint8), _p(eo, C.c_int), _p(out, C.c_int))
    return out[:n].copy()


def permute_table(perm, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos, tracked,
                  kp_mp):
    """The same map with its rows renumbered: new row i = old row perm[i] (perm covers every row).  -> dict of the
    renumbered arrays + `old_of_new` / `new_of_old` to translate results back."""
    perm = np.asarray(perm, np.int64)
    P = len(candidate)
    new_of_old = np.empty(P, np.int64)
    new_of_old[perm] = np.arange(P)
    deg = np.diff(edge_off)
    off = np.zeros(P + 1, np.int32)
    off[1:] = np.cumsum(deg[perm])
    other = np.zeros(len(edge_other), np.int32)
    ok = np.zeros(len(edge_ok), np.uint8)
    pos_new_of_old = np.zeros(max(len(edge_other), 1), np.int64)  # CSR position translation (kedge_me)
    for i, p in enumerate(perm):
        a, b = edge_off[p], edge_off[p + 1]
        o = np.asarray(edge_other[a:b], np.int64)
        other[off[i]:off[i + 1]] = np.where(o >= 0, new_of_old[np.maximum(o, 0)], -1)
        ok[off[i]:off[i + 1]] = edge_ok[a:b]
        pos_new_of_old[a:b] = np.arange(off[i], off[i + 1])
    km = np.asarray(kp_mp, np.int64)
    return dict(map_desc=np.ascontiguousarray(map_desc[perm]), candidate=np.asarray(candidate)[perm],
                observed=np.asarray(observed)[perm], bad=np.asarray(bad)[perm], edge_off=off, edge_other=other,
                edge_ok=ok, proj_uv=np.ascontiguousarray(proj_uv[perm]), view_cos=np.asarray(view_cos)[perm],
                tracked=np.asarray(tracked)[perm],
                kp_mp=np.where(km >= 0, new_of_old[np.maximum(km, 0)], km).astype(np.int32), old_of_new=perm,
                new_of_old=new_of_old, pos_new_of_old=pos_new_of_old)


def walk_order_permutation(candidate, bad, edge_off):
    """perm for permute_table: the candidates in the reference's walk order first, then all other rows."""
    order = candidate_order(candidate, bad, edge_off)
    rest = np.setdiff1d(np.arange(len(candidate)), order)
    return np.concatenate([order, rest]).astype(np.int64)


def search_for_initialization(cam, kx1, ky1, desc1, prev_matched, kx2, ky2, desc2, window=50, ratio=0.9):
    """The reference's own Matcher::SearchForInitialization on two Frames rebuilt from the flat arrays."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    kx1, ky1, d1, kx2, ky2, d2 = f32(kx1), f32(ky1), f32(desc1), f32(kx2), f32(ky2), f32(desc2)
    prev = f32(prev_matched).copy()
    n1 = len(kx1)
    m12 = np.full(max(n1, 1), -1, np.int32)
    nm = lib.ref_search_for_initialization(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), n1, _p(kx1),
                                           _p(ky1), _p(d1), _p(prev), len(kx2), _p(kx2), _p(ky2), _p(d2), int(window),
                                           C.c_float(ratio), _p(m12, C.c_int))
    return dict(nmatches=int(nm), matches12=m12[:n1], prev_matched=prev)


def search_for_triangulation(cam, R1, t1, R2, t2, pos1, desc1, node1, has_mp1, pos2, desc2, node2, has_mp2):
    """The reference's own Matcher::SearchForTriangulation with its own Pinhole / KannalaBrandt8 camera (by cam.fisheye) on
    two KeyFrames rebuilt from the flat arrays (poses = world -> camera).  -> dict(nmatches, match12, F12 (3x3),
    epipole (2,), R12 (3x3), t12 (3,)): what the reference's classes compute from the poses (Matcher.cpp:776-788,
    Pinhole.cpp:101-104)."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    p1, d1, p2, d2 = f32(pos1), f32(desc1), f32(pos2), f32(desc2)
    n1, n2 = len(p1), len(p2)
    m12 = np.full(max(n1, 1), -1, np.int32)
    F, R12 = np.zeros(9, np.float32), np.zeros(9, np.float32)
    ep, t12 = np.zeros(2, np.float32), np.zeros(3, np.float32)
    nm = lib.ref_search_for_triangulation(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye),
                                          _p(f32(R1).reshape(9)), _p(f32(t1)), _p(f32(R2).reshape(9)), _p(f32(t2)), n1,
                                          _p(p1), _p(d1), _p(i32(node1), C.c_int), _p(u8(has_mp1), C.c_ubyte), n2, _p(p2),
                                          _p(d2), _p(i32(node2), C.c_int), _p(u8(has_mp2), C.c_ubyte), _p(m12, C.c_int),
                                          _p(F), _p(ep), _p(R12), _p(t12))
    return dict(nmatches=int(nm), match12=m12[:n1], F12=F.reshape(3, 3), epipole=ep, R12=R12.reshape(3, 3), t12=t12)


def kb8_triangulate(cam, pos1, pos2, R12, t12):
    """The reference's own KannalaBrandt8::unproject / project / TriangulateMatches (JacobiSVD = the stand-in's).
    -> dict(value, r1, r2, uv1 = project(3 * r1), x3D)."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    r1, r2, uv1, x = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32), np.zeros(3, np.float32)
    lib.ref_kb8_triangulate.restype = C.c_float
    z = lib.ref_kb8_triangulate(_p(_cam_params(cam)), cam.width, cam.height, _p(f32(pos1)), _p(f32(pos2)),
                                _p(f32(R12).reshape(9)), _p(f32(t12)), _p(r1), _p(r2), _p(uv1), _p(x))
    return dict(value=float(z), r1=r1, r2=r2, uv1=uv1, x3D=x)


def check_in_frustum(cam, Rcw, tcw, Ow, world_pos, normal, min_dist, max_dist, cos_limit=0.5):
    """The reference's own Frame::CheckInFrustum (map/src/Frame.cpp:223-260) with its own Pinhole / KannalaBrandt8
    camera on real MapPoint objects; same result layout as oracle.post_ref.check_in_frustum plus `visible` (the
    IncreaseVisible count of :259)."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    wp, nr, mn, mx = f32(world_pos), f32(normal), f32(min_dist), f32(max_dist)
    m = len(mn)
    iv, out4, vis = np.zeros(max(m, 1), np.uint8), np.zeros((max(m, 1), 4), np.float32), np.zeros(max(m, 1), np.int32)
    lib.ref_check_in_frustum.restype = None
    lib.ref_check_in_frustum(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _p(f32(Rcw).reshape(9)),
                             _p(f32(tcw)), _p(f32(Ow)), m, _p(wp), _p(nr), _p(mn), _p(mx), C.c_float(cos_limit),
                             _p(iv, C.c_ubyte), _p(out4), _p(vis, C.c_int))
    return dict(in_view=iv[:m], proj_uv=out4[:m, :2].copy(), depth=out4[:m, 2].copy(), view_cos=out4[:m, 3].copy(),
                visible=vis[:m])


def search_by_projection(cam, mode, x, th, desc_dist=0.7):
    """The reference's own Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (mode 0) / (CurrentFrame, pKF,
    sAlreadyFound, th, descDist) (mode 1) / (pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (mode 2; x["scale"] = the
    similarity's scale) on objects rebuilt from synth.projection_inputs.
    -> dict(nmatches, kp_mp (source feature indices, -1 / -2 / -3), row_valid, proj_uv)"""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    n_src, n = len(x["state"]), len(x["kp_x"])
    km = np.ascontiguousarray(x["kp_mp"], np.int32).copy()
    uv, valid = np.zeros((max(n_src, 1), 2), np.float32), np.zeros(max(n_src, 1), np.uint8)
    nm = lib.ref_search_by_projection(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), int(mode),
                                      _p(f32(x["Rcw"]).reshape(9)), _p(f32(x["tcw"])), n_src, _p(f32(x["world_pos"])),
                                      _p(f32(x["mp_desc"])), _p(u8(x["state"]), C.c_ubyte),
                                      _p(u8(x["observed"]), C.c_ubyte), _p(f32(x["min_dist"])), _p(f32(x["max_dist"])), n,
                                      _p(f32(x["kp_x"])), _p(f32(x["kp_y"])), _p(f32(x["desc"])), _p(km, C.c_int),
                                      C.c_float(th), C.c_float(desc_dist), _p(uv), _p(valid, C.c_ubyte),
                                      _p(f32(x["normal"])), C.c_float(x.get("scale", 1.0)))
    return dict(nmatches=int(nm), kp_mp=km, row_valid=valid[:n_src], proj_uv=uv[:n_src])


def distinctive_descriptor(obs_desc, state=None, point_bad=False):
    """The reference's own MapPoint::ComputeDistinctiveDescriptors on a MapPoint observed by len(obs_desc) raw key frames
    (state per observation: 0 good, 1 key frame bad, 2 index -1).  -> mDescriptor (256,), all -1 when untouched."""
    lib = _lib("matcher")
    d = np.ascontiguousarray(obs_desc, np.float32).reshape(-1, 256)
    st = np.zeros(max(len(d), 1), np.uint8) if state is None else np.ascontiguousarray(state, np.uint8)
    out = np.zeros(256, np.float32)
    lib.ref_distinctive_descriptor.restype = None
    lib.ref_distinctive_descriptor(len(d), _p(d if len(d) else np.zeros((1, 256), np.float32)), _p(st, C.c_ubyte),
                                   int(point_bad), _p(out))
    return out


def search_by_sim3(cam, x, th):
    """The reference's own Matcher::SearchBySim3 on two raw key frames with real MapPoint objects (synth.sim3_inputs).
    -> dict(nfound, matches12 (KF2 feature per KF1 feature, -1 none), valid1 / uv1, valid2 / uv2: which features the loop
    heads let search and where they project into the other key frame)"""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    n1, n2 = len(x["pos1"]), len(x["pos2"])
    m12 = np.ascontiguousarray(x["matches12"], np.int32).copy()
    v1, v2 = np.zeros(max(n1, 1), np.uint8), np.zeros(max(n2, 1), np.uint8)
    uv1, uv2 = np.zeros((max(n1, 1), 2), np.float32), np.zeros((max(n2, 1), 2), np.float32)
    nf = lib.ref_search_by_sim3(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _p(f32(x["R1"]).reshape(9)),
                                _p(f32(x["t1"])), _p(f32(x["R2"]).reshape(9)), _p(f32(x["t2"])),
                                _p(f32(x["R12"]).reshape(9)), _p(f32(x["t12"])), C.c_float(float(x["s12"])), n1,
                                _p(f32(x["pos1"])), _p(f32(x["desc1"])), _p(u8(x["state1"]), C.c_ubyte),
                                _p(f32(x["world1"])), _p(f32(x["mp_desc1"])), _p(f32(x["min_dist1"])),
                                _p(f32(x["max_dist1"])), n2, _p(f32(x["pos2"])), _p(f32(x["desc2"])),
                                _p(u8(x["state2"]), C.c_ubyte), _p(f32(x["world2"])), _p(f32(x["mp_desc2"])),
                                _p(f32(x["min_dist2"])), _p(f32(x["max_dist2"])), _p(m12, C.c_int), C.c_float(th),
                                _p(v1, C.c_ubyte), _p(uv1), _p(v2, C.c_ubyte), _p(uv2))
    return dict(nfound=int(nf), matches12=m12[:n1], valid1=v1[:n1], uv1=uv1[:n1], valid2=v2[:n2], uv2=uv2[:n2])


def search_by_bow_kf_f(cam, desc_kf, node_kf, state_kf, desc_f, node_f, ratio):
    """The reference's own Matcher::SearchByBoW(KeyFrame*, Frame&, ...) (Matcher.cpp:393-477).  state_kf: 0 no map point,
    1 good, 2 bad.  -> dict(nmatches, f2kf): f2kf[i] = key-frame feature whose map point frame feature i received."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dk, df = f32(desc_kf), f32(desc_f)
    out = np.full(max(len(df), 1), -1, np.int32)
    nm = lib.ref_search_by_bow_kf_f(_p(_cam_params(cam)), cam.width, cam.height, len(dk), _p(dk),
                                    _p(np.ascontiguousarray(node_kf, np.int32), C.c_int),
                                    _p(np.ascontiguousarray(state_kf, np.uint8), C.c_ubyte), len(df), _p(df),
                                    _p(np.ascontiguousarray(node_f, np.int32), C.c_int), C.c_float(ratio),
                                    _p(out, C.c_int))
    return dict(nmatches=int(nm), f2kf=out[:len(df)])


def search_by_bow_kf_kf(cam, desc1, node1, state1, desc2, node2, state2, ratio):
    """The reference's own Matcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (Matcher.cpp:663-754).
    -> dict(nmatches, match12): match12[i1] = feature of KF2 whose map point vpMatches12[i1] is."""
    lib = _lib("matcher")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    d1, d2 = f32(desc1), f32(desc2)
    out = np.full(max(len(d1), 1), -1, np.int32)
    nm = lib.ref_search_by_bow_kf_kf(_p(_cam_params(cam)), cam.width, cam.height, len(d1), _p(d1),
                                     _p(np.ascontiguousarray(node1, np.int32), C.c_int),
                                     _p(np.ascontiguousarray(state1, np.uint8), C.c_ubyte), len(d2), _p(d2),
                                     _p(np.ascontiguousarray(node2, np.int32), C.c_int),
                                     _p(np.ascontiguousarray(state2, np.uint8), C.c_ubyte), C.c_float(ratio),
                                     _p(out, C.c_int))
    return dict(nmatches=int(nm), match12=out[:len(d1)])


# ---------------------------------------------------------------- include/ppg_shim.hpp executed on the reference's objects
def shim_available():
    return os.path.exists(ref_build.lib_path("shim")) and os.path.exists(ref_build.lib_path("matcher"))


def _weights_dir():
    return os.path.join(os.path.dirname(ref_build.HERE), "ppg_slam_b200", "weights")


def shim_extend_both(cam, map_desc, candidate, observed, bad, edge_off, edge_other, edge_ok, proj_uv, view_cos, tracked,
                     kp_x, kp_y, frame_desc, kp_mp, edge_start, edge_end, conn_off, conn_idx, th, ratio):
    """The same pointer graph through the reference's Matcher::ExtendMapMatches (host) and through
    ppg_shim::Matcher::ExtendMapMatches (flattening -> C ABI -> GPU -> write-back).  -> (reference result, shim result)."""
    lib = _lib("shim")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    md, cd, ob, bd = f32(map_desc), u8(candidate), u8(observed), u8(bad)
    eo, et, ek = i32(edge_off), i32(edge_other), u8(edge_ok)
    uv, vc, tr = f32(proj_uv), f32(view_cos), u8(tracked)
    kx, ky, fd, km = f32(kp_x), f32(kp_y), f32(frame_desc), i32(kp_mp)
    es, ee, co, ci = i32(edge_start), i32(edge_end), i32(conn_off), i32(conn_idx)
    P, n, ne = len(cd), len(kx), len(es)
    if len(et) == 0:
        et, ek = np.zeros(1, np.int32), np.zeros(1, np.uint8)
    if len(ci) == 0:
        ci = np.zeros(1, np.int32)
    if ne == 0:
        es, ee = np.zeros(1, np.int32), np.zeros(1, np.int32)
    order = walk_order_permutation(cd, bd, eo).astype(np.int32)
    nm = np.zeros(2, np.int32)
    kp2 = np.zeros((2, max(n, 1)), np.int32)
    me2 = np.zeros((2, max(ne, 1)), np.int32)
    tr2 = np.zeros((2, max(P, 1)), np.uint8)
    u8p, i32p = C.c_uint8, C.c_int
    rc = lib.shim_extend_both(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _weights_dir().encode(), P,
                              _p(md), _p(cd, u8p), _p(ob, u8p), _p(bd, u8p), _p(eo, i32p), _p(et, i32p), _p(ek, u8p),
                              _p(uv), _p(vc), _p(tr, u8p), n, _p(kx), _p(ky), _p(fd), _p(km, i32p), ne, _p(es, i32p),
                              _p(ee, i32p), _p(co, i32p), _p(ci, i32p), _p(order, i32p), len(order), C.c_float(th),
                              C.c_float(ratio), _p(nm, i32p), _p(kp2, i32p), _p(me2, i32p), _p(tr2, u8p))
    if rc != 0:
        raise RuntimeError("shim_extend_both failed (see stderr)")
    return tuple(dict(nmatches=int(nm[k]), kp_mp=kp2[k, :n], kedge_me=me2[k, :ne], tracked=tr2[k, :P]) for k in (0, 1))


def shim_init_both(cam, kx1, ky1, desc1, prev_matched, kx2, ky2, desc2, window, ratio):
    lib = _lib("shim")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    kx1, ky1, d1, kx2, ky2, d2, prev = f32(kx1), f32(ky1), f32(desc1), f32(kx2), f32(ky2), f32(desc2), f32(prev_matched)
    n1 = len(kx1)
    nm = np.zeros(2, np.int32)
    m12 = np.zeros((2, max(n1, 1)), np.int32)
    pv = np.zeros((2, max(n1, 1), 2), np.float32)
    rc = lib.shim_init_both(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _weights_dir().encode(), n1,
                            _p(kx1), _p(ky1), _p(d1), _p(prev), len(kx2), _p(kx2), _p(ky2), _p(d2), int(window),
                            C.c_float(ratio), _p(nm, C.c_int), _p(m12, C.c_int), _p(pv))
    if rc != 0:
        raise RuntimeError("shim_init_both failed (see stderr)")
    return tuple(dict(nmatches=int(nm[k]), matches12=m12[k, :n1], prev_matched=pv[k, :n1]) for k in (0, 1))


def shim_triangulation_both(cam, R1, t1, R2, t2, pos1, desc1, node1, has_mp1, pos2, desc2, node2, has_mp2):
    """The same two key frames through the reference's Matcher::SearchForTriangulation (host, its own Pinhole / KannalaBrandt8
    camera by cam.fisheye) and through ppg_shim::Matcher::SearchForTriangulation (poses -> F12 / R12 / t12 / epipole with the
    reference's classes -> C ABI -> GPU).  -> (reference result, shim result)."""
    lib = _lib("shim")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    p1, d1, p2, d2 = f32(pos1), f32(desc1), f32(pos2), f32(desc2)
    n1, n2 = len(p1), len(p2)
    nm = np.zeros(2, np.int32)
    m12 = np.zeros((2, max(n1, 1)), np.int32)
    rc = lib.shim_triangulation_both(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _weights_dir().encode(),
                                     _p(f32(R1).reshape(9)), _p(f32(t1)), _p(f32(R2).reshape(9)), _p(f32(t2)), n1, _p(p1),
                                     _p(d1), _p(i32(node1), C.c_int), _p(u8(has_mp1), C.c_ubyte), n2, _p(p2), _p(d2),
                                     _p(i32(node2), C.c_int), _p(u8(has_mp2), C.c_ubyte), _p(nm, C.c_int),
                                     _p(m12, C.c_int))
    if rc != 0:
        raise RuntimeError("shim_triangulation_both failed (rc %d, see stderr)" % rc)
    return tuple(dict(nmatches=int(nm[k]), match12=m12[k, :n1]) for k in (0, 1))


def shim_projection_both(cam, mode, x, th, desc_dist=0.7):
    """Matcher::SearchByProjection (mode 0 / 1, see search_by_projection) through the reference's host function and through
    ppg_shim::Matcher (projection tests with the reference's classes -> C ABI -> GPU) on identical objects.
    -> (reference result, shim result)"""
    lib = _lib("shim")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    n_src, n = len(x["state"]), len(x["kp_x"])
    km = np.ascontiguousarray(x["kp_mp"], np.int32)
    nm = np.zeros(2, np.int32)
    out = np.zeros((2, max(n, 1)), np.int32)
    rc = lib.shim_projection_both(_p(_cam_params(cam)), cam.width, cam.height, int(cam.fisheye), _weights_dir().encode(),
                                  int(mode), _p(f32(x["Rcw"]).reshape(9)), _p(f32(x["tcw"])), n_src,
                                  _p(f32(x["world_pos"])), _p(f32(x["mp_desc"])), _p(u8(x["state"]), C.c_ubyte),
                                  _p(u8(x["observed"]), C.c_ubyte), _p(f32(x["min_dist"])), _p(f32(x["max_dist"])), n,
                                  _p(f32(x["kp_x"])), _p(f32(x["kp_y"])), _p(f32(x["desc"])), _p(km, C.c_int),
                                  C.c_float(th), C.c_float(desc_dist), _p(nm, C.c_int), _p(out, C.c_int),
                                  _p(f32(x["normal"])), C.c_float(x.get("scale", 1.0)))
    if rc != 0:
        raise RuntimeError("shim_projection_both failed (rc %d, see stderr)" % rc)
    return tuple(dict(nmatches=int(nm[k]), kp_mp=out[k, :n]) for k in (0, 1))


def shim_bow_both(cam, kf_kf, desc1, node1, state1, desc2, node2, state2, ratio):
    """Matcher::SearchByBoW through the reference's host function and through ppg_shim::Matcher (GPU) on the same objects.
    kf_kf False: (KeyFrame, Frame) -> f2kf per frame feature; True: (KeyFrame, KeyFrame) -> match12 per KF1 feature.
    -> (reference result, shim result), each dict(nmatches, out)."""
    lib = _lib("shim")
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    d1, d2 = f32(desc1), f32(desc2)
    n = len(d1) if kf_kf else len(d2)
    nm = np.zeros(2, np.int32)
    out = np.zeros((2, max(n, 1)), np.int32)
    rc = lib.shim_bow_both(_p(_cam_params(cam)), cam.width, cam.height, _weights_dir().encode(), int(bool(kf_kf)), len(d1),
                           _p(d1), _p(i32(node1), C.c_int), _p(u8(state1), C.c_ubyte), len(d2), _p(d2),
                           _p(i32(node2), C.c_int), _p(u8(state2), C.c_ubyte), C.c_float(ratio), _p(nm, C.c_int),
                           _p(out, C.c_int))
    if rc != 0:
        raise RuntimeError("shim_bow_both failed (see stderr)")
    return tuple(dict(nmatches=int(nm[k]), out=out[k, :n]) for k in (0, 1))
