"""Known-answer test of the whole-function oracle of Matcher::ExtendMapMatches (oracle/ppg_oracle.c::
ppgo_extend_map_matches) against a literal restatement of matching/src/Matcher.cpp:203-381 on a Python POINTER graph
(MapPoint / MapEdge / Frame objects, deque, erase from lx / ly), i.e. written the way the reference is and sharing
only DescriptorDistance and GetFeaturesInArea with the oracle (both pinned by their own KATs)."""
from collections import deque

import numpy as np
import pytest

from oracle import post_ref as O
from ppg_slam_b200 import cameras, synth


class MapPoint:
    def __init__(self, row, desc, bad, nobs):
        self.row, self.desc, self.bad, self.nobs = row, desc, bad, nobs
        self.edges = []
        self.mbTrackInView = False
        self.mnTrackedbyFrame = -1
        self.mTrackProjX = self.mTrackProjY = self.mTrackViewCos = 0.0

    def isBad(self):
        return self.bad

    def Observations(self):
        return self.nobs

    def getEdges(self):
        return list(self.edges)


class MapEdge:
    def __init__(self, ps, pe, bad, valid, uid):
        self.mpMPs, self.mpMPe, self.bad, self.mbValid, self.uid = ps, pe, bad, valid, uid

    def isBad(self):
        return self.bad

    def theOtherPt(self, p):  # feature/src/PPGGraph.cpp:47-54
        if self.mpMPs is p:
            return self.mpMPe
        if self.mpMPe is p:
            return self.mpMPs
        return None


def py_extend(cam, F, vpMapPoints, th, ratio, TH_HIGH=0.8):
    """Matcher.cpp:203-381, statement by statement.  F: dict(mnId, kx, ky, desc, mvpMapPoints, mvpMapEdges,
    mvKeyEdges [(s, e)], mvConnected [[edge ids]])."""
    nmatches = 0
    cands = [p for p in vpMapPoints if not (p.isBad() or not p.mbTrackInView)]
    cands.sort(key=lambda p: -len(p.getEdges()))  # stable: ties keep vpMapPoints order
    for pMP in cands:
        if pMP.mnTrackedbyFrame == F["mnId"] or pMP.isBad():
            continue
        bestDist, bestDist2, bestIdx = 1e6, 1e6, -1
        r = np.float32(th)
        r = np.float32(np.float64(r) * (2.5 if np.float64(pMP.mTrackViewCos) > 0.998 else 4.0))
        vIndices = O.features_in_area(cam, F["kx"], F["ky"], pMP.mTrackProjX, pMP.mTrackProjY, r)
        if len(vIndices) == 0:
            continue
        for idx in vIndices:
            q = F["mvpMapPoints"][idx]
            if q is not None and q.Observations() > 0:
                continue
            dist = O.descriptor_distance(pMP.desc, F["desc"][idx])
            if dist < bestDist:
                bestDist2, bestDist, bestIdx = bestDist, dist, idx
            elif dist < bestDist2:
                bestDist2 = dist
        if bestDist > np.float32(TH_HIGH) and bestDist > float(np.float32(np.float32(ratio) * np.float32(bestDist2))):
            continue
        F["mvpMapPoints"][bestIdx] = pMP
        pMP.mnTrackedbyFrame = F["mnId"]
        nmatches += 1
        matchSeed = deque([bestIdx])
        while len(matchSeed) > 0:
            keyID = matchSeed.popleft()
            mapEdge_set = pMP.getEdges()
            keyEdge_set = F["mvConnected"][keyID]
            if len(mapEdge_set) == 0 or len(keyEdge_set) == 0:
                continue
            weight = np.full((len(mapEdge_set), len(keyEdge_set)), 1e6, np.float32)
            lx = [i for i, e in enumerate(mapEdge_set)
                  if not (e.isBad() or not e.mbValid or e.theOtherPt(pMP) is None)]
            ly = list(range(len(keyEdge_set)))

            def other_kp(j):
                s, e = F["mvKeyEdges"][keyEdge_set[j]]
                assert keyID in (s, e)
                return e if keyID == s else s

            for i in lx:
                for j in ly:
                    pMP_o = mapEdge_set[i].theOtherPt(pMP)
                    keyID_o = other_kp(j)
                    if pMP_o is F["mvpMapPoints"][keyID_o]:
                        weight[i, j] = -1
                    else:
                        weight[i, j] = O.descriptor_distance(pMP_o.desc, F["desc"][keyID_o])
            while lx and ly:
                minlx = minly = 0
                minWeight = np.float32(1e6)
                for i in range(len(lx)):
                    for j in range(len(ly)):
                        if weight[lx[i], ly[j]] < minWeight:
                            minWeight, minlx, minly = weight[lx[i], ly[j]], i, j
                if minWeight > np.float32(TH_HIGH):
                    break
                mi, kj = lx.pop(minlx), ly.pop(minly)
                pME = mapEdge_set[mi]
                keyEdgeID = keyEdge_set[kj]
                pMP_o = pME.theOtherPt(pMP)
                keyID_o = other_kp(kj)
                if pMP_o is None or pMP_o.isBad() or pMP_o.mnTrackedbyFrame == F["mnId"]:
                    continue
                F["mvpMapPoints"][keyID_o] = pMP_o
                F["mvpMapEdges"][keyEdgeID] = pME
                pMP_o.mnTrackedbyFrame = F["mnId"]
                matchSeed.append(keyID_o)
        nmatches += 1
    return nmatches


def random_frame_graph(rs, cam, n, n_edges):
    kx = rs.uniform(20, cam.width - 20, n).astype(np.float32)
    ky = rs.uniform(20, cam.height - 20, n).astype(np.float32)
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
    return kx, ky, fd.astype(np.float32), es, ee, off, idx, conn


def build_pointer_graph(inp, F_id=7):
    M = len(inp["map_desc"])
    mps = [MapPoint(p, inp["map_desc"][p], bool(inp["bad"][p]), 3 if inp["observed"][p] else 0) for p in range(M)]
    outside = MapPoint(-2, np.zeros(256, np.float32), False, 5)  # a map point that is not in the table
    uid = 0
    for p in range(M):
        mps[p].mbTrackInView = bool(inp["candidate"][p])
        mps[p].mnTrackedbyFrame = F_id if inp["tracked"][p] else -1
        mps[p].mTrackProjX, mps[p].mTrackProjY = inp["proj_uv"][p]
        mps[p].mTrackViewCos = inp["view_cos"][p]
        for k in range(inp["edge_off"][p], inp["edge_off"][p + 1]):
            o = inp["edge_other"][k]
            # theOtherPt(pMP) == nullptr <=> pMP is not an endpoint of the edge
            e = MapEdge(mps[p] if o >= 0 else outside, mps[o] if o >= 0 else outside, False,
                        bool(inp["edge_ok"][k]), k)
            mps[p].edges.append(e)
            uid += 1
    return mps, outside


@pytest.mark.parametrize("seed,clean", [(0, True), (1, False), (2, False), (3, False)])
def test_extend_map_matches_kat(seed, clean):
    cam = cameras.EUROC
    rs = np.random.RandomState(100 + seed)
    n, M = 90, 400
    kx, ky, fd, es, ee, coff, cidx, conn = random_frame_graph(rs, cam, n, 140)
    inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=10.0,
                              planted_frac=0.4, clean=clean)
    got = O.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                               inp["edge_other"], inp["edge_ok"], inp["proj_uv"], inp["view_cos"], inp["tracked"],
                               kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, th=10.0, ratio=0.8)
    mps, outside = build_pointer_graph(inp)
    F = dict(mnId=7, kx=kx, ky=ky, desc=fd, mvKeyEdges=list(zip(es.tolist(), ee.tolist())), mvConnected=conn,
             mvpMapPoints=[mps[r] if r >= 0 else (outside if r == -2 else None) for r in inp["kp_mp"]],
             mvpMapEdges=[None] * len(es))
    want_n = py_extend(cam, F, mps, 10.0, 0.8)
    assert got["nmatches"] == want_n
    want_kp = [(-1 if q is None else q.row) for q in F["mvpMapPoints"]]
    assert got["kp_mp"].tolist() == want_kp
    assert got["kedge_me"].tolist() == [(-1 if e is None else e.uid) for e in F["mvpMapEdges"]]
    assert got["tracked"].tolist() == [int(p.mnTrackedbyFrame == 7) for p in mps]
    # the case is not vacuous: direct matches, matches grown along edges, and consumed-but-rejected pairs
    assert want_n >= 20
    assert (got["kedge_me"] >= 0).sum() >= 5
    if not clean:
        assert (inp["kp_mp"] != got["kp_mp"]).sum() >= 10


@pytest.mark.parametrize("cam", [cameras.EUROC, cameras.TUMVI], ids=lambda c: c.name)
def test_check_in_frustum_kat(cam):
    """Frame::CheckInFrustum (map/src/Frame.cpp:223-260) restated with numpy float32 scalars, statement by statement."""
    import math
    f = np.float32
    g = synth.frustum_inputs(3, cam, 600)
    R, t, Ow = g["Rcw"][0], g["tcw"][0], g["Ow"][0]
    got = O.check_in_frustum(cam, R, t, Ow, g["world_pos"], g["normal"], g["min_dist"], g["max_dist"], 0.5)
    b = O.image_bounds(cam)
    fx, fy, cx, cy = f(cam.K[0]), f(cam.K[4]), f(cam.K[2]), f(cam.K[5])
    n_in = 0
    for j in range(len(g["min_dist"])):
        P, Pn = g["world_pos"][j], g["normal"][j]
        dot3 = lambda a, c: f(f(f(a[0] * c[0]) + f(a[1] * c[1])) + f(a[2] * c[2]))
        Pc = [f(dot3(R[i], P) + t[i]) for i in range(3)]
        want = (0, -1.0, -1.0, -1.0, 0.0)
        if not Pc[2] < 0:
            with np.errstate(all="ignore"):
                if not cam.fisheye:
                    u = f(f(f(fx * Pc[0]) / Pc[2]) + cx)
                    v = f(f(f(fy * Pc[1]) / Pc[2]) + cy)
                else:
                    x2y2 = f(f(Pc[0] * Pc[0]) + f(Pc[1] * Pc[1]))
                    theta = f(math.atan2(float(np.sqrt(x2y2)), float(Pc[2])))
                    psi = f(math.atan2(float(Pc[1]), float(Pc[0])))
                    t2 = f(theta * theta)
                    t3 = f(theta * t2)
                    t5 = f(t3 * t2)
                    t7 = f(t5 * t2)
                    t9 = f(t7 * t2)
                    D = [f(x) for x in cam.D]
                    r = f(f(f(f(theta + f(D[0] * t3)) + f(D[1] * t5)) + f(D[2] * t7)) + f(D[3] * t9))
                    u = f(float(f(fx * r)) * math.cos(float(psi)) + float(cx))
                    v = f(float(f(fy * r)) * math.sin(float(psi)) + float(cy))
            if u >= f(b.minX) and u < f(b.maxX) and v >= f(b.minY) and v < f(b.maxY):
                PO = [f(P[i] - Ow[i]) for i in range(3)]
                dist = f(np.sqrt(dot3(PO, PO)))
                if not (dist < g["min_dist"][j] or dist > g["max_dist"][j]):
                    vc = f(dot3(PO, Pn) / dist)
                    if not vc < f(0.5):
                        want = (1, u, v, dist, vc)
        assert got["in_view"][j] == want[0]
        for a, w in zip((got["proj_uv"][j, 0], got["proj_uv"][j, 1], got["depth"][j], got["view_cos"][j]), want[1:]):
            assert f(a).view(np.uint32) == f(w).view(np.uint32), (j, a, w)
        n_in += want[0]
    assert 60 < n_in < 500  # every exit of the function is taken by some point


@pytest.mark.parametrize("seed", list(range(20)))
def test_extend_map_matches_random_small_graphs(seed):
    """A sweep over small random configurations (sizes, edge densities, duplicate and dangling map edges, keypoints
    without key edges, thresholds): oracle == pointer-graph restatement, plus invariants of the function itself."""
    cam = cameras.EUROC
    rs = np.random.RandomState(1000 + seed)
    n = int(rs.randint(3, 45))
    M = int(rs.randint(5, 140))
    ne = int(rs.randint(0, min(3 * n, n * (n - 1) // 2) + 1))
    kx, ky, fd, es, ee, coff, cidx, conn = random_frame_graph(rs, cam, n, ne)
    if seed % 3 == 0:  # a tight cluster of keypoints: windows overlap, rows compete for the same keypoints
        kx = (300 + rs.uniform(0, 60, n)).astype(np.float32)
        ky = (200 + rs.uniform(0, 40, n)).astype(np.float32)
    th = float(rs.choice([3.0, 10.0, 15.0]))
    inp = synth.extend_inputs(seed, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=th,
                              planted_frac=float(rs.uniform(0.2, 0.9)), clean=bool(seed % 4 == 1))
    if seed % 5 == 2 and len(inp["edge_other"]) > 4:  # duplicate map edges (two edges to the same other point)
        inp["edge_other"][1::7] = inp["edge_other"][0::7][:len(inp["edge_other"][1::7])]
    ratio = float(rs.choice([0.6, 0.8, 0.95]))
    got = O.extend_map_matches(cam, inp["map_desc"], inp["candidate"], inp["observed"], inp["bad"], inp["edge_off"],
                               inp["edge_other"], inp["edge_ok"], inp["proj_uv"], inp["view_cos"], inp["tracked"],
                               kx, ky, fd, inp["kp_mp"], es, ee, coff, cidx, th=th, ratio=ratio)
    mps, outside = build_pointer_graph(inp)
    F = dict(mnId=7, kx=kx, ky=ky, desc=fd, mvKeyEdges=list(zip(es.tolist(), ee.tolist())), mvConnected=conn,
             mvpMapPoints=[mps[r] if r >= 0 else (outside if r == -2 else None) for r in inp["kp_mp"]],
             mvpMapEdges=[None] * len(es))
    want_n = py_extend(cam, F, mps, th, ratio)
    assert got["nmatches"] == want_n
    want_kp = [(-1 if q is None else q.row) for q in F["mvpMapPoints"]]
    assert got["kp_mp"].tolist() == want_kp
    assert got["kedge_me"].tolist() == [(-1 if e is None else e.uid) for e in F["mvpMapEdges"]]
    assert got["tracked"].tolist() == [int(p.mnTrackedbyFrame == 7) for p in mps]
    # invariants: tracking only grows; every newly assigned map point is tracked, not bad, and assigned once
    assert (got["tracked"] >= inp["tracked"]).all()
    new = got["kp_mp"][got["kp_mp"] != inp["kp_mp"]]
    assert (new >= 0).all() and got["tracked"][new].all() and not inp["bad"][new].any()
    assert got["nmatches"] % 2 == 0 and got["nmatches"] <= 2 * int(inp["candidate"].sum())
    newly_tracked = np.nonzero(got["tracked"] > inp["tracked"])[0]
    assert len(newly_tracked) >= got["nmatches"] // 2
