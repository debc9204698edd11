"""Seeded synthetic inputs.  Pure numpy so the GPU box reproduces them bit-for-bit.

Uniform noise gives the networks nothing to detect, so frames are structured: random filled
rotated rectangles + anti-aliased line segments on a flat background, light blur, sensor noise.
"""
import numpy as np


def _blur3(img, sigma=0.8):
    k = np.exp(-0.5 * (np.arange(-1, 2) / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img, 1, mode="edge")
    t = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * t[:-2] + k[1] * t[1:-1] + k[2] * t[2:]


def frame(seed, width, height, n_rect=40, n_line=30):
    """-> (height, width) uint8 structured frame."""
    rs = np.random.RandomState(seed)
    W, H = width, height
    img = np.full((H, W), float(rs.randint(60, 120)), dtype=np.float64)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    for _ in range(n_rect):
        cx, cy = rs.uniform(0, W), rs.uniform(0, H)
        w, h = rs.uniform(20, max(21.0, W / 4)), rs.uniform(20, max(21.0, H / 4))
        a = rs.uniform(0, np.pi)
        g = float(rs.randint(0, 256))
        ca, sa = np.cos(a), np.sin(a)
        u = (xx - cx) * ca + (yy - cy) * sa
        v = -(xx - cx) * sa + (yy - cy) * ca
        # soft edge of ~1 px so rectangle borders are anti-aliased
        cov = np.clip(w / 2 - np.abs(u) + 0.5, 0, 1) * np.clip(h / 2 - np.abs(v) + 0.5, 0, 1)
        img = img * (1 - cov) + g * cov
    for _ in range(n_line):
        x0, y0, x1, y1 = rs.uniform(0, W), rs.uniform(0, H), rs.uniform(0, W), rs.uniform(0, H)
        t = rs.uniform(1, 3)
        g = float(rs.randint(0, 256))
        dx, dy = x1 - x0, y1 - y0
        L2 = dx * dx + dy * dy + 1e-9
        s = np.clip(((xx - x0) * dx + (yy - y0) * dy) / L2, 0, 1)
        d = np.hypot(xx - (x0 + s * dx), yy - (y0 + s * dy))
        cov = np.clip(t / 2 - d + 0.5, 0, 1)
        img = img * (1 - cov) + g * cov
    img = _blur3(img)
    img = img + rs.normal(0, 3, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frames(seeds, width, height):
    return np.stack([frame(s, width, height) for s in seeds])


def association_inputs(seed, frame_desc, kp_xy, n_map, width, height, th=10.0, planted_frac=0.5):
    """Map-side inputs for the image<->map association.

    frame_desc (N,256) unit rows; kp_xy (N,2) undistorted positions.
    -> dict(map_desc (M,256) f32, proj_uv (M,2) f32, view_cos (M,) f32, n_edges (M,) i32)
    Planted rows = a frame descriptor + N(0,0.05) noise renormalised, projected near its keypoint.
    """
    rs = np.random.RandomState(seed)
    N = frame_desc.shape[0]
    M = n_map
    d = rs.normal(size=(M, 256)).astype(np.float32)
    uv = np.stack([rs.uniform(0, width, M), rs.uniform(0, height, M)], 1).astype(np.float32)
    view_cos = rs.uniform(0.9, 1.0, M).astype(np.float32)
    if N > 0:
        n_pl = min(int(M * planted_frac), M)
        rows = rs.choice(M, n_pl, replace=False)
        src = rs.randint(0, N, n_pl)
        d[rows] = frame_desc[src] + rs.normal(0, 0.05, size=(n_pl, 256)).astype(np.float32)
        r = np.where(view_cos[rows] > 0.998, 2.5, 4.0) * th
        uv[rows] = kp_xy[src] + (rs.uniform(-0.5, 0.5, size=(n_pl, 2)) * r[:, None]).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    n_edges = rs.randint(0, 6, M).astype(np.int32)
    return dict(map_desc=d.astype(np.float32), proj_uv=uv.astype(np.float32),
                view_cos=view_cos, n_edges=n_edges)


def extend_inputs(seed, frame_desc, kp_xy, kedge_start, kedge_end, n_map, width, height, th=10.0, planted_frac=0.5,
                  clean=False):
    """Map-side inputs for the whole Matcher::ExtendMapMatches (Matcher.cpp:203-381): association_inputs plus a map
    graph.  Planted rows copy a keypoint (descriptor + noise, projection near it); for every key edge of the frame
    whose two endpoints have planted rows, those rows are joined by a map edge (so seed growing has something to
    grow along), plus random edges between arbitrary rows.  Unless `clean`, the state a running system would have
    is mixed in: non-candidate rows, bad rows, unobserved rows, invalid / dangling edges, rows already tracked and
    keypoints that already hold a map point.
    -> dict(map_desc, proj_uv, view_cos, candidate, observed, bad, edge_off, edge_other, edge_ok, tracked, kp_mp,
            planted_rows, planted_src)"""
    rs = np.random.RandomState(seed)
    N, M = frame_desc.shape[0], n_map
    d = rs.normal(size=(M, 256)).astype(np.float32)
    uv = np.stack([rs.uniform(0, width, M), rs.uniform(0, height, M)], 1).astype(np.float32)
    view_cos = rs.uniform(0.9, 1.0, M).astype(np.float32)
    rows = np.zeros(0, np.int64)
    src = np.zeros(0, np.int64)
    if N > 0:
        n_pl = min(int(M * planted_frac), M)
        rows = rs.choice(M, n_pl, replace=False)
        src = rs.randint(0, N, n_pl)
        d[rows] = frame_desc[src] + rs.normal(0, 0.05, size=(n_pl, 256)).astype(np.float32)
        r = np.where(view_cos[rows] > 0.998, 2.5, 4.0) * th
        uv[rows] = kp_xy[src] + (rs.uniform(-0.5, 0.5, size=(n_pl, 2)) * r[:, None]).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    first_row = {}
    for rw, s in zip(rows.tolist(), src.tolist()):
        first_row.setdefault(s, rw)
    adj = [[] for _ in range(M)]  # (other, ok)
    for a, b in zip(np.asarray(kedge_start).tolist(), np.asarray(kedge_end).tolist()):
        if a in first_row and b in first_row and rs.rand() < 0.85:
            ra, rb = first_row[a], first_row[b]
            ok = 1 if (clean or rs.rand() < 0.93) else 0
            adj[ra].append((rb, ok))
            adj[rb].append((ra, ok))
    for _ in range(0 if clean else M // 4):
        ra, rb = rs.randint(0, M, 2)
        if ra == rb:
            continue
        ok = 1 if rs.rand() < 0.9 else 0
        adj[ra].append((rb if rs.rand() < 0.95 else -1, ok))  # -1: theOtherPt() == nullptr
        adj[rb].append((ra, ok))
    edge_off = np.zeros(M + 1, np.int32)
    edge_off[1:] = np.cumsum([len(a) for a in adj])
    edge_other = np.array([o for a in adj for o, _ in a], np.int32)
    edge_ok = np.array([k for a in adj for _, k in a], np.uint8)
    if clean:
        candidate, observed, bad = np.ones(M, np.uint8), np.ones(M, np.uint8), np.zeros(M, np.uint8)
        tracked, kp_mp = np.zeros(M, np.uint8), np.full(N, -1, np.int32)
    else:
        candidate = (rs.rand(M) < 0.9).astype(np.uint8)
        observed = (rs.rand(M) < 0.95).astype(np.uint8)
        bad = ((rs.rand(M) < 0.03) & (candidate == 0)).astype(np.uint8)
        tracked = (rs.rand(M) < 0.02).astype(np.uint8)
        kp_mp = np.full(N, -1, np.int32)
        pre = rs.rand(N) < 0.08
        kp_mp[pre] = rs.randint(0, M, int(pre.sum()))
        kp_mp[rs.rand(N) < 0.03] = -2
    return dict(map_desc=d.astype(np.float32), proj_uv=uv.astype(np.float32), view_cos=view_cos, candidate=candidate,
                observed=observed, bad=bad, edge_off=edge_off, edge_other=edge_other, edge_ok=edge_ok, tracked=tracked,
                kp_mp=kp_mp, planted_rows=rows.astype(np.int32), planted_src=src.astype(np.int32))


def extend_inputs_multi(seed, recs, n_map, width, height, th=10.0, rows_per_frame=None, n_slices=None):
    """ONE resident map for a batch of frames, every frame with its own local map: the table is cut into slices,
    slice f holds map points planted around the keypoints of frame f (descriptor + noise, map edges mirroring the
    frame's point-pair graph, as extend_inputs does for a single frame), and in frame f's projections those rows land
    near their keypoints while all other rows project somewhere else in the image.  Every frame's
    Matcher::ExtendMapMatches walk then accepts and grows a few hundred matches -- what a SLAM run looks like, where
    every frame matches its local map -- instead of only the first one.  `recs`: per-frame dicts with kp_x, kp_y, desc,
    edge_start, edge_end.  n_slices (default len(recs)) fixes the slice size when fewer frames than slices are given
    (the CPU arm of bench.py processes a sample of the batch against a table of the same shape).
    -> dict like extend_inputs (clean state) + proj_all (F, M, 2), vcos_all (F, M), slices [(row0, rows)]"""
    rs = np.random.RandomState(seed)
    F, M = len(recs), n_map
    n_slices = n_slices or F
    per = rows_per_frame or M // n_slices
    d = rs.normal(size=(M, 256)).astype(np.float32)
    view_cos = rs.uniform(0.9, 1.0, M).astype(np.float32)
    proj_all = np.stack([np.stack([rs.uniform(0, width, M), rs.uniform(0, height, M)], 1) for _ in range(F)]).astype(
        np.float32)
    adj = [[] for _ in range(M)]
    slices = []
    for f, r in enumerate(recs):
        r0 = f * per
        slices.append((r0, per))
        N = len(r["kp_x"])
        if N == 0:
            continue
        kp_xy = np.stack([r["kp_x"], r["kp_y"]], 1)
        src = rs.permutation(N)[:per] if N >= per else rs.randint(0, N, per)
        rows = r0 + np.arange(len(src))
        d[rows] = r["desc"][src] + rs.normal(0, 0.05, size=(len(src), 256)).astype(np.float32)
        rad = np.where(view_cos[rows] > 0.998, 2.5, 4.0) * th
        proj_all[f, rows] = kp_xy[src] + (rs.uniform(-0.5, 0.5, size=(len(src), 2)) * rad[:, None]).astype(np.float32)
        first_row = {}
        for rw, sidx in zip(rows.tolist(), src.tolist()):
            first_row.setdefault(sidx, rw)
        for a, b in zip(np.asarray(r["edge_start"]).tolist(), np.asarray(r["edge_end"]).tolist()):
            if a in first_row and b in first_row and rs.rand() < 0.85:
                ra, rb = first_row[a], first_row[b]
                adj[ra].append((rb, 1))
                adj[rb].append((ra, 1))
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    edge_off = np.zeros(M + 1, np.int32)
    edge_off[1:] = np.cumsum([len(a) for a in adj])
    edge_other = np.array([o for a in adj for o, _ in a], np.int32)
    edge_ok = np.array([k for a in adj for _, k in a], np.uint8)
    vcos_all = np.ascontiguousarray(np.broadcast_to(view_cos, (F, M)), np.float32)
    return dict(map_desc=d.astype(np.float32), proj_uv=proj_all[0].copy(), view_cos=view_cos,
                candidate=np.ones(M, np.uint8), observed=np.ones(M, np.uint8), bad=np.zeros(M, np.uint8),
                edge_off=edge_off, edge_other=edge_other, edge_ok=edge_ok, tracked=np.zeros(M, np.uint8),
                proj_all=proj_all, vcos_all=vcos_all, slices=slices)


def frustum_inputs(seed, cam, n_points, n_frames=1):
    """Map geometry + camera poses for Frame::CheckInFrustum: points scattered in front of (and partly behind /
    beside) a camera near the origin, mean viewing directions (MapPoint::GetNormal) roughly along the ray, scale-invariance distance bands, and small random
    rigid motions per frame.  -> dict(world_pos, normal, min_dist, max_dist, Rcw (F,3,3), tcw (F,3), Ow (F,3))"""
    rs = np.random.RandomState(seed)
    fx, fy, cx, cy = cam.K[0], cam.K[4], cam.K[2], cam.K[5]
    z = rs.uniform(1.0, 12.0, n_points)
    u = rs.uniform(-0.3 * cam.width, 1.3 * cam.width, n_points)
    v = rs.uniform(-0.3 * cam.height, 1.3 * cam.height, n_points)
    P = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    behind = rs.rand(n_points) < 0.05
    P[behind, 2] *= -1
    n = P / np.linalg.norm(P, axis=1, keepdims=True) + rs.normal(0, 0.6, P.shape)  # mean viewing direction
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = np.linalg.norm(P, axis=1)
    min_d = d * rs.uniform(0.3, 1.05, n_points)
    max_d = d * rs.uniform(0.95, 3.0, n_points)
    Rs, ts, Os = [], [], []
    for _ in range(n_frames):
        w = rs.normal(0, 0.03, 3)
        th = np.linalg.norm(w)
        k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        t = rs.normal(0, 0.1, 3)
        Rs.append(R)
        ts.append(t)
        Os.append(-R.T @ t)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(world_pos=f(P), normal=f(n), min_dist=f(min_d), max_dist=f(max_d), Rcw=f(np.stack(Rs)),
                tcw=f(np.stack(ts)), Ow=f(np.stack(Os)))


def two_view_inputs(seed, cam, n1=300, n2=320, n_nodes=40, frac_mp=0.2, noise_px=0.6, desc_noise=0.05, forward=False):
    """Two key frames (pinhole or KannalaBrandt8, by cam.fisheye) looking at the same random 3-D points
    (Matcher::SearchForTriangulation, Matcher.cpp:767-885):
    world -> camera poses, undistorted pixel positions, unit descriptors (those of a common point differ by desc_noise),
    one vocabulary node per feature (common points share it), a fraction of features already carrying a map point.
    Pixel noise around the epipolar test's 3.84 threshold so that both of its outcomes occur."""
    r = np.random.RandomState(seed)
    fx, fy, cx, cy = float(cam.K[0]), float(cam.K[4]), float(cam.K[2]), float(cam.K[5])

    def rot(ax, ang):
        ax = np.asarray(ax, np.float64)
        ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    R1, t1 = rot(r.randn(3), 0.05 * r.rand()), 0.1 * r.randn(3)
    R2, t2 = rot(r.randn(3), 0.15 * r.rand()), np.array([0.4, 0.05, 0.02]) * (1 + r.rand(3))
    if forward:  # motion along the optical axis: the epipole falls inside the image (the 10-pixel exclusion of :846-847)
        R2, t2 = rot(r.randn(3), 0.02 * r.rand()), np.array([0.01, 0.01, -0.8]) * (1 + 0.2 * r.rand(3))
    n_common = min(n1, n2) * 2 // 3
    X = np.stack([r.uniform(-3, 3, n_common), r.uniform(-2, 2, n_common), r.uniform(4, 12, n_common)], 1)

    def proj(R, t, P):
        Pc = P @ R.T + t
        if cam.fisheye:  # KannalaBrandt8::project (sensors/src/KannalaBrandt8.cpp:26-42): the positions are raw pixels
            th = np.arctan2(np.hypot(Pc[:, 0], Pc[:, 1]), Pc[:, 2])
            psi = np.arctan2(Pc[:, 1], Pc[:, 0])
            k = [float(v) for v in cam.D]
            rr = th + k[0] * th ** 3 + k[1] * th ** 5 + k[2] * th ** 7 + k[3] * th ** 9
            return np.stack([fx * rr * np.cos(psi) + cx, fy * rr * np.sin(psi) + cy], 1)
        return np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
    d_common = r.randn(n_common, 256)
    node_common = r.randint(0, n_nodes, n_common)
    out = {}
    for k, (R, t, n) in enumerate(((R1, t1, n1), (R2, t2, n2)), 1):
        pos = np.empty((n, 2))
        pos[:n_common] = proj(R, t, X) + noise_px * r.randn(n_common, 2) * r.choice([1.0, 3.0], (n_common, 1))
        pos[n_common:] = np.stack([r.uniform(0, cam.width, n - n_common), r.uniform(0, cam.height, n - n_common)], 1)
        d = np.empty((n, 256))
        d[:n_common] = d_common + desc_noise * 16 * r.randn(n_common, 256) * r.choice([0.2, 0.45, 1.2], (n_common, 1))
        d[n_common:] = r.randn(n - n_common, 256)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        node = np.empty(n, np.int32)
        node[:n_common] = node_common
        node[n_common:] = r.randint(0, n_nodes, n - n_common)
        perm = r.permutation(n)  # feature order is unrelated to the point identity
        out["pos%d" % k] = pos[perm].astype(np.float32)
        out["desc%d" % k] = d[perm].astype(np.float32)
        out["node%d" % k] = node[perm]
        out["has_mp%d" % k] = (r.rand(n) < frac_mp).astype(np.uint8)
        out["R%d" % k] = R.astype(np.float32)
        out["t%d" % k] = t.astype(np.float32)
    return out


def bow_pair_inputs(seed, n1=300, n2=280, n_nodes=12, frac_none=0.25, frac_bad=0.1, noise=0.35, dup=0.15):
    """Two feature sets for Matcher::SearchByBoW (Matcher.cpp:393-477, :663-754): unit descriptors, the second set holding
    noisy copies of part of the first (some of them twice -- near-duplicates that fail the ratio test or compete for the same
    feature), one vocabulary node per feature (a few features unlisted: -1), and a map-point state per feature (0 none,
    1 good, 2 bad)."""
    r = np.random.RandomState(seed)
    d1 = r.randn(n1, 256)
    node1 = r.randint(0, n_nodes, n1)
    n_common = min(n1, n2) * 3 // 5
    src = r.permutation(n1)[:n_common]
    d2 = r.randn(n2, 256)
    node2 = r.randint(0, n_nodes, n2)
    dst = r.permutation(n2)[:n_common]
    d2[dst] = d1[src] + noise * r.randn(n_common, 256) * r.choice([0.5, 1.0, 2.0], (n_common, 1))
    node2[dst] = node1[src]
    n_dup = int(dup * n_common)
    free = np.setdiff1d(np.arange(n2), dst)[:n_dup]
    if len(free):
        d2[free] = d2[dst[:len(free)]] + 0.15 * r.randn(len(free), 256)
        node2[free] = node2[dst[:len(free)]]
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    node1[r.rand(n1) < 0.03] = -1
    node2[r.rand(n2) < 0.03] = -1

    def states(n):
        u = r.rand(n)
        return np.where(u < frac_none, 0, np.where(u < frac_none + frac_bad, 2, 1)).astype(np.uint8)
    return dict(desc1=d1.astype(np.float32), node1=node1.astype(np.int32), state1=states(n1),
                desc2=d2.astype(np.float32), node2=node2.astype(np.int32), state2=states(n2))


def bow_rows(desc, node, state):
    """The rows Matcher::SearchByBoW visits: features with a good map point in FeatureVector order (node ascending, then
    index).  -> (feature index of every row, row descriptors, row nodes)"""
    idx = np.array([i for i in np.lexsort((np.arange(len(node)), node)) if node[i] >= 0 and state[i] == 1], np.int64)
    return idx, desc[idx], node[idx].astype(np.int32)


def projection_inputs(seed, cam, n_src=300, n=340, noise_px=2.0, desc_noise=0.35, frac_dup=0.25, frac_assigned=0.1):
    """A current frame and a source of map points (the last frame / a key frame) for Matcher::SearchByProjection
    (Matcher.cpp:31-87, :1337-1411): a pose, map points that project onto (or near, or next to) the current frame's
    keypoints -- several of them onto the SAME keypoint with near-identical descriptors, so that the order of the walk and
    the live occupancy decide --, points behind the camera and outside the image, source features without a point,
    outliers / bad points, unobserved points (they do not occupy), pre-assigned keypoints, distance bands."""
    r = np.random.RandomState(seed)
    fx, fy, cx, cy = float(cam.K[0]), float(cam.K[4]), float(cam.K[2]), float(cam.K[5])

    def proj(Pc):
        if cam.fisheye:
            th = np.arctan2(np.hypot(Pc[:, 0], Pc[:, 1]), Pc[:, 2])
            psi = np.arctan2(Pc[:, 1], Pc[:, 0])
            k = [float(v) for v in cam.D]
            rr = th + k[0] * th ** 3 + k[1] * th ** 5 + k[2] * th ** 7 + k[3] * th ** 9
            return np.stack([fx * rr * np.cos(psi) + cx, fy * rr * np.sin(psi) + cy], 1)
        return np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
    w = r.normal(0, 0.05, 3)
    ang = np.linalg.norm(w)
    k = w / ang
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    Rcw = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    tcw = r.normal(0, 0.2, 3)
    # camera-frame points spread over (and a little beyond) the field of view
    z = r.uniform(1.5, 10.0, n_src)
    span = 0.55 if not cam.fisheye else 0.9
    Pc = np.stack([r.uniform(-span, span, n_src) * z * cam.width / (2 * fx) * 2,
                   r.uniform(-span, span, n_src) * z * cam.height / (2 * fy) * 2, z], 1)
    n_dup = int(frac_dup * n_src)
    for i in range(n_src - n_dup, n_src):  # the same spot as an earlier point, a little deeper
        j = r.randint(0, n_src - n_dup)
        Pc[i] = Pc[j] * (1 + 0.02 * r.rand())
    behind = r.rand(n_src) < 0.04
    Pc[behind, 2] *= -1
    uv = proj(np.where(behind[:, None], Pc * [1, 1, -1], Pc))
    world = (Pc - tcw) @ Rcw  # Rcw^T (Pc - tcw)
    # current-frame keypoints: on the projections (with noise around the window radius), the rest anywhere
    kp = np.stack([r.uniform(0, cam.width, n), r.uniform(0, cam.height, n)], 1)
    kd = r.randn(n, 256)
    n_on = min(n, n_src - n_dup) * 3 // 4
    tgt = r.permutation(n_src - n_dup)[:n_on]
    kp[:n_on] = uv[tgt] + noise_px * r.randn(n_on, 2) * r.choice([0.5, 1.0, 4.0], (n_on, 1))
    kd /= np.linalg.norm(kd, axis=1, keepdims=True)
    md = r.randn(n_src, 256)
    md /= np.linalg.norm(md, axis=1, keepdims=True)
    md[tgt] = kd[:n_on] + desc_noise / 16 * r.randn(n_on, 256) * r.choice([0.3, 1.0, 2.2], (n_on, 1))
    for i in range(n_src - n_dup, n_src):
        j = int(np.argmin(np.abs(Pc[:n_src - n_dup] - Pc[i] / np.linalg.norm(Pc[i]) * np.linalg.norm(Pc[:n_src - n_dup], axis=1, keepdims=True)).sum(1)))
        md[i] = md[j] + 0.01 * r.randn(256)
    md /= np.linalg.norm(md, axis=1, keepdims=True)
    perm = r.permutation(n)
    kp, kd = kp[perm], kd[perm]
    u = r.rand(n_src)
    state = np.where(u < 0.08, 0, np.where(u < 0.16, 2, np.where(u < 0.2, 3, 1))).astype(np.uint8)
    observed = (r.rand(n_src) < 0.8).astype(np.uint8)
    d3 = np.linalg.norm(Pc, axis=1)
    min_d = d3 * r.uniform(0.3, 1.03, n_src)
    max_d = d3 * r.uniform(0.97, 3.0, n_src)
    kp_mp = np.full(n, -1, np.int32)
    pre = np.nonzero(r.rand(n) < frac_assigned)[0]
    for i in pre:
        c = r.randint(3)
        kp_mp[i] = -2 if c == 0 else (-3 if c == 1 else r.randint(n_src))
    for i in pre:  # a pre-assigned source feature must hold a point
        if kp_mp[i] >= 0 and state[kp_mp[i]] == 0:
            kp_mp[i] = -2
    f = lambda a: np.ascontiguousarray(a, np.float32)
    Ow = -Rcw.T @ tcw
    nrm = (world - Ow) / np.linalg.norm(world - Ow, axis=1, keepdims=True) + r.normal(0, 0.5, world.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)  # mean viewing direction, roughly along the ray (GetNormal)
    inside = (~behind) & (uv[:, 0] >= 0) & (uv[:, 0] < cam.width) & (uv[:, 1] >= 0) & (uv[:, 1] < cam.height)
    return dict(Rcw=f(Rcw), tcw=f(tcw), world_pos=f(world), mp_desc=f(md), state=state, observed=observed,
                min_dist=f(min_d), max_dist=f(max_d), kp_x=f(kp[:, 0]), kp_y=f(kp[:, 1]), desc=f(kd), kp_mp=kp_mp,
                uv_numpy=f(uv), inside_numpy=inside.astype(np.uint8), normal=f(nrm))  # the projection in numpy (no reference needed)


def projection_rows(x, mode, ref_row_valid, ref_proj_uv):
    """Flat source features -> the rows the C ABI / the oracle take (the features the reference's tests let through, in
    loop order) and CurrentFrame.mvpMapPoints recoded as rows.  -> dict(rows, map_desc, proj_uv, observed, kp_mp)"""
    rows = np.nonzero(ref_row_valid)[0].astype(np.int32)
    row_of = {int(s): k for k, s in enumerate(rows)}
    km = np.full(len(x["kp_mp"]), -1, np.int32)
    for i, v in enumerate(x["kp_mp"]):
        v = int(v)
        if mode >= 1:  # relocalisation (:1386) and the key-frame variant (:543) test the pointer alone
            km[i] = -1 if v == -1 else -2
        elif v >= 0:
            km[i] = row_of[v] if v in row_of else (-2 if x["observed"][v] else -1)
        else:
            km[i] = -2 if v == -2 else -1
    return dict(rows=rows, map_desc=x["mp_desc"][rows], proj_uv=ref_proj_uv[rows],
                observed=None if mode >= 1 else x["observed"][rows], kp_mp=km)


def projection_result(x, rows, kp_mp_rows):
    """kp_mp in rows after the call -> the reference's coding (source feature index; untouched keypoints keep theirs)."""
    out = np.array(x["kp_mp"], np.int32).copy()
    for i, v in enumerate(kp_mp_rows):
        if v >= 0:
            out[i] = rows[v]
    return out


def sim3_inputs(seed, cam, n1=300, n2=320, noise_px=1.0, desc_noise=0.3, scale=1.03):
    """Two key frames of two maps that see the same place (Matcher::SearchBySim3, Matcher.cpp:1149-1335): each feature may
    carry its OWN map point (the two maps are not merged yet) -- those of a common 3-D point lie close to each other in
    the world and have similar descriptors --, a similarity S12 (camera 2 -> camera 1) that is a little off the truth,
    bad points, features without a point, distance bands, a few entries of vpMatches12 already filled."""
    r = np.random.RandomState(seed)
    fx, fy, cx, cy = float(cam.K[0]), float(cam.K[4]), float(cam.K[2]), float(cam.K[5])

    def rot(ax, ang):
        ax = np.asarray(ax, np.float64)
        ax /= np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K

    def proj(Pc):
        if cam.fisheye:
            th = np.arctan2(np.hypot(Pc[:, 0], Pc[:, 1]), Pc[:, 2])
            psi = np.arctan2(Pc[:, 1], Pc[:, 0])
            k = [float(v) for v in cam.D]
            rr = th + k[0] * th ** 3 + k[1] * th ** 5 + k[2] * th ** 7 + k[3] * th ** 9
            return np.stack([fx * rr * np.cos(psi) + cx, fy * rr * np.sin(psi) + cy], 1)
        return np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
    poses = [(rot(r.randn(3), 0.05 * r.rand()), 0.1 * r.randn(3)),
             (rot(r.randn(3), 0.12 * r.rand()), np.array([0.3, 0.04, 0.02]) * (1 + r.rand(3)))]
    n_common = min(n1, n2) * 2 // 3
    span = 0.45 if not cam.fisheye else 0.8
    z = r.uniform(3.0, 11.0, n_common)
    X = np.stack([r.uniform(-span, span, n_common) * z * cam.width / fx, r.uniform(-span, span, n_common) * z * cam.height / fy,
                  z], 1)
    d_common = r.randn(n_common, 256)
    out = {}
    for k, ((R, t), n) in enumerate(zip(poses, (n1, n2)), 1):
        nc = min(n_common, n)
        Pc = np.empty((n, 3))
        Pc[:nc] = X[:nc] @ R.T + t
        zz = r.uniform(3.0, 11.0, n - nc)
        Pc[nc:] = np.stack([r.uniform(-span, span, n - nc) * zz * cam.width / fx,
                            r.uniform(-span, span, n - nc) * zz * cam.height / fy, zz], 1)
        pos = proj(Pc) + noise_px * r.randn(n, 2)
        d = np.empty((n, 256))
        d[:nc] = d_common[:nc] + desc_noise * r.randn(nc, 256) * r.choice([0.5, 1.0, 3.0], (nc, 1))
        d[nc:] = r.randn(n - nc, 256)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        world = (Pc - t) @ R + r.normal(0, 0.01, (n, 3))  # the key frame's own map point of the feature
        md = d + 0.02 * r.randn(n, 256)
        md /= np.linalg.norm(md, axis=1, keepdims=True)
        d3 = np.linalg.norm(Pc, axis=1)
        u = r.rand(n)
        perm = r.permutation(n)
        f = lambda a: np.ascontiguousarray(a[perm], np.float32)
        out.update({"R%d" % k: R.astype(np.float32), "t%d" % k: t.astype(np.float32), "pos%d" % k: f(pos),
                    "desc%d" % k: f(d), "state%d" % k: np.where(u < 0.12, 0, np.where(u < 0.18, 2, 1)).astype(np.uint8)[perm],
                    "world%d" % k: f(world), "mp_desc%d" % k: f(md),
                    "min_dist%d" % k: f(d3 * r.uniform(0.3, 1.0, n)), "max_dist%d" % k: f(d3 * r.uniform(1.0, 3.0, n))})
    (R1, t1), (R2, t2) = poses
    R12 = R1 @ R2.T
    t12 = t1 - R12 @ t2
    out["R12"], out["t12"] = R12.astype(np.float32), (t12 * scale + r.normal(0, 0.003, 3)).astype(np.float32)
    out["s12"] = np.float32(scale)
    m12 = np.full(n1, -1, np.int32)
    for i in np.nonzero(r.rand(n1) < 0.05)[0]:
        j = r.randint(n2)
        if out["state2"][j]:
            m12[i] = j
    out["matches12"] = m12
    return out
