"""Inputs of the association-level golden vectors (tests/golden/l2_association.npz).  `inputs()` generates them (used
only by tests/golden/make_golden_l2.py, which stores them in the fixture next to the oracle's outputs: numpy's float32
reductions may round differently on another CPU, so the tests never regenerate them); `load()` reads the fixture."""
import os

import numpy as np

from ppg_slam_b200 import cameras, synth, vocabulary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def inputs():
    cam = cameras.EUROC
    rs = np.random.RandomState(2024)
    n, n_edges, M = 100, 220, 400
    gx, gy = np.meshgrid(np.arange(8, cam.width - 8, 6), np.arange(8, cam.height - 8, 6))
    sel = rs.choice(gx.size, n, replace=False)
    kx = gx.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    ky = gy.ravel()[sel].astype(np.float32) + rs.uniform(-0.4, 0.4, n).astype(np.float32)
    voc = vocabulary.load_blob(os.path.join(ROOT, "ppg_slam_b200", "weights", "voc_euroc_9x3.bin"))
    leaves = np.nonzero(voc.word_id >= 0)[0]
    fd = voc.desc[leaves[rs.randint(0, len(leaves), n)]] + rs.normal(0, 0.05, (n, 256)).astype(np.float32)
    fd = (fd / np.linalg.norm(fd, axis=1, keepdims=True)).astype(np.float32)
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
    coff = np.zeros(n + 1, np.int32)
    coff[1:] = np.cumsum([len(c) for c in conn])
    cidx = np.array([e for c in conn for e in c], np.int32)
    ext = synth.extend_inputs(31, fd, np.stack([kx, ky], 1), es, ee, M, cam.width, cam.height, th=10.0,
                              planted_frac=0.5, clean=False)
    geo = synth.frustum_inputs(32, cam, M, n_frames=1)
    kd = fd[rs.randint(0, n // 3, 80)] + rs.normal(0, 0.03, (80, 256)).astype(np.float32)
    kd = (kd / np.linalg.norm(kd, axis=1, keepdims=True)).astype(np.float32)
    return dict(cam=cam, kx=kx, ky=ky, fd=fd, es=es, ee=ee, coff=coff, cidx=cidx, ext=ext, geo=geo, voc=voc, kd=kd)


EXT_KEYS = ("map_desc", "candidate", "observed", "bad", "edge_off", "edge_other", "edge_ok", "proj_uv", "view_cos",
            "tracked", "kp_mp")
GEO_KEYS = ("world_pos", "normal", "min_dist", "max_dist", "Rcw", "tcw", "Ow")
FRAME_KEYS = ("kx", "ky", "fd", "es", "ee", "coff", "cidx", "kd")


def pack(x):
    d = {"in_" + k: x[k] for k in FRAME_KEYS}
    d.update({"in_ext_" + k: x["ext"][k] for k in EXT_KEYS})
    d.update({"in_geo_" + k: x["geo"][k] for k in GEO_KEYS})
    return d


def load(path):
    g = np.load(path)
    x = {k: g["in_" + k] for k in FRAME_KEYS}
    x["ext"] = {k: g["in_ext_" + k] for k in EXT_KEYS}
    x["geo"] = {k: g["in_geo_" + k] for k in GEO_KEYS}
    x["cam"] = cameras.EUROC
    x["voc"] = vocabulary.load_blob(os.path.join(ROOT, "ppg_slam_b200", "weights", "voc_euroc_9x3.bin"))
    return x, g
