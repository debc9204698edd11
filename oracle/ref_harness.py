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
