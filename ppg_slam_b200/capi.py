"""ctypes binding of the C ABI in include/ppg_b200.h (libppg_b200.so).

This is the same binding a maintainer would write for any FFI host; tests and bench.py drive the
product exclusively through it.  There is no fallback: if the library is missing, loading raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libppg_b200.so")
DEFAULT_WEIGHTS = os.path.join(HERE, "weights", "ppg_weights.bin")

PPG_OK, PPG_ERR_ARG, PPG_ERR_CUDA, PPG_ERR_WEIGHTS, PPG_ERR_CAPACITY, PPG_ERR_NCCL = 0, -1, -2, -3, -4, -5


class PpgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ppg error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("K", C.c_float * 9),
                ("D", C.c_float * 4), ("fisheye", C.c_int), ("weights_path", C.c_char_p),
                ("junction_thresh", C.c_float), ("junction_nms_radius", C.c_int), ("junction_max_num", C.c_int),
                ("line_valid_thresh", C.c_float), ("line_valid_ratio", C.c_float), ("line_dist_thresh", C.c_float),
                ("heatmap_refine_sz", C.c_int), ("line_heatmap_thresh", C.c_float), ("line_inlier_rate", C.c_float),
                ("th_low", C.c_float), ("th_high", C.c_float), ("max_batch", C.c_int), ("max_edges", C.c_int),
                ("max_colines", C.c_int), ("max_map_points", C.c_int)]


class FrameOut(C.Structure):
    _fields_ = [("n_kp", C.c_int), ("n_edges", C.c_int), ("n_colines", C.c_int), ("status", C.c_uint32),
                ("n_candidates", C.c_int), ("n_pairs_tested_ok", C.c_int), ("n_candidate_lines", C.c_int),
                ("nms_rounds", C.c_int),
                ("kp_x", C.POINTER(C.c_float)), ("kp_y", C.POINTER(C.c_float)),
                ("kp_px", C.POINTER(C.c_int32)), ("kp_py", C.POINTER(C.c_int32)),
                ("kp_score", C.POINTER(C.c_float)), ("kp_xun", C.POINTER(C.c_float)),
                ("kp_yun", C.POINTER(C.c_float)), ("kp_out", C.POINTER(C.c_uint8)),
                ("edge_start", C.POINTER(C.c_int32)), ("edge_end", C.POINTER(C.c_int32)),
                ("edge_score", C.POINTER(C.c_float)), ("conn_off", C.POINTER(C.c_int32)),
                ("conn_idx", C.POINTER(C.c_int32)), ("col_off", C.POINTER(C.c_int32)),
                ("col_pairs", C.POINTER(C.c_int32)), ("desc", C.POINTER(C.c_float)), ("diag", C.c_int * 8)]


class AssocIn(C.Structure):
    _fields_ = [("n_kp", C.c_int), ("kp_x", C.POINTER(C.c_float)), ("kp_y", C.POINTER(C.c_float)),
                ("frame_desc", C.POINTER(C.c_float)), ("free_mask", C.POINTER(C.c_uint8)), ("n_rows", C.c_int),
                ("proj_uv", C.POINTER(C.c_float)), ("view_cos", C.POINTER(C.c_float)), ("th", C.c_float),
                ("ratio", C.c_float), ("mode", C.c_int), ("max_dist", C.c_float), ("e2_max", C.c_double),
                ("row_node", C.POINTER(C.c_int32)), ("kp_node", C.POINTER(C.c_int32))]


SEARCH_EXTEND_MAP, SEARCH_WINDOW, SEARCH_NODE = 0, 1, 2


class AssocOut(C.Structure):
    _fields_ = [("best_idx", C.POINTER(C.c_int32)), ("second_idx", C.POINTER(C.c_int32)),
                ("best_dist", C.POINTER(C.c_float)), ("second_dist", C.POINTER(C.c_float)),
                ("accept", C.POINTER(C.c_uint8))]


class MapGraph(C.Structure):
    _fields_ = [("n_points", C.c_int), ("candidate", C.POINTER(C.c_uint8)), ("observed", C.POINTER(C.c_uint8)),
                ("bad", C.POINTER(C.c_uint8)), ("edge_off", C.POINTER(C.c_int32)),
                ("edge_other", C.POINTER(C.c_int32)), ("edge_ok", C.POINTER(C.c_uint8))]


class ExtendIn(C.Structure):
    _fields_ = [("n_kp", C.c_int), ("kp_x", C.POINTER(C.c_float)), ("kp_y", C.POINTER(C.c_float)),
                ("frame_desc", C.POINTER(C.c_float)), ("kp_mp", C.POINTER(C.c_int32)), ("n_edges", C.c_int),
                ("edge_start", C.POINTER(C.c_int32)), ("edge_end", C.POINTER(C.c_int32)),
                ("conn_off", C.POINTER(C.c_int32)), ("conn_idx", C.POINTER(C.c_int32)),
                ("kedge_me", C.POINTER(C.c_int32)), ("proj_uv", C.POINTER(C.c_float)),
                ("view_cos", C.POINTER(C.c_float)), ("tracked", C.POINTER(C.c_uint8)), ("th", C.c_float),
                ("ratio", C.c_float)]


class ExtendOut(C.Structure):
    _fields_ = [("kp_mp", C.POINTER(C.c_int32)), ("kedge_me", C.POINTER(C.c_int32)),
                ("tracked", C.POINTER(C.c_uint8)), ("nmatches", C.c_int), ("status", C.c_uint32),
                ("n_kp", C.c_int), ("n_edges", C.c_int), ("n_accepted", C.c_int), ("n_grown", C.c_int),
                ("n_rescans", C.c_int), ("diag", C.c_int * 9)]


class VocabularyPod(C.Structure):
    _fields_ = [("k", C.c_int), ("L", C.c_int), ("scoring", C.c_int), ("weighting", C.c_int), ("n_nodes", C.c_int),
                ("dim", C.c_int), ("children", C.POINTER(C.c_int32)), ("word_id", C.POINTER(C.c_int32)),
                ("weight", C.POINTER(C.c_double)), ("desc", C.POINTER(C.c_float))]


class BowOut(C.Structure):
    _fields_ = [("n_features", C.c_int), ("n_bow", C.c_int), ("word_id", C.POINTER(C.c_int32)),
                ("word_weight", C.POINTER(C.c_double)), ("node_id", C.POINTER(C.c_int32)),
                ("bow_word", C.POINTER(C.c_int32)), ("bow_value", C.POINTER(C.c_double))]


class BowMatchIn(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("row_node", C.POINTER(C.c_int32)), ("n_kp", C.c_int),
                ("frame_desc", C.POINTER(C.c_float)), ("kp_node", C.POINTER(C.c_int32)), ("ratio", C.c_float),
                ("max_dist", C.c_float), ("strict", C.c_int)]


class BowMatchOut(C.Structure):
    _fields_ = [("kp_row", C.POINTER(C.c_int32)), ("nmatches", C.c_int), ("n_rescans", C.c_int)]


# every symbol include/ppg_b200.h declares (tests/test_abi.py checks the header against this list)
class InitMatchIn(C.Structure):
    _fields_ = [("n1", C.c_int), ("prev_matched", C.POINTER(C.c_float)), ("n2", C.c_int),
                ("kp2_x", C.POINTER(C.c_float)), ("kp2_y", C.POINTER(C.c_float)), ("desc2", C.POINTER(C.c_float)),
                ("window", C.c_int), ("ratio", C.c_float)]


class InitMatchOut(C.Structure):
    _fields_ = [("matches12", C.POINTER(C.c_int32)), ("prev_matched", C.POINTER(C.c_float)), ("nmatches", C.c_int),
                ("n_rescans", C.c_int)]


class TriangulationMatchIn(C.Structure):
    _fields_ = [("n1", C.c_int), ("n2", C.c_int), ("desc1", C.POINTER(C.c_float)), ("desc2", C.POINTER(C.c_float)),
                ("node1", C.POINTER(C.c_int32)), ("node2", C.POINTER(C.c_int32)),
                ("has_mp1", C.POINTER(C.c_uint8)), ("has_mp2", C.POINTER(C.c_uint8)),
                ("pos1", C.POINTER(C.c_float)), ("pos2", C.POINTER(C.c_float)),
                ("F12", C.c_float * 9), ("epipole", C.c_float * 2), ("th_low", C.c_float),
                ("camera_model", C.c_int), ("cam8", C.c_float * 8), ("R12", C.c_float * 9), ("t12", C.c_float * 3)]


class ProjectionMatchIn(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("proj_uv", C.POINTER(C.c_float)), ("observed", C.POINTER(C.c_uint8)),
                ("n", C.c_int), ("kp_x", C.POINTER(C.c_float)), ("kp_y", C.POINTER(C.c_float)),
                ("desc", C.POINTER(C.c_float)), ("kp_mp", C.POINTER(C.c_int32)), ("th", C.c_float),
                ("max_dist", C.c_float)]


class ProjectionMatchOut(C.Structure):
    _fields_ = [("kp_mp", C.POINTER(C.c_int32)), ("nmatches", C.c_int), ("n_rescans", C.c_int)]


class TriangulationMatchOut(C.Structure):
    _fields_ = [("match12", C.POINTER(C.c_int32)), ("nmatches", C.c_int)]


SYMBOLS = ["ppg_default_config", "ppg_create", "ppg_destroy", "ppg_last_error", "ppg_api_version", "ppg_extract",
           "ppg_upload_frames", "ppg_run", "ppg_download", "ppg_sync", "ppg_extract_from_maps", "ppg_get_maps",
           "ppg_selftest_conv", "ppg_get_layer_output", "ppg_set_profiling", "ppg_get_stage_times", "ppg_launch_count", "ppg_timer_start",
           "ppg_timer_stop", "ppg_upload_map", "ppg_associate", "ppg_assoc_stage", "ppg_assoc_run",
           "ppg_assoc_fetch", "ppg_assoc_run_frame", "ppg_assoc_stage_batch", "ppg_assoc_run_batch",
           "ppg_assoc_fetch_batch", "ppg_assoc_fallback_rows", "ppg_assoc_device_results",
           "ppg_distinctive_descriptors", "ppg_upload_map_distinctive", "ppg_stream", "ppg_upload_map_graph",
           "ppg_extend_map_matches", "ppg_extend_run_batch", "ppg_extend_fetch_batch", "ppg_upload_map_geometry",
           "ppg_assoc_stage_poses", "ppg_frustum_fetch", "ppg_upload_vocabulary", "ppg_bow_transform",
           "ppg_bow_run_batch", "ppg_bow_fetch_batch", "ppg_search_by_bow", "ppg_vocabulary_open",
           "ppg_vocabulary_close", "ppg_vocabulary_error", "ppg_load_vocabulary", "ppg_extract_async",
           "ppg_extract_wait", "ppg_host_alloc", "ppg_host_free", "ppg_host_register", "ppg_host_unregister",
           "ppg_assoc_stage_batch_async", "ppg_extend_fetch_batch_async", "ppg_extend_collect",
           "ppg_comm_unique_id", "ppg_comm_init", "ppg_comm_destroy", "ppg_assoc_allgather",
           "ppg_assoc_allgather_fetch", "ppg_record_bytes",
           "ppg_search_for_initialization", "ppg_search_for_triangulation", "ppg_search_by_projection"]

_lib = None


def load():
    """Loads libppg_b200.so (raises OSError if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ppg_last_error.restype = C.c_char_p
        lib.ppg_last_error.argtypes = [C.c_void_p]
        lib.ppg_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        lib.ppg_destroy.argtypes = [C.c_void_p]
        lib.ppg_destroy.restype = None
        lib.ppg_launch_count.restype = C.c_longlong
        lib.ppg_launch_count.argtypes = [C.c_void_p]
        lib.ppg_vocabulary_error.restype = C.c_char_p
        lib.ppg_vocabulary_close.restype = None
        lib.ppg_vocabulary_close.argtypes = [C.c_void_p]
        lib.ppg_stream.restype = C.c_void_p
        lib.ppg_stream.argtypes = [C.c_void_p]
        for name in ["ppg_extract", "ppg_upload_frames", "ppg_run", "ppg_download", "ppg_sync",
                     "ppg_extract_from_maps", "ppg_get_maps", "ppg_selftest_conv", "ppg_set_profiling",
                     "ppg_get_stage_times", "ppg_timer_start", "ppg_timer_stop", "ppg_upload_map", "ppg_associate",
                     "ppg_assoc_stage", "ppg_assoc_run", "ppg_assoc_fetch", "ppg_assoc_run_frame", "ppg_assoc_stage_batch", "ppg_assoc_run_batch",
           "ppg_assoc_fetch_batch",
                     "ppg_assoc_fallback_rows", "ppg_assoc_device_results", "ppg_assoc_stage_batch",
                     "ppg_assoc_run_batch", "ppg_assoc_fetch_batch", "ppg_distinctive_descriptors",
                     "ppg_upload_map_distinctive", "ppg_upload_map_graph", "ppg_extend_map_matches",
                     "ppg_extend_run_batch", "ppg_extend_fetch_batch", "ppg_upload_map_geometry",
                     "ppg_assoc_stage_poses", "ppg_frustum_fetch", "ppg_upload_vocabulary", "ppg_bow_transform",
                     "ppg_bow_run_batch", "ppg_bow_fetch_batch", "ppg_search_by_bow", "ppg_vocabulary_open",
                     "ppg_load_vocabulary", "ppg_extract_async", "ppg_extract_wait", "ppg_host_register",
                     "ppg_host_unregister", "ppg_assoc_stage_batch_async", "ppg_extend_fetch_batch_async",
                     "ppg_extend_collect", "ppg_comm_unique_id", "ppg_comm_init", "ppg_comm_destroy",
                     "ppg_assoc_allgather", "ppg_assoc_allgather_fetch", "ppg_search_for_initialization",
                     "ppg_search_for_triangulation", "ppg_search_by_projection"]:
            getattr(lib, name).restype = C.c_int
        lib.ppg_get_layer_output.restype = C.c_int
        lib.ppg_get_layer_output.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t,
                                             C.POINTER(C.c_size_t)]
        lib.ppg_record_bytes.restype = C.c_longlong
        lib.ppg_record_bytes.argtypes = [C.c_void_p]
        lib.ppg_host_alloc.restype = C.c_void_p
        lib.ppg_host_alloc.argtypes = [C.c_size_t]
        lib.ppg_host_free.restype = None
        lib.ppg_host_free.argtypes = [C.c_void_p]
        lib.ppg_host_register.argtypes = [C.c_void_p, C.c_size_t]
        lib.ppg_host_unregister.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


class _Pinned:
    """Owner of one ppg_host_alloc block; the numpy views keep it alive."""

    def __init__(self, nbytes):
        self.lib = load()
        self.ptr = self.lib.ppg_host_alloc(max(int(nbytes), 1))
        if not self.ptr:
            raise MemoryError("ppg_host_alloc(%d) failed" % nbytes)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.ppg_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_array(shape, dtype):
    """numpy array in page-locked host memory (cudaHostAlloc through the C ABI): the source / destination of the
    asynchronous copies of the pipelined calls."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    blk = _Pinned(n)
    buf = (C.c_uint8 * max(n, 1)).from_address(blk.ptr)
    a = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    _PINNED_KEEP[id(buf)] = blk  # the block must outlive every view; freed at interpreter exit or by drop_pinned()
    return a


_PINNED_KEEP = {}


def drop_pinned():
    _PINNED_KEEP.clear()


def comm_unique_id():
    """128 bytes for ppg_comm_init, generated on one rank and handed to the others by the host program."""
    buf = (C.c_uint8 * 128)()
    rc = load().ppg_comm_unique_id(buf)
    if rc != PPG_OK:
        raise PpgError(rc, "ppg_comm_unique_id failed (libnccl.so.2 missing?)")
    return bytes(buf)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def read_vocabulary(path):
    """The library's own host-side reader (ppg_vocabulary_open: DBoW3 binary + QuickLZ, or the PVOC blob).
    -> ppg_slam_b200.vocabulary.PodVocabulary.  No GPU involved."""
    from . import vocabulary
    lib = load()
    h, v = C.c_void_p(), VocabularyPod()
    rc = lib.ppg_vocabulary_open(os.fsencode(path), C.byref(h), C.byref(v))
    if rc != PPG_OK:
        raise PpgError(rc, lib.ppg_vocabulary_error().decode())
    try:
        n, k, dim = v.n_nodes, v.k, v.dim
        return vocabulary.PodVocabulary(
            k, v.L, v.scoring, v.weighting, np.ctypeslib.as_array(v.children, shape=(n, k)).copy(),
            np.ctypeslib.as_array(v.word_id, shape=(n,)).copy(), np.ctypeslib.as_array(v.weight, shape=(n,)).copy(),
            np.ctypeslib.as_array(v.desc, shape=(n, dim)).copy())
    finally:
        lib.ppg_vocabulary_close(h)


def _frame_to_dict(o):
    """Copies one ppg_frame_out record out of the ctx-owned pinned memory into numpy arrays."""
    n, E, nc = o.n_kp, o.n_edges, o.n_colines

    def arr(ptr, cnt, dt):
        if cnt <= 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ptr, shape=(cnt,)).astype(dt, copy=True)

    conn_off = arr(o.conn_off, n + 1, np.int32) if n > 0 else np.zeros(1, np.int32)
    col_off = arr(o.col_off, n + 1, np.int32) if n > 0 else np.zeros(1, np.int32)
    return dict(
        n_kp=n, n_edges=E, n_colines=nc, status=int(o.status), n_cand=o.n_candidates,
        n_pairs_ok=o.n_pairs_tested_ok, n_candidate_lines=o.n_candidate_lines, nms_rounds=o.nms_rounds,
        kp_x=arr(o.kp_x, n, np.float32), kp_y=arr(o.kp_y, n, np.float32),
        px=arr(o.kp_px, n, np.int32), py=arr(o.kp_py, n, np.int32), score=arr(o.kp_score, n, np.float32),
        xun=arr(o.kp_xun, n, np.float32), yun=arr(o.kp_yun, n, np.float32), out=arr(o.kp_out, n, np.uint8),
        edge_start=arr(o.edge_start, E, np.int32), edge_end=arr(o.edge_end, E, np.int32),
        edge_score=arr(o.edge_score, E, np.float32), conn_off=conn_off,
        conn_idx=arr(o.conn_idx, int(conn_off[-1]), np.int32), col_off=col_off,
        col_pairs=arr(o.col_pairs, 2 * nc, np.int32).reshape(-1, 2),
        desc=arr(o.desc, n * 256, np.float32).reshape(-1, 256), diag=[int(v) for v in o.diag])


class Extractor:
    """Host-side mirror of the reference's PPGExtractor (feature/include/PPGExtractor.h:34-148) over
    the C ABI: constructed from a camera + weight path, `run()` takes 8-bit gray frames and returns the
    keypoints / edges / descriptors record of PPGExtractor::run (PPGExtractor.cpp:118-147)."""

    def __init__(self, cam, weights=DEFAULT_WEIGHTS, device=0, max_batch=1, **over):
        lib = load()
        cfg = Config()
        lib.ppg_default_config(C.byref(cfg))
        cfg.device, cfg.width, cfg.height, cfg.fisheye = device, cam.width, cam.height, int(cam.fisheye)
        cfg.K[:] = cam.K
        cfg.D[:] = cam.D
        self._wp = os.fsencode(weights)
        cfg.weights_path = self._wp
        cfg.max_batch = max_batch
        for k, v in over.items():
            setattr(cfg, k, v)
        self.cfg, self.cam, self.lib = cfg, cam, lib
        self.W, self.H, self.max_batch = cam.width, cam.height, max_batch
        h = C.c_void_p()
        rc = lib.ppg_create(C.byref(cfg), C.byref(h))
        if rc != PPG_OK:
            raise PpgError(rc, lib.ppg_last_error(None).decode())
        self.h = h
        self._outs = (FrameOut * max_batch)()

    def close(self):
        if getattr(self, "h", None):
            self.lib.ppg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow_capacity=False):
        if rc == PPG_OK or (allow_capacity and rc == PPG_ERR_CAPACITY):
            return rc
        raise PpgError(rc, self.lib.ppg_last_error(self.h).decode())

    # ---- PPGExtractor::run
    def _frame_ptrs(self, frames):
        # rows may be padded (cv::Mat::step > cols): such views are passed as they are, with their row stride
        frames = [f if (isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.ndim == 2 and f.strides[1] == 1 and
                        f.strides[0] >= f.shape[1]) else np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
        for f in frames:
            if f.shape != (self.H, self.W):
                raise ValueError("frame shape %r != (%d, %d)" % (f.shape, self.H, self.W))
        n = len(frames)
        ptrs = (C.POINTER(C.c_uint8) * n)(*[f.ctypes.data_as(C.POINTER(C.c_uint8)) for f in frames])
        strides = (C.c_int * n)(*[f.strides[0] for f in frames])
        return frames, ptrs, strides, n

    def run(self, frames, allow_capacity=False):
        """frames: list of (H,W) uint8 arrays (<= max_batch).  -> list of per-frame dicts."""
        keep, ptrs, strides, n = self._frame_ptrs(frames)
        self._check(self.lib.ppg_extract(self.h, ptrs, strides, n, self._outs), allow_capacity)
        return [_frame_to_dict(self._outs[i]) for i in range(n)]

    # ---- pipelined form: enqueue, do something else (another ctx), wait
    def extract_async(self, frames):
        """frames should be views of pinned memory (capi.pinned_array) and must stay alive until extract_wait."""
        keep, ptrs, strides, n = self._frame_ptrs(frames)
        self._async_keep = (keep, ptrs, strides)
        self._check(self.lib.ppg_extract_async(self.h, ptrs, strides, n))
        return n

    def extract_wait(self, n, allow_capacity=False, as_dicts=True):
        self._check(self.lib.ppg_extract_wait(self.h, n, self._outs), allow_capacity)
        return [_frame_to_dict(self._outs[i]) for i in range(n)] if as_dicts else None

    def assoc_stage_batch_async(self, proj_uv, view_cos, th, ratio):
        """proj_uv (F, M, 2) / view_cos (F, M) float32 C-contiguous arrays in pinned memory (capi.pinned_array)."""
        F, M = view_cos.shape
        self._check(self.lib.ppg_assoc_stage_batch_async(self.h, F, M, _fp(proj_uv), _fp(view_cos), C.c_float(th),
                                                         C.c_float(ratio)))
        self._assoc_rows, self._assoc_frames = M, F
        self._batch_out = None

    def extend_fetch_batch_async(self, n_frames):
        self._check(self.lib.ppg_extend_fetch_batch_async(self.h, n_frames))

    def extend_collect(self, n_frames, as_dicts=True):
        """After extract_wait / sync: the results of the last extend_fetch_batch_async (no CUDA call)."""
        outs, res = self._extend_outs(n_frames)
        self._check(self.lib.ppg_extend_collect(self.h, n_frames, outs))
        if not as_dicts:
            return outs
        return [self._extend_result(outs[f], res[f], self._graph_points) for f in range(n_frames)]

    def record_bytes(self):
        return int(self.lib.ppg_record_bytes(self.h))

    def upload(self, frames):
        keep, ptrs, strides, n = self._frame_ptrs(frames)
        self._check(self.lib.ppg_upload_frames(self.h, ptrs, strides, n))
        self._check(self.lib.ppg_sync(self.h))
        return n

    def run_device(self, n):
        self._check(self.lib.ppg_run(self.h, n))

    def sync(self):
        self._check(self.lib.ppg_sync(self.h))

    def download(self, n, allow_capacity=False, as_dicts=True):
        self._check(self.lib.ppg_download(self.h, n, self._outs), allow_capacity)
        return [_frame_to_dict(self._outs[i]) for i in range(n)] if as_dicts else None

    def run_from_maps(self, prob, heat, desc_chw, allow_capacity=False):
        """Post-processing only, fed with reference dense maps (parity entry point)."""
        prob = np.ascontiguousarray(prob, np.float32).reshape(-1, self.H, self.W)
        heat = np.ascontiguousarray(heat, np.float32).reshape(-1, self.H, self.W)
        n = prob.shape[0]
        desc = np.ascontiguousarray(desc_chw, np.float32).reshape(n, 256, self.H // 8, self.W // 8)
        self._check(self.lib.ppg_extract_from_maps(self.h, _fp(prob), _fp(heat), _fp(desc), n, self._outs),
                    allow_capacity)
        return [_frame_to_dict(self._outs[i]) for i in range(n)]

    def get_maps(self, frame=0, feature=False):
        H, W = self.H, self.W
        prob, hr, hf = (np.empty((H, W), np.float32) for _ in range(3))
        desc = np.empty((256, H // 8, W // 8), np.float32)
        feat = np.empty((128, H // 8, W // 8), np.float32) if feature else None
        self._check(self.lib.ppg_get_maps(self.h, frame, _fp(prob), _fp(hr), _fp(hf), _fp(desc),
                                          _fp(feat) if feature else None))
        r = dict(prob=prob, heat=hr, heat_final=hf, desc=desc)
        if feature:
            r["feature"] = feat
        return r

    def layer_output(self, name, frame=0):
        """Raw output tensor of a tensor-core layer as bytes (validation aid)."""
        n = C.c_size_t(0)
        self._check(self.lib.ppg_get_layer_output(self.h, name.encode(), frame, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._check(self.lib.ppg_get_layer_output(self.h, name.encode(), frame, buf.ctypes.data_as(C.c_void_p), n.value,
                                                  C.byref(n)))
        return buf

    def selftest_conv(self):
        names = (C.c_char_p * 32)()
        d, r = (C.c_float * 32)(), (C.c_float * 32)()
        n = C.c_int(0)
        self._check(self.lib.ppg_selftest_conv(self.h, 32, names, d, r, C.byref(n)))
        return [(names[i].decode(), float(d[i]), float(r[i])) for i in range(n.value)]

    def set_profiling(self, on=True):
        self._check(self.lib.ppg_set_profiling(self.h, int(on)))

    def stage_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = C.c_int(0)
        self._check(self.lib.ppg_get_stage_times(self.h, 64, names, ms, C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def launch_count(self):
        return int(self.lib.ppg_launch_count(self.h))

    def timer_start(self):
        self._check(self.lib.ppg_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.lib.ppg_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    # ---- association (search core of Matcher::ExtendMapMatches, matching/src/Matcher.cpp:224-281)
    def upload_map(self, map_desc):
        m = np.ascontiguousarray(map_desc, np.float32)
        self._check(self.lib.ppg_upload_map(self.h, _fp(m), m.shape[0]))
        self._n_rows = m.shape[0]

    def _assoc_in(self, kp_x, kp_y, frame_desc, free_mask, proj_uv, view_cos, th, ratio, mode=0, max_dist=0.0,
                  e2_max=0.0, row_node=None, kp_node=None):
        a = AssocIn()
        keep = [np.ascontiguousarray(kp_x, np.float32), np.ascontiguousarray(kp_y, np.float32),
                np.ascontiguousarray(frame_desc, np.float32), np.ascontiguousarray(free_mask, np.uint8),
                np.ascontiguousarray(proj_uv, np.float32), np.ascontiguousarray(view_cos, np.float32)]
        a.n_kp = len(keep[0])
        a.kp_x, a.kp_y, a.frame_desc = _fp(keep[0]), _fp(keep[1]), _fp(keep[2])
        a.free_mask = keep[3].ctypes.data_as(C.POINTER(C.c_uint8))
        a.n_rows = len(keep[5])
        a.proj_uv, a.view_cos = _fp(keep[4]), _fp(keep[5])
        a.th, a.ratio = th, ratio
        a.mode, a.max_dist, a.e2_max = mode, max_dist, e2_max
        if row_node is not None:
            keep += [np.ascontiguousarray(row_node, np.int32), np.ascontiguousarray(kp_node, np.int32)]
            a.row_node = keep[-2].ctypes.data_as(C.POINTER(C.c_int32))
            a.kp_node = keep[-1].ctypes.data_as(C.POINTER(C.c_int32))
        return a, keep

    @staticmethod
    def _assoc_out(m):
        r = dict(best_idx=np.zeros(m, np.int32), second_idx=np.zeros(m, np.int32),
                 best_d=np.zeros(m, np.float32), second_d=np.zeros(m, np.float32), accept=np.zeros(m, np.uint8))
        o = AssocOut()
        o.best_idx = r["best_idx"].ctypes.data_as(C.POINTER(C.c_int32))
        o.second_idx = r["second_idx"].ctypes.data_as(C.POINTER(C.c_int32))
        o.best_dist, o.second_dist = _fp(r["best_d"]), _fp(r["second_d"])
        o.accept = r["accept"].ctypes.data_as(C.POINTER(C.c_uint8))
        return o, r

    def associate(self, kp_x, kp_y, frame_desc, free_mask, proj_uv, view_cos, th, ratio, mode=0, max_dist=0.0,
                  e2_max=0.0, row_node=None, kp_node=None):
        """mode 0: search core of ExtendMapMatches; mode 1 (SEARCH_WINDOW): best-only cores of SearchByProjection /
        Fuse -- r = th, accept = best <= max_dist, optional circular limit e2_max (include/ppg_b200.h)."""
        a, keep = self._assoc_in(kp_x, kp_y, frame_desc, free_mask, proj_uv, view_cos, th, ratio, mode, max_dist,
                                 e2_max, row_node, kp_node)
        o, r = self._assoc_out(a.n_rows)
        self._check(self.lib.ppg_associate(self.h, C.byref(a), C.byref(o)))
        return r

    def assoc_stage(self, kp_x, kp_y, frame_desc, free_mask, proj_uv, view_cos, th, ratio):
        a, keep = self._assoc_in(kp_x, kp_y, frame_desc, free_mask, proj_uv, view_cos, th, ratio)
        self._check(self.lib.ppg_assoc_stage(self.h, C.byref(a)))
        self._assoc_rows = a.n_rows

    def assoc_run(self):
        self._check(self.lib.ppg_assoc_run(self.h))

    def assoc_run_frame(self, frame):
        self._check(self.lib.ppg_assoc_run_frame(self.h, frame))

    def assoc_stage_batch(self, proj_uv, view_cos, th, ratio):
        """proj_uv (F, M, 2), view_cos (F, M): per-frame projections of the resident map points."""
        uv = np.ascontiguousarray(proj_uv, np.float32)
        vc = np.ascontiguousarray(view_cos, np.float32)
        F, M = vc.shape
        self._check(self.lib.ppg_assoc_stage_batch(self.h, F, M, _fp(uv), _fp(vc), C.c_float(th), C.c_float(ratio)))
        self._assoc_rows, self._assoc_frames = M, F
        self._batch_out = None

    def assoc_run_batch(self, n_frames):
        self._check(self.lib.ppg_assoc_run_batch(self.h, n_frames))

    def assoc_fetch_batch(self, n_frames):
        if getattr(self, "_batch_out", None) is None or len(self._batch_out[1]) != n_frames:
            outs = (AssocOut * n_frames)()
            res = []
            for f in range(n_frames):
                o, r = self._assoc_out(self._assoc_rows)
                outs[f] = o
                res.append(r)
            self._batch_out = (outs, res)
        outs, res = self._batch_out
        self._check(self.lib.ppg_assoc_fetch_batch(self.h, n_frames, outs))
        return res

    def assoc_fetch(self):
        o, r = self._assoc_out(self._assoc_rows)
        self._check(self.lib.ppg_assoc_fetch(self.h, C.byref(o)))
        return r

    # ---- the whole Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381): search + sequential walk + seed growing
    def upload_map_graph(self, candidate, observed, bad, edge_off, edge_other, edge_ok):
        """POD form of the map-point graph for the rows of the resident table (include/ppg_b200.h, ppg_map_graph)."""
        u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        keep = [np.ascontiguousarray(candidate, np.uint8), np.ascontiguousarray(observed, np.uint8),
                np.ascontiguousarray(bad, np.uint8), np.ascontiguousarray(edge_off, np.int32),
                np.ascontiguousarray(edge_other, np.int32), np.ascontiguousarray(edge_ok, np.uint8)]
        g = MapGraph()
        g.n_points = len(keep[0])
        g.candidate, g.observed, g.bad = (k.ctypes.data_as(u8p) for k in keep[:3])
        g.edge_off = keep[3].ctypes.data_as(i32p)
        g.edge_other = keep[4].ctypes.data_as(i32p) if keep[4].size else None
        g.edge_ok = keep[5].ctypes.data_as(u8p) if keep[5].size else None
        self._check(self.lib.ppg_upload_map_graph(self.h, C.byref(g)))
        self._graph_points = g.n_points

    @staticmethod
    def _extend_out(n_kp, n_edges, n_points):
        r = dict(kp_mp=np.zeros(max(n_kp, 1), np.int32), kedge_me=np.zeros(max(n_edges, 1), np.int32),
                 tracked=np.zeros(max(n_points, 1), np.uint8))
        o = ExtendOut()
        o.kp_mp = r["kp_mp"].ctypes.data_as(C.POINTER(C.c_int32))
        o.kedge_me = r["kedge_me"].ctypes.data_as(C.POINTER(C.c_int32))
        o.tracked = r["tracked"].ctypes.data_as(C.POINTER(C.c_uint8))
        return o, r

    @staticmethod
    def _extend_result(o, r, n_points):
        return dict(nmatches=o.nmatches, status=int(o.status), kp_mp=r["kp_mp"][:o.n_kp].copy(),
                    kedge_me=r["kedge_me"][:o.n_edges].copy(), tracked=r["tracked"][:n_points].copy(),
                    n_accepted=o.n_accepted, n_grown=o.n_grown, n_rescans=o.n_rescans,
                    diag=[int(v) for v in o.diag])

    def extend_map_matches(self, kp_x, kp_y, frame_desc, kp_mp, edge_start, edge_end, conn_off, conn_idx, proj_uv,
                           view_cos, tracked, th, ratio, kedge_me=None):
        """Matcher::ExtendMapMatches for one frame given in host memory (ppg_extend_in).  -> dict(nmatches, kp_mp,
        kedge_me, tracked, ...)."""
        i32p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        f32a = lambda a: np.ascontiguousarray(a, np.float32)
        i32a = lambda a: np.ascontiguousarray(a, np.int32)
        kx, ky, fd = f32a(kp_x), f32a(kp_y), f32a(frame_desc)
        uv, vc = (f32a(proj_uv), f32a(view_cos)) if proj_uv is not None else (None, None)
        es, ee, coff, cidx = i32a(edge_start), i32a(edge_end), i32a(conn_off), i32a(conn_idx)
        a = ExtendIn()
        a.n_kp, a.n_edges = len(kx), len(es)
        a.kp_x, a.kp_y, a.frame_desc = _fp(kx), _fp(ky), _fp(fd)
        keep = [kx, ky, fd, uv, vc, es, ee, coff, cidx]
        if kp_mp is not None:
            k = i32a(kp_mp)
            keep.append(k)
            a.kp_mp = k.ctypes.data_as(i32p)
        a.edge_start, a.edge_end = es.ctypes.data_as(i32p), ee.ctypes.data_as(i32p)
        a.conn_off, a.conn_idx = coff.ctypes.data_as(i32p), cidx.ctypes.data_as(i32p)
        if kedge_me is not None:
            k = i32a(kedge_me)
            keep.append(k)
            a.kedge_me = k.ctypes.data_as(i32p)
        if uv is not None:  # else: projections staged on the device by assoc_stage_poses
            a.proj_uv, a.view_cos = _fp(uv), _fp(vc)
        if tracked is not None:
            t = np.ascontiguousarray(tracked, np.uint8)
            keep.append(t)
            a.tracked = t.ctypes.data_as(u8p)
        a.th, a.ratio = th, ratio
        o, r = self._extend_out(a.n_kp, a.n_edges, self._graph_points)
        self._check(self.lib.ppg_extend_map_matches(self.h, C.byref(a), C.byref(o)))
        return self._extend_result(o, r, self._graph_points)

    def extend_run_batch(self, n_frames):
        """Every frame of the last extraction batch against the projections staged with assoc_stage_batch."""
        self._check(self.lib.ppg_extend_run_batch(self.h, n_frames))

    def _extend_outs(self, n_frames):
        if getattr(self, "_xbatch_out", None) is None or len(self._xbatch_out[1]) != n_frames:
            outs = (ExtendOut * n_frames)()
            res = []
            for f in range(n_frames):
                o, r = self._extend_out(self.cfg.junction_max_num, self.cfg.max_edges, self._graph_points)
                outs[f] = o
                res.append(r)
            self._xbatch_out = (outs, res)
        return self._xbatch_out

    def extend_fetch_batch(self, n_frames):
        outs, res = self._extend_outs(n_frames)
        self._check(self.lib.ppg_extend_fetch_batch(self.h, n_frames, outs))
        return [self._extend_result(outs[f], res[f], self._graph_points) for f in range(n_frames)]

    # ---- Frame::CheckInFrustum on the device (map/src/Frame.cpp:223-260)
    def upload_map_geometry(self, world_pos, normal, min_dist, max_dist):
        a = [np.ascontiguousarray(x, np.float32) for x in (world_pos, normal, min_dist, max_dist)]
        self._check(self.lib.ppg_upload_map_geometry(self.h, _fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(a[3]), len(a[2])))

    def assoc_stage_poses(self, Rcw, tcw, Ow, n_rows, cos_limit, th, ratio):
        """Rcw (F,3,3), tcw (F,3), Ow (F,3): projects the resident map points into every frame on the device and
        stages the result for assoc_run_batch / extend_run_batch."""
        R, t, o = (np.ascontiguousarray(x, np.float32) for x in (Rcw, tcw, Ow))
        F = t.reshape(-1, 3).shape[0]
        self._check(self.lib.ppg_assoc_stage_poses(self.h, F, n_rows, _fp(R), _fp(t), _fp(o), C.c_float(cos_limit),
                                                   C.c_float(th), C.c_float(ratio)))
        self._assoc_rows, self._assoc_frames = n_rows, F
        self._batch_out = None

    def frustum_fetch(self, n_frames):
        M = self._assoc_rows
        iv, uv = np.zeros((n_frames, M), np.uint8), np.zeros((n_frames, M, 2), np.float32)
        dp, vc = np.zeros((n_frames, M), np.float32), np.zeros((n_frames, M), np.float32)
        self._check(self.lib.ppg_frustum_fetch(self.h, n_frames, iv.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(uv),
                                               _fp(dp), _fp(vc)))
        return dict(in_view=iv, proj_uv=uv, depth=dp, view_cos=vc)

    # ---- bag of words: DBoW3::Vocabulary::transform (Frame::ComputeBoW, map/src/Frame.cpp:331-340)
    def load_vocabulary(self, path):
        """ppg_load_vocabulary: the reference's Vocabulary/voc_*.gz (or the exported blob) straight into the ctx."""
        self._check(self.lib.ppg_load_vocabulary(self.h, os.fsencode(path)))

    def upload_vocabulary(self, voc):
        """voc: ppg_slam_b200.vocabulary.PodVocabulary / Vocabulary (k, L, scoring, weighting, child_table(), ...)."""
        ch = np.ascontiguousarray(voc.child_table(), np.int32)
        wid = np.ascontiguousarray(voc.word_id, np.int32)
        w = np.ascontiguousarray(voc.weight, np.float64)
        d = np.ascontiguousarray(voc.desc, np.float32)
        v = VocabularyPod()
        v.k, v.L, v.scoring, v.weighting, v.n_nodes, v.dim = voc.k, voc.L, voc.scoring, voc.weighting, ch.shape[0], d.shape[1]
        v.children = ch.ctypes.data_as(C.POINTER(C.c_int32))
        v.word_id = wid.ctypes.data_as(C.POINTER(C.c_int32))
        v.weight = w.ctypes.data_as(C.POINTER(C.c_double))
        v.desc = _fp(d)
        self._check(self.lib.ppg_upload_vocabulary(self.h, C.byref(v)))

    @staticmethod
    def _bow_out(cap):
        r = dict(word=np.zeros(cap, np.int32), weight=np.zeros(cap, np.float64), node=np.zeros(cap, np.int32),
                 bow_word=np.zeros(cap, np.int32), bow_value=np.zeros(cap, np.float64))
        o = BowOut()
        o.word_id = r["word"].ctypes.data_as(C.POINTER(C.c_int32))
        o.word_weight = r["weight"].ctypes.data_as(C.POINTER(C.c_double))
        o.node_id = r["node"].ctypes.data_as(C.POINTER(C.c_int32))
        o.bow_word = r["bow_word"].ctypes.data_as(C.POINTER(C.c_int32))
        o.bow_value = r["bow_value"].ctypes.data_as(C.POINTER(C.c_double))
        return o, r

    @staticmethod
    def _bow_result(o, r):
        n, nb = o.n_features, o.n_bow
        return dict(word=r["word"][:n].copy(), weight=r["weight"][:n].copy(), node=r["node"][:n].copy(),
                    bow_word=r["bow_word"][:nb].copy(), bow_value=r["bow_value"][:nb].copy())

    def bow_transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.float32).reshape(-1, 256)
        o, r = self._bow_out(max(len(d), 1))
        self._check(self.lib.ppg_bow_transform(self.h, _fp(d), len(d), levelsup, C.byref(o)))
        return self._bow_result(o, r)

    def bow_run_batch(self, n_frames, levelsup=4):
        self._check(self.lib.ppg_bow_run_batch(self.h, n_frames, levelsup))

    def bow_fetch_batch(self, n_frames):
        outs = (BowOut * n_frames)()
        res = []
        for f in range(n_frames):
            o, r = self._bow_out(1024)
            outs[f] = o
            res.append(r)
        self._check(self.lib.ppg_bow_fetch_batch(self.h, n_frames, outs))
        return [self._bow_result(outs[f], res[f]) for f in range(n_frames)]

    def search_by_bow(self, row_node, frame_desc, kp_node, ratio, max_dist, strict=False):
        """Matcher::SearchByBoW whole (rows = the table uploaded with upload_map, in visiting order).
        -> dict(kp_row, nmatches, n_rescans)."""
        rn = np.ascontiguousarray(row_node, np.int32)
        fd = np.ascontiguousarray(frame_desc, np.float32).reshape(-1, 256)
        kn = np.ascontiguousarray(kp_node, np.int32)
        a = BowMatchIn()
        a.n_rows, a.n_kp = len(rn), len(kn)
        a.row_node = rn.ctypes.data_as(C.POINTER(C.c_int32))
        a.frame_desc, a.kp_node = _fp(fd), kn.ctypes.data_as(C.POINTER(C.c_int32))
        a.ratio, a.max_dist, a.strict = ratio, max_dist, int(strict)
        kr = np.full(max(len(kn), 1), -1, np.int32)
        o = BowMatchOut()
        o.kp_row = kr.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.ppg_search_by_bow(self.h, C.byref(a), C.byref(o)))
        return dict(kp_row=kr[:len(kn)].copy(), nmatches=o.nmatches, n_rescans=o.n_rescans)

    def search_for_initialization(self, desc1, prev_matched, kx2, ky2, desc2, window=50, ratio=0.9):
        """Matcher::SearchForInitialization (Matcher.cpp:582-651); F1's descriptors are uploaded as the table here.
        -> dict(nmatches, matches12, prev_matched, n_rescans)"""
        d1 = np.ascontiguousarray(desc1, np.float32)
        self.upload_map(d1)
        pm = np.ascontiguousarray(prev_matched, np.float32).copy()
        kx, ky, d2 = (np.ascontiguousarray(a, np.float32) for a in (kx2, ky2, desc2))
        a = InitMatchIn()
        a.n1, a.n2, a.window, a.ratio = len(d1), len(kx), int(window), ratio
        a.prev_matched, a.kp2_x, a.kp2_y, a.desc2 = _fp(pm), _fp(kx), _fp(ky), _fp(d2)
        m12 = np.full(max(len(d1), 1), -1, np.int32)
        o = InitMatchOut()
        o.matches12 = m12.ctypes.data_as(C.POINTER(C.c_int32))
        o.prev_matched = _fp(pm)
        self._check(self.lib.ppg_search_for_initialization(self.h, C.byref(a), C.byref(o)))
        return dict(nmatches=o.nmatches, matches12=m12[:len(d1)], prev_matched=pm, n_rescans=o.n_rescans)

    def search_by_projection(self, map_desc, proj_uv, observed, kp_x, kp_y, desc, kp_mp, th, max_dist):
        """Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (Matcher.cpp:31-87) / (CurrentFrame, pKF, sAlreadyFound,
        th, descDist) (:1337-1411) whole: rows = the projected map points in loop order (map_desc, proj_uv, observed or
        None), kp_mp = CurrentFrame.mvpMapPoints as rows (-1 / -2, or None).  -> dict(nmatches, kp_mp, n_rescans)"""
        f32 = lambda v: np.ascontiguousarray(v, np.float32)
        md, uv, kx, ky, d = f32(map_desc), f32(proj_uv), f32(kp_x), f32(kp_y), f32(desc)
        m, n = len(md), len(kx)
        if m > 0:
            self.upload_map(md)
        a = ProjectionMatchIn()
        a.n_rows, a.n, a.th, a.max_dist = m, n, th, max_dist
        a.proj_uv, a.kp_x, a.kp_y, a.desc = _fp(uv), _fp(kx), _fp(ky), _fp(d)
        keep = []
        if observed is not None:
            ob = np.ascontiguousarray(observed, np.uint8)
            keep.append(ob)
            a.observed = ob.ctypes.data_as(C.POINTER(C.c_uint8))
        if kp_mp is not None:
            km = np.ascontiguousarray(kp_mp, np.int32)
            keep.append(km)
            a.kp_mp = km.ctypes.data_as(C.POINTER(C.c_int32))
        res = np.full(max(n, 1), -1, np.int32)
        o = ProjectionMatchOut()
        o.kp_mp = res.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.ppg_search_by_projection(self.h, C.byref(a), C.byref(o)))
        return dict(nmatches=o.nmatches, kp_mp=res[:n], n_rescans=o.n_rescans)

    def search_for_triangulation(self, desc1, node1, has_mp1, pos1, desc2, node2, has_mp2, pos2, F12, epipole,
                                 th_low=0.7, kb8=None):
        """Matcher::SearchForTriangulation (Matcher.cpp:767-885).  kb8 = None: pinhole epipolar test from F12;
        kb8 = (cam8, R12, t12): KannalaBrandt8::epipolarConstrain (F12 unused).  -> dict(nmatches, match12)"""
        f32 = lambda v: np.ascontiguousarray(v, np.float32)
        d1, d2, p1, p2 = f32(desc1), f32(desc2), f32(pos1), f32(pos2)
        nd1, nd2 = np.ascontiguousarray(node1, np.int32), np.ascontiguousarray(node2, np.int32)
        m1, m2 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(has_mp2, np.uint8)
        a = TriangulationMatchIn()
        a.n1, a.n2, a.th_low = len(d1), len(d2), th_low
        a.desc1, a.desc2, a.pos1, a.pos2 = _fp(d1), _fp(d2), _fp(p1), _fp(p2)
        a.node1 = nd1.ctypes.data_as(C.POINTER(C.c_int32))
        a.node2 = nd2.ctypes.data_as(C.POINTER(C.c_int32))
        a.has_mp1 = m1.ctypes.data_as(C.POINTER(C.c_uint8))
        a.has_mp2 = m2.ctypes.data_as(C.POINTER(C.c_uint8))
        a.F12 = (C.c_float * 9)(*[float(v) for v in f32(F12).reshape(9)])
        a.epipole = (C.c_float * 2)(*[float(v) for v in f32(epipole).reshape(2)])
        if kb8 is not None:
            a.camera_model = 1
            a.cam8 = (C.c_float * 8)(*[float(v) for v in f32(kb8[0]).reshape(8)])
            a.R12 = (C.c_float * 9)(*[float(v) for v in f32(kb8[1]).reshape(9)])
            a.t12 = (C.c_float * 3)(*[float(v) for v in f32(kb8[2]).reshape(3)])
        m12 = np.full(max(len(d1), 1), -1, np.int32)
        o = TriangulationMatchOut()
        o.match12 = m12.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.ppg_search_for_triangulation(self.h, C.byref(a), C.byref(o)))
        return dict(nmatches=o.nmatches, match12=m12[:len(d1)])

    def distinctive_descriptors(self, desc, offsets, to_table=False):
        """MapPoint::ComputeDistinctiveDescriptors for a batch of map points (packed observation descriptors +
        offsets) -> BestIdx per point; to_table=True also makes the chosen rows the resident association table."""
        d = np.ascontiguousarray(desc, np.float32)
        off = np.ascontiguousarray(offsets, np.int32)
        out = np.zeros(len(off) - 1, np.int32)
        fn = self.lib.ppg_upload_map_distinctive if to_table else self.lib.ppg_distinctive_descriptors
        self._check(fn(self.h, _fp(d), off.ctypes.data_as(C.POINTER(C.c_int32)), len(off) - 1,
                       out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    # ---- row-sharded association: NCCL all-gather of the per-row records on the ctx stream (comm.cu)
    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.ppg_comm_init(self.h, buf, rank, world))
        self._comm_world = world

    def comm_destroy(self):
        self._check(self.lib.ppg_comm_destroy(self.h))

    def assoc_allgather(self, n_local, rows_per_rank):
        self._check(self.lib.ppg_assoc_allgather(self.h, n_local, rows_per_rank))
        self._gather_rows = rows_per_rank

    def assoc_allgather_fetch(self, records=True):
        """-> ((world, rows_per_rank, 5) int32 or None, device microseconds of the all-gather)."""
        us = C.c_float(0)
        rec = np.zeros((self._comm_world, self._gather_rows, 5), np.int32) if records else None
        self._check(self.lib.ppg_assoc_allgather_fetch(
            self.h, rec.ctypes.data_as(C.POINTER(C.c_int32)) if records else None, C.byref(us)))
        return rec, us.value

    def assoc_fallback_rows(self):
        n = C.c_int(0)
        self._check(self.lib.ppg_assoc_fallback_rows(self.h, C.byref(n)))
        return n.value

    def assoc_device_results(self):
        p = [C.c_void_p() for _ in range(5)]
        self._check(self.lib.ppg_assoc_device_results(self.h, *[C.byref(x) for x in p]))
        return [x.value for x in p]
