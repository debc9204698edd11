/*
 * ppg_b200.h -- C ABI of the B200-native PPG-SLAM front-end (libppg_b200.so).
 *
 * Drop-in boundary for two reference classes (all file:line relative to the PPG-SLAM tree):
 *   PPGExtractor   feature/include/PPGExtractor.h:34-148, feature/src/PPGExtractor.cpp:55-603
 *   Matcher        matching/include/Matcher.h:20-64, search core matching/src/Matcher.cpp:224-281
 * The header-only C++ shims (include/ppg_shim.hpp) rebuild those classes on top of these entry points.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (PPG_OK) or a negative
 * ppg_status, never throws, never aborts.  One ppg_ctx per GPU; calls on one ctx are serialised by the
 * caller (the reference extractor is single-threaded too); different ctxs are
 * independent.  There is NO CPU fallback: without a CUDA device ppg_create fails with PPG_ERR_CUDA.
 */
#ifndef PPG_B200_H
#define PPG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPG_API_VERSION 7
#define PPG_DESC_DIM 256 /* PPGExtractor::DESC_DIM_SIZE, PPGExtractor.cpp:44 */

typedef enum {
    PPG_OK = 0,
    PPG_ERR_ARG = -1,      /* bad argument / unsupported shape */
    PPG_ERR_CUDA = -2,     /* CUDA runtime or driver error (ppg_last_error has the string) */
    PPG_ERR_WEIGHTS = -3,  /* weight blob missing or malformed */
    PPG_ERR_CAPACITY = -4, /* a per-frame capacity (edges, colines, pair candidates) was exceeded;
                              the frame's `status` word says which.  Results for that frame are invalid. */
    PPG_ERR_NCCL = -5
} ppg_status;

/* bits of ppg_frame_out.status */
#define PPG_FRAME_OVF_ACCEPT 1u /* > acc_cap NMS survivors */
#define PPG_FRAME_OVF_PAIRS 2u  /* > pair_cap point pairs passed the 3-point heat test */
#define PPG_FRAME_OVF_DEGREE 4u /* a keypoint exceeded the adjacency capacity */
#define PPG_FRAME_OVF_EDGES 8u  /* > max_edges final edges */
#define PPG_FRAME_OVF_COLINE 16u

typedef struct ppg_ctx ppg_ctx;

/* Replaces the PPGExtractor constructor arguments (GeometricCamera*, net dir; PPGExtractor.cpp:55-107)
 * plus the class's static tunables (PPGExtractor.cpp:44-53) and Matcher::TH_LOW/TH_HIGH
 * (Matcher.cpp:12-13). */
typedef struct {
    int device;               /* CUDA ordinal (reference hard-codes cuda:0, PPGExtractor.cpp:41) */
    int width, height;        /* GeometricCamera::imWidth/imHeight; both must be multiples of 16 */
    float K[9];               /* GeometricCamera::toK(), row major */
    float D[4];               /* GeometricCamera::toD(): pinhole k1 k2 p1 p2, KB8 k0..k3 */
    int fisheye;              /* mnType == CAM_FISHEYE */
    const char* weights_path; /* flat export of the net .pt files (tools/export_weights.py) */
    float junction_thresh;    /* JUNCTION_THRESH 1/128 */
    int junction_nms_radius;  /* JUNCTION_NMS_RADIUS 4 (<= 8) */
    int junction_max_num;     /* JUNCTION_MAX_NUM 500 */
    float line_valid_thresh;  /* LINE_VALID_THRESH 0.01 */
    float line_valid_ratio;   /* LINE_VALID_RATIO 0.3 */
    float line_dist_thresh;   /* LINE_DISTTHRESH 2.0 */
    int heatmap_refine_sz;    /* HEATMAP_REFINE_SZ: only 16 is supported */
    float line_heatmap_thresh;/* LINE_HEATMAP_THRESH 0.2 */
    float line_inlier_rate;   /* LINE_INLIER_RATE 0.8 */
    float th_low, th_high;    /* Matcher::TH_LOW 0.7, TH_HIGH 0.8 */
    int max_batch;            /* frames per ppg_extract call (the reference is batch 1) */
    int max_edges;            /* capacity of the per-frame edge list */
    int max_colines;          /* capacity of the per-frame coline-pair list */
    int max_map_points;       /* capacity of the resident map-descriptor table (rows on this GPU) */
} ppg_config;

/* Output record of PPGExtractor::run (PPGExtractor.cpp:118-147) for one frame, SoA.
 * Pointers refer to ctx-owned pinned host memory, valid until the next call on the ctx.
 *   KeyPointEx  (sensors/include/GeometricCamera.h:22-37):
 *     mPos   = (kp_x, kp_y)   -- as run() returns it: pinhole mPos <- mPosUn (:141-145)
 *     mPosUn = (kp_xun, kp_yun), mfScore = kp_score, mbOut = kp_out
 *     mvConnected[i] = conn_idx[conn_off[i] .. conn_off[i+1])
 *     mvColine[i]    = pairs col_pairs[2k], col_pairs[2k+1] for k in [col_off[i], col_off[i+1])
 *   KeyEdge (feature/include/PPGGraph.h:32-55): startIdx, endIdx, lscore (isBad is always false)
 *   descriptors: n_kp x 256 fp32 rows (cv::Mat CV_32FC1 in the reference) */
typedef struct {
    int n_kp, n_edges, n_colines;
    uint32_t status;       /* 0, or PPG_FRAME_OVF_* bits */
    int n_candidates;      /* pixels >= junction_thresh (diagnostic) */
    int n_pairs_tested_ok; /* point pairs passing the 3-point heat test (diagnostic) */
    int n_candidate_lines; /* candidateLines.size() after the overlap filter (diagnostic) */
    int nms_rounds;        /* parallel NMS rounds used (diagnostic) */
    const float* kp_x;
    const float* kp_y;
    const int32_t* kp_px; /* detection pixel (integer mPos before :141-145) */
    const int32_t* kp_py;
    const float* kp_score;
    const float* kp_xun;
    const float* kp_yun;
    const uint8_t* kp_out;
    const int32_t* edge_start;
    const int32_t* edge_end;
    const float* edge_score;
    const int32_t* conn_off; /* n_kp + 1 */
    const int32_t* conn_idx;
    const int32_t* col_off;  /* n_kp + 1 */
    const int32_t* col_pairs;
    const float* desc;       /* n_kp x 256 */
    int diag[8];             /* diagnostics of the point-pair graph kernels: [0],[1],[3],[4] SM clock cycles / 16 of
                                setup, overlap filter, edges + adjacency, colinearity; [2] candidates with a block entry,
                                [5] candidates visited by the sequential pass, [6] of those with spilled lists */
} ppg_frame_out;

void ppg_default_config(ppg_config* cfg);
int ppg_create(const ppg_config* cfg, ppg_ctx** out);
void ppg_destroy(ppg_ctx* ctx);
/* Thread-local message of the last failure on this ctx (or of ppg_create when ctx is NULL). */
const char* ppg_last_error(const ppg_ctx* ctx);
int ppg_api_version(void);

/* ---- extraction: PPGExtractor::run for n_frames independent frames ------------------------------
 * gray[f] points to an 8-bit single-channel image with row stride stride[f] bytes (cv::Mat data/step;
 * PPGExtractor.cpp:121 asserts one channel).  Host memory.  Synchronous: copies in, runs the networks
 * and the post-processing on the GPU, copies the records out.  Returns PPG_ERR_CAPACITY if any frame's
 * status is non-zero (records of the other frames are still valid). */
int ppg_extract(ppg_ctx* ctx, const uint8_t* const* gray, const int* stride, int n_frames, ppg_frame_out* out);

/* The same three steps separately, so that a caller (bench.py) can time the device part alone or
 * overlap copies itself.  ppg_upload_frames: host -> device staging.  ppg_run: networks +
 * post-processing on frames already resident (asynchronous on the ctx stream).  ppg_download:
 * device -> pinned host, synchronises, fills `out`. */
int ppg_upload_frames(ppg_ctx* ctx, const uint8_t* const* gray, const int* stride, int n_frames);
int ppg_run(ppg_ctx* ctx, int n_frames);
int ppg_download(ppg_ctx* ctx, int n_frames, ppg_frame_out* out);
int ppg_sync(ppg_ctx* ctx);

/* ---- pipelined form (one host thread keeps several contexts busy) -----------------------------------------------
 * The reference's run() blocks on the GPU (PPGExtractor.cpp:125 `torch::cuda::synchronize`); a throughput caller wants
 * the copies of batch k + 1 behind the kernels of batch k instead.  ppg_extract_async enqueues host -> device copy,
 * networks, post-processing and the device -> host copy of the records on the ctx stream and returns at once;
 * ppg_extract_wait blocks until they are done and fills `out` (same records, same return codes as ppg_extract).
 * Frames that lie in pinned host memory (ppg_host_alloc, or the caller's own buffers after ppg_host_register, e.g. the
 * cv::Mat data of a camera ring buffer) are DMA'd from where they are; pageable frames still go through the ctx's
 * staging buffer, which costs a memcpy and a wait for the previous batch's DMA.  The frames must stay valid until
 * ppg_extract_wait / ppg_sync returns.  One batch in flight per ctx: the records live in the ctx's pinned mirror. */
int ppg_extract_async(ppg_ctx* ctx, const uint8_t* const* gray, const int* stride, int n_frames);
int ppg_extract_wait(ppg_ctx* ctx, int n_frames, ppg_frame_out* out);
long long ppg_record_bytes(const ppg_ctx* ctx); /* bytes of one frame's record block (what ppg_extract_async copies back) */
void* ppg_host_alloc(size_t bytes); /* pinned, portable across devices; NULL on failure */
void ppg_host_free(void* p);
int ppg_host_register(void* p, size_t bytes);
int ppg_host_unregister(void* p);

/* ---- parity / debug entry points ----------------------------------------------------------------
 * ppg_extract_from_maps: PPGExtractor::run minus the networks (detectKeyPoint, detectLines,
 * genPointDescriptor; PPGExtractor.cpp:126-146) fed with caller-supplied dense maps (host fp32):
 * prob n x H x W (softmax + pixel_shuffle junction map, :161-162), heat n x H x W (softmax[:,1] BEFORE
 * refine, :242), desc n x 256 x Hc x Wc (raw dense descriptors, CHW as LibTorch holds them). */
int ppg_extract_from_maps(ppg_ctx* ctx, const float* prob, const float* heat, const float* desc_chw, int n_frames,
                          ppg_frame_out* out);
/* Dense maps of frame `frame` of the last ppg_run/ppg_extract*, copied to host.  Any pointer may be
 * NULL.  prob/heat_raw/heat_final are H x W; desc_chw is 256 x Hc x Wc; feature_chw is 128 x Hc x Wc. */
int ppg_get_maps(ppg_ctx* ctx, int frame, float* prob, float* heat_raw, float* heat_final, float* desc_chw,
                 float* feature_chw);
/* Runs every tensor-core convolution layer of the last batch's frame 0 again through a plain fp32
 * CUDA-core convolution over the same fp16 operands and reports the max abs difference per layer.
 * names: up to max_layers pointers to static strings. */
int ppg_selftest_conv(ppg_ctx* ctx, int max_layers, const char** names, float* max_abs_diff, float* max_abs_ref,
                      int* n_layers);
/* Validation aid: the raw output tensor (NHWC, fp16 or fp32 as the layer stores it) of tensor-core layer `name`
 * ("conv1b", "conv2a", ...) for frame `frame` of the last batch.  Copies min(max_bytes, size) bytes, *bytes = size. */
int ppg_get_layer_output(ppg_ctx* ctx, const char* name, int frame, void* dst, size_t max_bytes, size_t* bytes);

/* Per-stage device times of the last ppg_run when profiling is on (CUDA events between launches on
 * the ctx stream).  names[i] are static strings; returns the stage count through n_stages. */
int ppg_set_profiling(ppg_ctx* ctx, int on);
int ppg_get_stage_times(ppg_ctx* ctx, int max_stages, const char** names, float* ms, int* n_stages);
/* Kernel launches issued on the ctx stream since creation (own kernels only). */
long long ppg_launch_count(const ppg_ctx* ctx);
/* CUDA-event stopwatch on the ctx stream: start/stop enqueue events, elapsed synchronises. */
int ppg_timer_start(ppg_ctx* ctx);
int ppg_timer_stop(ppg_ctx* ctx, float* ms);

/* ---- association: search core of Matcher::ExtendMapMatches (Matcher.cpp:224-281) ----------------
 * The map-descriptor table (MapPoint::GetDescriptor rows, feature/src/MapPoint.cpp:304-308) is kept
 * resident on the GPU.  `row0` is the global index of this ctx's first row when the table is row-sharded
 * across ctxs/GPUs (0 otherwise); indices returned are keypoint indices, rows are local. */
int ppg_upload_map(ppg_ctx* ctx, const float* map_desc, int n_rows);

typedef struct {
    int n_kp;                /* frame keypoints N */
    const float* kp_x;       /* mvKeysUn[i].mPos (Frame.cpp:138-156) */
    const float* kp_y;
    const float* frame_desc; /* N x 256 fp32, unit rows (Frame::mDescriptors) */
    const uint8_t* free_mask;/* N: 1 = keypoint may be matched (Matcher.cpp:253 inverted) */
    int n_rows;              /* map points M (<= rows uploaded) */
    const float* proj_uv;    /* M x 2: MapPoint::mTrackProjX/Y */
    const float* view_cos;   /* M: MapPoint::mTrackViewCos */
    float th;                /* search radius factor (Matcher.cpp:240-244: r = th * (cos>0.998 ? 2.5 : 4)) */
    float ratio;             /* Matcher::mfNNratio */
    int mode;                /* PPG_SEARCH_EXTEND_MAP (0) or PPG_SEARCH_WINDOW (1), see below */
    float max_dist;          /* mode 1: accept iff best_dist <= max_dist */
    double e2_max;           /* mode 1: > 0 -> candidates with ex*ex + ey*ey > e2_max are skipped */
    const int32_t* row_node; /* mode 2: M, FeatureVector node of every row's feature (-1 = not listed) */
    const int32_t* kp_node;  /* mode 2: N, FeatureVector node of every frame feature (ppg_bow_out.node_id) */
} ppg_assoc_in;

/* Search rules.  The window walk (Frame/KeyFrame::GetFeaturesInArea), the free mask, DescriptorDistance and the
 * first-minimum tie rule are shared; radius and acceptance differ:
 *   PPG_SEARCH_EXTEND_MAP  Matcher::ExtendMapMatches (Matcher.cpp:224-281): r = th * (viewCos > 0.998 ? 2.5 : 4),
 *                          accept = !(best > TH_HIGH && best > ratio * second)
 *   PPG_SEARCH_WINDOW      the best-only projection cores: r = th, accept = best <= max_dist
 *                            SearchByProjection(Cur, Last, th)                  :31-87      max_dist = TH_HIGH
 *                            SearchByProjection(F, KF, sFound, th, descDist)    :1337-1411  max_dist = descDist
 *                            Fuse(KF, vpMapPoints, th)                          :897-1036   max_dist = TH_LOW,
 *                                                                               e2_max = 5.99 (:1000-1005)
 *                            SearchByProjection(KF, Scw, vpPoints, vpMatched, th, ratioHamming)
 *                                                                               :479-568    max_dist = TH_LOW * ratioHamming,
 *                                                                               free_mask = !vpMatched[idx]
 *                            Fuse(KF, Scw, vpPoints, th, vpReplacePoint)        :1038-1135  max_dist = TH_LOW
 *                            SearchBySim3(KF1, KF2, vpMatches12, S12, th)       :1149-1335  both directions,
 *                                                                               max_dist = TH_HIGH, free_mask all ones
 *                          view_cos is ignored.
 *   PPG_SEARCH_NODE        the bag-of-words matchers: candidates = the frame features of the row's vocabulary node
 *                          (FeatureVector, ppg_bow_transform) in ascending index, no window (proj_uv / view_cos / th
 *                          unused, kp_x / kp_y may be zeros), accept = best <= max_dist && best < ratio * second
 *                            SearchByBoW(KF, F, vpMapPointMatches)              :393-477    rows = the keyframe's features
 *                                                                               that hold a good map point, max_dist = TH_LOW,
 *                                                                               free_mask = !vpMapPointMatches[idx]
 *                            SearchByBoW(KF1, KF2, vpMatches12)                 :663-754    max_dist = TH_LOW,
 *                                                                               free_mask = !vbMatched2[idx] && map point good */
enum { PPG_SEARCH_EXTEND_MAP = 0, PPG_SEARCH_WINDOW = 1, PPG_SEARCH_NODE = 2 };

typedef struct {       /* caller-allocated host arrays of n_rows */
    int32_t* best_idx;   /* -1 when the window is empty */
    int32_t* second_idx; /* -1 when fewer than two candidates */
    float* best_dist;    /* DescriptorDistance (MapPoint.cpp:22-29); 1e6 when absent */
    float* second_dist;
    uint8_t* accept;     /* mode 0: !(best > TH_HIGH && best > ratio*second), Matcher.cpp:276; mode 1: best <= max_dist */
} ppg_assoc_out;

/* Inputs/outputs in host memory; synchronous. */
int ppg_associate(ppg_ctx* ctx, const ppg_assoc_in* in, ppg_assoc_out* out);
/* Device-resident variant for timing: inputs staged by ppg_assoc_stage, kernels enqueued by
 * ppg_assoc_run (asynchronous), results fetched by ppg_assoc_fetch. */
int ppg_assoc_stage(ppg_ctx* ctx, const ppg_assoc_in* in);
int ppg_assoc_run(ppg_ctx* ctx);
int ppg_assoc_fetch(ppg_ctx* ctx, ppg_assoc_out* out);
/* Associates frame `frame` of the last extraction batch (descriptors and keypoints still on the
 * device) against the staged projections -- the extract+associate step of the benchmark. */
int ppg_assoc_run_frame(ppg_ctx* ctx, int frame);
/* Throughput form of the same step: every frame of the last extraction batch against the resident table in
 * ONE set of launches.  proj_uv is n_frames x n_rows x 2, view_cos n_frames x n_rows (host, each frame has its
 * own projections of the same map points); results come back per frame through outs[0..n_frames). */
int ppg_assoc_stage_batch(ppg_ctx* ctx, int n_frames, int n_rows, const float* proj_uv, const float* view_cos, float th,
                          float ratio);
int ppg_assoc_run_batch(ppg_ctx* ctx, int n_frames);
int ppg_assoc_fetch_batch(ppg_ctx* ctx, int n_frames, ppg_assoc_out* outs);
/* ppg_assoc_stage_batch without the host round trip: proj_uv / view_cos must lie in pinned host memory and stay valid
 * until the ctx is synchronised (PPG_ERR_ARG for pageable pointers); nothing is waited for. */
int ppg_assoc_stage_batch_async(ppg_ctx* ctx, int n_frames, int n_rows, const float* proj_uv, const float* view_cos,
                                float th, float ratio);
/* Rows whose tensor-core candidate filter could not guarantee the exact top-2 and were re-scored
 * over the whole window (diagnostic). */
int ppg_assoc_fallback_rows(ppg_ctx* ctx, int* n);

/* ---- map-descriptor table from raw observations: MapPoint::ComputeDistinctiveDescriptors ---------
 * (feature/src/MapPoint.cpp:234-302) for n_points map points at once.  desc = the observation descriptors of
 * all points packed back to back (rows of 256 floats, each point's rows in the order the caller iterates
 * mObservations), offsets[n_points + 1] = first row of every point.  best_idx[p] = BestIdx of point p (index
 * into its own list; 0 when no median distance is below 1.0, as in the reference).  At most
 * PPG_MAX_OBSERVATIONS rows per point (PPG_ERR_CAPACITY otherwise).
 * ppg_upload_map_distinctive additionally makes the chosen rows the resident association table (row p = the
 * representative descriptor of point p), i.e. ppg_upload_map without the host round trip; best_idx may be NULL. */
#define PPG_MAX_OBSERVATIONS 128
int ppg_distinctive_descriptors(ppg_ctx* ctx, const float* desc, const int32_t* offsets, int n_points,
                                int32_t* best_idx);
int ppg_upload_map_distinctive(ppg_ctx* ctx, const float* desc, const int32_t* offsets, int n_points,
                               int32_t* best_idx);

/* ---- Frame::CheckInFrustum (map/src/Frame.cpp:223-260) on the device ----------------------------------
 * The projections that ExtendMapMatches consumes (mbTrackInView, mTrackProjX/Y, mTrackDepth, mTrackViewCos) computed
 * from the map geometry and the frame poses, so that per frame only a 15-float pose crosses the bus instead of 12 bytes
 * per map point.  Geometry per resident row: MapPoint::GetWorldPos, GetNormal, GetMinDistanceInvariance,
 * GetMaxDistanceInvariance.  Pose per frame: Frame::mRcw (row major), mtcw, mOw.  Projection =
 * GeometricCamera::project of the ctx's camera (Pinhole.cpp:32-38 / KannalaBrandt8.cpp:44-59), image test =
 * GeometricCamera::IsInImage (GeometricCamera.cpp:21-24).  ppg_assoc_stage_poses replaces ppg_assoc_stage_batch:
 * rows that are not in view get no search window (Matcher.cpp:212).  MapPoint::IncreaseVisible (:259) stays with
 * the caller (ppg_frustum_fetch returns the flags). */
int ppg_upload_map_geometry(ppg_ctx* ctx, const float* world_pos, const float* normal, const float* min_dist,
                            const float* max_dist, int n_rows);
int ppg_assoc_stage_poses(ppg_ctx* ctx, int n_frames, int n_rows, const float* Rcw, const float* tcw, const float* Ow,
                          float cos_limit, float th, float ratio);
/* n_frames x n_rows each (any pointer may be NULL): mbTrackInView, mTrackProjX/Y (-1 when not in view), mTrackDepth
 * (-1), mTrackViewCos (0). */
int ppg_frustum_fetch(ppg_ctx* ctx, int n_frames, uint8_t* in_view, float* proj_uv, float* depth, float* view_cos);

/* ---- the whole Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381) on the GPU ---------------
 * Search core as above PLUS the sequential part: the walk over the candidate map points in the order of
 * getEdges().size() (descending; ties keep table order), the live "keypoint already holds an observed map point"
 * test (:253), the assignment (:279-281) and the seed growing over (map edges of pMP) x (key edges of the matched
 * keypoint) with its greedy minimum-weight assignment (:287-377).  Nothing of it runs on the host.
 *
 * The pointer graph is passed in POD form.  The resident table (ppg_upload_map / ppg_upload_map_distinctive) holds
 * n_points rows: the candidate map points AND every map point reachable as theOtherPt() of one of their edges (its
 * descriptor is read at :330).  Per row:
 *   candidate  !isBad() && mbTrackInView (:210-215)        observed  Observations() > 0 (:253)
 *   bad        isBad() (:227, :364)
 *   edge_off[n_points + 1], edge_other[], edge_ok[]  CSR of MapPoint::getEdges() in vector order: the row of
 *              theOtherPt(pMP) (-1 = nullptr) and !isBad() && mbValid (:311) of every edge.  A map edge is named by
 *              its CSR position in the results.
 * Capacities: at most PPG_EXTEND_MAX_DEGREE edges per map point and key edges per keypoint, and (valid map edges) x
 * (key edges) <= PPG_EXTEND_MAX_WEIGHTS per seed; beyond that the frame's status gets PPG_EXTEND_OVF and the call
 * returns PPG_ERR_CAPACITY.  Requires ratio < 1 (with ratio >= 1 the reference itself writes mvpMapPoints[-1]). */
#define PPG_EXTEND_MAX_DEGREE 128
#define PPG_EXTEND_MAX_WEIGHTS 4096
#define PPG_EXTEND_OVF 1u
typedef struct {
    int n_points;
    const uint8_t* candidate;
    const uint8_t* observed;
    const uint8_t* bad;
    const int32_t* edge_off;
    const int32_t* edge_other;
    const uint8_t* edge_ok;
} ppg_map_graph;
int ppg_upload_map_graph(ppg_ctx* ctx, const ppg_map_graph* graph);

typedef struct {
    int n_kp;                 /* Frame::N */
    const float* kp_x;        /* mvKeysUn[i].mPos */
    const float* kp_y;
    const float* frame_desc;  /* N x 256 */
    const int32_t* kp_mp;     /* N: F.mvpMapPoints[i] as a table row, -1 = nullptr, -2 = a map point outside the table
                                 that has observations; NULL = all -1 */
    int n_edges;              /* F.mvKeyEdges.size() (<= ppg_config.max_edges) */
    const int32_t* edge_start;/* KeyEdge::startIdx / endIdx */
    const int32_t* edge_end;
    const int32_t* conn_off;  /* n_kp + 1: CSR of mvKeysUn[i].mvConnected */
    const int32_t* conn_idx;
    const int32_t* kedge_me;  /* n_edges: F.mvpMapEdges[e] as a CSR position, -1 = nullptr; NULL = all -1 */
    const float* proj_uv;     /* n_points x 2: mTrackProjX/Y (read for candidate rows only) */
    const float* view_cos;    /* n_points.  Both NULL: use the projections ppg_assoc_stage_poses left on the device
                                 (Frame::CheckInFrustum there; rows not in view are no candidates) */
    const uint8_t* tracked;   /* n_points: mnTrackedbyFrame == F.mnId; NULL = none */
    float th, ratio;
} ppg_extend_in;

typedef struct {        /* caller-allocated */
    int32_t* kp_mp;     /* n_kp: F.mvpMapPoints after the call (batch form: room for junction_max_num) */
    int32_t* kedge_me;  /* n_edges: F.mvpMapEdges after the call (batch form: room for max_edges) */
    uint8_t* tracked;   /* n_points, may be NULL */
    int nmatches;       /* the reference's return value (two increments per accepted map point, :281 and :378) */
    uint32_t status;    /* 0 or PPG_EXTEND_OVF */
    int n_kp, n_edges;  /* entries written to kp_mp / kedge_me */
    int n_accepted;     /* map points accepted by the window search (diagnostic) */
    int n_grown;        /* map points matched by seed growing (diagnostic) */
    int n_rescans;      /* rows whose stored candidate list had to be rebuilt from the whole window (diagnostic) */
    int diag[9];        /* walk kernel: [0] evaluation rounds, [1..5] SM clock cycles / 16 of setup, chunk loads,
                           evaluation, event handling, seed growing; [6] seeds popped; [7] cycles / 16 of the weight matrices
                           (part of seed growing); [8] seeds skipped because no other endpoint of pMP was left to match */
} ppg_extend_out;

/* One frame, everything in host memory; synchronous. */
int ppg_extend_map_matches(ppg_ctx* ctx, const ppg_extend_in* in, ppg_extend_out* out);
/* Every frame of the last extraction batch (keypoints, descriptors and point-pair graph still on the device, no map
 * point assigned yet) against the projections staged with ppg_assoc_stage_batch (n_rows = n_points); asynchronous. */
int ppg_extend_run_batch(ppg_ctx* ctx, int n_frames);
int ppg_extend_fetch_batch(ppg_ctx* ctx, int n_frames, ppg_extend_out* outs);
/* The fetch split for pipelined callers: ..._async enqueues the device -> pinned-host copies of the results behind the
 * kernels; after ppg_extract_wait / ppg_sync, ppg_extend_collect hands them out (no CUDA call, no wait). */
int ppg_extend_fetch_batch_async(ppg_ctx* ctx, int n_frames);
int ppg_extend_collect(ppg_ctx* ctx, int n_frames, ppg_extend_out* outs);

/* ---- bag of words: DBoW3::Vocabulary::transform (Frame::ComputeBoW, map/src/Frame.cpp:331-340) --------
 * The vocabulary is the k-ary tree of the reference's Vocabulary/voc_*_9x3.gz (DBoW3 binary; tools/export_vocabulary.py
 * converts it to these flat arrays): children[n_nodes][k] in DBoW3's visiting order (-1 padded, leaves all -1),
 * word_id[n_nodes] (-1 for inner nodes), weight[n_nodes] (the leaf's idf), desc[n_nodes][256].  transform:
 *   per feature: descend from the root taking at every level the FIRST child with the least DescManip::distance
 *   (sum of float (a_i - b_i)^2 accumulated in double in index order); word / weight = the leaf's; FeatureVector node =
 *   the node reached at level L - levelsup, the root when that is <= 0 (the reference: L = 3, levelsup = 4);
 *   BowVector = the weights of the features with weight > 0 summed per word in feature order, then divided by their
 *   L2 norm (scoring 1), L1 norm (0, 2, 3, 4) or, for DOT_PRODUCT (5), by the number of words.
 * Weighting must be TF_IDF (0) or TF (1). */
typedef struct {
    int k, L, scoring, weighting, n_nodes, dim;
    const int32_t* children;
    const int32_t* word_id;
    const double* weight;
    const float* desc;
} ppg_vocabulary;
int ppg_upload_vocabulary(ppg_ctx* ctx, const ppg_vocabulary* voc);

/* Host-side reader (no GPU needed): the reference's Vocabulary/voc_*_9x3.gz as System loads them with
 * DBoW3::Vocabulary::load (DBoW3 binary stream, QuickLZ level-1 chunks) or the flat blob of tools/export_vocabulary.py.
 * `view` points into memory owned by the handle until ppg_vocabulary_close.  ppg_load_vocabulary = open + upload. */
typedef struct ppg_voc_file ppg_voc_file;
int ppg_vocabulary_open(const char* path, ppg_voc_file** out, ppg_vocabulary* view);
void ppg_vocabulary_close(ppg_voc_file* voc);
const char* ppg_vocabulary_error(void); /* why the last ppg_vocabulary_open on this thread failed */
int ppg_load_vocabulary(ppg_ctx* ctx, const char* path);

typedef struct {          /* caller-allocated (any pointer may be NULL) */
    int n_features;       /* out: features transformed */
    int n_bow;            /* out: entries of the BowVector */
    int32_t* word_id;     /* n_features: WordId of every feature */
    double* word_weight;  /* n_features: its weight */
    int32_t* node_id;     /* n_features: FeatureVector node of the feature, -1 when its weight is <= 0 (not listed) */
    int32_t* bow_word;    /* n_bow: BowVector keys, ascending */
    double* bow_value;    /* n_bow: BowVector values */
} ppg_bow_out;
/* One frame, descriptors (n_features x 256) in host memory; synchronous. */
int ppg_bow_transform(ppg_ctx* ctx, const float* desc, int n_features, int levelsup, ppg_bow_out* out);
/* Every frame of the last extraction batch (descriptors still on the device); asynchronous + fetch. */
int ppg_bow_run_batch(ppg_ctx* ctx, int n_frames, int levelsup);
int ppg_bow_fetch_batch(ppg_ctx* ctx, int n_frames, ppg_bow_out* outs);

/* ---- the whole Matcher::SearchByBoW on the GPU (matching/src/Matcher.cpp:393-477 and :663-754) ------------
 * The keyframe features that hold a good map point are the rows (their descriptors uploaded with ppg_upload_map, in
 * the reference's visiting order: FeatureVector node ascending, then feature index -- with the reference's
 * levelsup = 4 every feature is under the root, i.e. plain feature order); the frame / second keyframe gives its
 * descriptors and the node of every feature (ppg_bow_out.node_id; -1 = not listed, or for SearchByBoW(KF1, KF2) a
 * feature without a good map point, :707-711).  A row's candidates are the features of its node that no earlier row
 * has taken (vpMapPointMatches[idx] / vbMatched2[idx], live); accept = best <= max_dist (strict = 0, :456) or
 * best < max_dist (strict = 1, :733), and best < ratio * second.  kp_row[i] = the row matched to feature i or -1. */
typedef struct {
    int n_rows;
    const int32_t* row_node;
    int n_kp;
    const float* frame_desc;
    const int32_t* kp_node;
    float ratio, max_dist;
    int strict;
} ppg_bow_match_in;
typedef struct {
    int32_t* kp_row; /* n_kp, caller-allocated */
    int nmatches;    /* the reference's return value */
    int n_rescans;   /* diagnostic, as in ppg_extend_out */
} ppg_bow_match_out;
int ppg_search_by_bow(ppg_ctx* ctx, const ppg_bow_match_in* in, ppg_bow_match_out* out);

/* ---- the whole Matcher::SearchForInitialization on the GPU (matching/src/Matcher.cpp:582-651; called from
 * system/src/Tracking.cpp:525 with windowSize 50).  The n1 descriptors of F1 are the resident table (ppg_upload_map);
 * for every F1 feature in index order: the radius-`window` box around prev_matched[i1] in F2
 * (Frame::GetFeaturesInArea), best / second best DescriptorDistance over the features not matched yet, accept iff
 * best <= th_low && best < ratio * second (the reference's vector<int> vMatchedDistance makes a matched F2 feature
 * unavailable for good, see oracle/ppg_oracle.c).  Needs no map graph. */
typedef struct {
    int n1;                    /* F1.mvKeysUn.size() = rows uploaded with ppg_upload_map (F1.mDescriptors) */
    const float* prev_matched; /* n1 x 2: vbPrevMatched */
    int n2;
    const float* kp2_x;        /* n2: F2.mvKeysUn[i].mPos */
    const float* kp2_y;
    const float* desc2;        /* n2 x 256: F2.mDescriptors */
    int window;                /* windowSize */
    float ratio;               /* Matcher::mfNNratio */
} ppg_init_match_in;
typedef struct {
    int32_t* matches12;  /* n1: vnMatches12 (index in F2 or -1), or NULL */
    float* prev_matched; /* n1 x 2: vbPrevMatched after the update of :644-647 (may alias the input), or NULL */
    int nmatches;
    int n_rescans;
} ppg_init_match_out;
int ppg_search_for_initialization(ppg_ctx* ctx, const ppg_init_match_in* in, ppg_init_match_out* out);

/* ---- the whole Matcher::SearchForTriangulation on the GPU (matching/src/Matcher.cpp:767-885; called from
 * LocalMapping to find untracked feature pairs of two key frames), for both camera models of the reference.  For every
 * feature i1 of KF1 without a map point: over the features i2 of KF2 under the same vocabulary node (FeatureVector,
 * ascending index) without a map point, the smallest DescriptorDistance <= th_low (a tie takes the later one, :842) among
 * those that lie at least 10 px from the epipole (:846) and pass mpCamera->epipolarConstrain (:848):
 *   camera_model 0, Pinhole (sensors/src/Pinhole.cpp:98-114): dsqr < 3.84 from kp1's epipolar line, closed form from F12;
 *   camera_model 1, KannalaBrandt8 (sensors/src/KannalaBrandt8.cpp:167-236): unproject both pixels, parallax test,
 *     triangulation (null vector of the 4 x 4 DLT matrix), positive depth and reprojection error < 5.991 px^2 in both
 *     images, from cam8 / R12 / t12.
 * The reference never sets vbMatched2, so the features of KF1 are independent of one another: one warp per feature.
 * F12, R12, t12 and the epipole depend on the two poses only; the caller computes them with the reference's own
 * expressions (Matcher.cpp:776-788, Pinhole.cpp:101-104; ppg_shim.hpp does). */
typedef struct {
    int n1, n2;
    const float* desc1;      /* n1 x 256: pKF1->mDescriptors */
    const float* desc2;      /* n2 x 256 */
    const int32_t* node1;    /* n1: the node of pKF1->mFeatVec that lists the feature, -1 if none */
    const int32_t* node2;    /* n2 */
    const uint8_t* has_mp1;  /* n1: pKF1->GetMapPoint(i) != NULL */
    const uint8_t* has_mp2;  /* n2 */
    const float* pos1;       /* n1 x 2: pKF1->mvKeysUn[i].mPos */
    const float* pos2;       /* n2 x 2 */
    float F12[9];            /* row-major K1^-T [t12]x R12 K2^-1 (camera_model 0) */
    float epipole[2];        /* mpCamera->project(T2w * Cw) */
    float th_low;            /* Matcher::TH_LOW */
    int camera_model;        /* 0 Pinhole, 1 KannalaBrandt8 (since API version 7) */
    float cam8[8];           /* camera_model 1: fx fy cx cy k0 k1 k2 k3 (GeometricCamera::mvParameters) */
    float R12[9];            /* camera_model 1: row-major T12.rotationMatrix(), Matcher.cpp:786 */
    float t12[3];            /* camera_model 1: T12.translation(), :787 */
} ppg_triangulation_match_in;
typedef struct {
    int32_t* match12; /* n1, caller-allocated: index in KF2 or -1 (vMatches12) */
    int nmatches;
} ppg_triangulation_match_out;
int ppg_search_for_triangulation(ppg_ctx* ctx, const ppg_triangulation_match_in* in, ppg_triangulation_match_out* out);

/* ---- the whole Matcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th) on the GPU
 * (matching/src/Matcher.cpp:31-87; MSTracking::TrackWithMotionModel, system/src/Tracking.cpp:811/817, every tracked
 * frame) and the whole Matcher::SearchByProjection(Frame &CurrentFrame, KeyFrame*, sAlreadyFound, th, descDist)
 * (:1337-1411; relocalisation, Tracking.cpp:1297/1311).  Both are sequential: an accepted match occupies its keypoint
 * for the later map points.  The rows are the map points that passed the reference's projection tests (:38-56 /
 * :1347-1371: the caller evaluates them with the reference's own camera, ppg_shim.hpp does), in loop order; their
 * descriptors (pMP->GetDescriptor()) are the resident table (ppg_upload_map, n_rows rows).  For every row: the window of
 * radius th around proj_uv in the current frame (Frame::GetFeaturesInArea), the smallest DescriptorDistance over the
 * window's keypoints that are not occupied (first minimum in visiting order), accepted iff <= max_dist (Matcher::TH_HIGH
 * for the motion model, descDist for relocalisation); the keypoint then holds the row.
 * kp_mp: CurrentFrame.mvpMapPoints as rows of this table -- -1 none (or a map point without observations, which does
 * not occupy, :71-73), -2 a map point outside the table that occupies, >= 0 a row (occupies iff observed[row]).
 * The relocalisation variant tests the pointer alone (:1386): pass observed = NULL (all ones) and -2 for every
 * assigned keypoint.  Matcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming) (:479-568,
 * loop closing) has the same structure over a key frame's keypoints (vpMatched[idx] occupies, :543) with max_dist =
 * TH_LOW * ratioHamming: the same call. */
typedef struct {
    int n_rows;
    const float* proj_uv;    /* n_rows x 2 */
    const uint8_t* observed; /* n_rows: pMP->Observations() > 0; NULL = all */
    int n;                   /* keypoints of CurrentFrame */
    const float* kp_x;       /* n: mvKeysUn[i].mPos[0] */
    const float* kp_y;
    const float* desc;       /* n x 256: CurrentFrame.mDescriptors */
    const int32_t* kp_mp;    /* n, or NULL = all -1 */
    float th;                /* window radius */
    float max_dist;          /* TH_HIGH / descDist */
} ppg_projection_match_in;
typedef struct {
    int32_t* kp_mp; /* n, caller-allocated: CurrentFrame.mvpMapPoints after the call (same coding) */
    int nmatches;
    int n_rescans;  /* diagnostics: rows whose stored window list ran dry and were rescanned */
} ppg_projection_match_out;
int ppg_search_by_projection(ppg_ctx* ctx, const ppg_projection_match_in* in, ppg_projection_match_out* out);

/* Device pointers of the staged association results (n_rows each), for the sharded all-gather that
 * the multi-GPU host layer issues through NCCL (ppg_slam_b200/sharded.py). */
int ppg_assoc_device_results(ppg_ctx* ctx, void** best_idx, void** second_idx, void** best_dist, void** second_dist,
                             void** accept);
void* ppg_stream(ppg_ctx* ctx); /* cudaStream_t of the ctx */

/* ---- row-sharded association over the GPUs of one box (BASELINE.json config 5) -------------------------------------
 * Every rank holds M / world rows of the map table (ppg_upload_map of its rows) and the whole frame; after
 * ppg_assoc_stage + ppg_assoc_run (or ppg_assoc_run_frame) on its rows, ppg_assoc_allgather packs the per-row records
 * {best_idx, second_idx, best_dist bits, second_dist bits, accept} (5 x int32) and enqueues ONE ncclAllGather of them
 * on the ctx stream behind the kernels -- no host synchronisation in between.  rows_per_rank is the send count common
 * to all ranks (>= n_local; the records past n_local are {-1, -1, 0, 0, 0}).  ppg_assoc_allgather_fetch waits and
 * returns the [world][rows_per_rank][5] table and the device time of the collective.
 * The communicator: rank 0 calls ppg_comm_unique_id and hands the 128 bytes to the other ranks by whatever means the
 * host program has (bench.py: torch.distributed broadcast), then every rank calls ppg_comm_init.  NCCL is loaded at
 * run time (libnccl.so.2); PPG_ERR_NCCL if it is missing or a call fails. */
int ppg_comm_unique_id(void* id128);
int ppg_comm_init(ppg_ctx* ctx, const void* id128, int rank, int world);
int ppg_comm_destroy(ppg_ctx* ctx);
int ppg_assoc_allgather(ppg_ctx* ctx, int n_local, int rows_per_rank);
int ppg_assoc_allgather_fetch(ppg_ctx* ctx, int32_t* records, float* gather_us);

#ifdef __cplusplus
}
#endif
#endif /* PPG_B200_H */
