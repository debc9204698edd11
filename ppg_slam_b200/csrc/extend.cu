// The whole Matcher::ExtendMapMatches (matching/src/Matcher.cpp:203-381) on the GPU.
//
// The reference walks the candidate map points one after the other because every accepted match changes what
// the later ones may take (Matcher.cpp:253) and seed growing assigns further keypoints on the way (:287-377).
// Split used here:
//
//   X1 extend_lists_kernel (grid-wide, one warp per candidate map point and frame)
//        everything that does NOT depend on the walk: the keypoints of the search window (Frame::GetFeaturesInArea,
//        Frame.cpp:262-315) with their exact DescriptorDistance, sorted by (distance, visiting order) -- the first
//        X_LIST of them are stored.  Whatever the walk has taken by the time it reaches a row, that row's best /
//        second best are the first two FREE entries of its list (strict-< update at :262-271 = first minimum in
//        visiting order), so the walk never computes a window distance again (unless fewer than two free entries
//        are left of a list that was cut, which falls back to a CTA-wide rescan of the window).
//   X2 extend_walk_kernel (one CTA per frame)
//        the sequential part.  256 threads hold 256 consecutive rows of the sorted order and evaluate them against
//        the live occupancy in shared memory; the first row that is ACCEPTED is an event (rows that are rejected or
//        empty change nothing and cost nothing), processed by the whole CTA: assignment, then seed growing -- the
//        weight matrix (:324-340) is computed by 8 warps, one exact distance each, the greedy minimum-weight
//        assignment (:342-374) by one warp.  The rows after the event are re-evaluated and so on.
//
// Results are bit-identical to the CPU restatement of the reference function (tests/test_gpu_extend.py).
// Built with -fmad=false.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "assoc.cuh"
#include "ctx.cuh"
#include "once.cuh"

namespace ppg {

constexpr int X_LIST = 32;                           // window candidates stored per row
constexpr int X_PRE = 8;                             // map edges of a row preloaded with its chunk (row header layout)
constexpr int X_LCAP = PPG_EXTEND_MAX_DEGREE;        // map edges per map point / key edges per keypoint
constexpr int X_WCAP = PPG_EXTEND_MAX_WEIGHTS;       // weight-matrix entries per seed
constexpr int X_THREADS = 256, X_WARPS = X_THREADS / 32;
constexpr int X_NOEVENT = 1 << 20;
constexpr int G_CELLS = 64 * 48;                     // FRAME_GRID_COLS x FRAME_GRID_ROWS (GeometricCamera.h:79-80)
constexpr int G_STRIDE = (G_CELLS + 1 + 7) / 8 * 8;  // cell-start table of a frame, padded to whole 16-byte vectors

// Where the point-pair graph of frame f lives: staged arrays (one frame) or the extraction output blocks.
struct FrameGraphSrc {
    const uint8_t *es, *ee, *coff, *cidx, *ne;
    size_t stride;
    int ne_val;
    __device__ const int* es_of(int f) const { return reinterpret_cast<const int*>(es + f * stride); }
    __device__ const int* ee_of(int f) const { return reinterpret_cast<const int*>(ee + f * stride); }
    __device__ const int* coff_of(int f) const { return reinterpret_cast<const int*>(coff + f * stride); }
    __device__ const int* cidx_of(int f) const { return reinterpret_cast<const int*>(cidx + f * stride); }
    __device__ int ne_of(int f) const { return ne ? *reinterpret_cast<const int*>(ne + f * stride) : ne_val; }
};

enum { XR_NMATCHES = 0, XR_STATUS, XR_ACCEPTED, XR_GROWN, XR_RESCANS, XR_NKP, XR_NEDGES, XR_ROUNDS,
       XR_T_SETUP, XR_T_CHUNK, XR_T_EVAL, XR_T_EVENT, XR_T_SEED, XR_SEEDS, XR_T_WEIGHTS, XR_SEEDS_SKIPPED,
       XR_WORDS = 16 };  // XR_T_*: SM cycles / 16

struct ExtendState {
    // map graph (shared by all frames)
    int P = 0, nc = 0, edge_cap = 0;
    uint8_t *observed = nullptr, *bad = nullptr, *edge_ok = nullptr;
    int *edge_off = nullptr, *edge_other = nullptr, *order = nullptr;
    int4* row_hdr = nullptr;  // [max_rows][3]
    // per-row candidate lists, indexed by position in the sorted order: [bcap][max_rows][X_LIST]
    uint16_t* l_idx = nullptr;
    float* l_d = nullptr;
    uint8_t* l_cnt = nullptr;  // [bcap][max_rows] window size, saturated at 255
    // per-frame state
    uint8_t* tracked = nullptr;  // [bcap][max_rows]
    int* kp_mp = nullptr;        // [bcap][ncap]
    int* kedge_me = nullptr;     // [bcap][ecap]
    int* result = nullptr;       // [bcap][XR_WORDS]
    int ecap = 0;
    // staged point-pair graph of one frame (ppg_extend_map_matches)
    int *g_es = nullptr, *g_ee = nullptr, *g_coff = nullptr, *g_cidx = nullptr;
    // node mode (SearchByBoW): no map graph -> zero CSR / flags, identity order, every row "observed"
    int *zero_i = nullptr, *ident = nullptr, *row_node = nullptr, *kp_node = nullptr;
    uint8_t* ones_u8 = nullptr;
    uint8_t* proj_obs = nullptr;  // [max_rows] Observations() > 0 of the rows of ppg_search_by_projection
    // Frame grid of every frame of the batch in GetFeaturesInArea's visiting order (grid_index_kernel)
    float2* gs_xy = nullptr;      // [bcap][ncap] positions by rank
    uint16_t* gs_idx = nullptr;   // [bcap][ncap] keypoint index by rank
    uint16_t* gs_start = nullptr; // [bcap][G_STRIDE] first rank of every cell (cx * 48 + cy), [G_CELLS] = total
    // pinned mirrors for the fetch
    int *h_kp_mp = nullptr, *h_kedge_me = nullptr, *h_result = nullptr;
    uint8_t* h_tracked = nullptr;
};

namespace {

// ------------------------------------------------------------------------------------------------
struct ListParams {
    int nc, max_rows, ncap;
    FrameSrc src;
    const int* order;
    const RowParam* rowp;
    const float* map_f32;
    const uint32_t* kinfo;
    const uint32_t* korder;
    uint16_t* l_idx;
    float* l_d;
    uint8_t* l_cnt;
    // node mode (Matcher::SearchByBoW): the "window" of a row = the frame features of its FeatureVector node, in
    // ascending index; no geometry
    int node_mode;
    const int* row_node;
    const int* kp_node;
    // spatial mode: the frame grid by rank (grid_index_kernel)
    const float2* gs_xy;
    const uint16_t* gs_idx;
    const uint16_t* gs_start;
};

// The frame grid of Frame::AssignFeaturesToGrid (Frame.cpp:138-156) as a CSR in GetFeaturesInArea's visiting order
// (Frame.cpp:294-312: cell columns ix ascending, rows iy ascending, the keypoints of a cell in insertion = index order):
// rank-sorted positions / indices of the indexable, free keypoints and the first rank of every cell.  A search window
// then is, per cell column, ONE contiguous interval of ranks.  One CTA per frame; kinfo from prep_frame_kernel.
__global__ void __launch_bounds__(256) grid_index_kernel(const FrameSrc src, int ncap, const uint32_t* __restrict__ kinfo,
                                                         float2* __restrict__ gs_xy, uint16_t* __restrict__ gs_idx,
                                                         uint16_t* __restrict__ gs_start) {
    __shared__ uint32_t cnt[G_CELLS];
    __shared__ uint16_t start[G_CELLS + 1];
    __shared__ uint32_t wsum[8];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(src.n_of(f), ncap);
    const uint32_t* info = kinfo + (size_t)f * ncap;
    for (int k = tid; k < G_CELLS; k += 256) cnt[k] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const uint32_t v = info[i];
        if (v & 0x10000u) atomicAdd(&cnt[(v & 0xffu) * 48 + ((v >> 8) & 0xffu)], 1u);
    }
    __syncthreads();
    {   // exclusive scan of the 3072 counters: 12 per thread
        uint32_t loc[12], sum = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            loc[k] = sum;
            sum += cnt[tid * 12 + k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const uint32_t t = __shfl_up_sync(AFULL, inc, m);
            if (lane >= m) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t base = inc - sum;
        for (int w = 0; w < warp; w++) base += wsum[w];
#pragma unroll
        for (int k = 0; k < 12; k++) start[tid * 12 + k] = (uint16_t)(base + loc[k]);
        if (tid == 255) start[G_CELLS] = (uint16_t)(base + sum);
    }
    __syncthreads();
    for (int k = tid; k < G_CELLS; k += 256) cnt[k] = 0;  // now: keypoints of the cell placed so far
    __syncthreads();
    if (warp == 0) {  // ascending index, 32 at a time: rank = first rank of the cell + earlier keypoints of the same cell
        const float* kx = src.kx_of(f);
        const float* ky = src.ky_of(f);
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const uint32_t v = i < n ? info[i] : 0u;
            const bool ok = (v & 0x10000u) != 0u;
            const uint32_t cell = (v & 0xffu) * 48 + ((v >> 8) & 0xffu);
            const unsigned grp = __match_any_sync(AFULL, ok ? cell : (uint32_t)(G_CELLS + lane));
            const int off = __popc(grp & ((1u << lane) - 1u));
            uint32_t before = 0;
            if (ok) before = cnt[cell];
            __syncwarp();
            if (ok && off == 0) cnt[cell] = before + __popc(grp);
            __syncwarp();
            if (ok) {
                const size_t o = (size_t)f * ncap + start[cell] + before + off;
                gs_xy[o] = make_float2(kx[i], ky[i]);
                gs_idx[o] = (uint16_t)i;
            }
        }
    }
    for (int k = tid; k < G_STRIDE; k += 256) gs_start[(size_t)f * G_STRIDE + k] = start[min(k, G_CELLS)];
}

// DescriptorDistance with both rows already in registers (same operation order as exact_distance).
__device__ __forceinline__ float exact_distance_rr(const float (&av)[8], const float (&bv)[8]) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float d = av[k] - bv[k];
        s = s + d * d;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s = s + __shfl_xor_sync(AFULL, s, m);
    return sqrtf(s);
}

// Four DescriptorDistances against the same row at once, with the SAME sums as exact_distance_rr: the xor butterfly adds
// own + partner at every level and float addition is commutative, so a lane may as well keep only the distances its
// half of the partners is responsible for -- two after the first level, one after the second -- which takes 6 shuffles
// and 6 additions instead of 20 + 20 and one square root instead of four; 4 shuffles hand the results to every lane.
__device__ __forceinline__ void exact_distance4_rr(const float (&av)[8], const float (&bv)[4][8], int lane, float (&d)[4]) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        s[u] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float t = av[k] - bv[u][k];
            s[u] = s[u] + t * t;
        }
    }
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    float k0 = h16 ? s[2] : s[0], k1 = h16 ? s[3] : s[1];
    const float g0 = h16 ? s[0] : s[2], g1 = h16 ? s[1] : s[3];
    k0 = k0 + __shfl_xor_sync(AFULL, g0, 16);
    k1 = k1 + __shfl_xor_sync(AFULL, g1, 16);
    float v = h8 ? k1 : k0;
    const float g = h8 ? k0 : k1;
    v = v + __shfl_xor_sync(AFULL, g, 8);
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) v = v + __shfl_xor_sync(AFULL, v, m);
    v = sqrtf(v);  // lanes 0-7: distance 0, 8-15: 1, 16-23: 2, 24-31: 3
#pragma unroll
    for (int u = 0; u < 4; u++) d[u] = __shfl_sync(AFULL, v, 8 * u);
}

constexpr int XL_ROWS = 64;  // rows per CTA (8 per warp): the frame's keypoint table is staged once for all of them

// grid (ceil(nc / XL_ROWS), frames).  The window scan reads the keypoint table from shared memory; the hits of a row
// are collected first and their descriptor rows are then fetched four at a time (the first version took one
// dependent round of global loads per 32 keypoints scanned and one per hit: 0.39 ms per 32 frames x 8192 rows).
__global__ void __launch_bounds__(256, 4) extend_lists_kernel(const ListParams p) {
    __shared__ uint32_t s_ord[1024];
    __shared__ uint16_t s_hit[8][1024];
    // node mode: the node of every feature; spatial mode: the frame grid by rank (positions, indices, cell starts)
    __shared__ __align__(16) uint8_t s_mode[1024 * 8 + 1024 * 2 + G_STRIDE * 2];
    uint32_t* s_info = reinterpret_cast<uint32_t*>(s_mode);
    float2* s_sxy = reinterpret_cast<float2*>(s_mode);
    uint16_t* s_sidx = reinterpret_cast<uint16_t*>(s_mode + 1024 * 8);
    uint16_t* s_start = s_sidx + 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, f = blockIdx.y;
    const int n = min(p.src.n_of(f), p.ncap);
    // this warp's eight rows: ids and window parameters in lanes 0-7, fetched next to the staging loads below (two
    // dependent round trips per CTA instead of two per row)
    int pre_row = -1, pre_node = -1;
    RowParam pre_rp;
    pre_rp.u = pre_rp.v = pre_rp.r = pre_rp.na2 = 0.f;
    pre_rp.cells = 0xffffffffu;
    if (lane < XL_ROWS / 8) {
        const int ql = blockIdx.x * XL_ROWS + lane * 8 + warp;
        if (ql < p.nc) {
            pre_row = p.order[ql];
            if (p.node_mode)
                pre_node = p.row_node[pre_row];
            else
                pre_rp = p.rowp[(size_t)f * p.max_rows + pre_row];
        }
    }
    if (p.node_mode) {
        for (int i = threadIdx.x; i < n; i += 256) {
            s_info[i] = (uint32_t)p.kp_node[i];
            s_ord[i] = (uint32_t)i;  // vIndicesF is filled in feature order
        }
    } else {
        const uint32_t* korder = p.korder + (size_t)f * p.ncap;
        const float2* gxy = p.gs_xy + (size_t)f * p.ncap;
        const uint16_t* gidx = p.gs_idx + (size_t)f * p.ncap;
        const uint16_t* gst = p.gs_start + (size_t)f * G_STRIDE;
        const int ng = gst[G_CELLS];  // indexable, free keypoints
        for (int k = threadIdx.x; k < G_STRIDE / 8; k += 256)
            reinterpret_cast<uint4*>(s_start)[k] = reinterpret_cast<const uint4*>(gst)[k];
        for (int i = threadIdx.x; i < n; i += 256) s_ord[i] = korder[i];
        for (int i = threadIdx.x; i < ng; i += 256) {
            s_sxy[i] = gxy[i];
            s_sidx[i] = gidx[i];
        }
    }
    __syncthreads();
    const float* fdesc = p.src.desc_of(f);
    for (int r = 0; r < XL_ROWS / 8; r++) {
        const int q = blockIdx.x * XL_ROWS + r * 8 + warp;
        if (q >= p.nc) break;
        const int row = __shfl_sync(AFULL, pre_row, r);
        const int row_next = r + 1 < XL_ROWS / 8 ? __shfl_sync(AFULL, pre_row, (r + 1) & 31) : -1;  // -1 past the end
        const size_t ol = (size_t)f * p.max_rows + q;
        RowParam rp;
        rp.u = __shfl_sync(AFULL, pre_rp.u, r);
        rp.v = __shfl_sync(AFULL, pre_rp.v, r);
        rp.r = __shfl_sync(AFULL, pre_rp.r, r);
        rp.cells = __shfl_sync(AFULL, pre_rp.cells, r);
        const int my_node = __shfl_sync(AFULL, pre_node, r);
        if (p.node_mode) rp.cells = my_node < 0 ? 0xffffffffu : 0u;
        // the sorted list lives across the warp: lane k holds its k-th entry (empty = +inf)
        float ld = INFINITY;
        uint32_t lo = 0xffffffffu;
        int li = -1;
        int nh = 0;
        if (rp.cells != 0xffffffffu) {
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = p.map_f32[(size_t)row * 256 + lane + 32 * k];
            const uint32_t cx0 = rp.cells & 0xff, cx1 = (rp.cells >> 8) & 0xff, cy0 = (rp.cells >> 16) & 0xff,
                           cy1 = rp.cells >> 24;
            if (p.node_mode) {
                // the features of the row's node, four groups of 32 per iteration (Matcher.cpp:417-418)
                for (int c0 = 0; c0 < n; c0 += 128) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = c0 + 32 * u + lane;
                        const bool in = c < n && (int)s_info[c] == my_node;
                        const unsigned mask = __ballot_sync(AFULL, in);
                        if (in) s_hit[warp][nh + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)c;
                        nh += __popc(mask);
                    }
                }
            } else {
                // Frame::GetFeaturesInArea (Frame.cpp:262-315) over the grid: the cells [cx0, cx1] x [cy0, cy1] of one
                // column are one interval of ranks; a lane takes a column and walks its interval (about one keypoint per
                // column for a 80 x 80 px window), the distance test of :305-309 decides.  The order in which the hits
                // are collected does not matter: the list below is sorted by (distance, visiting order).  (The first
                // version tested all ~350 keypoints of the frame for every map point: 40 % of the kernel's instructions.)
                const int ncol = (int)cx1 - (int)cx0 + 1;
                for (int c0 = 0; c0 < ncol; c0 += 32) {
                    int rk = 0, rk_end = 0;  // my column's interval of ranks
                    if (c0 + lane < ncol) {
                        const int col = ((int)cx0 + c0 + lane) * 48;
                        rk = s_start[col + cy0];
                        rk_end = s_start[col + cy1 + 1];
                    }
                    while (__any_sync(AFULL, rk < rk_end)) {
                        bool in = false;
                        if (rk < rk_end) {
                            const float2 pt = s_sxy[rk];
                            in = fabsf(pt.x - rp.u) < rp.r && fabsf(pt.y - rp.v) < rp.r;
                        }
                        const unsigned mask = __ballot_sync(AFULL, in);
                        if (in) s_hit[warp][nh + __popc(mask & ((1u << lane) - 1u))] = s_sidx[rk];
                        nh += __popc(mask);
                        rk++;
                    }
                }
            }
            __syncwarp();
            // the next row's map descriptor on its way to L1 while this row's hits are scored (8 lines of 128 bytes)
            if (row_next >= 0 && lane < 8)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.map_f32 + (size_t)row_next * 256 + lane * 32));
            for (int h0 = 0; h0 < nh; h0 += 4) {
                int cc[4];
                float b[4][8];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    cc[u] = s_hit[warp][min(h0 + u, nh - 1)];
#pragma unroll
                    for (int k = 0; k < 8; k++) b[u][k] = fdesc[(size_t)cc[u] * 256 + lane + 32 * k];
                }
                // and the next four hits' rows (4 x 8 lines, one per lane) while these four are reduced and inserted
                if (h0 + 4 + (lane >> 3) < nh)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(
                        fdesc + (size_t)s_hit[warp][h0 + 4 + (lane >> 3)] * 256 + (lane & 7) * 32));
                float d4[4];
                exact_distance4_rr(a, b, lane, d4);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (h0 + u >= nh) break;  // warp-uniform
                    const float d = d4[u];
                    const uint32_t ord = s_ord[cc[u]];
                    // entries that stay in front of the new one form a prefix of the lanes
                    const bool before = ld < d || (ld == d && lo < ord);
                    const int pos = __popc(__ballot_sync(AFULL, before));
                    const float ud = __shfl_up_sync(AFULL, ld, 1);
                    const uint32_t uo = __shfl_up_sync(AFULL, lo, 1);
                    const int ui = __shfl_up_sync(AFULL, li, 1);
                    if (lane == pos) {
                        ld = d; lo = ord; li = cc[u];
                    } else if (lane > pos) {
                        ld = ud; lo = uo; li = ui;
                    }
                }
            }
            __syncwarp();
        }
        {
            p.l_idx[ol * X_LIST + lane] = li < 0 ? (uint16_t)0xffff : (uint16_t)li;
            p.l_d[ol * X_LIST + lane] = ld;
        }
        if (lane == 0) p.l_cnt[ol] = (uint8_t)min(nh, 255);
    }
}

// ------------------------------------------------------------------------------------------------
struct WalkParams {
    int nc, P, max_rows, ncap, ecap;
    FrameSrc src;
    FrameGraphSrc gsrc;
    const int* order;
    const int4* row_hdr;  // [nc][3] or null (then order / edge_off / edge_other / edge_ok are read per row)
    const RowParam* rowp;
    const float* map_f32;
    const uint32_t* kinfo;
    const uint32_t* korder;
    const uint16_t* l_idx;
    const float* l_d;
    const uint8_t* l_cnt;
    const uint8_t *observed, *bad, *edge_ok;
    const int *edge_off, *edge_other;
    uint8_t* tracked;  // in (when has_state) / out
    int* kp_mp;        // in (when has_state) / out
    int* kedge_me;     // in / out (initialised by the host side)
    int* result;
    float ratio, th_high;
    int has_state;
    // node mode (Matcher::SearchByBoW, Matcher.cpp:393-477 / :663-754): accept = best <= (or <) max_dist && best <
    // ratio * second; every matched feature is taken (vpMapPointMatches / vbMatched2), no map edges
    int node_mode, strict;
    float max_dist;
    const int* row_node;
    const int* kp_node;
    // and_rule (Matcher::SearchForInitialization, Matcher.cpp:582-651): spatial windows like ExtendMapMatches, but the
    // accept rule of the node mode (best <= max_dist && best < ratio * second) and every matched feature is taken
    int and_rule;
    // best_only (Matcher::SearchByProjection x 2, Matcher.cpp:31-87 / :1337-1411): spatial windows, accept = best <= max_dist,
    // no second best; every accepted row takes its keypoint
    int best_only;
};

struct WalkShared {
    int kpmp[1024];      // F.mvpMapPoints as table rows
    uint8_t occ[1024];   // mvpMapPoints[i] && Observations() > 0   (Matcher.cpp:253)
    float w[X_WCAP];     // weight matrix of the current seed, [lx position][key edge]
    int po[X_LCAP];      // other map point of the valid map edges of pMP (lx)
    int lxi[X_LCAP];     // their position in getEdges()
    uint16_t taken[X_LCAP + 2];    // keypoints that became occupied during the last event
    int queue[X_LCAP + 2];         // matchSeed: every push marks one more endpoint of pMP tracked
    uint16_t cl[1024];   // window of a rescanned row
    float t2d[X_WARPS][2];
    uint32_t t2o[X_WARPS][2];
    int t2i[X_WARPS][2];
    int wfirst[2][X_WARPS];  // double-buffered by iteration parity: one barrier per evaluation round is enough
    int ev[8];           // row, act, best idx, first map edge, map edges
    int cln, nlx, qn, freed, ntaken;
    int res[XR_WORDS];
};

__device__ __forceinline__ bool trk_get(const uint32_t* trk, int r) { return (trk[r >> 5] >> (r & 31)) & 1u; }
__device__ __forceinline__ void trk_set(uint32_t* trk, int r) { trk[r >> 5] |= 1u << (r & 31); }

__global__ void __launch_bounds__(X_THREADS, 1) extend_walk_kernel(const WalkParams p) {
    extern __shared__ __align__(16) uint8_t xsm[];
    WalkShared& S = *reinterpret_cast<WalkShared*>(xsm);
    const int tw = (p.P + 31) >> 5;
    uint32_t* trk = reinterpret_cast<uint32_t*>(xsm + ((sizeof(WalkShared) + 15) & ~size_t(15)));
    uint32_t* obs = trk + tw;   // Observations() > 0 of every table row
    uint32_t* badb = obs + tw;  // isBad()
    uint16_t* s_coff = reinterpret_cast<uint16_t*>(badb + tw);  // [ncap + 1] CSR of mvConnected
    uint16_t* s_cko = s_coff + p.ncap + 2;                      // [2 ecap] other keypoint of every CSR entry
    uint16_t* s_cke = s_cko + 2 * p.ecap;                       // [2 ecap] its key edge
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t_begin = clock64();
    const int n = min(p.src.n_of(f), p.ncap);
    const int ne = min(p.gsrc.ne_of(f), p.ecap);
    const float* fdesc = p.src.desc_of(f);
    int* kedge_me = p.kedge_me + (size_t)f * p.ecap;
    uint8_t* tracked_g = p.tracked + (size_t)f * p.max_rows;
    int* kpmp_g = p.kp_mp + (size_t)f * p.ncap;

    // ---- initial state
    for (int w = tid; w < tw; w += X_THREADS) {
        uint32_t tb = 0, ob = 0, bb = 0;
#pragma unroll
        for (int b = 0; b < 32; b++) {  // unconditional loads (clamped) so that all of them are in flight together
            const int r = min(w * 32 + b, p.P - 1);
            const uint32_t t = tracked_g[r], o = p.observed[r], d = p.bad[r];
            if (w * 32 + b < p.P) {
                if (p.has_state && t) tb |= 1u << b;
                if (o) ob |= 1u << b;
                if (d) bb |= 1u << b;
            }
        }
        trk[w] = tb;
        obs[w] = ob;
        badb[w] = bb;
    }
    {   // the frame's point-pair graph: every mvConnected entry with KeyEdge::theOtherPid resolved
        const int* es = p.gsrc.es_of(f);
        const int* ee = p.gsrc.ee_of(f);
        const int* coff = p.gsrc.coff_of(f);
        const int* cidx = p.gsrc.cidx_of(f);
        for (int i = tid; i <= n && n > 0; i += X_THREADS) s_coff[i] = (uint16_t)min(coff[i], 2 * p.ecap);
        for (int i = tid; i < n; i += X_THREADS) {
            const int k0 = coff[i], k1 = min(coff[i + 1], 2 * p.ecap);
            for (int k = k0; k < k1; k++) {
                const int e = cidx[k];
                const int s0 = es[e], e0 = ee[e];
                s_cko[k] = (uint16_t)(s0 == i ? e0 : s0);
                s_cke[k] = (uint16_t)e;
            }
        }
    }
    if (tid < XR_WORDS) S.res[tid] = 0;
    if (tid == 0) S.freed = S.ntaken = 0;
    __syncthreads();
    for (int i = tid; i < p.ncap; i += X_THREADS) {
        int r = -1;
        if (p.has_state && i < n) r = kpmp_g[i];
        S.kpmp[i] = r;
        S.occ[i] = r >= 0 ? trk_get(obs, r) : (r == -2);
    }
    __syncthreads();

    int round = 0;
    long long tc = clock64(), t_chunk = 0, t_eval = 0, t_event = 0, t_seed = 0, t_weights = 0;
    if (tid == 0) S.res[XR_T_SETUP] = (int)((tc - t_begin) >> 4);
    for (int base = 0; base < p.nc; base += X_THREADS) {
        // ---- static part of the 256 rows of this chunk: my row's stored window list (thread-local, indexed
        // dynamically: L1-resident local memory) and its first map edges (registers)
        const int pos = base + tid;
        int row = -1, cnt = 0, my_me0 = 0, my_nme = 0;
        uint32_t li[X_LIST / 2];  // keypoints of my list, two per word: registers (only unrolled, static indexing)
        uint32_t lim[X_LIST / 2];  // the same for the dynamic look-up of the best entry: thread-local memory
        float ld[X_LIST];          // exact distances, thread-local memory
        int pe[X_PRE];  // theOtherPt row of my first map edges, -1 when the edge is not usable (:312-318)
#pragma unroll
        for (int k = 0; k < X_PRE; k++) pe[k] = -1;
#pragma unroll
        for (int k = 0; k < X_LIST / 2; k++) li[k] = 0xffffffffu;
        if (pos < p.nc) {
            if (!p.row_hdr) row = p.order[pos];
            const size_t ol = (size_t)f * p.max_rows + pos;
            const uint4* pi = reinterpret_cast<const uint4*>(p.l_idx + ol * X_LIST);
            const float4* pd = reinterpret_cast<const float4*>(p.l_d + ol * X_LIST);
            uint4 iv[X_LIST / 8];
            float4 dv[X_LIST / 4];
#pragma unroll
            for (int k = 0; k < X_LIST / 8; k++) iv[k] = pi[k];
#pragma unroll
            for (int k = 0; k < X_LIST / 4; k++) dv[k] = pd[k];
            cnt = p.l_cnt[ol];
            if (p.row_hdr) {
                // row, first map edge, edge count and the first eight usable other endpoints in one 48-byte record
                // per sorted position (built when the map graph is uploaded): one round trip instead of three
                // dependent ones (order -> edge_off -> edge_other / edge_ok)
                const int4* hp = p.row_hdr + (size_t)pos * 3;
                const int4 h0 = hp[0], h1 = hp[1], h2 = hp[2];
                row = h0.x;
                my_me0 = h0.y;
                my_nme = h0.z;
                pe[0] = h0.w; pe[1] = h1.x; pe[2] = h1.y; pe[3] = h1.z;
                pe[4] = h1.w; pe[5] = h2.x; pe[6] = h2.y; pe[7] = h2.z;
            } else {
                my_me0 = p.edge_off[row];
                my_nme = p.edge_off[row + 1] - my_me0;
#pragma unroll
                for (int k = 0; k < X_PRE; k++)
                    if (k < my_nme) pe[k] = p.edge_ok[my_me0 + k] ? p.edge_other[my_me0 + k] : -1;
            }
#pragma unroll
            for (int k = 0; k < X_LIST / 8; k++) {
                li[4 * k] = iv[k].x; li[4 * k + 1] = iv[k].y; li[4 * k + 2] = iv[k].z; li[4 * k + 3] = iv[k].w;
            }
#pragma unroll
            for (int k = 0; k < X_LIST / 4; k++) {
                ld[4 * k] = dv[k].x; ld[4 * k + 1] = dv[k].y; ld[4 * k + 2] = dv[k].z; ld[4 * k + 3] = dv[k].w;
            }
            if (base + X_THREADS + tid < p.nc) {  // the next chunk's list on its way to L2 while this one is walked
                const size_t nl = ol + X_THREADS;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.l_idx + nl * X_LIST));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.l_d + nl * X_LIST));
            }
        }
#pragma unroll
        for (int k = 0; k < X_LIST / 2; k++) lim[k] = li[k];
        const int stored = cnt < X_LIST ? cnt : X_LIST;
        const int wmax = __reduce_max_sync(AFULL, stored);  // longest list of the warp: the unrolled loops stop there
        const uint32_t valid = stored >= 32 ? AFULL : (1u << stored) - 1u;
        // fm: which entries of my list are FREE keypoints.  Built once per chunk from the occupancy table; after that
        // every event publishes the keypoints it took (S.taken) and each thread strikes them out of its mask with
        // register compares -- no walk over the list, no dependent shared-memory look-ups per evaluation round.
        // S.freed counts the rare opposite case (an unobserved map point written over an observed one): rebuild.
        auto build_mask = [&]() -> uint32_t {
            uint32_t fm = 0;
#pragma unroll
            for (int k = 0; k < X_LIST / 2; k++) {
                if (2 * k >= wmax) break;
                const uint32_t w = li[k];
                if (2 * k < stored && !S.occ[w & 0xffffu]) fm |= 1u << (2 * k);
                if (2 * k + 1 < stored && !S.occ[w >> 16]) fm |= 1u << (2 * k + 1);
            }
            return fm;
        };
        int my_freed = S.freed;
        uint32_t fm = build_mask();
        bool fresh = true;  // fm was built from the table in this round: S.taken is already in it
        int cursor = 0;
        {
            const long long t = clock64();
            t_chunk += t - tc;
            tc = t;
        }
        while (true) {
            // ---- what would the reference do with my row in the current state?  (Matcher.cpp:229-277)
            int act = 0, bidx = -1;
            if (row >= 0 && tid >= cursor && cnt > 0 && !trk_get(trk, row)) {
                const int fr = S.freed;
                if (fr != my_freed) {
                    my_freed = fr;
                    fm = build_mask();
                } else if (!fresh) {
                    const int nt = S.ntaken;
                    for (int t = 0; t < nt; t++) {
                        const uint32_t kp = S.taken[t], kp2 = kp | (kp << 16);
#pragma unroll
                        for (int k = 0; k < X_LIST / 2; k++) {
                            if (2 * k >= wmax) break;
                            const uint32_t x = li[k] ^ kp2;
                            if ((x & 0xffffu) == 0) fm &= ~(1u << (2 * k));
                            if ((x >> 16) == 0) fm &= ~(1u << (2 * k + 1));
                        }
                    }
                }
                fm &= valid;
                const int nfree = __popc(fm);
                // (best_only needs the first free entry alone: the entries cut off the sorted list are all farther)
                if (nfree >= 2 || (nfree == 1 && (cnt <= X_LIST || p.best_only))) {
                    const int k1 = __ffs(fm) - 1;
                    const uint32_t rest = fm & (fm - 1);
                    const float b1 = ld[k1], b2 = rest ? ld[__ffs(rest) - 1] : 1e6f;
                    bidx = (int)((lim[k1 >> 1] >> ((k1 & 1) * 16)) & 0xffffu);
                    if (p.best_only)
                        act = b1 <= p.max_dist ? 1 : 0;  // :82 / :1401
                    else if (p.node_mode || p.and_rule)
                        act = ((p.strict ? b1 < p.max_dist : b1 <= p.max_dist) && b1 < p.ratio * b2) ? 1 : 0;  // :456-458
                    else
                        act = !(b1 > p.th_high && b1 > p.ratio * b2) ? 1 : 0;  // :276
                } else if (cnt > X_LIST) {
                    act = 2;  // the list was cut and fewer than two free entries are left of it: rescan the window
                }
                // nfree == 0 with the whole window stored: bestIdx stays -1, the row is rejected (ratio < 1)
            }
            fresh = false;
            const unsigned m = __ballot_sync(AFULL, act != 0);
            if (lane == 0) S.wfirst[round & 1][warp] = m ? warp * 32 + __ffs(m) - 1 : X_NOEVENT;
            __syncthreads();
            int first = X_NOEVENT;
#pragma unroll
            for (int w = 0; w < X_WARPS; w++) first = min(first, S.wfirst[round & 1][w]);
            round++;
            {
                const long long t = clock64();
                t_eval += t - tc;
                tc = t;
            }
            if (first == X_NOEVENT) {
                break;
            }
            if (tid == first) {
                // every thread is past its evaluation (barrier above): the state may change now
                S.ev[0] = row;
                S.ev[1] = act;
                S.ev[2] = bidx;
                S.ev[3] = my_me0;
                S.ev[4] = my_nme;
                // lx: the valid map edges of pMP (:312-318); mapEdge_set is pMP's for every seed of the event
                int nlx = -1;
                if (my_nme <= X_PRE) {
                    nlx = 0;
#pragma unroll
                    for (int k = 0; k < X_PRE; k++)
                        if (pe[k] >= 0) {
                            S.po[nlx] = pe[k];
                            S.lxi[nlx] = k;
                            nlx++;
                        }
                }
                S.nlx = nlx;  // -1: more edges than were preloaded, built from global memory below
                S.ntaken = 0;
                if (act == 1) {  // F.mvpMapPoints[bestIdx] = pMP (:279-281); bestIdx was free
                    S.kpmp[bidx] = row;
                    S.occ[bidx] = trk_get(obs, row);
                    if (S.occ[bidx]) S.taken[S.ntaken++] = (uint16_t)bidx;
                    trk_set(trk, row);
                    S.res[XR_NMATCHES] += 2;  // :281 and :378
                    S.res[XR_ACCEPTED]++;
                    S.queue[0] = bidx;
                    S.qn = 1;
                }
            }
            __syncthreads();
            const int erow = S.ev[0];
            // ---- a cut list ran dry: best / second best over the whole window with the live occupancy
            if (S.ev[1] == 2) {
                if (tid == 0) S.cln = 0;
                __syncthreads();
                const uint32_t* korder = p.korder + (size_t)f * p.ncap;
                if (p.node_mode) {
                    const int my_node = p.row_node[erow];
                    for (int c = tid; c < n; c += X_THREADS)
                        if (p.kp_node[c] == my_node && !S.occ[c]) S.cl[atomicAdd(&S.cln, 1)] = (uint16_t)c;
                } else {
                    const RowParam rp = p.rowp[(size_t)f * p.max_rows + erow];
                    const float* kx = p.src.kx_of(f);
                    const float* ky = p.src.ky_of(f);
                    const uint32_t* kinfo = p.kinfo + (size_t)f * p.ncap;
                    for (int c = tid; c < n; c += X_THREADS)
                        if (in_window(rp, kinfo[c], kx[c], ky[c], 0.0) && !S.occ[c])
                            S.cl[atomicAdd(&S.cln, 1)] = (uint16_t)c;
                }
                __syncthreads();
                float a[8];
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] = p.map_f32[(size_t)erow * 256 + lane + 32 * k];
                float b1 = 1e6f, b2 = 1e6f;
                uint32_t o1 = 0xffffffffu, o2 = 0xffffffffu;
                int i1 = -1, i2 = -1;
                const int ncl = S.cln;
                for (int k = warp; k < ncl; k += X_WARPS) {
                    const int cc = S.cl[k];
                    const float dd = exact_distance(a, fdesc + (size_t)cc * 256, lane);
                    top2_update(dd, p.node_mode ? (uint32_t)cc : korder[cc], cc, b1, o1, i1, b2, o2, i2);
                }
                if (lane == 0) {
                    S.t2d[warp][0] = b1; S.t2o[warp][0] = o1; S.t2i[warp][0] = i1;
                    S.t2d[warp][1] = b2; S.t2o[warp][1] = o2; S.t2i[warp][1] = i2;
                }
                __syncthreads();
                if (tid == 0) {
                    b1 = b2 = 1e6f;
                    o1 = o2 = 0xffffffffu;
                    i1 = i2 = -1;
                    for (int w = 0; w < X_WARPS; w++)
                        for (int k = 0; k < 2; k++)
                            if (S.t2i[w][k] >= 0) top2_update(S.t2d[w][k], S.t2o[w][k], S.t2i[w][k], b1, o1, i1, b2, o2, i2);
                    const bool acc = i1 >= 0 && (p.best_only ? b1 <= p.max_dist
                                                 : (p.node_mode || p.and_rule) ? ((p.strict ? b1 < p.max_dist : b1 <= p.max_dist) &&
                                                                b1 < p.ratio * b2)
                                                             : !(b1 > p.th_high && b1 > p.ratio * b2));
                    S.ev[1] = acc ? 1 : 0;
                    S.ev[2] = i1;
                    S.res[XR_RESCANS]++;
                    if (acc) {
                        S.kpmp[i1] = erow;
                        S.occ[i1] = trk_get(obs, erow);
                        if (S.occ[i1]) S.taken[S.ntaken++] = (uint16_t)i1;
                        trk_set(trk, erow);
                        S.res[XR_NMATCHES] += 2;
                        S.res[XR_ACCEPTED]++;
                        S.queue[0] = i1;
                        S.qn = 1;
                    }
                }
                __syncthreads();
            }
            {
                const long long t = clock64();
                t_event += t - tc;
                tc = t;
            }
            // ---- accepted: seed growing (:287-377).  Without map edges every seed is skipped at :300-301.
            const int me0 = S.ev[3], nme = S.ev[4];
            if (S.ev[1] == 1 && nme > 0) {
                if (nme > X_LCAP) {
                    if (tid == 0) S.res[XR_STATUS] |= PPG_EXTEND_OVF;
                } else {
                    if (S.nlx < 0) {  // uniform: S.nlx is rewritten only behind the barrier below
                        __syncthreads();
                        if (warp == 0) {
                            int nlx = 0;
                            for (int i0 = 0; i0 < nme; i0 += 32) {
                                const int i = i0 + lane;
                                int other = -1;
                                bool ok = false;
                                if (i < nme) {
                                    other = p.edge_other[me0 + i];
                                    ok = p.edge_ok[me0 + i] != 0 && other >= 0;
                                }
                                const unsigned mk = __ballot_sync(AFULL, ok);
                                if (ok) {
                                    const int k = nlx + __popc(mk & ((1u << lane) - 1u));
                                    S.po[k] = other;
                                    S.lxi[k] = i;
                                }
                                nlx += __popc(mk);
                            }
                            if (lane == 0) S.nlx = nlx;
                        }
                        __syncthreads();
                    }
                    const int nlx0 = S.nlx;
                    int qh = 0;
                    while (nlx0 > 0) {  // lx empty: the assignment loop never runs (:342)
                        if (qh >= S.qn) break;
                        const int keyID = S.queue[qh++];
                        if (tid == 0) S.res[XR_SEEDS]++;
                        const int ke0 = s_coff[keyID], nke = s_coff[keyID + 1] - ke0;
                        if (nke == 0) continue;  // :300-301
                        // The assignment loop changes the frame only through map points that are neither bad nor
                        // tracked (:364-365); once every other endpoint of pMP's edges is tracked -- the usual state
                        // after the first seed of an event -- a seed can only strike pairs out of its own lists, so
                        // its weight matrix and its loop are skipped altogether.  Every warp evaluates the same
                        // shared-memory state: the decision is uniform without a barrier.
                        {
                            bool live = false;
                            for (int i0 = 0; i0 < nlx0; i0 += 32) {
                                const int i = i0 + lane;
                                bool l = false;
                                if (i < nlx0) {
                                    const int po = S.po[i];
                                    l = !(trk_get(badb, po) || trk_get(trk, po));
                                }
                                live |= __ballot_sync(AFULL, l) != 0u;
                            }
                            if (!live) {
                                if (tid == 0) S.res[XR_SEEDS_SKIPPED]++;
                                continue;
                            }
                        }
                        if (nke > X_LCAP || nlx0 * nke > X_WCAP) {
                            if (tid == 0) S.res[XR_STATUS] |= PPG_EXTEND_OVF;
                            continue;
                        }
                        // weight matrix (:324-340): one warp per entry, four entries of a warp in flight
                        const int tot = nlx0 * nke;
                        const long long tw0 = clock64();
                        for (int q0 = warp; q0 < tot; q0 += 4 * X_WARPS) {
                            float av[4][8], bv[4][8];
                            bool same[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int q = min(q0 + u * X_WARPS, tot - 1);
                                const int i = q / nke, j = q - i * nke;
                                const int po = S.po[i], ko = s_cko[ke0 + j];
                                same[u] = po == S.kpmp[ko];
                                if (!same[u]) {
#pragma unroll
                                    for (int k = 0; k < 8; k++) {
                                        av[u][k] = p.map_f32[(size_t)po * 256 + lane + 32 * k];
                                        bv[u][k] = fdesc[(size_t)ko * 256 + lane + 32 * k];
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < 8; k++) av[u][k] = bv[u][k] = 0.f;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int q = q0 + u * X_WARPS;
                                if (q >= tot) break;  // warp-uniform
                                const float w = same[u] ? -1.f : exact_distance_rr(av[u], bv[u]);
                                if (lane == 0) S.w[q] = w;
                            }
                        }
                        __syncthreads();
                        t_weights += clock64() - tw0;
                        // greedy minimum-weight assignment (:342-374), one warp.  lx.erase / ly.erase keep the order of
                        // the remaining entries, so "first minimum over the current lx x ly" = the alive entry with the
                        // least (weight, original i, original j): no list is rebuilt, rows / columns are just struck out.
                        if (warp == 0) {
                            uint32_t ra[X_LCAP / 32], ca[X_LCAP / 32];  // alive map edges (lx) / key edges (ly)
#pragma unroll
                            for (int k = 0; k < X_LCAP / 32; k++) {
                                ra[k] = nlx0 >= 32 * (k + 1) ? AFULL : (nlx0 > 32 * k ? (1u << (nlx0 - 32 * k)) - 1u : 0u);
                                ca[k] = nke >= 32 * (k + 1) ? AFULL : (nke > 32 * k ? (1u << (nke - 32 * k)) - 1u : 0u);
                            }
                            auto alive = [&](const uint32_t(&m)[X_LCAP / 32], int i) -> bool {
                                uint32_t w = m[0];
#pragma unroll
                                for (int k = 1; k < X_LCAP / 32; k++) w = (i >> 5) == k ? m[k] : w;
                                return (w >> (i & 31)) & 1u;
                            };
                            auto strike = [&](uint32_t(&m)[X_LCAP / 32], int i) {
#pragma unroll
                                for (int k = 0; k < X_LCAP / 32; k++)
                                    if ((i >> 5) == k) m[k] &= ~(1u << (i & 31));
                            };
                            // the at-sign snapshot of :327-331: pMP_o == F.mvpMapPoints[keyID_o] is part of the weights above
                            const bool small = tot <= 32;
                            const int my_i = lane / nke, my_j = lane - my_i * nke;
                            const float my_w = (small && lane < tot) ? S.w[lane] : 1e6f;
                            int nlx = nlx0, nly = nke;
                            int qn_r = S.qn, nt_r = S.ntaken, freed_r = S.freed, grown_r = S.res[XR_GROWN];
                            while (nlx > 0 && nly > 0) {
                                float bw = 1e6f;
                                int bq = 0x7fffffff;
                                if (small) {
                                    if (lane < tot && alive(ra, my_i) && alive(ca, my_j) && my_w < bw) {
                                        bw = my_w;
                                        bq = lane;
                                    }
                                } else {
                                    for (int q = lane; q < tot; q += 32) {
                                        const int i = q / nke, j = q - i * nke;
                                        if (!alive(ra, i) || !alive(ca, j)) continue;
                                        const float w = S.w[q];
                                        if (w < bw) {  // strict <: the first minimum in (i, j) order wins
                                            bw = w;
                                            bq = q;
                                        }
                                    }
                                }
                                // warp argmin of (weight, q) with two hardware reductions: weights are -1 or >= 0, so
                                // their bit patterns (shifted by one, -1 -> 0) order like the values
                                const uint32_t wkey = bq == 0x7fffffff ? 0xffffffffu
                                                                       : (bw < 0.f ? 0u : __float_as_uint(bw) + 1u);
                                const uint32_t mkey = __reduce_min_sync(AFULL, wkey);
                                if (mkey == 0xffffffffu) break;
                                bw = mkey == 0u ? -1.f : __uint_as_float(mkey - 1u);
                                if (bw > p.th_high) break;  // :357-358
                                bq = (int)__reduce_min_sync(AFULL, wkey == mkey ? (uint32_t)bq : 0xffffffffu);
                                const int mi = bq / nke, kj = bq - mi * nke;
                                strike(ra, mi);
                                strike(ca, kj);
                                nlx--;
                                nly--;
                                if (lane == 0) {
                                    const int po = S.po[mi], ko = s_cko[ke0 + kj];
                                    if (!(trk_get(badb, po) || trk_get(trk, po))) {  // :364-365
                                        S.kpmp[ko] = po;                             // :366
                                        const bool nocc = trk_get(obs, po), wocc = S.occ[ko] != 0;
                                        if (wocc && !nocc) freed_r++;  // an observed map point was overwritten
                                        if (!wocc && nocc && nt_r < X_LCAP + 2) S.taken[nt_r++] = (uint16_t)ko;
                                        S.occ[ko] = nocc;
                                        kedge_me[s_cke[ke0 + kj]] = me0 + S.lxi[mi];
                                        trk_set(trk, po);
                                        grown_r++;
                                        if (qn_r < X_LCAP + 2) S.queue[qn_r++] = ko;
                                    }
                                }
                            }
                            if (lane == 0) {  // counters kept in registers over the picks of the seed
                                S.qn = qn_r;
                                S.ntaken = nt_r;
                                S.freed = freed_r;
                                S.res[XR_GROWN] = grown_r;
                            }
                        }
                        __syncthreads();
                    }
                }
            }
            cursor = first + 1;
            {
                const long long t = clock64();
                t_seed += t - tc;
                tc = t;
            }
            // no barrier here: whatever changed the state above is followed by one, and the next writes to wfirst / ev
            // come after the next iteration's barrier, which every thread reaches only after its reads of this one
        }
    }

    // ---- results
    for (int i = tid; i < p.ncap; i += X_THREADS) kpmp_g[i] = S.kpmp[i];
    for (int r = tid; r < p.P; r += X_THREADS) tracked_g[r] = trk_get(trk, r) ? 1 : 0;
    if (tid == 0) {
        S.res[XR_NKP] = n;
        S.res[XR_NEDGES] = ne;
        S.res[XR_ROUNDS] = round;
        S.res[XR_T_CHUNK] = (int)(t_chunk >> 4);
        S.res[XR_T_EVAL] = (int)(t_eval >> 4);
        S.res[XR_T_EVENT] = (int)(t_event >> 4);
        S.res[XR_T_SEED] = (int)(t_seed >> 4);
        S.res[XR_T_WEIGHTS] = (int)(t_weights >> 4);
    }
    __syncthreads();
    if (tid < XR_WORDS) p.result[f * XR_WORDS + tid] = S.res[tid];
}

size_t walk_smem(int P, int ncap, int ecap) {
    return ((sizeof(WalkShared) + 15) & ~size_t(15)) + (size_t)((P + 31) / 32) * 12 + (size_t)(ncap + 2) * 2 +
           (size_t)ecap * 8 + 16;
}

int ensure_extend(ppg_ctx* c) {
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    if (s->ext) return PPG_OK;
    if (s->ncap > 1024) return set_err(c, PPG_ERR_ARG, "extend: keypoint capacity above 1024");
    ExtendState* x = new ExtendState();
    s->ext = x;
    const size_t R = s->max_rows, N = s->ncap, B = s->bcap;
    x->ecap = c->post.lay.max_edges;
    if (x->ecap > 32767 || walk_smem(s->max_rows, s->ncap, x->ecap) > 200 * 1024)
        return set_err(c, PPG_ERR_ARG, "extend: max_edges / max_map_points too large for the walk kernel's shared memory");
    const size_t E = x->ecap;
    PPG_CUDA(c, dalloc(&x->observed, R));
    PPG_CUDA(c, dalloc(&x->bad, R));
    PPG_CUDA(c, dalloc(&x->edge_off, R + 1));
    PPG_CUDA(c, dalloc(&x->order, R));
    PPG_CUDA(c, dalloc(&x->row_hdr, R * 3));
    PPG_CUDA(c, dalloc(&x->l_idx, B * R * X_LIST));
    PPG_CUDA(c, dalloc(&x->l_d, B * R * X_LIST));
    PPG_CUDA(c, dalloc(&x->l_cnt, B * R));
    PPG_CUDA(c, dalloc(&x->tracked, B * R));
    PPG_CUDA(c, dalloc(&x->kp_mp, B * N));
    PPG_CUDA(c, dalloc(&x->kedge_me, B * E));
    PPG_CUDA(c, dalloc(&x->result, B * XR_WORDS));
    PPG_CUDA(c, dalloc(&x->g_es, E));
    PPG_CUDA(c, dalloc(&x->g_ee, E));
    PPG_CUDA(c, dalloc(&x->g_coff, N + 1));
    PPG_CUDA(c, dalloc(&x->g_cidx, 2 * E));
    PPG_CUDA(c, dalloc(&x->zero_i, R + N + 2));
    PPG_CUDA(c, cudaMemset(x->zero_i, 0, (R + N + 2) * 4));
    PPG_CUDA(c, dalloc(&x->ones_u8, R));
    PPG_CUDA(c, dalloc(&x->proj_obs, R));
    PPG_CUDA(c, dalloc(&x->gs_xy, B * N));
    PPG_CUDA(c, dalloc(&x->gs_idx, B * N));
    PPG_CUDA(c, dalloc(&x->gs_start, B * G_STRIDE));
    PPG_CUDA(c, cudaMemset(x->ones_u8, 1, R));
    PPG_CUDA(c, dalloc(&x->row_node, R));
    PPG_CUDA(c, dalloc(&x->kp_node, N));
    PPG_CUDA(c, dalloc(&x->ident, R));
    {
        std::vector<int> id(R);
        std::iota(id.begin(), id.end(), 0);
        PPG_CUDA(c, cudaMemcpy(x->ident, id.data(), R * 4, cudaMemcpyHostToDevice));
    }
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&x->h_kp_mp), B * N * 4));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&x->h_kedge_me), B * E * 4));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&x->h_result), B * XR_WORDS * 4));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&x->h_tracked), B * R));
    static bool attr_done[64];
    static std::mutex attr_mu;
    PPG_CUDA(c, once_per_device(attr_done, attr_mu, [] {
                 // four CTAs of the lists kernel per SM (64 registers, 37 KB of static shared memory each): its stalls
                 // are dependency and L2 latency with ~5 warps per scheduler (ncu), so residency is what it needs
                 cudaError_t e = cudaFuncSetAttribute(extend_lists_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                      (int)cudaSharedmemCarveoutMaxShared);
                 if (e != cudaSuccess) return e;
                 return cudaFuncSetAttribute(extend_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
             }));
    return PPG_OK;
}

FrameGraphSrc staged_graph(const ExtendState* x, int ne) {
    FrameGraphSrc g;
    g.es = reinterpret_cast<const uint8_t*>(x->g_es);
    g.ee = reinterpret_cast<const uint8_t*>(x->g_ee);
    g.coff = reinterpret_cast<const uint8_t*>(x->g_coff);
    g.cidx = reinterpret_cast<const uint8_t*>(x->g_cidx);
    g.ne = nullptr;
    g.stride = 0;
    g.ne_val = ne;
    return g;
}

FrameGraphSrc extracted_graph(const ppg_ctx* c) {
    const OutLayout& L = c->post.lay;
    const uint8_t* blk = c->d_out;
    FrameGraphSrc g;
    g.es = blk + L.edge_s;
    g.ee = blk + L.edge_e;
    g.coff = blk + L.conn_off;
    g.cidx = blk + L.conn_idx;
    g.ne = blk + L.hdr + HDR_NEDGES * sizeof(int);
    g.stride = L.total;
    g.ne_val = 0;
    return g;
}

// grid_index_kernel (spatial modes) + extend_lists_kernel for `frames` frames on the ctx stream
void launch_lists(ppg_ctx* c, ListParams lp, int frames) {
    ExtendState* x = c->assoc->ext;
    lp.gs_xy = x->gs_xy;
    lp.gs_idx = x->gs_idx;
    lp.gs_start = x->gs_start;
    if (!lp.node_mode) {
        grid_index_kernel<<<frames, 256, 0, c->st>>>(lp.src, lp.ncap, lp.kinfo, x->gs_xy, x->gs_idx, x->gs_start);
        c->launches++;
    }
    extend_lists_kernel<<<dim3((lp.nc + XL_ROWS - 1) / XL_ROWS, frames), 256, 0, c->st>>>(lp);
    c->launches++;
}

// prep + lists + walk on the ctx stream
int run_extend(ppg_ctx* c, const FrameSrc& src, const FrameGraphSrc& gsrc, int frames, int has_state) {
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    int rc = assoc_prep(c, src, frames);
    if (rc != PPG_OK) return rc;
    if (x->nc > 0) {
        ListParams lp;
        lp.nc = x->nc;
        lp.max_rows = s->max_rows;
        lp.ncap = s->ncap;
        lp.src = src;
        lp.order = x->order;
        lp.rowp = s->rowp;
        lp.map_f32 = s->map_f32;
        lp.kinfo = s->kinfo;
        lp.korder = s->korder;
        lp.l_idx = x->l_idx;
        lp.l_d = x->l_d;
        lp.l_cnt = x->l_cnt;
        lp.node_mode = 0;
        lp.row_node = lp.kp_node = nullptr;
        launch_lists(c, lp, frames);
        stage_mark(c, "extend.lists");
    }
    WalkParams wp;
    wp.nc = x->nc;
    wp.P = x->P;
    wp.max_rows = s->max_rows;
    wp.ncap = s->ncap;
    wp.ecap = x->ecap;
    wp.src = src;
    wp.gsrc = gsrc;
    wp.order = x->order;
    wp.row_hdr = x->row_hdr;
    wp.rowp = s->rowp;
    wp.map_f32 = s->map_f32;
    wp.kinfo = s->kinfo;
    wp.korder = s->korder;
    wp.l_idx = x->l_idx;
    wp.l_d = x->l_d;
    wp.l_cnt = x->l_cnt;
    wp.observed = x->observed;
    wp.bad = x->bad;
    wp.edge_ok = x->edge_ok;
    wp.edge_off = x->edge_off;
    wp.edge_other = x->edge_other;
    wp.tracked = x->tracked;
    wp.kp_mp = x->kp_mp;
    wp.kedge_me = x->kedge_me;
    wp.result = x->result;
    wp.ratio = s->ratio;
    wp.th_high = c->cfg.th_high;
    wp.has_state = has_state;
    wp.node_mode = 0;
    wp.strict = 0;
    wp.max_dist = 0.f;
    wp.row_node = wp.kp_node = nullptr;
    wp.and_rule = 0;
    wp.best_only = 0;
    extend_walk_kernel<<<frames, X_THREADS, walk_smem(x->P, s->ncap, x->ecap), c->st>>>(wp);
    c->launches++;
    stage_mark(c, "extend.walk");
    PPG_CUDA(c, cudaGetLastError());
    return PPG_OK;
}

int check_status(ppg_ctx* c, const int* res, int frames) {
    for (int f = 0; f < frames; f++)
        if (res[f * XR_WORDS + XR_STATUS] & PPG_EXTEND_OVF)
            return set_err(c, PPG_ERR_CAPACITY,
                           "ExtendMapMatches: a map point / keypoint has more edges than PPG_EXTEND_MAX_DEGREE or a "
                           "seed's weight matrix exceeds PPG_EXTEND_MAX_WEIGHTS (frame " + std::to_string(f) + ")");
    return PPG_OK;
}

void fill_out(const ExtendState* x, const AssocState* s, int f, ppg_extend_out* o) {
    const int* r = x->h_result + f * XR_WORDS;
    o->nmatches = r[XR_NMATCHES];
    o->status = (uint32_t)r[XR_STATUS];
    o->n_kp = r[XR_NKP];
    o->n_edges = r[XR_NEDGES];
    o->n_accepted = r[XR_ACCEPTED];
    o->n_grown = r[XR_GROWN];
    o->n_rescans = r[XR_RESCANS];
    for (int k = 0; k < 9; k++) o->diag[k] = r[XR_ROUNDS + k];
    if (o->kp_mp) memcpy(o->kp_mp, x->h_kp_mp + (size_t)f * s->ncap, (size_t)o->n_kp * 4);
    if (o->kedge_me) memcpy(o->kedge_me, x->h_kedge_me + (size_t)f * x->ecap, (size_t)o->n_edges * 4);
    if (o->tracked) memcpy(o->tracked, x->h_tracked + (size_t)f * x->P, (size_t)x->P);
}

int fetch_extend_async(ppg_ctx* c, int frames) {
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    PPG_CUDA(c, cudaMemcpyAsync(x->h_result, x->result, (size_t)frames * XR_WORDS * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->h_kp_mp, x->kp_mp, (size_t)frames * s->ncap * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->h_kedge_me, x->kedge_me, (size_t)frames * x->ecap * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpy2DAsync(x->h_tracked, (size_t)x->P, x->tracked, (size_t)s->max_rows, (size_t)x->P, frames,
                                  cudaMemcpyDeviceToHost, c->st));
    return PPG_OK;
}

int fetch_extend(ppg_ctx* c, int frames) {
    const int rc = fetch_extend_async(c, frames);
    if (rc != PPG_OK) return rc;
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    return PPG_OK;
}

}  // namespace

void extend_destroy(AssocState* s) {
    ExtendState* x = s->ext;
    if (!x) return;
    void* bufs[] = {x->row_hdr, x->observed, x->bad, x->edge_ok, x->edge_off, x->edge_other, x->order, x->l_idx, x->l_d, x->l_cnt,
                    x->tracked, x->kp_mp, x->kedge_me, x->result, x->g_es, x->g_ee, x->g_coff, x->g_cidx,
                    x->zero_i, x->ident, x->row_node, x->kp_node, x->ones_u8, x->proj_obs, x->gs_xy, x->gs_idx, x->gs_start};
    for (void* b : bufs)
        if (b) cudaFree(b);
    void* hbufs[] = {x->h_kp_mp, x->h_kedge_me, x->h_result, x->h_tracked};
    for (void* b : hbufs)
        if (b) cudaFreeHost(b);
    delete x;
    s->ext = nullptr;
}

}  // namespace ppg

using namespace ppg;

extern "C" {

int ppg_upload_map_graph(ppg_ctx* c, const ppg_map_graph* g) {
    if (!c || !g || g->n_points < 1 || !g->candidate || !g->observed || !g->bad || !g->edge_off)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_map_graph: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = ensure_extend(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    const int P = g->n_points;
    if (P > s->n_rows) return set_err(c, PPG_ERR_ARG, "ppg_upload_map_graph: upload the descriptors of all n_points rows first");
    const int ne = g->edge_off[P];
    if (g->edge_off[0] != 0 || ne < 0 || (ne > 0 && (!g->edge_other || !g->edge_ok)))
        return set_err(c, PPG_ERR_ARG, "ppg_upload_map_graph: malformed edge CSR");
    for (int p = 0; p < P; p++)
        if (g->edge_off[p + 1] < g->edge_off[p]) return set_err(c, PPG_ERR_ARG, "ppg_upload_map_graph: malformed edge CSR");
    for (int k = 0; k < ne; k++)
        if (g->edge_other[k] >= P) return set_err(c, PPG_ERR_ARG, "ppg_upload_map_graph: edge endpoint outside the table");
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    if (ne > x->edge_cap) {
        if (x->edge_other) cudaFree(x->edge_other);
        if (x->edge_ok) cudaFree(x->edge_ok);
        x->edge_other = nullptr;
        x->edge_ok = nullptr;
        x->edge_cap = 0;
        const size_t cap = (size_t)ne + ne / 2 + 1024;
        PPG_CUDA(c, dalloc(&x->edge_other, cap));
        PPG_CUDA(c, dalloc(&x->edge_ok, cap));
        x->edge_cap = (int)cap;
    }
    // std::sort by getEdges().size() descending over the trackable points (Matcher.cpp:210-224); ties keep table order
    std::vector<int> order;
    order.reserve(P);
    for (int p = 0; p < P; p++)
        if (g->candidate[p] && !g->bad[p]) order.push_back(p);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return g->edge_off[a + 1] - g->edge_off[a] > g->edge_off[b + 1] - g->edge_off[b];
    });
    PPG_CUDA(c, cudaMemcpyAsync(x->observed, g->observed, (size_t)P, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->bad, g->bad, (size_t)P, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->edge_off, g->edge_off, ((size_t)P + 1) * 4, cudaMemcpyHostToDevice, c->st));
    if (ne > 0) {
        PPG_CUDA(c, cudaMemcpyAsync(x->edge_other, g->edge_other, (size_t)ne * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(x->edge_ok, g->edge_ok, (size_t)ne, cudaMemcpyHostToDevice, c->st));
    }
    std::vector<int32_t> hdr(order.size() * 12, -1);
    for (size_t q = 0; q < order.size(); q++) {
        const int r = order[q], e0 = g->edge_off[r], ne_r = g->edge_off[r + 1] - e0;
        int32_t* hq = &hdr[q * 12];
        hq[0] = r;
        hq[1] = e0;
        hq[2] = ne_r;
        for (int k = 0; k < X_PRE && k < ne_r; k++) hq[3 + k] = g->edge_ok[e0 + k] ? g->edge_other[e0 + k] : -1;
    }
    if (!order.empty()) {
        PPG_CUDA(c, cudaMemcpyAsync(x->order, order.data(), order.size() * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(x->row_hdr, hdr.data(), hdr.size() * 4, cudaMemcpyHostToDevice, c->st));
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    x->P = P;
    x->nc = (int)order.size();
    return PPG_OK;
}

int ppg_extend_map_matches(ppg_ctx* c, const ppg_extend_in* in, ppg_extend_out* out) {
    if (!c || !in || !out) return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: null argument");
    if (!c->assoc || !c->assoc->ext || c->assoc->ext->P < 1)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: upload the map graph first");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    const int N = in->n_kp, E = in->n_edges, P = x->P;
    if (N < 0 || N > s->ncap || E < 0 || E > x->ecap)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: too many keypoints / key edges");
    if (!(in->ratio < 1.0f)) return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: ratio must be below 1");
    if (N > 0 && (!in->kp_x || !in->kp_y || !in->frame_desc || !in->conn_off))
        return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: null frame arrays");
    if (E > 0 && (!in->edge_start || !in->edge_end || !in->conn_idx))
        return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: null key-edge arrays");
    if (N > 0) {
        if (in->conn_off[0] != 0 || in->conn_off[N] < 0 || in->conn_off[N] > 2 * x->ecap)
            return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: malformed mvConnected CSR");
        for (int i = 0; i < N; i++)
            if (in->conn_off[i + 1] < in->conn_off[i])
                return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: malformed mvConnected CSR");
        for (int k = 0; k < in->conn_off[N]; k++) {
            const int e = in->conn_idx[k];
            if (e < 0 || e >= E) return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: mvConnected names a missing edge");
        }
        for (int e = 0; e < E; e++)
            if (in->edge_start[e] < 0 || in->edge_start[e] >= N || in->edge_end[e] < 0 || in->edge_end[e] >= N)
                return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: key edge endpoint out of range");
        if (in->kp_mp)
            for (int i = 0; i < N; i++)
                if (in->kp_mp[i] < -2 || in->kp_mp[i] >= P)
                    return set_err(c, PPG_ERR_ARG, "ppg_extend_map_matches: kp_mp row out of range");
    }
    int rc;
    if (!in->proj_uv && !in->view_cos) {
        // projections already on the device: Frame::CheckInFrustum ran there (ppg_assoc_stage_poses, one frame)
        if (!s->use_in_view || s->staged_rows != P || s->staged_frames < 1)
            return set_err(c, PPG_ERR_ARG,
                           "ppg_extend_map_matches: proj_uv / view_cos are null but no poses were staged for the "
                           "n_points rows (ppg_assoc_stage_poses)");
        s->th = in->th;
        s->ratio = in->ratio;
        s->mode = 0;
    } else if ((rc = assoc_stage_rows(c, 1, P, in->proj_uv, in->view_cos, in->th, in->ratio)) != PPG_OK) {
        return rc;
    }
    if (N > 0) {
        PPG_CUDA(c, cudaMemcpyAsync(s->kx, in->kp_x, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->ky, in->kp_y, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->fdesc, in->frame_desc, (size_t)N * 1024, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(x->g_coff, in->conn_off, ((size_t)N + 1) * 4, cudaMemcpyHostToDevice, c->st));
        if (in->conn_off[N] > 0)
            PPG_CUDA(c, cudaMemcpyAsync(x->g_cidx, in->conn_idx, (size_t)in->conn_off[N] * 4, cudaMemcpyHostToDevice, c->st));
        if (in->kp_mp)
            PPG_CUDA(c, cudaMemcpyAsync(x->kp_mp, in->kp_mp, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        else
            PPG_CUDA(c, cudaMemsetAsync(x->kp_mp, 0xff, (size_t)N * 4, c->st));
    }
    if (E > 0) {
        PPG_CUDA(c, cudaMemcpyAsync(x->g_es, in->edge_start, (size_t)E * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(x->g_ee, in->edge_end, (size_t)E * 4, cudaMemcpyHostToDevice, c->st));
        if (in->kedge_me)
            PPG_CUDA(c, cudaMemcpyAsync(x->kedge_me, in->kedge_me, (size_t)E * 4, cudaMemcpyHostToDevice, c->st));
        else
            PPG_CUDA(c, cudaMemsetAsync(x->kedge_me, 0xff, (size_t)E * 4, c->st));
    }
    if (in->tracked)
        PPG_CUDA(c, cudaMemcpyAsync(x->tracked, in->tracked, (size_t)P, cudaMemcpyHostToDevice, c->st));
    else
        PPG_CUDA(c, cudaMemsetAsync(x->tracked, 0, (size_t)P, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));  // the caller's arrays are pageable
    s->staged_n = N;
    FrameSrc src = assoc_staged_src(s);
    src.free_mask = s->ones;  // the lists hold every indexable keypoint of the window; occupancy is live in the walk
    if ((rc = run_extend(c, src, staged_graph(x, E), 1, 1)) != PPG_OK) return rc;
    if ((rc = fetch_extend(c, 1)) != PPG_OK) return rc;
    fill_out(x, s, 0, out);
    return check_status(c, x->h_result, 1);
}

int ppg_extend_run_batch(ppg_ctx* c, int n_frames) {
    if (!c || !c->assoc || !c->assoc->ext || c->assoc->ext->P < 1)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_run_batch: upload the map graph first");
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    if (n_frames < 1 || n_frames > c->maxB || n_frames > s->staged_frames || s->staged_rows != x->P)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_run_batch: stage projections of all n_points rows for every frame first");
    if (!(s->ratio < 1.0f)) return set_err(c, PPG_ERR_ARG, "ppg_extend_run_batch: ratio must be below 1");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaMemsetAsync(x->kedge_me, 0xff, (size_t)n_frames * x->ecap * 4, c->st));
    s->mode = 0;
    return run_extend(c, assoc_extracted_src(c, 0), extracted_graph(c), n_frames, 0);
}

int ppg_extend_fetch_batch(ppg_ctx* c, int n_frames, ppg_extend_out* outs) {
    if (!c || !c->assoc || !c->assoc->ext || !outs || n_frames < 1 || n_frames > c->assoc->bcap)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_fetch_batch: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = fetch_extend(c, n_frames);
    if (rc != PPG_OK) return rc;
    for (int f = 0; f < n_frames; f++) fill_out(c->assoc->ext, c->assoc, f, &outs[f]);
    return check_status(c, c->assoc->ext->h_result, n_frames);
}

int ppg_extend_fetch_batch_async(ppg_ctx* c, int n_frames) {
    if (!c || !c->assoc || !c->assoc->ext || n_frames < 1 || n_frames > c->assoc->bcap)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_fetch_batch_async: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    return fetch_extend_async(c, n_frames);
}

int ppg_extend_collect(ppg_ctx* c, int n_frames, ppg_extend_out* outs) {
    if (!c || !c->assoc || !c->assoc->ext || !outs || n_frames < 1 || n_frames > c->assoc->bcap)
        return set_err(c, PPG_ERR_ARG, "ppg_extend_collect: bad arguments");
    for (int f = 0; f < n_frames; f++) fill_out(c->assoc->ext, c->assoc, f, &outs[f]);
    return check_status(c, c->assoc->ext->h_result, n_frames);
}

// Matcher::SearchByBoW whole (Matcher.cpp:393-477 / :663-754) through the same two kernels: lists over the features of
// the row's vocabulary node, walk with the live vpMapPointMatches / vbMatched2 state, no map graph.
int ppg_search_by_bow(ppg_ctx* c, const ppg_bow_match_in* in, ppg_bow_match_out* out) {
    if (!c || !in || !out || !in->row_node) return set_err(c, PPG_ERR_ARG, "ppg_search_by_bow: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = ensure_extend(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    const int M = in->n_rows, N = in->n_kp;
    if (M < 1 || M > s->n_rows) return set_err(c, PPG_ERR_ARG, "ppg_search_by_bow: upload the n_rows descriptors first");
    if (N < 0 || N > s->ncap || (N > 0 && (!in->frame_desc || !in->kp_node)))
        return set_err(c, PPG_ERR_ARG, "ppg_search_by_bow: bad frame arrays");
    int last = -1;
    for (int m = 0; m < M; m++) {  // the reference walks the FeatureVector nodes in ascending order
        if (in->row_node[m] < 0) continue;
        if (in->row_node[m] < last) return set_err(c, PPG_ERR_ARG, "ppg_search_by_bow: rows must be ordered by node");
        last = in->row_node[m];
    }
    PPG_CUDA(c, cudaMemcpyAsync(x->row_node, in->row_node, (size_t)M * 4, cudaMemcpyHostToDevice, c->st));
    if (N > 0) {
        PPG_CUDA(c, cudaMemcpyAsync(x->kp_node, in->kp_node, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->fdesc, in->frame_desc, (size_t)N * 1024, cudaMemcpyHostToDevice, c->st));
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));  // pageable sources
    s->staged_n = N;
    FrameSrc src = assoc_staged_src(s);
    src.free_mask = s->ones;
    ListParams lp{};
    lp.nc = M;
    lp.max_rows = s->max_rows;
    lp.ncap = s->ncap;
    lp.src = src;
    lp.order = x->ident;
    lp.map_f32 = s->map_f32;
    lp.l_idx = x->l_idx;
    lp.l_d = x->l_d;
    lp.l_cnt = x->l_cnt;
    lp.node_mode = 1;
    lp.row_node = x->row_node;
    lp.kp_node = x->kp_node;
    launch_lists(c, lp, 1);
    stage_mark(c, "bow_match.lists");
    WalkParams wp{};
    wp.nc = M;
    wp.P = M;
    wp.max_rows = s->max_rows;
    wp.ncap = s->ncap;
    wp.ecap = x->ecap;
    wp.src = src;
    FrameGraphSrc g{};
    g.coff = reinterpret_cast<const uint8_t*>(x->zero_i);  // no key edges
    g.es = g.ee = g.cidx = g.coff;
    g.ne_val = 0;
    wp.gsrc = g;
    wp.order = x->ident;
    wp.map_f32 = s->map_f32;
    wp.korder = s->korder;
    wp.l_idx = x->l_idx;
    wp.l_d = x->l_d;
    wp.l_cnt = x->l_cnt;
    wp.observed = x->ones_u8;
    wp.bad = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_ok = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_off = x->zero_i;
    wp.edge_other = x->zero_i;
    wp.tracked = x->tracked;
    wp.kp_mp = x->kp_mp;
    wp.kedge_me = x->kedge_me;
    wp.result = x->result;
    wp.ratio = in->ratio;
    wp.th_high = c->cfg.th_high;
    wp.has_state = 0;
    wp.node_mode = 1;
    wp.strict = in->strict;
    wp.max_dist = in->max_dist;
    wp.row_node = x->row_node;
    wp.kp_node = x->kp_node;
    extend_walk_kernel<<<1, X_THREADS, walk_smem(M, s->ncap, x->ecap), c->st>>>(wp);
    stage_mark(c, "bow_match.walk");
    c->launches += 1;  // the walk (launch_lists counts its own)
    PPG_CUDA(c, cudaGetLastError());
    PPG_CUDA(c, cudaMemcpyAsync(x->h_result, x->result, XR_WORDS * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->h_kp_mp, x->kp_mp, (size_t)s->ncap * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    out->nmatches = x->h_result[XR_ACCEPTED];
    out->n_rescans = x->h_result[XR_RESCANS];
    if (out->kp_row && N > 0) memcpy(out->kp_row, x->h_kp_mp, (size_t)N * 4);
    return PPG_OK;
}

// Matcher::SearchForInitialization whole (Matcher.cpp:582-651) through the same two kernels: the rows are the features
// of F1 in index order (their descriptors are the resident table: ppg_upload_map), the window of a row is the radius-
// windowSize box around vbPrevMatched[i1] in F2 (prep_rows_kernel, PPG_SEARCH_WINDOW), the walk keeps vnMatches21 live
// (a matched F2 feature is never offered again: the vector<int> quirk, see the oracle) and accepts with
// best <= TH_LOW && best < ratio * second.
int ppg_search_for_initialization(ppg_ctx* c, const ppg_init_match_in* in, ppg_init_match_out* out) {
    if (!c || !in || !out || !in->prev_matched) return set_err(c, PPG_ERR_ARG, "ppg_search_for_initialization: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = ensure_extend(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    const int M = in->n1, N = in->n2;
    if (M < 1 || M > s->n_rows) return set_err(c, PPG_ERR_ARG, "ppg_search_for_initialization: upload the n1 descriptors of F1 first (ppg_upload_map)");
    if (N < 0 || N > s->ncap || (N > 0 && (!in->kp2_x || !in->kp2_y || !in->desc2)))
        return set_err(c, PPG_ERR_ARG, "ppg_search_for_initialization: bad F2 arrays");
    if (in->window < 0) return set_err(c, PPG_ERR_ARG, "ppg_search_for_initialization: negative window");
    std::vector<float> zeros((size_t)M, 0.f);
    if ((rc = assoc_stage_rows(c, 1, M, in->prev_matched, zeros.data(), (float)in->window, in->ratio)) != PPG_OK) return rc;
    s->mode = PPG_SEARCH_WINDOW;  // r = th for every row (GetFeaturesInArea(x, y, windowSize), :596)
    if (N > 0) {
        PPG_CUDA(c, cudaMemcpyAsync(s->kx, in->kp2_x, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->ky, in->kp2_y, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->fdesc, in->desc2, (size_t)N * 1024, cudaMemcpyHostToDevice, c->st));
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));  // pageable sources
    s->staged_n = N;
    FrameSrc src = assoc_staged_src(s);
    src.free_mask = s->ones;
    if ((rc = assoc_prep(c, src, 1)) != PPG_OK) return rc;
    ListParams lp{};
    lp.nc = M;
    lp.max_rows = s->max_rows;
    lp.ncap = s->ncap;
    lp.src = src;
    lp.order = x->ident;
    lp.rowp = s->rowp;
    lp.map_f32 = s->map_f32;
    lp.kinfo = s->kinfo;
    lp.korder = s->korder;
    lp.l_idx = x->l_idx;
    lp.l_d = x->l_d;
    lp.l_cnt = x->l_cnt;
    lp.node_mode = 0;
    launch_lists(c, lp, 1);
    stage_mark(c, "init_match.lists");
    WalkParams wp{};
    wp.nc = M;
    wp.P = M;
    wp.max_rows = s->max_rows;
    wp.ncap = s->ncap;
    wp.ecap = x->ecap;
    wp.src = src;
    FrameGraphSrc g{};
    g.coff = reinterpret_cast<const uint8_t*>(x->zero_i);  // no key edges
    g.es = g.ee = g.cidx = g.coff;
    g.ne_val = 0;
    wp.gsrc = g;
    wp.order = x->ident;
    wp.rowp = s->rowp;
    wp.map_f32 = s->map_f32;
    wp.kinfo = s->kinfo;
    wp.korder = s->korder;
    wp.l_idx = x->l_idx;
    wp.l_d = x->l_d;
    wp.l_cnt = x->l_cnt;
    wp.observed = x->ones_u8;  // every match takes its feature
    wp.bad = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_ok = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_off = x->zero_i;
    wp.edge_other = x->zero_i;
    wp.tracked = x->tracked;
    wp.kp_mp = x->kp_mp;
    wp.kedge_me = x->kedge_me;
    wp.result = x->result;
    wp.ratio = in->ratio;
    wp.th_high = c->cfg.th_high;
    wp.has_state = 0;
    wp.node_mode = 0;
    wp.and_rule = 1;
    wp.best_only = 0;
    wp.strict = 0;
    wp.max_dist = c->cfg.th_low;
    extend_walk_kernel<<<1, X_THREADS, walk_smem(M, s->ncap, x->ecap), c->st>>>(wp);
    stage_mark(c, "init_match.walk");
    c->launches += 1;
    PPG_CUDA(c, cudaGetLastError());
    PPG_CUDA(c, cudaMemcpyAsync(x->h_result, x->result, XR_WORDS * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->h_kp_mp, x->kp_mp, (size_t)s->ncap * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    out->nmatches = x->h_result[XR_ACCEPTED];
    out->n_rescans = x->h_result[XR_RESCANS];
    if (out->matches12) {
        for (int i = 0; i < M; i++) out->matches12[i] = -1;
        for (int j = 0; j < N; j++)
            if (x->h_kp_mp[j] >= 0 && x->h_kp_mp[j] < M) out->matches12[x->h_kp_mp[j]] = j;  // vnMatches21 -> vnMatches12
    }
    if (out->prev_matched) {  // :644-647
        if (out->prev_matched != in->prev_matched) memcpy(out->prev_matched, in->prev_matched, (size_t)M * 8);
        for (int j = 0; j < N; j++)
            if (x->h_kp_mp[j] >= 0 && x->h_kp_mp[j] < M) {
                out->prev_matched[2 * x->h_kp_mp[j]] = in->kp2_x[j];
                out->prev_matched[2 * x->h_kp_mp[j] + 1] = in->kp2_y[j];
            }
    }
    return PPG_OK;
}

// Matcher::SearchByProjection(CurrentFrame, LastFrame, th) (Matcher.cpp:31-87) and SearchByProjection(CurrentFrame, pKF,
// sAlreadyFound, th, descDist) (:1337-1411) whole through the same two kernels: the rows are the projected map points in
// loop order (their descriptors are the resident table), the window of a row is the radius-th box around its projection
// (prep_rows_kernel, PPG_SEARCH_WINDOW), the walk keeps CurrentFrame.mvpMapPoints live and accepts with best <= max_dist.
int ppg_search_by_projection(ppg_ctx* c, const ppg_projection_match_in* in, ppg_projection_match_out* out) {
    if (!c || !in || !out || !out->kp_mp) return set_err(c, PPG_ERR_ARG, "ppg_search_by_projection: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = ensure_extend(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    ExtendState* x = s->ext;
    const int M = in->n_rows, N = in->n;
    if (N < 0 || N > s->ncap || (N > 0 && (!in->kp_x || !in->kp_y || !in->desc)))
        return set_err(c, PPG_ERR_ARG, "ppg_search_by_projection: bad keypoint arrays");
    if (in->kp_mp)
        for (int i = 0; i < N; i++)
            if (in->kp_mp[i] < -2 || in->kp_mp[i] >= M)
                return set_err(c, PPG_ERR_ARG, "ppg_search_by_projection: kp_mp row out of range");
    out->nmatches = out->n_rescans = 0;
    if (M == 0 || N == 0) {  // the reference's loop body never reaches a window
        for (int i = 0; i < N; i++) out->kp_mp[i] = in->kp_mp ? in->kp_mp[i] : -1;
        return PPG_OK;
    }
    if (M < 0 || M > s->n_rows || !in->proj_uv)
        return set_err(c, PPG_ERR_ARG, "ppg_search_by_projection: upload the n_rows descriptors first (ppg_upload_map)");
    if (!(in->th >= 0.f)) return set_err(c, PPG_ERR_ARG, "ppg_search_by_projection: negative window");
    std::vector<float> zeros((size_t)M, 0.f);
    if ((rc = assoc_stage_rows(c, 1, M, in->proj_uv, zeros.data(), in->th, 1.f)) != PPG_OK) return rc;
    s->mode = PPG_SEARCH_WINDOW;  // r = th for every row (GetFeaturesInArea(u, v, th), :58)
    PPG_CUDA(c, cudaMemcpyAsync(s->kx, in->kp_x, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(s->ky, in->kp_y, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(s->fdesc, in->desc, (size_t)N * 1024, cudaMemcpyHostToDevice, c->st));
    if (in->kp_mp)
        PPG_CUDA(c, cudaMemcpyAsync(x->kp_mp, in->kp_mp, (size_t)N * 4, cudaMemcpyHostToDevice, c->st));
    else
        PPG_CUDA(c, cudaMemsetAsync(x->kp_mp, 0xff, (size_t)N * 4, c->st));
    PPG_CUDA(c, cudaMemsetAsync(x->tracked, 0, (size_t)M, c->st));
    const uint8_t* obs = x->ones_u8;
    if (in->observed) {  // per-row Observations() > 0: borrowed from the map-graph flags of this ctx
        PPG_CUDA(c, cudaMemcpyAsync(x->proj_obs, in->observed, (size_t)M, cudaMemcpyHostToDevice, c->st));
        obs = x->proj_obs;
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));  // pageable sources
    s->staged_n = N;
    FrameSrc src = assoc_staged_src(s);
    src.free_mask = s->ones;
    if ((rc = assoc_prep(c, src, 1)) != PPG_OK) return rc;
    ListParams lp{};
    lp.nc = M;
    lp.max_rows = s->max_rows;
    lp.ncap = s->ncap;
    lp.src = src;
    lp.order = x->ident;
    lp.rowp = s->rowp;
    lp.map_f32 = s->map_f32;
    lp.kinfo = s->kinfo;
    lp.korder = s->korder;
    lp.l_idx = x->l_idx;
    lp.l_d = x->l_d;
    lp.l_cnt = x->l_cnt;
    lp.node_mode = 0;
    launch_lists(c, lp, 1);
    stage_mark(c, "proj_match.lists");
    WalkParams wp{};
    wp.nc = M;
    wp.P = M;
    wp.max_rows = s->max_rows;
    wp.ncap = s->ncap;
    wp.ecap = x->ecap;
    wp.src = src;
    FrameGraphSrc g{};
    g.coff = reinterpret_cast<const uint8_t*>(x->zero_i);  // no key edges
    g.es = g.ee = g.cidx = g.coff;
    g.ne_val = 0;
    wp.gsrc = g;
    wp.order = x->ident;
    wp.rowp = s->rowp;
    wp.map_f32 = s->map_f32;
    wp.kinfo = s->kinfo;
    wp.korder = s->korder;
    wp.l_idx = x->l_idx;
    wp.l_d = x->l_d;
    wp.l_cnt = x->l_cnt;
    wp.observed = obs;
    wp.bad = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_ok = reinterpret_cast<const uint8_t*>(x->zero_i);
    wp.edge_off = x->zero_i;
    wp.edge_other = x->zero_i;
    wp.tracked = x->tracked;
    wp.kp_mp = x->kp_mp;
    wp.kedge_me = x->kedge_me;
    wp.result = x->result;
    wp.ratio = 1.f;
    wp.th_high = c->cfg.th_high;
    wp.has_state = 1;
    wp.best_only = 1;
    wp.max_dist = in->max_dist;
    extend_walk_kernel<<<1, X_THREADS, walk_smem(M, s->ncap, x->ecap), c->st>>>(wp);
    stage_mark(c, "proj_match.walk");
    c->launches += 1;
    PPG_CUDA(c, cudaGetLastError());
    PPG_CUDA(c, cudaMemcpyAsync(x->h_result, x->result, XR_WORDS * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(x->h_kp_mp, x->kp_mp, (size_t)N * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    out->nmatches = x->h_result[XR_ACCEPTED];
    out->n_rescans = x->h_result[XR_RESCANS];
    memcpy(out->kp_mp, x->h_kp_mp, (size_t)N * 4);
    return PPG_OK;
}

}  // extern "C"
