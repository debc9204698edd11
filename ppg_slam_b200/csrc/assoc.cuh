// Shared by assoc.cu (search core) and extend.cu (whole Matcher::ExtendMapMatches): association state of a ctx,
// the window test of Frame::GetFeaturesInArea and DescriptorDistance in the summation order fixed with the oracle.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ctx.cuh"

namespace ppg {

struct ExtendState;

bool make_kmajor_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16);

constexpr int A_BM = 128, A_BN = 128, A_STAGES = 4, A_STAGE_BYTES = 32768, A_THREADS = 320, A_TOPK = 4;
constexpr int A_QCAP = 32;  // per-thread queue of window hits awaiting the exact mask + top-4 insertion
constexpr unsigned AFULL = 0xffffffffu;

struct RowParam {  // per map point
    float u, v, r, na2;
    uint32_t cells;  // minCx | maxCx << 8 | minCy << 16 | maxCy << 24 ; 0xffffffff = empty window
};

// Where the keypoints / descriptors of frame f live: staged arrays (one frame) or the extraction output blocks
// of the last batch (byte stride = one output block).
struct FrameSrc {
    const uint8_t *kx, *ky, *desc, *free_mask, *n;
    size_t stride, free_stride;
    int n_val;  // used when n == nullptr
    __host__ __device__ const float* kx_of(int f) const { return reinterpret_cast<const float*>(kx + f * stride); }
    __host__ __device__ const float* ky_of(int f) const { return reinterpret_cast<const float*>(ky + f * stride); }
    __host__ __device__ const float* desc_of(int f) const { return reinterpret_cast<const float*>(desc + f * stride); }
    __host__ __device__ const uint8_t* free_of(int f) const { return free_mask + f * free_stride; }
    __device__ int n_of(int f) const { return n ? *reinterpret_cast<const int*>(n + f * stride) : n_val; }
};

struct AssocState {
    int max_rows = 0, n_rows = 0, ncap = 0;  // ncap: keypoint capacity per frame (multiple of 128)
    int bcap = 1;                            // frames per batched call
    // map side (shared by all frames)
    float* map_f32 = nullptr;
    __nv_bfloat16* map_bf = nullptr;
    float* map_n2 = nullptr;
    CUtensorMap mapA, mapB;
    // frame side: staged single frame, or the extraction output blocks of the last batch
    float *kx = nullptr, *ky = nullptr, *fdesc = nullptr;
    uint8_t *free_mask = nullptr, *ones = nullptr;
    float* fn2 = nullptr;            // [bcap][ncap]
    __nv_bfloat16* f_bf = nullptr;   // [bcap][ncap][256]
    uint32_t* kinfo = nullptr;       // [bcap][ncap]  cx | cy << 8 | ok << 16
    uint32_t* korder = nullptr;      // [bcap][ncap]  (cx*48+cy) << 16 | i : GetFeaturesInArea visiting order
    float* nbmax = nullptr;          // [bcap] max squared norm of the frame descriptors (float bits, atomicMax)
    int staged_n = 0;
    // row side, [bcap][max_rows]
    float *proj = nullptr, *vcos = nullptr;
    RowParam* rowp = nullptr;
    float th = 0.f, ratio = 0.f;
    int mode = 0;            // PPG_SEARCH_EXTEND_MAP / PPG_SEARCH_WINDOW
    float max_dist = 0.f;    // mode 1 acceptance threshold
    double e2_max = 0.0;     // mode 1 circular limit (Fuse), 0 = none
    int staged_rows = 0, staged_frames = 0;
    // results, [bcap][max_rows]
    int* cand = nullptr;  // x4
    float* guard = nullptr;
    int *best_idx = nullptr, *second_idx = nullptr;
    float *best_d = nullptr, *second_d = nullptr;
    uint8_t* accept = nullptr;
    int* fallback = nullptr;
    uint8_t* h_res = nullptr;  // pinned staging for fetch: 5 planes [bcap][max_rows] (4 x 4 bytes, 1 x 1 byte)
    size_t h_res_bytes = 0;
    // Frame::CheckInFrustum on the device (ppg_upload_map_geometry / ppg_assoc_stage_poses): map geometry, poses,
    // mbTrackInView / mTrackDepth per (frame, row); proj and vcos above are then written by frustum_kernel
    int *row_node = nullptr, *kp_node = nullptr;  // PPG_SEARCH_NODE: vocabulary node of every row / keypoint
    float *wpos = nullptr, *nrm = nullptr, *dmin = nullptr, *dmax = nullptr;  // [max_rows] x 3, 3, 1, 1
    float* poses = nullptr;                                                   // [bcap][16]: Rcw 9, tcw 3, Ow 3
    uint8_t* in_view = nullptr;                                               // [bcap][max_rows]
    float* depth = nullptr;                                                   // [bcap][max_rows]
    int geo_rows = 0;
    bool use_in_view = false;  // rows staged through poses: out-of-view rows have no search window
    ExtendState* ext = nullptr;  // whole ExtendMapMatches (extend.cu), allocated on first use
    float* h_stage = nullptr;  // pinned staging for the per-frame projections: [bcap][max_rows] x (2 + 1) floats
};

struct GridParam {
    int minX, minY;
    float wInv, hInv;
};

// e2_max > 0 (Fuse, Matcher.cpp:1000-1005): candidates farther than sqrt(e2_max) from the projection are skipped;
// float e2 compared with the double literal, as the reference does.
__device__ __forceinline__ bool in_window(const RowParam& p, uint32_t info, float x, float y, double e2_max) {
    if (!(info & 0x10000u) || p.cells == 0xffffffffu) return false;
    const uint32_t cx = info & 0xff, cy = (info >> 8) & 0xff;
    if (cx < (p.cells & 0xff) || cx > ((p.cells >> 8) & 0xff) || cy < ((p.cells >> 16) & 0xff) || cy > (p.cells >> 24))
        return false;
    if (!(fabsf(x - p.u) < p.r && fabsf(y - p.v) < p.r)) return false;  // Frame.cpp:305-309
    if (e2_max > 0.0) {
        const float ex = p.u - x, ey = p.v - y;
        const float e2 = ex * ex + ey * ey;
        if ((double)e2 > e2_max) return false;
    }
    return true;
}

// DescriptorDistance in the fixed order shared with the oracle (ppgo_descriptor_distance): lane l sums
// elements l, l+32, ... in order, then an xor butterfly 16,8,4,2,1; every lane ends with the same value.
// `av` = the map row's elements lane, lane + 32, ... held in registers across the candidates of a row.
__device__ __forceinline__ float exact_distance(const float (&av)[8], const float* __restrict__ b, int lane) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float d = av[k] - b[lane + 32 * k];
        s = s + d * d;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s = s + __shfl_xor_sync(AFULL, s, m);
    return sqrtf(s);
}

// best = first minimum, second = first minimum of the rest, both in GetFeaturesInArea order
// (equivalent to the strict-< update at Matcher.cpp:262-271).
__device__ __forceinline__ void top2_update(float d, uint32_t ord, int idx, float& b1, uint32_t& o1, int& i1, float& b2,
                                            uint32_t& o2, int& i2) {
    if (d < b1 || (d == b1 && ord < o1)) {
        b2 = b1; o2 = o1; i2 = i1;
        b1 = d; o1 = ord; i1 = idx;
    } else if (d < b2 || (d == b2 && ord < o2)) {
        b2 = d; o2 = ord; i2 = idx;
    }
}

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

// host side, assoc.cu
int assoc_ensure_state(ppg_ctx* c);
FrameSrc assoc_staged_src(const AssocState* s);
FrameSrc assoc_extracted_src(const ppg_ctx* c, int first);
int assoc_stage_rows(ppg_ctx* c, int frames, int n_rows, const float* proj_uv, const float* view_cos, float th,
                      float ratio, bool pinned_src = false);
// prep_frame_kernel + prep_rows_kernel for `frames` frames on the ctx stream (kinfo / korder / rowp)
int assoc_prep(ppg_ctx* c, const FrameSrc& src, int frames);
void extend_destroy(AssocState* s);  // extend.cu

}  // namespace ppg
