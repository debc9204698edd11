// Post-processing of the dense network outputs: keypoints, heat-map refine/remap, point-pair graph,
// descriptors.  CUDA restatement of feature/src/PPGExtractor.cpp:158-589; kernels in post.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppg {

// Per-frame output block (device + pinned host mirror), offsets in bytes from the block start.
struct OutLayout {
    int max_kp, max_edges, max_col;
    size_t hdr, kp_x, kp_y, px, py, score, xun, yun, kout, edge_s, edge_e, edge_score, conn_off, conn_idx, col_off,
        col_pairs, desc, total;
    size_t small_total;  // everything before desc (copied separately from the descriptors)
};
OutLayout make_out_layout(int max_kp, int max_edges, int max_col);

enum HdrField {
    HDR_NKP = 0,
    HDR_NEDGES = 1,
    HDR_NCOL = 2,
    HDR_STATUS = 3,
    HDR_NCAND = 4,
    HDR_NACC = 5,
    HDR_NPASS = 6,
    HDR_NLINES = 7,
    HDR_NMS_ROUNDS = 8,
    HDR_DIAG = 9,  // .. 15: phase clocks of lines_kernel
    HDR_WORDS = 16
};

enum FrameStatus { ST_OVF_ACCEPT = 1, ST_OVF_PAIRS = 2, ST_OVF_DEGREE = 4, ST_OVF_EDGES = 8, ST_OVF_COLINE = 16 };

constexpr int POST_MAX_KP = 1024;  // hard upper bound on junction_max_num

struct PostParams {
    int B, H, W, Hc, Wc;
    int fisheye, do_remap;
    float junction_thresh;
    int nms_radius, max_kp;
    float line_valid_thresh, line_valid_ratio, line_dist_thresh, line_heatmap_thresh, line_inlier_rate;
    float inv_scale;
    // dense maps, [B][H*W] unless noted
    const float* prob;
    const float* heat_raw;
    float* heat_ref;    // refined
    float* heat_final;  // after remap (== heat_ref when !do_remap)
    const float* desc;  // [B][Hc][Wc][256] fp32 NHWC
    // init-time tables
    const float2* undist_lut;  // [H*W] (xun, yun) of every integer pixel
    const int2* remap_lut;     // [H*W] (iy * W + ix, fx | fy << 5 | inside bits << 10) of rint(map*32), see remap_kernel
    // scratch, per frame
    uint8_t* state;     // [B][H*W]   0 none, 1 undecided candidate, 2 accepted, 3 suppressed
    uint8_t* state2;    // [B][H*W/4] the same, 2 bits per pixel (shared-memory NMS variant)
    int nms_smem;       // 1: state2 + nms_smem_kernel, 0: state + nms_global_kernel
    int scan_fused;     // 1: candidates / state / counters were written by the junction head's epilogue (conv_tc.cu)
    uint32_t* cand;     // [B][H*W]   pixel indices of in-border candidates (unordered)
    int* counters;      // [B][8]     0: candidates in border, 1: pixels >= threshold, 2: inter_pool entries used
    int acc_cap;        // NMS survivors that can be ranked (power of two)
    uint32_t* pair_bits;  // [B][max_kp][pair_words] bit j of row i: pair (i,j) passed the 3-point test
    int pair_words;
    uint32_t* sym_bits;   // [B][max_kp][pair_words] the same matrix made symmetric (both endpoints' rows)
    int* row_cnt;       // [B][max_kp]
    uint16_t* row_prefix;  // [B][max_kp][pair_words] set bits of the row before each word
    int* row_off;       // [B][max_kp + 1] candidate id of the first pair of each row
    int pair_cap, deg_cap;
    uint32_t* c_se;     // [B][pair_cap] candidate lines: s | e << 16
    float *c_dist, *c_dirf, *c_dirb;
    uint16_t* inter;    // [B][pair_cap][32] overlap-filter interactions (16 per endpoint)
    uint32_t* inter_cnt;  // [B][pair_cap] count at s | count at e << 16
    uint32_t* inter_off;  // [B][pair_cap] offset of a spilled list in inter_pool, ~0 = inline
    uint16_t* inter_pool; // [B][pool_cap] spilled lists (s side then e side); counters[b][2] = used
    int pool_cap;
    uint32_t* alive_g;  // [B][max_kp][pair_words] symmetric matrix of the lines alive after the overlap filter
    float* l_score;     // [B][pair_cap]
    int* l_edge;        // [B][pair_cap]
    uint8_t* out;       // [B][lay.total]
    OutLayout lay;
};

size_t post_nms_smem(const PostParams& p);
void post_plan_nms(PostParams& p);  // sets nms_smem and acc_cap from H, W
size_t post_lines_smem(const PostParams& p);         // lines_graph_kernel
size_t post_lines_filter_smem(const PostParams& p);  // lines_filter_kernel
size_t post_lines_fixed_smem(int max_kp, int pair_words);
cudaError_t post_init_attrs(const PostParams& p);
// Optional per-kernel stage marks of a profiling run (api.cu records a CUDA event after each named kernel).
struct PostMark {
    void (*fn)(void* user, const char* name) = nullptr;
    void* user = nullptr;
    void operator()(const char* name) const {
        if (fn) fn(user, name);
    }
};
// each returns the number of kernels it launched through *launches (added)
cudaError_t post_keypoints_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark = {});  // scan + NMS + top-k
cudaError_t post_heat_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark = {});       // refine (+ remap)
cudaError_t post_lines_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark = {});      // point-pair graph
cudaError_t post_desc_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark = {});       // sampling + L2 norm

}  // namespace ppg
