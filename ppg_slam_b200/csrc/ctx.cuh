// Internal context of libppg_b200.so (shared by api.cu and assoc.cu).
#pragma once
#include <map>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/ppg_b200.h"
#include "conv_tc.cuh"
#include "post.cuh"

namespace ppg {

struct TcLayerInfo {
    const char* name;
    ConvLayer L;
    const __half* in;   // NHWC fp16 input
    __half* w;          // [taps][N][Cin] fp16
    float* bias;        // [N]
    int H, W, cin, cout, N, taps, mode, relu;
    void* out;
    int out_ld;
};

struct AssocState;  // assoc.cu
struct BowState;    // bow.cu
struct CommState;   // comm.cu

}  // namespace ppg

struct ppg_ctx {
    ppg_config cfg;
    std::string weights_path;
    mutable std::string err;
    int dev = 0, num_sms = 0;
    cudaStream_t st = nullptr;
    // small batches: the three heads and the post-processing branches that do not depend on one another run on side
    // streams (fork / join with events, captured into the CUDA graph as parallel branches)
    cudaStream_t st2 = nullptr, st3 = nullptr;
    cudaEvent_t ev_feat = nullptr, ev_heat = nullptr, ev_dmap = nullptr, ev_kp = nullptr, ev_desc = nullptr;
    bool use_fork = true;  // PPG_FORK=0 disables
    int H = 0, W = 0, Hc = 0, Wc = 0, maxB = 0;
    long long launches = 0;
    int last_batch = 0;

    // weights
    float *w1a = nullptr, *b1a = nullptr;                    // conv1a fp32
    float *we3 = nullptr, *be3 = nullptr, *we1 = nullptr, *be1b = nullptr;  // edge tail
    std::vector<ppg::TcLayerInfo> tc;

    // activations (NHWC fp16 unless noted)
    uint8_t* gray = nullptr;
    uint8_t* h_gray = nullptr;  // pinned staging
    __half *a1 = nullptr, *a2 = nullptr, *a3 = nullptr, *a4 = nullptr, *a5 = nullptr, *a6 = nullptr, *a7 = nullptr,
           *feat = nullptr, *p1 = nullptr, *d1 = nullptr, *e1 = nullptr, *e2 = nullptr;
    float *jlogits = nullptr, *desc = nullptr;  // fp32 NHWC
    float *prob = nullptr, *heat_raw = nullptr, *heat_ref = nullptr, *heat_final = nullptr;
    float *prob_in = nullptr, *heat_in = nullptr, *desc_in = nullptr;  // ppg_extract_from_maps inputs
    bool maps_from_caller = false;

    // post-processing
    ppg::PostParams post;
    float2* undist_lut = nullptr;
    int2* remap_lut = nullptr;
    uint8_t* d_out = nullptr;
    uint8_t* h_out = nullptr;  // pinned mirror
    // image bounds / grid of Frame (GeometricCamera.cpp:26-61)
    int minX = 0, minY = 0, maxX = 0, maxY = 0;
    float wInv = 0.f, hInv = 0.f;

    // profiling
    bool profiling = false;
    // CUDA graphs of the launch sequence of ppg_run, one per batch size (captured on the second call with that size;
    // PPG_GRAPH=0 disables).  Profiling runs bypass them (the stage events are not part of the graph).
    bool use_graph = true;
    bool graph_large = false;  // PPG_GRAPH_LARGE=1: graphs for batches > 8 as well
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int> graph_launches;  // kernels per replay, for launch_count
    std::map<int, int> run_calls;
    bool fuse_scan = false;   // PPG_FUSE_SCAN=1: threshold scan inside convPb's epilogue (measured: no gain)
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> ev_names;
    int n_ev = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;

    ppg::AssocState* assoc = nullptr;
    ppg::BowState* bow = nullptr;
    ppg::CommState* comm = nullptr;  // NCCL communicator of the row-sharded association (comm.cu)
};

namespace ppg {
int set_err(const ppg_ctx* c, int code, const std::string& msg);
int cuda_fail(const ppg_ctx* c, cudaError_t e, const char* what);
void assoc_destroy(ppg_ctx* c);
void bow_destroy(ppg_ctx* c);
void comm_destroy(ppg_ctx* c);
// profiling runs: records a CUDA event named after the stage that just ended on the ctx stream (api.cu)
void stage_mark(ppg_ctx* c, const char* name);
}  // namespace ppg

#define PPG_CUDA(c, call)                                               \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ppg::cuda_fail((c), e__, #call); \
    } while (0)
