// Matcher::SearchForTriangulation (matching/src/Matcher.cpp:767-885) for the pinhole camera on the GPU.
//
// The reference walks the FeatureVectors of two key frames node by node and, for every feature of KF1 without a map
// point, keeps the best feature of KF2 under the same node that passes the descriptor threshold, the epipole exclusion and
// the epipolar test (sensors/src/Pinhole.cpp:98-114).  Its "already matched" flag vbMatched2 is never set, so the
// features of KF1 do not interact: one warp per feature, the lanes share one exact DescriptorDistance per candidate (the
// fixed summation order of assoc.cuh, as in every other matcher here) and evaluate the scalar tests redundantly.
// Compiled with -fmad=false: a, b, c, num, den are float expressions evaluated left to right without contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "assoc.cuh"
#include "ctx.cuh"

namespace ppg {

namespace {

struct TriangParams {
    int n1, n2;
    const float *desc1, *desc2, *pos1, *pos2;
    const int *node1, *node2;
    const uint8_t *mp1, *mp2;
    float F[9], ep0, ep1, th_low;
    int* match12;
};

__global__ void __launch_bounds__(256) triangulation_match_kernel(const TriangParams p) {
    const int lane = threadIdx.x & 31;
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i1 >= p.n1) return;
    const int nd = p.node1[i1];
    int best_idx = -1;
    if (!p.mp1[i1] && nd >= 0) {  // :812-818
        float av[8];
#pragma unroll
        for (int k = 0; k < 8; k++) av[k] = p.desc1[(size_t)i1 * 256 + lane + 32 * k];
        const float x1 = p.pos1[2 * i1], y1 = p.pos1[2 * i1 + 1];
        // epipolar line of kp1 in the second image, Pinhole.cpp:106-108
        const float a = x1 * p.F[0] + y1 * p.F[3] + p.F[6];
        const float b = x1 * p.F[1] + y1 * p.F[4] + p.F[7];
        const float c = x1 * p.F[2] + y1 * p.F[5] + p.F[8];
        const float den = a * a + b * b;
        float best = p.th_low;
        for (int c0 = 0; c0 < p.n2; c0 += 32) {
            const int i2l = c0 + lane;
            const bool cand = i2l < p.n2 && p.node2[i2l] == nd && !p.mp2[i2l];  // same node, :831-836
            unsigned mask = __ballot_sync(AFULL, cand);
            while (mask) {  // ascending index = the order of f2it->second
                const int i2 = c0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const float dist = exact_distance(av, p.desc2 + (size_t)i2 * 256, lane);
                if (dist > p.th_low || dist > best) continue;  // :842
                const float x2 = p.pos2[2 * i2], y2 = p.pos2[2 * i2 + 1];
                const float ex = p.ep0 - x2, ey = p.ep1 - y2;
                if (sqrtf(ex * ex + ey * ey) < 10.0f) continue;  // :846-847
                if (den == 0) continue;                          // Pinhole.cpp:111-112
                const float num = a * x2 + b * y2 + c;
                const float dsqr = num * num / den;
                if ((double)dsqr < 3.84) {  // :113
                    best_idx = i2;
                    best = dist;
                }
            }
        }
    }
    if (lane == 0) p.match12[i1] = best_idx;
}

}  // namespace

}  // namespace ppg

using namespace ppg;

extern "C" int ppg_search_for_triangulation(ppg_ctx* c, const ppg_triangulation_match_in* in,
                                            ppg_triangulation_match_out* out) {
    if (!c || !in || !out || !out->match12) return set_err(c, PPG_ERR_ARG, "ppg_search_for_triangulation: null argument");
    const int n1 = in->n1, n2 = in->n2;
    if (n1 < 0 || n2 < 0) return set_err(c, PPG_ERR_ARG, "ppg_search_for_triangulation: negative size");
    if ((n1 > 0 && (!in->desc1 || !in->node1 || !in->has_mp1 || !in->pos1)) ||
        (n2 > 0 && (!in->desc2 || !in->node2 || !in->has_mp2 || !in->pos2)))
        return set_err(c, PPG_ERR_ARG, "ppg_search_for_triangulation: null array");
    out->nmatches = 0;
    if (n1 == 0) return PPG_OK;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    // one stream-ordered allocation for everything (this runs once per key-frame pair in LocalMapping, not per frame)
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_d1 = 0, o_d2 = o_d1 + up((size_t)n1 * 1024), o_p1 = o_d2 + up((size_t)n2 * 1024),
                 o_p2 = o_p1 + up((size_t)n1 * 8), o_n1 = o_p2 + up((size_t)n2 * 8), o_n2 = o_n1 + up((size_t)n1 * 4),
                 o_m1 = o_n2 + up((size_t)n2 * 4), o_m2 = o_m1 + up((size_t)n1), o_out = o_m2 + up((size_t)n2),
                 total = o_out + up((size_t)n1 * 4);
    uint8_t* d = nullptr;
    PPG_CUDA(c, cudaMallocAsync(reinterpret_cast<void**>(&d), total, c->st));
    auto h2d = [&](size_t off, const void* src, size_t bytes) {
        return bytes ? cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, c->st) : cudaSuccess;
    };
    cudaError_t e = h2d(o_d1, in->desc1, (size_t)n1 * 1024);
    if (e == cudaSuccess) e = h2d(o_d2, in->desc2, (size_t)n2 * 1024);
    if (e == cudaSuccess) e = h2d(o_p1, in->pos1, (size_t)n1 * 8);
    if (e == cudaSuccess) e = h2d(o_p2, in->pos2, (size_t)n2 * 8);
    if (e == cudaSuccess) e = h2d(o_n1, in->node1, (size_t)n1 * 4);
    if (e == cudaSuccess) e = h2d(o_n2, in->node2, (size_t)n2 * 4);
    if (e == cudaSuccess) e = h2d(o_m1, in->has_mp1, (size_t)n1);
    if (e == cudaSuccess) e = h2d(o_m2, in->has_mp2, (size_t)n2);
    if (e == cudaSuccess) {
        TriangParams p{};
        p.n1 = n1;
        p.n2 = n2;
        p.desc1 = reinterpret_cast<const float*>(d + o_d1);
        p.desc2 = reinterpret_cast<const float*>(d + o_d2);
        p.pos1 = reinterpret_cast<const float*>(d + o_p1);
        p.pos2 = reinterpret_cast<const float*>(d + o_p2);
        p.node1 = reinterpret_cast<const int*>(d + o_n1);
        p.node2 = reinterpret_cast<const int*>(d + o_n2);
        p.mp1 = d + o_m1;
        p.mp2 = d + o_m2;
        for (int i = 0; i < 9; i++) p.F[i] = in->F12[i];
        p.ep0 = in->epipole[0];
        p.ep1 = in->epipole[1];
        p.th_low = in->th_low;
        p.match12 = reinterpret_cast<int*>(d + o_out);
        triangulation_match_kernel<<<(n1 + 7) / 8, 256, 0, c->st>>>(p);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out->match12, d + o_out, (size_t)n1 * 4, cudaMemcpyDeviceToHost, c->st);
    const cudaError_t e2 = cudaStreamSynchronize(c->st);
    cudaFreeAsync(d, c->st);
    if (e != cudaSuccess) return cuda_fail(c, e, "ppg_search_for_triangulation");
    if (e2 != cudaSuccess) return cuda_fail(c, e2, "ppg_search_for_triangulation");
    int nm = 0;
    for (int i = 0; i < n1; i++) nm += out->match12[i] >= 0;
    out->nmatches = nm;
    return PPG_OK;
}
