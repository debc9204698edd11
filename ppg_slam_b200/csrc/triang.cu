// Matcher::SearchForTriangulation (matching/src/Matcher.cpp:767-885) on the GPU, for both camera models.
//
// The reference walks the FeatureVectors of two key frames node by node and, for every feature of KF1 without a map
// point, keeps the best feature of KF2 under the same node that passes the descriptor threshold, the epipole exclusion and
// mpCamera->epipolarConstrain: the closed-form epipolar-line distance of the pinhole camera (sensors/src/Pinhole.cpp:98-114)
// or the two-view triangulation of the KannalaBrandt8 camera (sensors/src/KannalaBrandt8.cpp:167-236: unproject by Newton
// iteration, parallax test, null vector of the 4 x 4 DLT matrix, positive depths, reprojection errors).  Its "already
// matched" flag vbMatched2 is never set, so the features of KF1 do not interact: one warp per feature, the lanes share one
// exact DescriptorDistance per candidate (the fixed summation order of assoc.cuh, as in every other matcher here) and
// evaluate the scalar tests redundantly.
// Compiled with -fmad=false: every float / double expression is evaluated left to right without contraction, which makes
// the kernel bit-identical to the test suite's CPU restatement (float transcendentals as the correctly rounded float of
// the double routine; the null vector by the same cyclic Jacobi iteration in double).
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
    int model;  // 0 Pinhole, 1 KannalaBrandt8
    float cam8[8], R12[9], t12[3];
    int* match12;
};

__device__ __forceinline__ float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// KannalaBrandt8::project(const Eigen::Vector3f&), KannalaBrandt8.cpp:44-59 (as frustum_kernel of assoc.cu)
__device__ void kb8_project(const float* cam8, const float* Pc, float& u, float& v) {
    const float x2y2 = Pc[0] * Pc[0] + Pc[1] * Pc[1];
    const float theta = (float)atan2((double)sqrtf(x2y2), (double)Pc[2]);
    const float psi = (float)atan2((double)Pc[1], (double)Pc[0]);
    const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2;
    const float theta7 = theta5 * theta2, theta9 = theta7 * theta2;
    const float r = theta + cam8[4] * theta3 + cam8[5] * theta5 + cam8[6] * theta7 + cam8[7] * theta9;
    u = (float)((double)(cam8[0] * r) * cos((double)psi) + (double)cam8[2]);
    v = (float)((double)(cam8[1] * r) * sin((double)psi) + (double)cam8[3]);
}

// KannalaBrandt8::unproject, :62-91
__device__ void kb8_unproject(const float* cam8, float px, float py, float* out3) {
    const float pw0 = (px - cam8[2]) / cam8[0], pw1 = (py - cam8[3]) / cam8[1];
    float scale = 1.f;
    float theta_d = sqrtf(pw0 * pw0 + pw1 * pw1);
    theta_d = fminf(fmaxf((float)(-3.1415926535897932384626433832795 / 2.0), theta_d),
                    (float)(3.1415926535897932384626433832795 / 2.0));
    if ((double)theta_d > 1e-8) {
        float theta = theta_d;
        for (int j = 0; j < 10; j++) {
            const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2,
                        theta8 = theta4 * theta4;
            const float k0_theta2 = cam8[4] * theta2, k1_theta4 = cam8[5] * theta4;
            const float k2_theta6 = cam8[6] * theta6, k3_theta8 = cam8[7] * theta8;
            const float theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                    (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < 1e-6f) break;
        }
        scale = (float)tan((double)theta) / theta_d;
    }
    out3[0] = pw0 * scale;
    out3[1] = pw1 * scale;
    out3[2] = 1.f;
}

// Right singular vector of the smallest singular value of a row-major 4 x 4 float matrix: cyclic Jacobi on A^T A in
// double, operation for operation as the CPU restatement's ppgo_null_vector4 (it stands in for Eigen::JacobiSVD at
// KannalaBrandt8.cpp:233-234: same vector, the sign cancels in the division by its last component).
__device__ void null_vector4(const float* A, float* v4) {
    double M[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double acc = (double)A[i] * (double)A[j];
#pragma unroll
            for (int k = 1; k < 4; k++) acc = acc + (double)A[4 * k + i] * (double)A[4 * k + j];
            M[i][j] = acc;
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) off = off + fabs(M[p][q]);
        if (off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                const double apq = M[p][q];
                if (apq == 0.0) continue;
                const double theta = (M[q][q] - M[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double mkp = M[k][p], mkq = M[k][q];
                    M[k][p] = c * mkp - sn * mkq;
                    M[k][q] = sn * mkp + c * mkq;
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double mpk = M[p][k], mqk = M[q][k];
                    M[p][k] = c * mpk - sn * mqk;
                    M[q][k] = sn * mpk + c * mqk;
                }
            }
    }
    double dmin = M[0][0];
#pragma unroll
    for (int i = 0; i < 4; i++) v4[i] = (float)V[i][0];
#pragma unroll
    for (int j = 1; j < 4; j++)
        if (M[j][j] < dmin) {
            dmin = M[j][j];
#pragma unroll
            for (int i = 0; i < 4; i++) v4[i] = (float)V[i][j];
        }
}

// KannalaBrandt8::epipolarConstrain = TriangulateMatches(...) > 0.0001f, :167-222, r1 = unproject(kp1.mPos) hoisted
__device__ bool kb8_epipolar_constrain(const TriangParams& p, const float* r1, float x1, float y1, float x2, float y2) {
    float r2[3], r21[3];
    kb8_unproject(p.cam8, x2, y2, r2);
#pragma unroll
    for (int i = 0; i < 3; i++) r21[i] = dot3(p.R12 + 3 * i, r2);  // :181
    const float cosParallaxRays = dot3(r1, r21) / (sqrtf(dot3(r1, r1)) * sqrtf(dot3(r21, r21)));
    if ((double)cosParallaxRays > 0.9998) return false;  // :183
    float R21[9], T2[12];  // Tcw2 = [R21 | -R21 * t12], :198-200; Tcw1 = [I | 0]
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R21[3 * i + j] = p.R12[3 * j + i];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) T2[4 * i + j] = R21[3 * i + j];
        T2[4 * i + 3] = ((-R21[3 * i]) * p.t12[0] + (-R21[3 * i + 1]) * p.t12[1]) + (-R21[3 * i + 2]) * p.t12[2];
    }
    float A[16];  // :227-231
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float t1r0 = j == 0 ? 1.f : 0.f, t1r1 = j == 1 ? 1.f : 0.f, t1r2 = j == 2 ? 1.f : 0.f;
        A[j] = r1[0] * t1r2 - t1r0;
        A[4 + j] = r1[1] * t1r2 - t1r1;
        A[8 + j] = r2[0] * T2[8 + j] - T2[j];
        A[12 + j] = r2[1] * T2[8 + j] - T2[4 + j];
    }
    float h[4];
    null_vector4(A, h);  // :233-234
    const float x3D[3] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]};  // :235
    const float z1 = x3D[2];
    if (z1 <= 0) return false;  // :205
    const float z2 = dot3(R21 + 6, x3D) + T2[11];
    if (z2 <= 0) return false;  // :209
    float u, v;
    kb8_project(p.cam8, x3D, u, v);
    const float e10 = u - x1, e11 = v - y1;
    if ((double)(e10 * e10 + e11 * e11) > 5.991) return false;  // :213-214
    float x3D2[3];
#pragma unroll
    for (int i = 0; i < 3; i++) x3D2[i] = dot3(R21 + 3 * i, x3D) + T2[4 * i + 3];
    kb8_project(p.cam8, x3D2, u, v);
    const float e20 = u - x2, e21 = v - y2;
    if ((double)(e20 * e20 + e21 * e21) > 5.991) return false;  // :218-219
    return z1 > 0.0001f;  // :171
}

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
        float r1[3] = {0.f, 0.f, 1.f};
        if (p.model == 1) kb8_unproject(p.cam8, x1, y1, r1);
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
                if (p.model == 1) {
                    if (kb8_epipolar_constrain(p, r1, x1, y1, x2, y2)) {  // :848
                        best_idx = i2;
                        best = dist;
                    }
                    continue;
                }
                if (den == 0) continue;  // Pinhole.cpp:111-112
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
    if (in->camera_model != 0 && in->camera_model != 1)
        return set_err(c, PPG_ERR_ARG, "ppg_search_for_triangulation: camera_model must be 0 (Pinhole) or 1 (KannalaBrandt8)");
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
        p.model = in->camera_model;
        for (int i = 0; i < 8; i++) p.cam8[i] = in->cam8[i];
        for (int i = 0; i < 9; i++) p.R12[i] = in->R12[i];
        for (int i = 0; i < 3; i++) p.t12[i] = in->t12[i];
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
