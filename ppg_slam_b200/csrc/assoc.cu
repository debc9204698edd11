// Image<->map descriptor association: the search core of Matcher::ExtendMapMatches
// (matching/src/Matcher.cpp:224-281) for every map point at once, frame state frozen.
//
//   K14a prep     map / frame descriptors fp32 -> bf16 + squared norms; Frame::PosInGrid cells of the
//                 keypoints (map/src/Frame.cpp:317-327); per-map-point search window and the grid-cell
//                 range Frame::GetFeaturesInArea would visit (Frame.cpp:262-315)
//   K14b gemm     brute-force (M x 256) . (256 x N) bf16 GEMM on tcgen05/TMEM fed by TMA; the epilogue
//                 turns dots into approximate squared distances, applies the window mask and keeps the
//                 per-row 4 best candidates + the 5th best value (the guard)
//   K14c rescore  exact fp32 DescriptorDistance (feature/src/MapPoint.cpp:22-29) of the <= 4 candidates
//                 in the fixed summation order the oracle defines, best / second best with the reference's
//                 tie rule (first in GetFeaturesInArea order), ratio test (Matcher.cpp:276).  If the guard
//                 cannot prove that no other candidate can beat the second best (bf16 error bound), the row
//                 is re-scored exactly over its whole window on the GPU (counted in fallback_rows).
// Built with -fmad=false (bit-exact distances).
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "assoc.cuh"
#include "ctx.cuh"
#include "once.cuh"
#include "ptx.cuh"

namespace ppg {
namespace {

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 + squared norm, one warp per row.  Rows >= n are zero-filled (inert GEMM columns).
__device__ __forceinline__ void prep_row(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                         float* __restrict__ n2, int row, int n, int lane, float* nbmax) {
    float v[8];
    if (row < n) {
        const float4 a = *reinterpret_cast<const float4*>(src + (size_t)row * 256 + lane * 8);
        const float4 b = *reinterpret_cast<const float4*>(src + (size_t)row * 256 + lane * 8 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = 0.f;
    }
    float s = 0.f;
    __nv_bfloat162 o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        s += v[2 * k] * v[2 * k] + v[2 * k + 1] * v[2 * k + 1];
    }
    *reinterpret_cast<uint4*>(dst + (size_t)row * 256 + lane * 8) = *reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(AFULL, s, m);
    if (lane == 0) {
        n2[row] = s;
        if (nbmax && row < n) atomicMax(reinterpret_cast<int*>(nbmax), __float_as_int(s));
    }
}

__global__ void __launch_bounds__(256) prep_map_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       float* __restrict__ n2, int n, int rows_padded) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_padded) return;
    prep_row(src, dst, n2, row, n, threadIdx.x & 31, nullptr);
}


// Frame side of one call, grid (ncap/8, frames): descriptors -> bf16 + norms; Frame::PosInGrid
// (Frame.cpp:317-327) of every keypoint + the free mask (Matcher.cpp:253).
__global__ void __launch_bounds__(256) prep_frame_kernel(const FrameSrc src, int ncap, GridParam g,
                                                         __nv_bfloat16* __restrict__ f_bf, float* __restrict__ fn2,
                                                         uint32_t* __restrict__ kinfo, uint32_t* __restrict__ korder,
                                                         float* __restrict__ nbmax) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int n = min(src.n_of(f), ncap);
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= ncap) return;
    prep_row(src.desc_of(f), f_bf + (size_t)f * ncap * 256, fn2 + (size_t)f * ncap, row, n, lane, nbmax + f);
    if (lane == 0) {
        uint32_t info = 0, order = 0xffffffffu;
        if (row < n) {
            const float x = src.kx_of(f)[row], y = src.ky_of(f)[row];
            const int px = (int)roundf((x - (float)g.minX) * g.wInv);
            const int py = (int)roundf((y - (float)g.minY) * g.hInv);
            const bool indexable = !(px < 0 || px >= 64 || py < 0 || py >= 48);
            if (indexable) {
                info = (uint32_t)px | ((uint32_t)py << 8) | ((src.free_of(f)[row] ? 1u : 0u) << 16);
                order = ((uint32_t)(px * 48 + py) << 16) | (uint32_t)row;
            }
        }
        kinfo[(size_t)f * ncap + row] = info;
        korder[(size_t)f * ncap + row] = order;
    }
}

// Frame::CheckInFrustum (map/src/Frame.cpp:223-260) of every resident map point under the pose of every frame, with
// Pinhole::project (sensors/src/Pinhole.cpp:32-38) / KannalaBrandt8::project (KannalaBrandt8.cpp:44-59) and
// GeometricCamera::IsInImage (GeometricCamera.cpp:21-24).  grid (rows/256, frames).  The arithmetic order is the one the
// CPU restatement fixes: (a0*b0 + a1*b1) + a2*b2 for the rows of Rcw * P and the dot products, no FMA (-fmad=false);
// KB8: atan2f / cos / sin through the double routines, rounded once.
struct FrustumParam {
    float fx, fy, cx, cy, k0, k1, k2, k3;
    float minX, maxX, minY, maxY;
    int fisheye;
    float cos_limit;
};
__device__ __forceinline__ float dot3(const float* a, float b0, float b1, float b2) {
    return (a[0] * b0 + a[1] * b1) + a[2] * b2;
}
__global__ void frustum_kernel(const float* __restrict__ wpos, const float* __restrict__ nrm,
                               const float* __restrict__ dmin, const float* __restrict__ dmax,
                               const float* __restrict__ poses, int rows, int max_rows, FrustumParam fp,
                               float* __restrict__ proj, float* __restrict__ vcos, float* __restrict__ depth,
                               uint8_t* __restrict__ in_view) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (m >= rows) return;
    const size_t o = (size_t)f * max_rows + m;
    const float* T = poses + f * 16;
    const float P0 = wpos[3 * m], P1 = wpos[3 * m + 1], P2 = wpos[3 * m + 2];
    float u = -1.f, v = -1.f, dep = -1.f, vc = 0.f;  // :226-228
    bool ok = false;
    const float Pc0 = dot3(T, P0, P1, P2) + T[9];
    const float Pc1 = dot3(T + 3, P0, P1, P2) + T[10];
    const float Pc2 = dot3(T + 6, P0, P1, P2) + T[11];
    if (!(Pc2 < 0.0f)) {  // :234
        float pu, pv;
        if (!fp.fisheye) {
            pu = fp.fx * Pc0 / Pc2 + fp.cx;
            pv = fp.fy * Pc1 / Pc2 + fp.cy;
        } else {
            const float x2y2 = Pc0 * Pc0 + Pc1 * Pc1;
            const float theta = (float)atan2((double)sqrtf(x2y2), (double)Pc2);
            const float psi = (float)atan2((double)Pc1, (double)Pc0);
            const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2;
            const float theta7 = theta5 * theta2, theta9 = theta7 * theta2;
            const float r = theta + fp.k0 * theta3 + fp.k1 * theta5 + fp.k2 * theta7 + fp.k3 * theta9;
            pu = (float)((double)(fp.fx * r) * cos((double)psi) + (double)fp.cx);
            pv = (float)((double)(fp.fy * r) * sin((double)psi) + (double)fp.cy);
        }
        if (pu >= fp.minX && pu < fp.maxX && pv >= fp.minY && pv < fp.maxY) {  // :238
            const float PO0 = P0 - T[12], PO1 = P1 - T[13], PO2 = P2 - T[14];
            const float dist = sqrtf((PO0 * PO0 + PO1 * PO1) + PO2 * PO2);
            if (!(dist < dmin[m] || dist > dmax[m])) {  // :245
                const float c = ((PO0 * nrm[3 * m] + PO1 * nrm[3 * m + 1]) + PO2 * nrm[3 * m + 2]) / dist;
                if (!(c < fp.cos_limit)) {  // :250
                    ok = true;
                    u = pu;
                    v = pv;
                    dep = dist;
                    vc = c;
                }
            }
        }
    }
    proj[2 * o] = u;
    proj[2 * o + 1] = v;
    vcos[o] = vc;
    depth[o] = dep;
    in_view[o] = ok ? 1 : 0;
}

// Search window of each map point: r (Matcher.cpp:240-244) and the cell range of GetFeaturesInArea
// (Frame.cpp:270-292) including its early returns.  grid (rows/256, frames).
__global__ void prep_rows_kernel(const float* __restrict__ proj, const float* __restrict__ vcos,
                                 const float* __restrict__ n2, int rows, int max_rows, float th, int mode,
                                 GridParam g, RowParam* __restrict__ rp, const uint8_t* __restrict__ in_view) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (m >= rows) return;
    const size_t o = (size_t)f * max_rows + m;
    RowParam p;
    p.u = proj[2 * o];
    p.v = proj[2 * o + 1];
    float r = th;  // PPG_SEARCH_WINDOW: r = th (Matcher.cpp:56, :1378, :971)
    if (mode == 0) {  // ExtendMapMatches :240-244
        if ((double)vcos[o] > 0.998)
            r = (float)((double)r * 2.5);
        else
            r = (float)((double)r * 4.0);
    }
    p.r = r;
    p.na2 = n2[m];
    bool empty = in_view != nullptr && !in_view[o];  // !mbTrackInView: not a candidate (Matcher.cpp:212)
    int x0 = (int)floorf((p.u - (float)g.minX - r) * g.wInv);
    if (x0 < 0) x0 = 0;
    if (x0 >= 64) empty = true;
    int x1 = (int)ceilf((p.u - (float)g.minX + r) * g.wInv);
    if (x1 > 63) x1 = 63;
    if (x1 < 0) empty = true;
    int y0 = (int)floorf((p.v - (float)g.minY - r) * g.hInv);
    if (y0 < 0) y0 = 0;
    if (y0 >= 48) empty = true;
    int y1 = (int)ceilf((p.v - (float)g.minY + r) * g.hInv);
    if (y1 > 47) y1 = 47;
    if (y1 < 0) empty = true;
    p.cells = empty ? 0xffffffffu : ((uint32_t)x0 | ((uint32_t)x1 << 8) | ((uint32_t)y0 << 16) | ((uint32_t)y1 << 24));
    rp[o] = p;
}

// ------------------------------------------------------------------------------------------------
// K14b.  GEMM view: D[map row, keypoint] = <a, b>.  M tile = 128 map rows (TMEM lanes), N tile = 128
// keypoints, K = 256 = 4 chunks of 64 bf16 (one 128-byte swizzled smem row per descriptor per chunk).
// warp 0: TMA producer, warp 1: MMA issuer / TMEM owner, warps 2-5: epilogue (one map row per thread).
// Work items = (frame, M tile); each CTA takes a contiguous range so that the per-frame keypoint table in
// shared memory is reloaded only when the frame changes.
struct GemmParams {
    int rows, max_rows, ncap, frames;
    FrameSrc src;
    const RowParam* rowp;
    const float* fn2;
    const uint32_t* kinfo;
    int* cand;
    float* guard;
    double e2_max;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(A_THREADS, 1)
assoc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)A_STAGES * A_STAGE_BYTES);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    // keypoint tables of the current frame: (x, y) and (|b|^2, info bits), [ncap] each
    float2* kxy = reinterpret_cast<float2*>(tmem_slot + 4);
    float2* kzi = kxy + p.ncap;
    float2* qbuf = kzi + p.ncap;                              // [A_QCAP][256] (d2, column) hit queues
    float* mbuf = reinterpret_cast<float*>(qbuf + A_QCAP * 256);  // [9][128] top-4 lists of column half 1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.rows + A_BM - 1) / A_BM;
    const int total = m_tiles * p.frames;
    const int per = (total + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * per, t1 = min(total, t0 + per);

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; i++) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 8);  // 2 column halves x 4 lane quarters
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&mapA);
        ptx::prefetch_tmap(&mapB);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // warps 0 and 1: the whole warp runs the (warp-uniform) loop and one elected lane issues the TMA / MMA
    // instructions, so that their operands stay in uniform registers (see conv_tc.cu)
    if (warp == 0) {
        uint32_t it = 0;
        for (int t = t0; t < t1; t++) {
            const int f = t / m_tiles, mt = t - f * m_tiles;
            const int n_tiles = (min(p.src.n_of(f), p.ncap) + A_BN - 1) / A_BN;
            for (int nt = 0; nt < n_tiles; nt++)
                for (int kc = 0; kc < 4; kc++, it++) {
                    const uint32_t s = it % A_STAGES, ph = (it / A_STAGES) & 1;
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&full[s], A_STAGE_BYTES);
                        uint8_t* a = smem + (size_t)s * A_STAGE_BYTES;
                        ptx::tma_load_2d(a, &mapA, &full[s], kc * 64, mt * A_BM);
                        ptx::tma_load_2d(a + 16384, &mapB, &full[s], kc * 64, f * p.ncap + nt * A_BN);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::make_idesc_f16(A_BM, A_BN, 1);
        uint32_t it = 0, lt = 0;
        for (int t = t0; t < t1; t++) {
            const int f = t / m_tiles;
            const int n_tiles = (min(p.src.n_of(f), p.ncap) + A_BN - 1) / A_BN;
            for (int nt = 0; nt < n_tiles; nt++, lt++) {
                const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
                ptx::mbar_wait(&tempty[acc], aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * A_BN;
                for (int kc = 0; kc < 4; kc++, it++) {
                    const uint32_t s = it % A_STAGES, ph = (it / A_STAGES) & 1;
                    ptx::mbar_wait(&full[s], ph);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint32_t a_addr = ptx::smem_u32(smem + (size_t)s * A_STAGE_BYTES);
                        const uint64_t adesc = ptx::make_sw128_desc(a_addr);
                        const uint64_t bdesc = ptx::make_sw128_desc(a_addr + 16384);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ptx::umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kc | k) != 0));
                        ptx::umma_commit(&empty[s]);
                        if (kc == 3) ptx::umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
        // One map row per thread and column half.  The first version ran the whole window mask and the top-4
        // insertion on every (row, keypoint) element and took 0.87 ms per 32 frames (the MMAs need 0.05 ms).  Now an
        // element costs four compares against a slightly widened window box; the ~2 % that pass are queued (d2,
        // column) in shared memory and get the exact mask (in_window) and the insertion once per row.  Both halves
        // keep a top-4 list + guard under the total order (d2, column); half 1 hands its list to half 0, which merges
        // and stores.  That order makes the result independent of how the columns are split or queued.
        const int q = warp & 3, grp = (warp - 2) >> 2;
        const int rloc = q * 32 + lane, etid = threadIdx.x - 64;  // 0..255
        uint32_t lt = 0;
        int cur_f = -1, n = 0;
        for (int t = t0; t < t1; t++) {
            const int f = t / m_tiles, mt = t - f * m_tiles;
            if (f != cur_f) {  // (re)load the frame's keypoint tables; all eight epilogue warps take this branch together
                epi_bar();
                n = min(p.src.n_of(f), p.ncap);
                const float* kx = p.src.kx_of(f);
                const float* ky = p.src.ky_of(f);
                const int ncols = (n + A_BN - 1) / A_BN * A_BN;
                for (int i = etid; i < ncols; i += 256) {
                    // padding columns sit far outside every window box
                    float2 xy = make_float2(-1e30f, -1e30f), zi = make_float2(0.f, 0.f);
                    if (i < n) {
                        xy = make_float2(kx[i], ky[i]);
                        zi = make_float2(p.fn2[(size_t)f * p.ncap + i], __uint_as_float(p.kinfo[(size_t)f * p.ncap + i]));
                    }
                    kxy[i] = xy;
                    kzi[i] = zi;
                }
                epi_bar();
                cur_f = f;
            }
            const int n_tiles = (n + A_BN - 1) / A_BN;
            const int row = mt * A_BM + rloc;
            RowParam rp;
            rp.u = rp.v = rp.r = rp.na2 = 0.f;
            rp.cells = 0xffffffffu;
            if (row < p.rows) rp = p.rowp[(size_t)f * p.max_rows + row];
            // conservative box around |x - u| < r, |y - v| < r (0.01 px covers the rounding of x - u for |x| < 2^13)
            const bool live = rp.cells != 0xffffffffu;
            const float xlo = live ? rp.u - rp.r - 0.01f : 1e30f, xhi = rp.u + rp.r + 0.01f;
            const float ylo = rp.v - rp.r - 0.01f, yhi = rp.v + rp.r + 0.01f;
            float bd[A_TOPK];
            int bi[A_TOPK];
#pragma unroll
            for (int k = 0; k < A_TOPK; k++) {
                bd[k] = INFINITY;
                bi[k] = 0x7fffffff;
            }
            float a5 = INFINITY;
            auto insert = [&](float d2, int col) {
                if (d2 < bd[A_TOPK - 1] || (d2 == bd[A_TOPK - 1] && col < bi[A_TOPK - 1])) {
                    a5 = fminf(a5, bd[A_TOPK - 1]);  // the displaced 4th best (min: a5 may be the -inf overflow mark)
                    float cd = d2;
                    int ci = col;
#pragma unroll
                    for (int k = 0; k < A_TOPK; k++) {
                        if (cd < bd[k] || (cd == bd[k] && ci < bi[k])) {
                            const float td = bd[k];
                            const int ti = bi[k];
                            bd[k] = cd;
                            bi[k] = ci;
                            cd = td;
                            ci = ti;
                        }
                    }
                } else if (d2 < a5) {
                    a5 = d2;
                }
            };
            // Hits are queued as (raw accumulator, column) and drained (exact mask + insertion) whenever fewer than 16
            // slots are left before a 16-column chunk, and at the end of the row.  The scan loop is deliberately NOT
            // fully unrolled and the drain exists once: the first queue version inlined it 64 times and ran at 0.07
            // IPC on instruction-cache misses (stall_no_instruction 9.9 per issue in the ncu capture).
            int qcnt = 0;
            auto drain = [&]() {
#pragma unroll 1
                for (int sidx = 0; sidx < qcnt; sidx++) {
                    const float2 e = qbuf[sidx * 256 + etid];
                    const int col = __float_as_int(e.y);
                    const float2 xy = kxy[col], zi = kzi[col];
                    if (in_window(rp, __float_as_uint(zi.y), xy.x, xy.y, p.e2_max))
                        insert((rp.na2 + zi.x) - 2.0f * e.x, col);
                }
                qcnt = 0;
            };
            for (int nt = 0; nt <= n_tiles; nt++) {
                if (nt == n_tiles) {  // same call site for the final drain: one copy of the code
                    drain();
                    break;
                }
                const uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
                lt++;
                ptx::mbar_wait(&tfull[acc], aph);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * A_BN + grp * 64;
                const int cbase = nt * A_BN + grp * 64;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t rr[16];
                    __syncwarp();  // the drain below is per-lane divergent; tcgen05.ld is warp-collective
                    ptx::tmem_ld16(taddr + c0, rr);
                    ptx::tmem_ld_wait();
                    if (qcnt > A_QCAP - 16) drain();
                    const float4* xy4 = reinterpret_cast<const float4*>(kxy + cbase + c0);
#pragma unroll
                    for (int j2 = 0; j2 < 8; j2++) {
                        const float4 v = xy4[j2];  // (x, y) of columns cbase + c0 + 2 j2, + 1
                        const bool h0 = (v.x > xlo) & (v.x < xhi) & (v.y > ylo) & (v.y < yhi);
                        const bool h1 = (v.z > xlo) & (v.z < xhi) & (v.w > ylo) & (v.w < yhi);
                        if (h0) {
                            qbuf[qcnt * 256 + etid] =
                                make_float2(__uint_as_float(rr[2 * j2]), __int_as_float(cbase + c0 + 2 * j2));
                            qcnt++;
                        }
                        if (h1) {
                            qbuf[qcnt * 256 + etid] =
                                make_float2(__uint_as_float(rr[2 * j2 + 1]), __int_as_float(cbase + c0 + 2 * j2 + 1));
                            qcnt++;
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            }
            // merge the two column halves (half 1 -> half 0)
            if (grp == 1) {
#pragma unroll
                for (int k = 0; k < A_TOPK; k++) {
                    mbuf[k * 128 + rloc] = bd[k];
                    mbuf[(A_TOPK + k) * 128 + rloc] = __int_as_float(bi[k]);
                }
                mbuf[2 * A_TOPK * 128 + rloc] = a5;
            }
            epi_bar();
            if (grp == 0) {
#pragma unroll
                for (int k = 0; k < A_TOPK; k++) {
                    const int ci = __float_as_int(mbuf[(A_TOPK + k) * 128 + rloc]);
                    if (ci != 0x7fffffff) insert(mbuf[k * 128 + rloc], ci);
                }
                a5 = fminf(a5, mbuf[2 * A_TOPK * 128 + rloc]);
                if (row < p.rows) {
                    const size_t o = (size_t)f * p.max_rows + row;
                    *reinterpret_cast<int4*>(p.cand + o * 4) =
                        make_int4(bi[0] == 0x7fffffff ? -1 : bi[0], bi[1] == 0x7fffffff ? -1 : bi[1],
                                  bi[2] == 0x7fffffff ? -1 : bi[2], bi[3] == 0x7fffffff ? -1 : bi[3]);
                    p.guard[o] = a5;
                }
            }
            epi_bar();  // mbuf is free again
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
struct RescoreParams {
    int rows, max_rows, ncap;
    FrameSrc src;
    const RowParam* rowp;
    const float* map_f32;
    const uint32_t* kinfo;
    const uint32_t* korder;
    const int* cand;
    const float* guard;
    const float* nbmax;
    float ratio, th_high;
    int mode;         // 0: ExtendMapMatches rule, 1: best <= max_dist
    float max_dist;
    double e2_max;
    int force_exact;  // 1: ignore the GEMM candidates and score every window exactly (validation, node mode)
    const int* row_node;  // mode 2 (SearchByBoW): vocabulary node of every row / keypoint, candidates = same node
    const int* kp_node;
    int *best_idx, *second_idx;
    float *best_d, *second_d;
    uint8_t* accept;
    int* fallback;
};

__global__ void __launch_bounds__(256) assoc_rescore_kernel(const RescoreParams p) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    if (row >= p.rows) return;
    const size_t o = (size_t)f * p.max_rows + row;
    const int n = min(p.src.n_of(f), p.ncap);
    const RowParam rp = p.rowp[o];
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = p.map_f32[(size_t)row * 256 + lane + 32 * k];
    const float* fdesc = p.src.desc_of(f);
    const uint32_t* korder = p.korder + (size_t)f * p.ncap;
    float b1 = 1e6f, b2 = 1e6f;  // Matcher.cpp:248-249 initial values
    uint32_t o1 = 0xffffffffu, o2 = 0xffffffffu;
    int i1 = -1, i2 = -1;
    bool exact_all = p.force_exact != 0;
    if (!exact_all) {
        const int4 c4 = *reinterpret_cast<const int4*>(p.cand + o * 4);
        const int cs[4] = {c4.x, c4.y, c4.z, c4.w};
        // the four candidates' rows are fetched and reduced together (independent chains; one at a time the kernel
        // was bound by 4 dependent rounds of load latency per row), then ranked in order
        float dk[4];
        uint32_t ok[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ck = cs[k] < 0 ? 0 : cs[k];
            dk[k] = exact_distance(a, fdesc + (size_t)ck * 256, lane);
            ok[k] = korder[ck];
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (cs[k] >= 0) top2_update(dk[k], ok[k], cs[k], b1, o1, i1, b2, o2, i2);
        const float g = p.guard[o];
        if (g < INFINITY) {
            // bf16 operands: |dot_bf16 - dot| <= (2^-8 + 2^-16) |a||b| (+ fp32 accumulation); in squared distance x2
            const float delta = 0.0080f * sqrtf(rp.na2 * p.nbmax[f]) + 2e-4f;
            const float e2 = b2 * b2;
            if (!(g - delta > e2 + delta)) exact_all = true;
        }
    }
    if (exact_all) {
        if (!p.force_exact && lane == 0) atomicAdd(p.fallback, 1);
        b1 = b2 = 1e6f;
        o1 = o2 = 0xffffffffu;
        i1 = i2 = -1;
        const float* kx = p.src.kx_of(f);
        const float* ky = p.src.ky_of(f);
        const uint32_t* kinfo = p.kinfo + (size_t)f * p.ncap;
        const int my_node = p.mode == 2 ? p.row_node[row] : -1;
        const uint8_t* freem = p.src.free_of(f);
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int c = c0 + lane;
            bool in = false;
            if (c < n) {
                if (p.mode == 2)  // features of the same vocabulary node, in vIndicesF order (Matcher.cpp:432-437)
                    in = my_node >= 0 && p.kp_node[c] == my_node && freem[c] != 0;
                else
                    in = in_window(rp, kinfo[c], kx[c], ky[c], p.e2_max);
            }
            unsigned mask = __ballot_sync(AFULL, in);
            while (mask) {
                const int cc = c0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const float d = exact_distance(a, fdesc + (size_t)cc * 256, lane);
                top2_update(d, p.mode == 2 ? (uint32_t)cc : korder[cc], cc, b1, o1, i1, b2, o2, i2);
            }
        }
    }
    if (lane == 0) {
        p.best_idx[o] = i1;
        p.second_idx[o] = i2;
        p.best_d[o] = b1;
        p.second_d[o] = b2;
        uint8_t acc = 0;
        if (i1 >= 0) {
            if (p.mode == 2)
                acc = b1 <= p.max_dist && b1 < p.ratio * b2;  // Matcher.cpp:456-458, :733-735
            else if (p.mode == 1)
                acc = b1 <= p.max_dist;  // Matcher.cpp:78, :1399, :1016
            else
                acc = !(b1 > p.th_high && b1 > p.ratio * b2);  // Matcher.cpp:276
        }
        p.accept[o] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (feature/src/MapPoint.cpp:234-302): one CTA per map point.
//   pairwise DescriptorDistance of its n observations (one warp per pair, the oracle's summation order) -> D in
//   shared memory; thread i finds the median of row i (the value of rank (int)(0.5 (n-1)), by counting) ; thread 0
//   takes the first row whose median is below the running best (start 1.0f, strict <, :279-291).
constexpr int DD_MAX = PPG_MAX_OBSERVATIONS;

__global__ void __launch_bounds__(256) distinctive_kernel(const float* __restrict__ desc, const int* __restrict__ offsets,
                                                          int n_points, int* __restrict__ best_idx,
                                                          int* __restrict__ err) {
    extern __shared__ float dd_smem[];  // D [n][n], then median [n]
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int o0 = offsets[p], n = offsets[p + 1] - o0;
    if (n <= 0 || n > DD_MAX) {
        if (tid == 0) {
            best_idx[p] = n <= 0 ? -1 : 0;
            if (n > DD_MAX) atomicExch(err, 1);
        }
        return;
    }
    float* D = dd_smem;
    float* med = dd_smem + n * n;
    const float* base = desc + (size_t)o0 * 256;
    for (int i = tid; i < n; i += blockDim.x) D[i * n + i] = 0.f;  // :270
    // pair (i, j), i < j, enumerated row-major; one warp per pair
    const int npairs = n * (n - 1) / 2;
    for (int q = warp; q < npairs; q += nw) {
        // row i such that i*(2n-i-1)/2 <= q
        int i = 0, rem = q;
        while (rem >= n - 1 - i) {
            rem -= n - 1 - i;
            i++;
        }
        const int j = i + 1 + rem;
        float av[8];
#pragma unroll
        for (int k = 0; k < 8; k++) av[k] = base[(size_t)i * 256 + lane + 32 * k];
        const float d = exact_distance(av, base + (size_t)j * 256, lane);
        if (lane == 0) {
            D[i * n + j] = d;
            D[j * n + i] = d;
        }
    }
    __syncthreads();
    const int k = (int)(0.5 * (double)(n - 1));  // :285
    for (int i = tid; i < n; i += blockDim.x) {
        const float* row = D + i * n;
        float m = 0.f;
        for (int c = 0; c < n; c++) {  // the value v with #(x < v) <= k < #(x <= v) is the element of rank k
            const float v = row[c];
            int lt = 0, le = 0;
            for (int x = 0; x < n; x++) {
                lt += row[x] < v;
                le += row[x] <= v;
            }
            if (lt <= k && k < le) {
                m = v;
                break;
            }
        }
        med[i] = m;
    }
    __syncthreads();
    if (tid == 0) {
        float best_median = 1.0f;
        int best = 0;
        for (int i = 0; i < n; i++)
            if (med[i] < best_median) {
                best_median = med[i];
                best = i;
            }
        best_idx[p] = best;
    }
}

// map_f32[p] = observation offsets[p] + best_idx[p]
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ desc, const int* __restrict__ offsets,
                                                          const int* __restrict__ best_idx, int n_points,
                                                          float* __restrict__ dst) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n_points) return;
    const int b = best_idx[p];
    const float4* src = reinterpret_cast<const float4*>(desc + (size_t)(offsets[p] + (b < 0 ? 0 : b)) * 256);
    float4* d = reinterpret_cast<float4*>(dst + (size_t)p * 256);
    d[lane] = b < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : src[lane];
    d[lane + 32] = b < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : src[lane + 32];
}

int gemm_smem(const AssocState* s) {
    return A_STAGES * A_STAGE_BYTES + 1024 + 20 * 8 + 16 + s->ncap * 16 + A_QCAP * 256 * 8 + 9 * 128 * 4 + 64;
}

}  // namespace

int assoc_ensure_state(ppg_ctx* c) {
    if (c->assoc) return PPG_OK;
    AssocState* s = new AssocState();
    c->assoc = s;
    s->max_rows = c->cfg.max_map_points > 0 ? c->cfg.max_map_points : 65536;
    s->ncap = 1024;
    s->bcap = c->maxB;
    const size_t R = s->max_rows, N = s->ncap, B = s->bcap;
    PPG_CUDA(c, dalloc(&s->map_f32, R * 256));
    PPG_CUDA(c, dalloc(&s->map_bf, R * 256));
    PPG_CUDA(c, dalloc(&s->map_n2, R));
    PPG_CUDA(c, dalloc(&s->kx, N));
    PPG_CUDA(c, dalloc(&s->ky, N));
    PPG_CUDA(c, dalloc(&s->fdesc, N * 256));
    PPG_CUDA(c, dalloc(&s->free_mask, N));
    PPG_CUDA(c, dalloc(&s->ones, N));
    PPG_CUDA(c, cudaMemset(s->ones, 1, N));
    PPG_CUDA(c, dalloc(&s->fn2, B * N));
    PPG_CUDA(c, dalloc(&s->f_bf, B * N * 256));
    PPG_CUDA(c, dalloc(&s->kinfo, B * N));
    PPG_CUDA(c, dalloc(&s->korder, B * N));
    PPG_CUDA(c, dalloc(&s->nbmax, B));
    PPG_CUDA(c, dalloc(&s->proj, B * R * 2));
    PPG_CUDA(c, dalloc(&s->vcos, B * R));
    PPG_CUDA(c, dalloc(&s->rowp, B * R));
    PPG_CUDA(c, dalloc(&s->cand, B * R * 4));
    PPG_CUDA(c, dalloc(&s->guard, B * R));
    PPG_CUDA(c, dalloc(&s->best_idx, B * R));
    PPG_CUDA(c, dalloc(&s->second_idx, B * R));
    PPG_CUDA(c, dalloc(&s->best_d, B * R));
    PPG_CUDA(c, dalloc(&s->second_d, B * R));
    PPG_CUDA(c, dalloc(&s->accept, B * R));
    PPG_CUDA(c, dalloc(&s->fallback, 1));
    PPG_CUDA(c, dalloc(&s->row_node, R));
    PPG_CUDA(c, dalloc(&s->kp_node, N));
    PPG_CUDA(c, dalloc(&s->wpos, R * 3));
    PPG_CUDA(c, dalloc(&s->nrm, R * 3));
    PPG_CUDA(c, dalloc(&s->dmin, R));
    PPG_CUDA(c, dalloc(&s->dmax, R));
    PPG_CUDA(c, dalloc(&s->poses, B * 16));
    PPG_CUDA(c, dalloc(&s->in_view, B * R));
    PPG_CUDA(c, dalloc(&s->depth, B * R));
    s->h_res_bytes = B * R * 17;
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&s->h_res), s->h_res_bytes));
    PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&s->h_stage), B * R * 3 * sizeof(float)));
    if (!make_kmajor_map(&s->mapA, s->map_bf, R, 256, A_BM, true) ||
        !make_kmajor_map(&s->mapB, s->f_bf, B * N, 256, A_BN, true))
        return set_err(c, PPG_ERR_CUDA, "cuTensorMapEncodeTiled failed for the association operands");
    PPG_CUDA(c, cudaFuncSetAttribute(assoc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    return PPG_OK;
}

FrameSrc assoc_staged_src(const AssocState* s) {
    FrameSrc f;
    f.kx = reinterpret_cast<const uint8_t*>(s->kx);
    f.ky = reinterpret_cast<const uint8_t*>(s->ky);
    f.desc = reinterpret_cast<const uint8_t*>(s->fdesc);
    f.free_mask = s->free_mask;
    f.n = nullptr;
    f.stride = 0;
    f.free_stride = 0;
    f.n_val = s->staged_n;
    return f;
}

// Frames first..first+frames-1 of the last extraction batch: mvKeysUn[i].mPos as run() returns it (kp_x, kp_y)
// and the descriptors, still on the device.
FrameSrc assoc_extracted_src(const ppg_ctx* c, int first) {
    const OutLayout& L = c->post.lay;
    const uint8_t* blk = c->d_out + (size_t)first * L.total;
    FrameSrc f;
    f.kx = blk + L.kp_x;
    f.ky = blk + L.kp_y;
    f.desc = blk + L.desc;
    f.free_mask = c->assoc->ones;
    f.n = blk + L.hdr + HDR_NKP * sizeof(int);
    f.stride = L.total;
    f.free_stride = 0;
    f.n_val = 0;
    return f;
}

int assoc_prep(ppg_ctx* c, const FrameSrc& src, int frames) {
    AssocState* s = c->assoc;
    const int rows = s->staged_rows;
    GridParam g{c->minX, c->minY, c->wInv, c->hInv};
    PPG_CUDA(c, cudaMemsetAsync(s->nbmax, 0, 4 * frames, c->st));
    prep_frame_kernel<<<dim3(s->ncap / 8, frames), 256, 0, c->st>>>(src, s->ncap, g, s->f_bf, s->fn2, s->kinfo,
                                                                    s->korder, s->nbmax);
    prep_rows_kernel<<<dim3((rows + 255) / 256, frames), 256, 0, c->st>>>(s->proj, s->vcos, s->map_n2, rows,
                                                                          s->max_rows, s->th, s->mode, g, s->rowp,
                                                                          s->use_in_view ? s->in_view : nullptr);
    c->launches += 2;
    PPG_CUDA(c, cudaGetLastError());
    stage_mark(c, "assoc.prep");
    return PPG_OK;
}

namespace {

// prep + gemm + rescore on the ctx stream for `frames` frames; results land in slots 0..frames-1.
int run_assoc(ppg_ctx* c, const FrameSrc& src, int frames, int force_exact) {
    AssocState* s = c->assoc;
    const int rows = s->staged_rows;
    if (rows < 1 || rows > s->n_rows) return set_err(c, PPG_ERR_ARG, "association: stage rows first (<= uploaded rows)");
    if (frames < 1 || frames > s->bcap) return set_err(c, PPG_ERR_ARG, "association: bad frame count");
    PPG_CUDA(c, cudaMemsetAsync(s->fallback, 0, 4, c->st));
    int rc = assoc_prep(c, src, frames);
    if (rc != PPG_OK) return rc;
    if (!force_exact && s->mode != 2) {
        GemmParams gp;
        gp.rows = rows;
        gp.max_rows = s->max_rows;
        gp.ncap = s->ncap;
        gp.frames = frames;
        gp.src = src;
        gp.rowp = s->rowp;
        gp.fn2 = s->fn2;
        gp.kinfo = s->kinfo;
        gp.cand = s->cand;
        gp.guard = s->guard;
        gp.e2_max = s->mode == 1 ? s->e2_max : 0.0;
        const int total = (rows + A_BM - 1) / A_BM * frames;
        const int grid = total < c->num_sms ? total : c->num_sms;
        assoc_gemm_kernel<<<grid, A_THREADS, gemm_smem(s), c->st>>>(s->mapA, s->mapB, gp);
        c->launches++;
        stage_mark(c, "assoc.gemm(top4)");
    }
    RescoreParams rp;
    rp.rows = rows;
    rp.max_rows = s->max_rows;
    rp.ncap = s->ncap;
    rp.src = src;
    rp.rowp = s->rowp;
    rp.map_f32 = s->map_f32;
    rp.kinfo = s->kinfo;
    rp.korder = s->korder;
    rp.cand = s->cand;
    rp.guard = s->guard;
    rp.nbmax = s->nbmax;
    rp.ratio = s->ratio;
    rp.th_high = c->cfg.th_high;
    rp.mode = s->mode;
    rp.max_dist = s->max_dist;
    rp.e2_max = s->mode == 1 ? s->e2_max : 0.0;
    rp.force_exact = (force_exact || s->mode == 2) ? 1 : 0;
    rp.row_node = s->row_node;
    rp.kp_node = s->kp_node;
    rp.best_idx = s->best_idx;
    rp.second_idx = s->second_idx;
    rp.best_d = s->best_d;
    rp.second_d = s->second_d;
    rp.accept = s->accept;
    rp.fallback = s->fallback;
    assoc_rescore_kernel<<<dim3((rows + 7) / 8, frames), 256, 0, c->st>>>(rp);
    c->launches++;
    stage_mark(c, "assoc.rescore");
    PPG_CUDA(c, cudaGetLastError());
    return PPG_OK;
}

// Results of `frames` slots -> pinned planes with ONE strided copy per array (the first version issued 5 copies per
// frame: 160 driver calls per batch-32 fetch); plane k of slot f starts at h_res + plane_off[k] + f * R * elem.
int fetch_slots(ppg_ctx* c, int frames) {
    AssocState* s = c->assoc;
    const size_t R = s->staged_rows, M = s->max_rows, F = frames, B = s->bcap;
    uint8_t* h = s->h_res;
    const void* src[5] = {s->best_idx, s->second_idx, s->best_d, s->second_d, s->accept};
    for (int k = 0; k < 5; k++) {
        const size_t es = k < 4 ? 4 : 1;
        PPG_CUDA(c, cudaMemcpy2DAsync(h + (size_t)k * B * M * 4, R * es, src[k], M * es, R * es, F,
                                      cudaMemcpyDeviceToHost, c->st));
    }
    return PPG_OK;
}

void unpack_slot(ppg_ctx* c, int slot, ppg_assoc_out* out) {
    AssocState* s = c->assoc;
    const size_t R = s->staged_rows, M = s->max_rows, B = s->bcap;
    const uint8_t* h = s->h_res;
    if (out->best_idx) memcpy(out->best_idx, h + 0 * B * M * 4 + (size_t)slot * R * 4, R * 4);
    if (out->second_idx) memcpy(out->second_idx, h + 1 * B * M * 4 + (size_t)slot * R * 4, R * 4);
    if (out->best_dist) memcpy(out->best_dist, h + 2 * B * M * 4 + (size_t)slot * R * 4, R * 4);
    if (out->second_dist) memcpy(out->second_dist, h + 3 * B * M * 4 + (size_t)slot * R * 4, R * 4);
    if (out->accept) memcpy(out->accept, h + 4 * B * M * 4 + (size_t)slot * R, R);
}

}  // namespace

void assoc_destroy(ppg_ctx* c) {
    AssocState* s = c->assoc;
    if (!s) return;
    void* bufs[] = {s->map_f32, s->map_bf, s->map_n2, s->kx, s->ky, s->fdesc, s->fn2, s->free_mask, s->ones, s->f_bf,
                    s->kinfo, s->korder, s->nbmax, s->proj, s->vcos, s->rowp, s->cand, s->guard,
                    s->best_idx, s->second_idx, s->best_d, s->second_d, s->accept, s->fallback,
                    s->wpos, s->nrm, s->dmin, s->dmax, s->poses, s->in_view, s->depth, s->row_node, s->kp_node};
    for (void* b : bufs)
        if (b) cudaFree(b);
    extend_destroy(s);
    if (s->h_res) cudaFreeHost(s->h_res);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    delete s;
    c->assoc = nullptr;
}

int assoc_stage_rows(ppg_ctx* c, int frames, int n_rows, const float* proj_uv, const float* view_cos, float th,
                      float ratio, bool pinned_src) {
    AssocState* s = c->assoc;
    if (n_rows < 1 || n_rows > s->n_rows || !proj_uv || !view_cos)
        return set_err(c, PPG_ERR_ARG, "association: n_rows must be in [1, uploaded rows]");
    if (frames < 1 || frames > s->bcap) return set_err(c, PPG_ERR_ARG, "association: bad frame count");
    const float* hp = proj_uv;
    const float* hv = view_cos;
    if (!pinned_src) {
        // the caller's arrays are pageable: one memcpy into pinned staging and one strided DMA per array instead of two
        // driver-staged copies per frame.  The previous batch's DMA out of the staging buffer must be done first.
        PPG_CUDA(c, cudaStreamSynchronize(c->st));
        float* sp = s->h_stage;
        float* sv = s->h_stage + (size_t)s->bcap * s->max_rows * 2;
        memcpy(sp, proj_uv, (size_t)frames * n_rows * 8);
        memcpy(sv, view_cos, (size_t)frames * n_rows * 4);
        hp = sp;
        hv = sv;
    }
    PPG_CUDA(c, cudaMemcpy2DAsync(s->proj, (size_t)s->max_rows * 8, hp, (size_t)n_rows * 8, (size_t)n_rows * 8, frames,
                                  cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpy2DAsync(s->vcos, (size_t)s->max_rows * 4, hv, (size_t)n_rows * 4, (size_t)n_rows * 4, frames,
                                  cudaMemcpyHostToDevice, c->st));
    s->staged_rows = n_rows;
    s->staged_frames = frames;
    s->th = th;
    s->ratio = ratio;
    s->mode = 0;  // ppg_assoc_stage overrides from ppg_assoc_in
    s->max_dist = 0.f;
    s->e2_max = 0.0;
    s->use_in_view = false;
    return PPG_OK;
}

}  // namespace ppg

using namespace ppg;

extern "C" {

int ppg_upload_map(ppg_ctx* c, const float* map_desc, int n_rows) {
    if (!c || !map_desc || n_rows < 1) return set_err(c, PPG_ERR_ARG, "ppg_upload_map: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    if (n_rows > s->max_rows) return set_err(c, PPG_ERR_ARG, "ppg_upload_map: more rows than max_map_points");
    PPG_CUDA(c, cudaMemcpyAsync(s->map_f32, map_desc, (size_t)n_rows * 1024, cudaMemcpyHostToDevice, c->st));
    prep_map_kernel<<<(n_rows + 7) / 8, 256, 0, c->st>>>(s->map_f32, s->map_bf, s->map_n2, n_rows, n_rows);
    c->launches++;
    PPG_CUDA(c, cudaGetLastError());
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    s->n_rows = n_rows;
    return PPG_OK;
}

int ppg_upload_map_geometry(ppg_ctx* c, const float* world_pos, const float* normal, const float* min_dist,
                            const float* max_dist, int n_rows) {
    if (!c || !world_pos || !normal || !min_dist || !max_dist || n_rows < 1)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_map_geometry: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    if (n_rows > s->max_rows) return set_err(c, PPG_ERR_ARG, "ppg_upload_map_geometry: more rows than max_map_points");
    PPG_CUDA(c, cudaMemcpyAsync(s->wpos, world_pos, (size_t)n_rows * 12, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(s->nrm, normal, (size_t)n_rows * 12, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(s->dmin, min_dist, (size_t)n_rows * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(s->dmax, max_dist, (size_t)n_rows * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    s->geo_rows = n_rows;
    return PPG_OK;
}

int ppg_assoc_stage_poses(ppg_ctx* c, int n_frames, int n_rows, const float* Rcw, const float* tcw, const float* Ow,
                          float cos_limit, float th, float ratio) {
    if (!c || !c->assoc || !Rcw || !tcw || !Ow) return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage_poses: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    AssocState* s = c->assoc;
    if (n_rows < 1 || n_rows > s->n_rows || n_rows > s->geo_rows)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage_poses: upload descriptors and geometry of all n_rows rows first");
    if (n_frames < 1 || n_frames > s->bcap) return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage_poses: bad frame count");
    // 64 bytes per frame from pageable memory: the driver stages such a copy before it returns, so no pinned buffer
    // (and no synchronisation with the previous batch) is needed
    std::vector<float> hv((size_t)n_frames * 16, 0.f);
    float* h = hv.data();
    for (int f = 0; f < n_frames; f++) {
        memcpy(h + 16 * f, Rcw + 9 * f, 36);
        memcpy(h + 16 * f + 9, tcw + 3 * f, 12);
        memcpy(h + 16 * f + 12, Ow + 3 * f, 12);
    }
    PPG_CUDA(c, cudaMemcpyAsync(s->poses, h, (size_t)n_frames * 64, cudaMemcpyHostToDevice, c->st));
    FrustumParam fp;
    fp.fx = c->cfg.K[0];
    fp.fy = c->cfg.K[4];
    fp.cx = c->cfg.K[2];
    fp.cy = c->cfg.K[5];
    fp.k0 = c->cfg.D[0];
    fp.k1 = c->cfg.D[1];
    fp.k2 = c->cfg.D[2];
    fp.k3 = c->cfg.D[3];
    fp.minX = (float)c->minX;
    fp.maxX = (float)c->maxX;
    fp.minY = (float)c->minY;
    fp.maxY = (float)c->maxY;
    fp.fisheye = c->cfg.fisheye;
    fp.cos_limit = cos_limit;
    frustum_kernel<<<dim3((n_rows + 255) / 256, n_frames), 256, 0, c->st>>>(s->wpos, s->nrm, s->dmin, s->dmax, s->poses,
                                                                             n_rows, s->max_rows, fp, s->proj, s->vcos,
                                                                             s->depth, s->in_view);
    c->launches++;
    PPG_CUDA(c, cudaGetLastError());
    stage_mark(c, "assoc.frustum");
    s->staged_rows = n_rows;
    s->staged_frames = n_frames;
    s->th = th;
    s->ratio = ratio;
    s->mode = 0;
    s->max_dist = 0.f;
    s->e2_max = 0.0;
    s->use_in_view = true;
    return PPG_OK;
}

int ppg_frustum_fetch(ppg_ctx* c, int n_frames, uint8_t* in_view, float* proj_uv, float* depth, float* view_cos) {
    if (!c || !c->assoc || n_frames < 1 || n_frames > c->assoc->staged_frames || !c->assoc->use_in_view)
        return set_err(c, PPG_ERR_ARG, "ppg_frustum_fetch: stage poses first");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    AssocState* s = c->assoc;
    const size_t R = s->staged_rows, M = s->max_rows, F = n_frames;
    if (in_view) PPG_CUDA(c, cudaMemcpy2DAsync(in_view, R, s->in_view, M, R, F, cudaMemcpyDeviceToHost, c->st));
    if (proj_uv) PPG_CUDA(c, cudaMemcpy2DAsync(proj_uv, R * 8, s->proj, M * 8, R * 8, F, cudaMemcpyDeviceToHost, c->st));
    if (depth) PPG_CUDA(c, cudaMemcpy2DAsync(depth, R * 4, s->depth, M * 4, R * 4, F, cudaMemcpyDeviceToHost, c->st));
    if (view_cos) PPG_CUDA(c, cudaMemcpy2DAsync(view_cos, R * 4, s->vcos, M * 4, R * 4, F, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    return PPG_OK;
}

int ppg_assoc_stage(ppg_ctx* c, const ppg_assoc_in* in) {
    if (!c || !in) return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    if (in->n_kp < 0 || in->n_kp > s->ncap) return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage: too many keypoints");
    if (in->mode == PPG_SEARCH_WINDOW && !in->view_cos)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage: view_cos must point to n_rows floats (ignored in mode 1)");
    if (in->mode == PPG_SEARCH_NODE) {
        if (!in->row_node || (in->n_kp > 0 && !in->kp_node) || in->n_rows < 1 || in->n_rows > s->n_rows)
            return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage: node mode needs row_node, kp_node and 1 <= n_rows <= uploaded rows");
        // no projections in this mode: the row parameters are prepared from zeros and never used
        std::vector<float> zeros((size_t)in->n_rows * 2, 0.f);
        if ((rc = assoc_stage_rows(c, 1, in->n_rows, zeros.data(), zeros.data(), in->th, in->ratio)) != PPG_OK) return rc;
        PPG_CUDA(c, cudaMemcpyAsync(s->row_node, in->row_node, (size_t)in->n_rows * 4, cudaMemcpyHostToDevice, c->st));
        if (in->n_kp > 0)
            PPG_CUDA(c, cudaMemcpyAsync(s->kp_node, in->kp_node, (size_t)in->n_kp * 4, cudaMemcpyHostToDevice, c->st));
    } else if ((rc = assoc_stage_rows(c, 1, in->n_rows, in->proj_uv, in->view_cos, in->th, in->ratio)) != PPG_OK)
        return rc;
    if (in->n_kp > 0 && in->kp_x && in->kp_y && in->frame_desc) {
        PPG_CUDA(c, cudaMemcpyAsync(s->kx, in->kp_x, (size_t)in->n_kp * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->ky, in->kp_y, (size_t)in->n_kp * 4, cudaMemcpyHostToDevice, c->st));
        PPG_CUDA(c, cudaMemcpyAsync(s->fdesc, in->frame_desc, (size_t)in->n_kp * 1024, cudaMemcpyHostToDevice, c->st));
        if (in->free_mask)
            PPG_CUDA(c, cudaMemcpyAsync(s->free_mask, in->free_mask, (size_t)in->n_kp, cudaMemcpyHostToDevice, c->st));
        else
            PPG_CUDA(c, cudaMemsetAsync(s->free_mask, 1, (size_t)in->n_kp, c->st));
    }
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    s->staged_n = in->n_kp;
    if (in->mode != PPG_SEARCH_EXTEND_MAP && in->mode != PPG_SEARCH_WINDOW && in->mode != PPG_SEARCH_NODE)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage: unknown search mode");
    s->mode = in->mode;
    s->max_dist = in->max_dist;
    s->e2_max = in->e2_max;
    return PPG_OK;
}

int ppg_assoc_stage_batch(ppg_ctx* c, int n_frames, int n_rows, const float* proj_uv, const float* view_cos, float th,
                          float ratio) {
    if (!c) return PPG_ERR_ARG;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    if ((rc = assoc_stage_rows(c, n_frames, n_rows, proj_uv, view_cos, th, ratio, false)) != PPG_OK) return rc;
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    return PPG_OK;
}

int ppg_assoc_stage_batch_async(ppg_ctx* c, int n_frames, int n_rows, const float* proj_uv, const float* view_cos,
                                float th, float ratio) {
    if (!c) return PPG_ERR_ARG;
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    cudaPointerAttributes a, b;
    if (!proj_uv || !view_cos || cudaPointerGetAttributes(&a, proj_uv) != cudaSuccess ||
        cudaPointerGetAttributes(&b, view_cos) != cudaSuccess || a.type != cudaMemoryTypeHost ||
        b.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_stage_batch_async: projections must lie in pinned host memory");
    }
    return assoc_stage_rows(c, n_frames, n_rows, proj_uv, view_cos, th, ratio, true);
}

int ppg_assoc_run(ppg_ctx* c) {
    if (!c || !c->assoc) return set_err(c, PPG_ERR_ARG, "ppg_assoc_run: nothing staged");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    return run_assoc(c, assoc_staged_src(c->assoc), 1, 0);
}

int ppg_assoc_run_frame(ppg_ctx* c, int frame) {
    if (!c || !c->assoc || frame < 0 || frame >= c->maxB)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_run_frame: bad frame or nothing staged");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    return run_assoc(c, assoc_extracted_src(c, frame), 1, 0);
}

int ppg_assoc_run_batch(ppg_ctx* c, int n_frames) {
    if (!c || !c->assoc || n_frames < 1 || n_frames > c->maxB || n_frames > c->assoc->staged_frames)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_run_batch: stage projections for every frame first");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    return run_assoc(c, assoc_extracted_src(c, 0), n_frames, 0);
}

int ppg_assoc_fetch(ppg_ctx* c, ppg_assoc_out* out) {
    if (!c || !c->assoc || !out) return set_err(c, PPG_ERR_ARG, "ppg_assoc_fetch: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = fetch_slots(c, 1);
    if (rc != PPG_OK) return rc;
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    unpack_slot(c, 0, out);
    return PPG_OK;
}

int ppg_assoc_fetch_batch(ppg_ctx* c, int n_frames, ppg_assoc_out* outs) {
    if (!c || !c->assoc || !outs || n_frames < 1 || n_frames > c->assoc->bcap)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_fetch_batch: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = fetch_slots(c, n_frames);
    if (rc != PPG_OK) return rc;
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    for (int f = 0; f < n_frames; f++) unpack_slot(c, f, &outs[f]);
    return PPG_OK;
}

int ppg_associate(ppg_ctx* c, const ppg_assoc_in* in, ppg_assoc_out* out) {
    int rc = ppg_assoc_stage(c, in);
    if (rc != PPG_OK) return rc;
    if ((rc = ppg_assoc_run(c)) != PPG_OK) return rc;
    return ppg_assoc_fetch(c, out);
}

static int distinctive_impl(ppg_ctx* c, const float* desc, const int32_t* offsets, int n_points, int32_t* best_idx,
                            bool to_table) {
    if (!c || !desc || !offsets || n_points < 1) return set_err(c, PPG_ERR_ARG, "distinctive descriptors: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    int rc = assoc_ensure_state(c);
    if (rc != PPG_OK) return rc;
    AssocState* s = c->assoc;
    if (to_table && n_points > s->max_rows)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_map_distinctive: more points than max_map_points");
    const size_t total = (size_t)offsets[n_points];
    if (offsets[0] != 0 || total < 1) return set_err(c, PPG_ERR_ARG, "distinctive descriptors: offsets must start at 0");
    float* d_desc = nullptr;
    int *d_off = nullptr, *d_best = nullptr, *d_err = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_desc);
        cudaFree(d_off);
        cudaFree(d_best);
        cudaFree(d_err);
    };
    cudaError_t e = dalloc(&d_desc, total * 256);
    if (e == cudaSuccess) e = dalloc(&d_off, (size_t)n_points + 1);
    if (e == cudaSuccess) e = dalloc(&d_best, (size_t)n_points);
    if (e == cudaSuccess) e = dalloc(&d_err, 1);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, desc, total * 1024, cudaMemcpyHostToDevice, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, offsets, ((size_t)n_points + 1) * 4, cudaMemcpyHostToDevice, c->st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_err, 0, 4, c->st);
    if (e == cudaSuccess) {
        constexpr int smem = (DD_MAX * DD_MAX + DD_MAX) * 4;
        static bool attr_done[64];
        static std::mutex attr_mu;
        e = once_per_device(attr_done, attr_mu, [] {
            return cudaFuncSetAttribute(distinctive_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        });
        if (e == cudaSuccess) {
            distinctive_kernel<<<n_points, 256, smem, c->st>>>(d_desc, d_off, n_points, d_best, d_err);
            c->launches++;
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess && to_table) {
        gather_rows_kernel<<<(n_points + 7) / 8, 256, 0, c->st>>>(d_desc, d_off, d_best, n_points, s->map_f32);
        prep_map_kernel<<<(n_points + 7) / 8, 256, 0, c->st>>>(s->map_f32, s->map_bf, s->map_n2, n_points, n_points);
        c->launches += 2;
        e = cudaGetLastError();
    }
    int h_err = 0;
    std::vector<int32_t> tmp;
    if (e == cudaSuccess && best_idx) e = cudaMemcpyAsync(best_idx, d_best, (size_t)n_points * 4, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, c->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
    cleanup();
    if (e != cudaSuccess) return set_err(c, PPG_ERR_CUDA, std::string("distinctive descriptors: ") + cudaGetErrorString(e));
    if (h_err) return set_err(c, PPG_ERR_CAPACITY, "a map point has more than PPG_MAX_OBSERVATIONS observations");
    if (to_table) s->n_rows = n_points;
    return PPG_OK;
}

int ppg_distinctive_descriptors(ppg_ctx* c, const float* desc, const int32_t* offsets, int n_points, int32_t* best_idx) {
    if (!best_idx) return set_err(c, PPG_ERR_ARG, "ppg_distinctive_descriptors: best_idx is null");
    return distinctive_impl(c, desc, offsets, n_points, best_idx, false);
}

int ppg_upload_map_distinctive(ppg_ctx* c, const float* desc, const int32_t* offsets, int n_points, int32_t* best_idx) {
    return distinctive_impl(c, desc, offsets, n_points, best_idx, true);
}

int ppg_assoc_fallback_rows(ppg_ctx* c, int* n) {
    if (!c || !c->assoc || !n) return set_err(c, PPG_ERR_ARG, "ppg_assoc_fallback_rows: null argument");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    PPG_CUDA(c, cudaMemcpy(n, c->assoc->fallback, 4, cudaMemcpyDeviceToHost));
    return PPG_OK;
}

int ppg_assoc_device_results(ppg_ctx* c, void** best_idx, void** second_idx, void** best_dist, void** second_dist,
                             void** accept) {
    if (!c || !c->assoc) return set_err(c, PPG_ERR_ARG, "ppg_assoc_device_results: nothing staged");
    AssocState* s = c->assoc;
    if (best_idx) *best_idx = s->best_idx;
    if (second_idx) *second_idx = s->second_idx;
    if (best_dist) *best_dist = s->best_d;
    if (second_dist) *second_dist = s->second_d;
    if (accept) *accept = s->accept;
    return PPG_OK;
}

}  // extern "C"
