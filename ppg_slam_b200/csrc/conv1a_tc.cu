// conv1a (u8 -> x/255 -> conv3x3 1->64 + bias + ReLU, PPGExtractor.cpp:151-152) as a K = 16 tcgen05 GEMM.
//
// The layer is 9 MACs per output: as fp32 FMAs it is issue bound at 0.40 ms per 32 frames against an HBM-write
// floor of 0.23 ms (1.48 GB of fp16 output), and legacy mma.sync is even slower on B200.  On the
// 5th-gen tensor cores the whole layer is two M128 N64 K16 instructions per 128 pixels:
//   A [128 pixels][16]: the 9 taps of the pixel as EXACT u8 integers in fp16, a 1.0 in column 9, zeros above --
//     built by the 128 worker threads (one pixel each) straight into the no-swizzle K-major core-matrix layout;
//   B [64 channels][16]: tap weights / 255 and the bias (column 9) as fp16 hi + lo halves -> two accumulating MMAs
//     (|w| reaches 197, so a single fp16 rounding of the weights would not do, SURVEY s.7);
//   fp32 accumulation in TMEM: the accumulator is the pre-activation, ~1e-7 relative from the fp32 FMA chain.
// The same 128 threads then read their TMEM lane, apply ReLU and store the pixel's 128-byte fp16 row (a warp
// writes 4 KB contiguous).  Warp 4 issues the MMAs.  Four CTAs per SM (128 TMEM columns each) hide the
// build -> MMA -> epilogue latency of one another.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "net_direct.cuh"
#include "once.cuh"
#include "ptx.cuh"

namespace ppg {

namespace {

constexpr int C1_THREADS = 160;
constexpr int C1_A_BYTES = 4096;  // [2 k-cores][128 rows][16 B]
constexpr int C1_B_BYTES = 2048;  // [2 k-cores][64 rows][16 B]

// No-swizzle K-major operand: core matrix = 8 rows x 16 bytes, contiguous (128 B); `lbo` = byte distance between the
// core matrices adjacent in K, `sbo` = between 8-row groups along M / N.
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version (sm_100)
    return d;                // layout type 0 = no swizzle
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(C1_THREADS) conv1a_tc_kernel(const uint8_t* __restrict__ gray, const float* __restrict__ w,
                                                               const float* __restrict__ bias, __half* __restrict__ out,
                                                               int H, int W, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t c1_smem[];
    uint8_t* sA = c1_smem;                       // 2 stages
    uint8_t* sBhi = c1_smem + 2 * C1_A_BYTES;
    uint8_t* sBlo = sBhi + C1_B_BYTES;
    uint64_t* afull = reinterpret_cast<uint64_t*>(sBlo + C1_B_BYTES);
    uint64_t* tfull = afull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
    uint8_t* sStage = c1_smem + 16384;  // 4 warps x 4 KB output staging
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    const int HW = H * W;

    if (t == 0) {
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&afull[s], 128);
            ptx::mbar_init(&tfull[s], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 4) {
        ptx::tmem_alloc(tmem_slot, 128);
        ptx::tmem_relinquish();
    }
    if (t < 64) {  // weights of channel t: k-core 0 = taps 0..7, k-core 1 = tap 8, bias, zeros
        float v[16];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = w[t * 9 + k] * (1.0f / 255.0f);
        v[9] = bias[t];
#pragma unroll
        for (int k = 10; k < 16; k++) v[k] = 0.f;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const __half h0 = __float2half_rn(v[2 * k]), h1 = __float2half_rn(v[2 * k + 1]);
            const __half l0 = __float2half_rn(v[2 * k] - __half2float(h0));
            const __half l1 = __float2half_rn(v[2 * k + 1] - __half2float(h1));
            hi[k] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            lo[k] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(sBhi + t * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sBhi + 1024 + t * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(sBlo + t * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(sBlo + 1024 + t * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 4) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0);
        const uint32_t a_lbo = 2048u, a_sbo = 128u;
        const uint32_t b_lbo = 1024u, b_sbo = 128u;
        const uint64_t bhi = make_nosw_desc(ptx::smem_u32(sBhi), b_lbo, b_sbo);
        const uint64_t blo = make_nosw_desc(ptx::smem_u32(sBlo), b_lbo, b_sbo);
        for (int i = 0; i < my_tiles; i++) {
            const int s = i & 1;
            ptx::mbar_wait(&afull[s], (i >> 1) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t adesc = make_nosw_desc(ptx::smem_u32(sA + s * C1_A_BYTES), a_lbo, a_sbo);
                ptx::umma_f16(tmem_base + s * 64, adesc, bhi, idesc, 0u);
                ptx::umma_f16(tmem_base + s * 64, adesc, blo, idesc, 1u);
                ptx::umma_commit(&tfull[s]);
            }
            __syncwarp();
        }
    } else {
        // ===================== workers: thread t <-> pixel t of the tile <-> TMEM lane t =====================
        // software pipeline per iteration i: store A(i) from the taps fetched one iteration ago, fetch the taps of
        // tile i+1 (global-load latency then overlaps the epilogue), drain the accumulator of tile i-1
        auto fetch = [&](int i, uint32_t (&u)[9]) {
#pragma unroll
            for (int k = 0; k < 9; k++) u[k] = 0u;
            if (i >= my_tiles) return;
            const long long gp = (long long)(blockIdx.x + i * gridDim.x) * 128 + t;
            const int n = (int)(gp / HW), pix = (int)(gp - (long long)n * HW);
            const int y = pix / W, x = pix - y * W;
            const uint8_t* g = gray + (size_t)n * HW;
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) u[k] = __ldg(g + (size_t)yy * W + xx);
            }
        };
        uint32_t u[9];
        fetch(0, u);
        for (int i = 0; i <= my_tiles; i++) {
            if (i < my_tiles) {
                const int s = i & 1;
                uint32_t hk[9];
#pragma unroll
                for (int k = 0; k < 9; k++) hk[k] = (uint32_t)__half_as_ushort(__ushort2half_rn((unsigned short)u[k]));
                uint8_t* a = sA + s * C1_A_BYTES;
                *reinterpret_cast<uint4*>(a + t * 16) = make_uint4(hk[0] | (hk[1] << 16), hk[2] | (hk[3] << 16),
                                                                   hk[4] | (hk[5] << 16), hk[6] | (hk[7] << 16));
                *reinterpret_cast<uint4*>(a + 2048 + t * 16) = make_uint4(hk[8] | 0x3C000000u, 0u, 0u, 0u);  // tap 8, 1.0
                ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
                ptx::tc_fence_before();     // orders this thread's earlier tcgen05.ld of the accumulator being reused
                ptx::mbar_arrive(&afull[s]);
                fetch(i + 1, u);
            }
            if (i > 0) {
                const int j = i - 1, s = j & 1;
                ptx::mbar_wait(&tfull[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                uint32_t r[64];
                ptx::tmem_ld64(tmem_base + ((uint32_t)(warp * 32) << 16) + s * 64, r);
                ptx::tmem_ld_wait();
                // Each thread holds one pixel's 128-byte row; stored directly, every STG.128 would touch 32 different
                // lines (1024 LSU wavefronts per tile -- the first version ran at 0.52 ms on exactly that).  The rows
                // go through a per-warp 4 KB staging tile (16-byte chunks XOR-swizzled by the row) so that one store
                // instruction writes 512 contiguous bytes = 4 whole pixel rows.
                uint8_t* stg = sStage + warp * 4096;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    uint4 v;
                    v.x = pack_h2(fmaxf(__uint_as_float(r[8 * q + 0]), 0.f), fmaxf(__uint_as_float(r[8 * q + 1]), 0.f));
                    v.y = pack_h2(fmaxf(__uint_as_float(r[8 * q + 2]), 0.f), fmaxf(__uint_as_float(r[8 * q + 3]), 0.f));
                    v.z = pack_h2(fmaxf(__uint_as_float(r[8 * q + 4]), 0.f), fmaxf(__uint_as_float(r[8 * q + 5]), 0.f));
                    v.w = pack_h2(fmaxf(__uint_as_float(r[8 * q + 6]), 0.f), fmaxf(__uint_as_float(r[8 * q + 7]), 0.f));
                    *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = v;
                }
                __syncwarp();
                const long long gp0 = (long long)(blockIdx.x + j * gridDim.x) * 128 + warp * 32;  // first pixel of the warp
                uint4* o = reinterpret_cast<uint4*>(out + gp0 * 64);
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int pr = 4 * it + (lane >> 3), c = lane & 7;  // pixel row of the warp tile, 16-byte chunk
                    o[it * 32 + lane] = *reinterpret_cast<const uint4*>(stg + pr * 128 + ((c ^ (pr & 7)) << 4));
                }
                __syncwarp();
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 128);
    }
}

// ------------------------------------------------------------------------------------------------
// Edge decoder tail on tcgen05: conv3x3 16->16 (BN folded) + ReLU + pixel_shuffle(2) + conv1x1 4->2 + softmax[:,1]
// (EdgeHeatmap.pt tail, PPGExtractor.cpp:154, :242).  The mma.sync version spent ~40 cycles per legacy HMMA
// (0.12 ms per 32 frames for 4.3 MB/frame of traffic).  Here one output tile of 8 x 16 low-res pixels is nine
// M128 N16 K16 instructions whose A operands are windows into ONE halo tile, as in conv_tc2_kernel, but in the
// no-swizzle layout: the halo is stored as two planes (channels 0-7 / 8-15) of [18 x 10 pixels][16 B], so the 8
// pixels of an MMA row group are one contiguous 128-byte core matrix for ANY tap offset (start address = 16-byte
// granular), the next row group is one halo row further (SBO = 160 B) and the second K half one plane further
// (LBO = 2880 B).  Worker threads load the halo (3 x 16 B each), the MMA warp issues, the workers' epilogue does
// bias/ReLU/pixel-shuffle/1x1/softmax for their pixel and writes the 2 x 2 full-resolution scores.
constexpr int ET_THREADS = 160, ET_TW = 8, ET_TH = 16, ET_HP = (ET_TW + 2) * (ET_TH + 2);  // 180 halo pixels
constexpr int ET_PLANE = ET_HP * 16, ET_A_BYTES = 2 * ET_PLANE;                              // 2880, 5760

__global__ void __launch_bounds__(ET_THREADS) edge_tail_tc_kernel(const __half* __restrict__ in, const float* __restrict__ w3,
                                                                  const float* __restrict__ b3, const float* __restrict__ w1,
                                                                  const float* __restrict__ b1, float* __restrict__ heat,
                                                                  int Hh, int Wh, int tiles_x, int tiles_y, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t et_smem[];
    uint8_t* sA = et_smem;                    // 2 stages x 5760 B
    uint8_t* sB = et_smem + 2 * ET_A_BYTES;   // 9 taps x [2 k-cores][16 rows][16 B] = 9 x 512 B
    uint64_t* afull = reinterpret_cast<uint64_t*>(sB + 9 * 512);
    uint64_t* tfull = afull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
    const int warp = threadIdx.x >> 5, t = threadIdx.x;

    if (t == 0) {
        for (int s = 0; s < 2; s++) {
            ptx::mbar_init(&afull[s], 128);
            ptx::mbar_init(&tfull[s], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 4) {
        ptx::tmem_alloc(tmem_slot, 32);
        ptx::tmem_relinquish();
    }
    // weights: w3 [16 co][9 taps][16 ci] fp32 -> per tap [k-core][co][8 ci] fp16
    for (int i = t; i < 9 * 2 * 16; i += ET_THREADS) {
        const int tap = i / 32, kc = (i >> 4) & 1, co = i & 15;
        const float* src = w3 + ((size_t)co * 9 + tap) * 16 + kc * 8;
        *reinterpret_cast<uint4*>(sB + tap * 512 + kc * 256 + co * 16) =
            make_uint4(pack_h2(src[0], src[1]), pack_h2(src[2], src[3]), pack_h2(src[4], src[5]), pack_h2(src[6], src[7]));
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto decode = [&](int i, int& n, int& y0, int& x0) {
        const int tile = blockIdx.x + i * gridDim.x, per = tiles_x * tiles_y;
        n = tile / per;
        const int r = tile - n * per, ty = r / tiles_x;
        y0 = ty * ET_TH;
        x0 = (r - ty * tiles_x) * ET_TW;
    };

    if (warp == 4) {
        const uint32_t idesc = ptx::make_idesc_f16(128, 16, 0);
        const uint32_t b_addr = ptx::smem_u32(sB);
        for (int i = 0; i < my_tiles; i++) {
            const int s = i & 1;
            ptx::mbar_wait(&afull[s], (i >> 1) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t a_addr = ptx::smem_u32(sA + s * ET_A_BYTES);
#pragma unroll
                for (int tap = 0; tap < 9; tap++) {
                    const uint64_t adesc =
                        make_nosw_desc(a_addr + (uint32_t)((tap / 3) * (ET_TW + 2) + tap % 3) * 16u, ET_PLANE, (ET_TW + 2) * 16);
                    const uint64_t bdesc = make_nosw_desc(b_addr + tap * 512, 256, 128);
                    ptx::umma_f16(tmem_base + s * 16, adesc, bdesc, idesc, (uint32_t)(tap != 0));
                }
                ptx::umma_commit(&tfull[s]);
            }
            __syncwarp();
        }
    } else {
        float bias_r[16], s1[8];
#pragma unroll
        for (int c = 0; c < 16; c++) bias_r[c] = b3[c];
#pragma unroll
        for (int c = 0; c < 8; c++) s1[c] = w1[c];
        const float sb10 = b1[0], sb11 = b1[1];
        const int H = Hh * 2, W = Wh * 2;
        // halo chunks of this thread: c = t, t + 128, t + 256 (< 360): plane c / 180, halo pixel c % 180
        auto fetch = [&](int i, uint4 (&v)[3]) {
#pragma unroll
            for (int e = 0; e < 3; e++) v[e] = make_uint4(0u, 0u, 0u, 0u);
            if (i >= my_tiles) return;
            int n, y0, x0;
            decode(i, n, y0, x0);
            const __half* src = in + (size_t)n * Hh * Wh * 16;
#pragma unroll
            for (int e = 0; e < 3; e++) {
                const int c = t + e * 128;
                if (c < 2 * ET_HP) {
                    const int plane = c / ET_HP, hp = c - plane * ET_HP;
                    const int hy = hp / (ET_TW + 2), hx = hp - hy * (ET_TW + 2);
                    const int y = y0 - 1 + hy, x = x0 - 1 + hx;
                    if (y >= 0 && y < Hh && x >= 0 && x < Wh)
                        v[e] = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y * Wh + x) * 16 + plane * 8));
                }
            }
        };
        uint4 hv[3];
        fetch(0, hv);
        for (int i = 0; i <= my_tiles; i++) {
            if (i < my_tiles) {
                const int s = i & 1;
                uint8_t* a = sA + s * ET_A_BYTES;
#pragma unroll
                for (int e = 0; e < 3; e++) {
                    const int c = t + e * 128;
                    if (c < 2 * ET_HP) *reinterpret_cast<uint4*>(a + c * 16) = hv[e];  // plane-major = chunk index order
                }
                ptx::fence_proxy_async();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&afull[s]);
                fetch(i + 1, hv);
            }
            if (i > 0) {
                const int j = i - 1, s = j & 1;
                ptx::mbar_wait(&tfull[s], (j >> 1) & 1);
                ptx::tc_fence_after();
                uint32_t r[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + s * 16, r);
                ptx::tmem_ld_wait();
                int n, y0, x0;
                decode(j, n, y0, x0);
                const int y = y0 + (t >> 3), x = x0 + (t & 7);  // TMEM lane = 8 * row + column of the tile
                if (y < Hh && x < Wh) {
                    float* dst = heat + (size_t)n * H * W;
#pragma unroll
                    for (int ii = 0; ii < 2; ii++) {
                        float hvv[2];
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            // pixel_shuffle(2): full-res channel c at (2y+ii, 2x+jj) = low-res channel 4c + 2ii + jj
                            float l0 = sb10, l1 = sb11;
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const int ch = 4 * c + 2 * ii + jj;
                                const float v = fmaxf(__uint_as_float(r[ch]) + bias_r[ch], 0.f);
                                l0 = fmaf(s1[c], v, l0);
                                l1 = fmaf(s1[4 + c], v, l1);
                            }
                            const float m = fmaxf(l0, l1);
                            const float e0 = expf(l0 - m), e1 = expf(l1 - m);
                            hvv[jj] = e1 / (e0 + e1);  // softmax(dim=1)[:,1], PPGExtractor.cpp:242
                        }
                        *reinterpret_cast<float2*>(dst + (size_t)(2 * y + ii) * W + 2 * x) = make_float2(hvv[0], hvv[1]);
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 32);
    }
}

}  // namespace

cudaError_t conv1a_tc_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                             cudaStream_t st) {
    // 56 KB of dynamic shared memory per CTA caps the residency at 4 CTAs per SM = 4 x 128 TMEM columns
    constexpr int SMEM = 56 * 1024;
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        return cudaFuncSetAttribute(conv1a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    });
    if (attr_err != cudaSuccess) return attr_err;
    if (((long long)H * W) % 128 != 0) return cudaErrorInvalidValue;  // whole 128-pixel tiles (W, H multiples of 16)
    const int total = (int)((long long)B * H * W / 128);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = total < sms * 4 ? total : sms * 4;
    if (grid <= 0) return cudaSuccess;
    conv1a_tc_kernel<<<grid, C1_THREADS, SMEM, st>>>(gray, w, bias, out, H, W, total);
    return cudaGetLastError();
}

cudaError_t edge_tail_tc_launch(const __half* in, const float* w3, const float* b3, const float* w1, const float* b1,
                                float* heat, int B, int Hh, int Wh, cudaStream_t st) {
    // 28 KB of dynamic shared memory per CTA -> 8 CTAs per SM (32 TMEM columns each)
    constexpr int SMEM = 28 * 1024;
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        return cudaFuncSetAttribute(edge_tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    });
    if (attr_err != cudaSuccess) return attr_err;
    const int tiles_x = (Wh + ET_TW - 1) / ET_TW, tiles_y = (Hh + ET_TH - 1) / ET_TH, total = tiles_x * tiles_y * B;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = total < sms * 8 ? total : sms * 8;
    if (grid <= 0) return cudaSuccess;
    edge_tail_tc_kernel<<<grid, ET_THREADS, SMEM, st>>>(in, w3, b3, w1, b1, heat, Hh, Wh, tiles_x, tiles_y, total);
    return cudaGetLastError();
}

}  // namespace ppg
