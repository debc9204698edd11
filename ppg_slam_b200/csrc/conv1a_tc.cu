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
                                                               int H, int W, int total_tiles, int swap_lbo_sbo) {
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
        const uint32_t a_lbo = swap_lbo_sbo ? 128u : 2048u, a_sbo = swap_lbo_sbo ? 2048u : 128u;
        const uint32_t b_lbo = swap_lbo_sbo ? 128u : 1024u, b_sbo = swap_lbo_sbo ? 1024u : 128u;
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

}  // namespace

// -> false when the shape is not covered (the frame must be a whole number of 128-pixel tiles)
bool conv1a_tc_supported(int H, int W) { return ((long long)H * W) % 128 == 0; }

cudaError_t conv1a_tc_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                             cudaStream_t st) {
    // 56 KB of dynamic shared memory per CTA caps the residency at 4 CTAs per SM = 4 x 128 TMEM columns
    constexpr int SMEM = 56 * 1024;
    static bool attr = false;
    static int swap = 0;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv1a_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        if (const char* s = getenv("PPG_C1_SWAP")) swap = atoi(s);
        attr = true;
    }
    const int total = (int)((long long)B * H * W / 128);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = total < sms * 4 ? total : sms * 4;
    if (grid <= 0) return cudaSuccess;
    conv1a_tc_kernel<<<grid, C1_THREADS, SMEM, st>>>(gray, w, bias, out, H, W, total, swap);
    return cudaGetLastError();
}

}  // namespace ppg
