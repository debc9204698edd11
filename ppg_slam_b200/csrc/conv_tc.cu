// tcgen05 implicit-GEMM convolution kernel -- see conv_tc.cuh for the design.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace ppg {

namespace {

struct TileCoord {
    int n, y0, x0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, int tile) {
    int per = p.tiles_x * p.tiles_y;
    TileCoord t;
    t.n = tile / per;
    int r = tile - t.n * per;
    int ty = r / p.tiles_x;
    t.y0 = ty * CONV_TILE_H;
    t.x0 = (r - ty * p.tiles_x) * CONV_TILE_W;
    return t;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const ConvTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles.
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages;
    const uint32_t stage_bytes = CONV_A_BYTES + (uint32_t)p.N * 128u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 4);  // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&mapA);
        ptx::prefetch_tmap(&mapB);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) sbias[i] = p.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nsteps = p.taps * p.cin_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                TileCoord t = decode_tile(p, tile);
                for (int tap = 0; tap < p.taps; tap++) {
                    int dy = 0, dx = 0;
                    if (p.taps == 9) {
                        dy = tap / 3 - 1;
                        dx = tap - (tap / 3) * 3 - 1;
                    }
                    for (int cc = 0; cc < p.cin_chunks; cc++, it++) {
                        uint32_t s = it % S, ph = (it / S) & 1;
                        ptx::mbar_wait(&empty[s], ph ^ 1);
                        ptx::mbar_expect_tx(&full[s], stage_bytes);
                        uint8_t* a = smem + (size_t)s * stage_bytes;
                        ptx::tma_load_4d(a, &mapA, &full[s], cc * 64, t.x0 + dx, t.y0 + dy, t.n);
                        ptx::tma_load_2d(a + CONV_A_BYTES, &mapB, &full[s], cc * 64, tap * p.N);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_f16(128, p.N, 0);
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
                uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
                ptx::mbar_wait(&tempty[acc], aph ^ 1);
                ptx::tc_fence_after();
                uint32_t d_tmem = tmem_base + acc * 256;
                for (int step = 0; step < nsteps; step++, it++) {
                    uint32_t s = it % S, ph = (it / S) & 1;
                    ptx::mbar_wait(&full[s], ph);
                    ptx::tc_fence_after();
                    uint32_t a_addr = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                    uint64_t adesc = ptx::make_sw128_desc(a_addr);
                    uint64_t bdesc = ptx::make_sw128_desc(a_addr + CONV_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++)  // 4 x (K = 16) per 64-channel chunk: +32 B in the swizzle atom
                        ptx::umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((step | k) != 0));
                    ptx::umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
                }
                ptx::umma_commit(&tfull[acc]);  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue (4 warps, TMEM lane quarter = warp & 3) =====================
        const int q = warp & 3;
        const int row = q * 32 + lane, h = row >> 4, w = row & 15;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
            uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
            TileCoord t = decode_tile(p, tile);
            const int y = t.y0 + h, x = t.x0 + w;
            const bool inb = (y < p.H) && (x < p.W);
            ptx::mbar_wait(&tfull[acc], aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
            for (int c0 = 0; c0 < p.N; c0 += 16) {
                uint32_t r[16];
                ptx::tmem_ld16(taddr + c0, r);
                ptx::tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    v[j] = __uint_as_float(r[j]) + sbias[c0 + j];
                    if (p.relu) v[j] = fmaxf(v[j], 0.f);
                }
                if (p.mode == EPI_F16) {
                    if (inb) {
                        __half* o = reinterpret_cast<__half*>(p.out) +
                                    ((size_t)(t.n * p.H + y) * p.W + x) * p.out_ld + c0;
                        uint4 u0 = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                              pack_half2(v[6], v[7]));
                        uint4 u1 = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]),
                                              pack_half2(v[12], v[13]), pack_half2(v[14], v[15]));
                        reinterpret_cast<uint4*>(o)[0] = u0;
                        reinterpret_cast<uint4*>(o)[1] = u1;
                    }
                } else if (p.mode == EPI_F16_POOL) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));   // x neighbour
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 16));  // y neighbour
                    }
                    const int Ho = p.H >> 1, Wo = p.W >> 1, yo = y >> 1, xo = x >> 1;
                    if ((lane & 17) == 0 && yo < Ho && xo < Wo) {
                        __half* o = reinterpret_cast<__half*>(p.out) +
                                    ((size_t)(t.n * Ho + yo) * Wo + xo) * p.out_ld + c0;
                        uint4 u0 = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                              pack_half2(v[6], v[7]));
                        uint4 u1 = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]),
                                              pack_half2(v[12], v[13]), pack_half2(v[14], v[15]));
                        reinterpret_cast<uint4*>(o)[0] = u0;
                        reinterpret_cast<uint4*>(o)[1] = u1;
                    }
                } else if (p.mode == EPI_F16_PS2) {
                    // out[2y+i][2x+j][c] = in[y][x][4c + 2i + j]  (torch.pixel_shuffle(2))
                    if (inb) {
                        const int Ho = p.H * 2, Wo = p.W * 2;
#pragma unroll
                        for (int i = 0; i < 2; i++)
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                const int o4 = 2 * i + jj;
                                uint2 u = make_uint2(pack_half2(v[o4], v[4 + o4]), pack_half2(v[8 + o4], v[12 + o4]));
                                __half* o = reinterpret_cast<__half*>(p.out) +
                                            ((size_t)(t.n * Ho + 2 * y + i) * Wo + 2 * x + jj) * p.out_ld + (c0 >> 2);
                                *reinterpret_cast<uint2*>(o) = u;
                            }
                    }
                } else {  // EPI_F32
                    if (inb) {
                        float* o = reinterpret_cast<float*>(p.out) +
                                   ((size_t)(t.n * p.H + y) * p.W + x) * p.out_ld + c0;
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            reinterpret_cast<float4*>(o)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

void conv_tc_plan(ConvLayer& L, int maxB, int H, int W, int cin, int cout_padded, int taps, int mode, int relu,
                  const float* bias, void* out, int out_ld) {
    ConvTcParams& p = L.p;
    p.B = maxB;
    p.H = H;
    p.W = W;
    p.cin_chunks = cin / 64;
    p.taps = taps;
    p.N = cout_padded;
    p.mode = mode;
    p.relu = relu;
    p.tiles_x = (W + CONV_TILE_W - 1) / CONV_TILE_W;
    p.tiles_y = (H + CONV_TILE_H - 1) / CONV_TILE_H;
    p.total_tiles = maxB * p.tiles_x * p.tiles_y;
    p.bias = bias;
    p.out = out;
    p.out_ld = out_ld;
    const int stage_bytes = CONV_A_BYTES + cout_padded * 128;
    int S = (200 * 1024) / stage_bytes;
    if (S > 8) S = 8;
    p.stages = S;
    L.smem_bytes = S * stage_bytes + 1024 /*align*/ + 20 * 8 + 16 + 256 * 4 + 64;
    L.cin = cin;
    L.cout = cout_padded;
}

cudaError_t conv_tc_launch(const ConvLayer& L, int batch, int num_sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ConvTcParams p = L.p;
    p.B = batch;
    p.total_tiles = batch * p.tiles_x * p.tiles_y;
    int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    conv_tc_kernel<<<grid, CONV_THREADS, L.smem_bytes, st>>>(L.mapA, L.mapB, p);
    return cudaGetLastError();
}

}  // namespace ppg
