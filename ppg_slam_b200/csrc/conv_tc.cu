// tcgen05 implicit-GEMM convolution kernel -- see conv_tc.cuh for the design.
#include "conv_tc.cuh"
#include "once.cuh"
#include "ptx.cuh"

#include <stdlib.h>

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
    t.y0 = ty * (128 >> p.tile_w_log2);
    t.x0 = (r - ty * p.tiles_x) << p.tile_w_log2;
    return t;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}


// Epilogue of one 16-channel group of one output pixel (one thread = one pixel = one TMEM lane).
// XL / YL: lane-xor distance of the x / y neighbour inside the warp (for the fused 2x2 max-pool).
template <int XL, int YL>
__device__ __forceinline__ void epilogue_store(const ConvTcParams& p, const ConvBias& cb, const uint32_t (&r)[16],
                                               int c0, int n, int y, int x, bool inb, int lane) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        v[j] = __uint_as_float(r[j]) + cb.v[c0 + j];
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.mode == EPI_F16) {
        if (inb) {
            __half* o = reinterpret_cast<__half*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * p.out_ld + c0;
            uint4 u0 = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                  pack_half2(v[6], v[7]));
            uint4 u1 = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]), pack_half2(v[12], v[13]),
                                  pack_half2(v[14], v[15]));
            reinterpret_cast<uint4*>(o)[0] = u0;
            reinterpret_cast<uint4*>(o)[1] = u1;
        }
    } else if (p.mode == EPI_F16_POOL) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], XL));  // x neighbour
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], YL));  // y neighbour
        }
        const int Ho = p.H >> 1, Wo = p.W >> 1, yo = y >> 1, xo = x >> 1;
        if ((lane & (XL | YL)) == 0 && yo < Ho && xo < Wo) {
            __half* o = reinterpret_cast<__half*>(p.out) + ((size_t)(n * Ho + yo) * Wo + xo) * p.out_ld + c0;
            uint4 u0 = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                  pack_half2(v[6], v[7]));
            uint4 u1 = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]), pack_half2(v[12], v[13]),
                                  pack_half2(v[14], v[15]));
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
                                ((size_t)(n * Ho + 2 * y + i) * Wo + 2 * x + jj) * p.out_ld + (c0 >> 2);
                    *reinterpret_cast<uint2*>(o) = u;
                }
        }
    } else {  // EPI_F32
        if (inb) {
            float* o = reinterpret_cast<float*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * p.out_ld + c0;
#pragma unroll
            for (int g = 0; g < 4; g++)
                reinterpret_cast<float4*>(o)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
    }
}

// Epilogue of 64 channels of one output pixel (halo kernel).  No shared-memory traffic at all: the bias comes from
// the constant bank, and the fused 2x2 max-pool is a reduce-scatter over the 4 lanes of a pool window (48 SHFL per
// thread instead of 128): after the exchange with the x neighbour a lane keeps 32 of the 64 channels, after the y
// neighbour 16, and every lane stores the 16 channels it ends up with.  max commutes with the monotone
// x -> fp16(relu(x)), so pooling the fp32 sums (bias already added) first gives bit-identical results.
template <int XL, int YL>
__device__ __forceinline__ void epilogue64(const ConvTcParams& p, const ConvBias& cb, const uint32_t (&r)[64], int c0,
                                           int n, int y, int x, bool inb, int lane) {
    float v[64];
#pragma unroll
    for (int j = 0; j < 64; j++) v[j] = __uint_as_float(r[j]) + cb.v[c0 + j];
    if (p.mode == EPI_F16_POOL) {
        const bool hx = (lane & XL) != 0, hy = (lane & YL) != 0;
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float keep = hx ? v[32 + j] : v[j], send = hx ? v[j] : v[32 + j];
            a[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, XL));
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float keep = hy ? a[16 + j] : a[j], send = hy ? a[j] : a[16 + j];
            o[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, YL));
            if (p.relu) o[j] = fmaxf(o[j], 0.f);
        }
        const int Ho = p.H >> 1, Wo = p.W >> 1, yo = y >> 1, xo = x >> 1;
        if (yo < Ho && xo < Wo) {
            __half* dst = reinterpret_cast<__half*>(p.out) + ((size_t)(n * Ho + yo) * Wo + xo) * p.out_ld + c0 +
                          (hx ? 32 : 0) + (hy ? 16 : 0);
            reinterpret_cast<uint4*>(dst)[0] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]),
                                                          pack_half2(o[4], o[5]), pack_half2(o[6], o[7]));
            reinterpret_cast<uint4*>(dst)[1] = make_uint4(pack_half2(o[8], o[9]), pack_half2(o[10], o[11]),
                                                          pack_half2(o[12], o[13]), pack_half2(o[14], o[15]));
        }
        return;
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = fmaxf(v[j], 0.f);
    }
    if (!inb) return;
    if (p.mode == EPI_F16) {
        __half* dst = reinterpret_cast<__half*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * p.out_ld + c0;
#pragma unroll
        for (int g = 0; g < 8; g++)
            reinterpret_cast<uint4*>(dst)[g] =
                make_uint4(pack_half2(v[8 * g], v[8 * g + 1]), pack_half2(v[8 * g + 2], v[8 * g + 3]),
                           pack_half2(v[8 * g + 4], v[8 * g + 5]), pack_half2(v[8 * g + 6], v[8 * g + 7]));
    } else if (p.mode == EPI_F16_PS2) {
        // out[2y+i][2x+j][c] = in[y][x][4c + 2i + j]  (torch.pixel_shuffle(2)); 16 output channels per 64 inputs
        const int Ho = p.H * 2, Wo = p.W * 2;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                const int s4 = 2 * i + jj;
                __half* dst = reinterpret_cast<__half*>(p.out) +
                              ((size_t)(n * Ho + 2 * y + i) * Wo + 2 * x + jj) * p.out_ld + (c0 >> 2);
                reinterpret_cast<uint4*>(dst)[0] =
                    make_uint4(pack_half2(v[s4], v[4 + s4]), pack_half2(v[8 + s4], v[12 + s4]),
                               pack_half2(v[16 + s4], v[20 + s4]), pack_half2(v[24 + s4], v[28 + s4]));
                reinterpret_cast<uint4*>(dst)[1] =
                    make_uint4(pack_half2(v[32 + s4], v[36 + s4]), pack_half2(v[40 + s4], v[44 + s4]),
                               pack_half2(v[48 + s4], v[52 + s4]), pack_half2(v[56 + s4], v[60 + s4]));
            }
    } else {  // EPI_F32
        float* dst = reinterpret_cast<float*>(p.out) + ((size_t)(n * p.H + y) * p.W + x) * p.out_ld + c0;
#pragma unroll
        for (int g = 0; g < 16; g++)
            reinterpret_cast<float4*>(dst)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const ConvTcParams p, const __grid_constant__ ConvBias cb) {
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
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nsteps = p.taps * p.cin_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop (warp-uniform control flow and operands); only the TMA / MMA instructions
        // themselves sit under elect.sync.  With `if (lane == 0)` around the loop the compiler cannot prove the
        // operands uniform and wraps every UTMALDG / UTCHMMA in an R2UR + ELECT + BRA.U.ANY loop (~50 cycles per
        // MMA in the first version's SASS, more than the 32-cycle MMA itself).
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
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(&full[s], stage_bytes);
                        uint8_t* a = smem + (size_t)s * stage_bytes;
                        ptx::tma_load_4d(a, &mapA, &full[s], cc * 64, t.x0 + dx, t.y0 + dy, t.n);
                        ptx::tma_load_2d(a + CONV_A_BYTES, &mapB, &full[s], cc * 64, tap * p.N);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
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
                if (ptx::elect_one()) {
                    uint32_t a_addr = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                    uint64_t adesc = ptx::make_sw128_desc(a_addr);
                    uint64_t bdesc = ptx::make_sw128_desc(a_addr + CONV_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++)  // 4 x (K = 16) per 64-channel chunk: +32 B in the swizzle atom
                        ptx::umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((step | k) != 0));
                    ptx::umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
                    if (step == nsteps - 1) ptx::umma_commit(&tfull[acc]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (2 groups of 4 warps, TMEM lane quarter = warp & 3) =====================
        // Group g drains accumulator g, i.e. every second tile: one warp per SM sub-partition runs the dependent
        // ld -> bias -> pool -> store chain at low IPC, and with a single group that chain, not the MMAs, set the
        // tile period (tensor pipe 37 % active in the ncu capture).
        const int q = warp & 3, grp = (warp - 2) >> 2;
        const int row = q * 32 + lane, h = row >> p.tile_w_log2, w = row & ((1 << p.tile_w_log2) - 1);
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
            uint32_t acc = lt & 1, aph = (lt >> 1) & 1;
            if ((int)acc != grp) continue;
            TileCoord t = decode_tile(p, tile);
            const int y = t.y0 + h, x = t.x0 + w;
            const bool inb = (y < p.H) && (x < p.W);
            ptx::mbar_wait(&tfull[acc], aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
            if (p.mode == EPI_SOFTMAX_D2S) {
                // one thread = one coarse cell: all 65 logits are in its TMEM lane, so the softmax needs no exchange
                // and the 8 x 8 probabilities go straight to the full-resolution map (no logits round trip, no
                // separate depth-to-space kernel)
                uint32_t r[64], r2[16];
                ptx::tmem_ld64(taddr, r);
                ptx::tmem_ld16(taddr + 64, r2);
                ptx::tmem_ld_wait();
                float v[64];
                float m = __uint_as_float(r2[0]) + cb.v[64];  // dustbin channel
                const float dust = m;
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    v[j] = __uint_as_float(r[j]) + cb.v[j];
                    m = fmaxf(m, v[j]);
                }
                float sum = expf(dust - m);
#pragma unroll
                for (int j = 0; j < 64; j++) {
                    v[j] = expf(v[j] - m);
                    sum += v[j];
                }
#pragma unroll
                for (int j = 0; j < 64; j++) v[j] = v[j] / sum;
                const int Wf = p.W * 8, Hf = p.H * 8;
                if (inb) {
                    float* dst = reinterpret_cast<float*>(p.out) + ((size_t)t.n * Hf + 8 * y) * Wf + 8 * x;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float4* o = reinterpret_cast<float4*>(dst + (size_t)i * Wf);
                        o[0] = make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]);
                        o[1] = make_float4(v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
                    }
                }
                if (p.scan_cand) {
                    // threshold scan (PPGExtractor.cpp:168-176): `!(score < thresh)` as the reference's `continue`
                    // on `<`; pixels closer than R to the border can never be accepted nor suppress (:190-193)
                    const int R = p.scan_radius;
                    const size_t HWf = (size_t)Hf * Wf;
                    int c = 0, call = 0;
                    if (inb) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int py = 8 * y + i;
                            const bool yin = py >= R && py <= Hf - R - 1;
                            uint32_t bits = 0;  // 2 bits per pixel of the row
                            uint32_t b_lo = 0, b_hi = 0;  // one byte per pixel
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int px = 8 * x + j;
                                const bool pass = !(v[8 * i + j] < p.scan_thresh);
                                const bool keep = pass && yin && px >= R && px <= Wf - R - 1;
                                call += pass;
                                c += keep;
                                if (keep) {
                                    bits |= 1u << (2 * j);
                                    if (j < 4) b_lo |= 1u << (8 * j); else b_hi |= 1u << (8 * (j - 4));
                                }
                            }
                            const size_t pix0 = (size_t)py * Wf + 8 * x;
                            if (p.scan_state2)
                                *reinterpret_cast<uint16_t*>(p.scan_state2 + (size_t)t.n * (HWf / 4) + pix0 / 4) = (uint16_t)bits;
                            else
                                *reinterpret_cast<uint2*>(p.scan_state + (size_t)t.n * HWf + pix0) = make_uint2(b_lo, b_hi);
                        }
                    }
                    // warp-aggregated append (the whole warp is here: the mode test is uniform)
                    int inc = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, inc, d);
                        if (lane >= d) inc += o;
                    }
                    const int tot = __shfl_sync(0xffffffffu, inc, 31);
                    int call_w = call;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) call_w += __shfl_xor_sync(0xffffffffu, call_w, d);
                    int base = 0;
                    if (lane == 0) {
                        if (tot) base = atomicAdd(&p.scan_counters[t.n * 8 + 0], tot);
                        if (call_w) atomicAdd(&p.scan_counters[t.n * 8 + 1], call_w);
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (c) {
                        uint32_t* cl = p.scan_cand + (size_t)t.n * HWf + base + inc - c;
                        const int R2 = R;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int py = 8 * y + i;
                            const bool yin = py >= R2 && py <= Hf - R2 - 1;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int px = 8 * x + j;
                                if (!(v[8 * i + j] < p.scan_thresh) && yin && px >= R2 && px <= Wf - R2 - 1)
                                    *cl++ = (uint32_t)(py * Wf + px);
                            }
                        }
                    }
                }
            } else if ((p.N & 63) == 0) {
                // 64 columns per tcgen05.ld; the pool is the reduce-scatter of epilogue64 (48 SHFL per 64 channels
                // instead of 128 -- SHFL shares the shared-memory data pipe with the tensor core's operand reads)
                for (int c0 = 0; c0 < p.N; c0 += 64) {
                    uint32_t r[64];
                    ptx::tmem_ld64(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    epilogue64<1, 16>(p, cb, r, c0, t.n, y, x, inb, lane);
                }
            } else {
                for (int c0 = 0; c0 < p.N; c0 += 16) {
                    uint32_t r[16];
                    ptx::tmem_ld16(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    epilogue_store<1, 16>(p, cb, r, c0, t.n, y, x, inb, lane);
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

// ------------------------------------------------------------------------------------------------
// v2 "halo" kernel for 3x3 convolutions with Cin = 64 (conv1b, conv2a/b, conv3a, edge block 1).
// The generic kernel re-fetches the shifted A tile for each of the 9 taps (9 x 16 KB + 9 x N x 128 B per
// tile through TMA); its ncu capture shows it bound by that traffic (tensor pipe 19 % active).  Here
//   * all 9 weight taps stay resident in shared memory for the whole kernel (9 x N x 128 B),
//   * one TMA box per tile brings the (8+2) x (16+2) pixel halo, one 128-byte swizzled row per pixel,
//   * the A operand of tap (dy,dx) is a descriptor into that halo: start = pixel (dy+1, dx+1), rows of a
//     group = 8 consecutive pixels in x, group stride (SBO) = one halo row.
// Output tile = 8 (x) x 16 (y) pixels, TMEM lane = 8*y + x.
struct Conv2Smem {
    uint8_t* w;      // resident weights
    uint8_t* halo;   // S stages
    uint64_t *full, *empty, *tfull, *tempty, *wbar;
    uint32_t* tmem_slot;
    float* sbias;
};

// flags: bit 0 = issue the MMAs of two consecutive tiles interleaved (independent accumulators back to back).
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const ConvTcParams p, const __grid_constant__ ConvBias cb, const int halo_pitch,
                const int halo_stage_bytes, const int flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages;
    const uint32_t w_bytes = 9u * (uint32_t)p.N * 128u;
    uint8_t* sw = smem;
    uint8_t* shalo = smem + w_bytes;  // w_bytes is a multiple of 1024 (N % 8 == 0)
    uint64_t* full = reinterpret_cast<uint64_t*>(shalo + (size_t)S * halo_stage_bytes);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 4;
    uint64_t* wbar = tempty + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = p.N <= 64 ? 256u : 512u;  // four accumulators of N columns
    const bool pair_mode = (flags & 1) != 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 4; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 4);
        }
        ptx::mbar_init(wbar, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&mapA);
        ptx::prefetch_tmap(&mapB);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t halo_bytes = (uint32_t)halo_pitch * (CONV2_TILE_H + 2) * 128u;

    auto decode = [&](int tile, int& n, int& y0, int& x0) {
        const int per = p.tiles_x * p.tiles_y;
        n = tile / per;
        const int r = tile - n * per, ty = r / p.tiles_x;
        y0 = ty * CONV2_TILE_H;
        x0 = (r - ty * p.tiles_x) * CONV2_TILE_W;
    };
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; local index it -> halo stage it % S,
    // accumulator it & 3 (phase (it >> 2) & 1), epilogue group it & 1
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // warps 0 and 1: the whole warp runs the loop, one elected lane issues (see conv_tc_kernel)
    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(wbar, w_bytes);
            for (int tap = 0; tap < 9; tap++) ptx::tma_load_2d(sw + (size_t)tap * p.N * 128, &mapB, wbar, 0, tap * p.N);
        }
        __syncwarp();
        for (int it = 0; it < my_tiles; it++) {
            int n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, n, y0, x0);
            const uint32_t s = it % S, ph = (it / S) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(&full[s], halo_bytes);
                ptx::tma_load_4d(shalo + (size_t)s * halo_stage_bytes, &mapA, &full[s], 0, x0 - 1, y0 - 1, n);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            const uint32_t idesc = ptx::make_idesc_f16(128, p.N, 0);
            ptx::mbar_wait(wbar, 0);
            ptx::tc_fence_after();
            const uint32_t w_addr = ptx::smem_u32(sw);
            const int ntaps = 9;
            const int step = pair_mode ? 2 : 1;
            for (int it = 0; it < my_tiles; it += step) {
                const int cnt = (pair_mode && it + 1 < my_tiles) ? 2 : 1;
                uint32_t d_tmem[2], h_addr[2], st[2];
                for (int u = 0; u < cnt; u++) {
                    const uint32_t t = it + u, acc = t & 3, aph = (t >> 2) & 1;
                    st[u] = t % S;
                    ptx::mbar_wait(&tempty[acc], aph ^ 1);
                    ptx::mbar_wait(&full[st[u]], (t / S) & 1);
                    d_tmem[u] = tmem_base + acc * (uint32_t)p.N;
                    h_addr[u] = ptx::smem_u32(shalo + (size_t)st[u] * halo_stage_bytes);
                }
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                for (int tap = 0; tap < ntaps; tap++) {
                    const uint32_t toff = (uint32_t)((tap / 3) * halo_pitch + (tap % 3)) * 128u;
                    const uint64_t bdesc = ptx::make_sw128_desc(w_addr + (uint32_t)tap * p.N * 128u);
                    const uint64_t adesc0 = ptx::make_sw128_desc_ex(h_addr[0] + toff, (uint32_t)halo_pitch * 128u, 0u);
                    if (cnt == 2) {
                        // the two tiles' MMAs alternate: consecutive instructions accumulate into different TMEM
                        // regions, so neither waits for the previous one's accumulator update
                        const uint64_t adesc1 =
                            ptx::make_sw128_desc_ex(h_addr[1] + toff, (uint32_t)halo_pitch * 128u, 0u);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            ptx::umma_f16(d_tmem[0], adesc0 + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((tap | k) != 0));
                            ptx::umma_f16(d_tmem[1], adesc1 + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((tap | k) != 0));
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            ptx::umma_f16(d_tmem[0], adesc0 + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((tap | k) != 0));
                    }
                }
                for (int u = 0; u < cnt; u++) {
                    ptx::umma_commit(&empty[st[u]]);
                    ptx::umma_commit(&tfull[(it + u) & 3]);
                }
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3, grp = (warp - 2) >> 2;  // two epilogue groups, see conv_tc_kernel
        const int row = q * 32 + lane, h = row >> 3, w = row & 7;
        for (int it = grp; it < my_tiles; it += 2) {
            const uint32_t acc = it & 3, aph = (it >> 2) & 1;
            int n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, n, y0, x0);
            const int y = y0 + h, x = x0 + w;
            const bool inb = (y < p.H) && (x < p.W);
            ptx::mbar_wait(&tfull[acc], aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.N;
            for (int c0 = 0; c0 < p.N; c0 += 64) {
                uint32_t r[64];
                ptx::tmem_ld64(taddr + c0, r);
                ptx::tmem_ld_wait();
                epilogue64<1, 8>(p, cb, r, c0, n, y, x, inb, lane);
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
        ptx::tmem_dealloc(tmem_base, tmem_cols);
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
    p.bias = bias;
    p.out = out;
    p.out_ld = out_ld;
    p.scan_cand = nullptr;
    p.scan_counters = nullptr;
    p.scan_state2 = nullptr;
    p.scan_state = nullptr;
    p.scan_thresh = 0.f;
    p.scan_radius = 0;
    L.cin = cin;
    L.cout = cout_padded;
    p.tile_w_log2 = 4;
    L.v3 = 0;
    L.wgt = nullptr;
    // Kernel choice.  PPG_CONV_KERNEL (A/B comparison only; the settings differ in the fp32 summation order of the taps, i.e. in the last bits):
    //   8: as 7, and conv3a (64 -> 128) and edge block 0 (128 -> 256 + pixel shuffle) run on conv_t128.cu as well --
    //      correct but not faster (0.12 / 0.125 ms against 0.10 / 0.095 ms per 32 frames: with 36 instructions per tile the
    //      epilogue -- twice the two-byte stores for the pixel shuffle -- is no longer hidden behind the tensor core);
    //   unset / 7: as 6, convDb (1x1 256 -> 256, fp32 output: 0.086 against 0.105 ms, its stores are 128 contiguous bytes
    //      per warp instead of 32 scattered 16-byte pieces) on conv_t128.cu, and conv2a (64 -> 64, no pool) runs on the transposed kernel too: with both epilogue groups
    //      draining every tile it takes 0.21-0.23 ms per 32 frames against 0.24 ms on the halo kernel;
    //   6: as 5, and the 3x3 layers with Cin = 128 (plain or pooled fp16 output) run on the transposed kernel of conv_t128.cu instead of the generic one;
    //   5: as 3, and conv1a is computed inside conv1b's producer warps (api.cu, conv_t64.cu);
    //   3: transposed kernel (conv_t64.cu) for the pooled 3x3 64 -> 64 layers (conv1b, conv2b: 90 % / 87 % of the
    //      tensor pipe), halo kernel for the other Cin = 64 layers;
    //   4: transposed kernel for conv2a too;  2: halo kernel for every Cin = 64 layer (the round-1 configuration);
    //   1: generic kernel everywhere.
    // Measured on B200 (tools/conv_variants.py): UMMA applies the 128-byte swizzle XOR on absolute shared-memory
    // address bits, so a descriptor may start at any 128-byte row of a swizzled tile with base_offset = 0.
    int kmode = 7;
    if (const char* e = getenv("PPG_CONV_KERNEL")) kmode = atoi(e);
    if (kmode >= 3 && conv_t64_applies(cin, cout_padded, taps, mode) && (mode == EPI_F16_POOL || kmode == 4 || kmode >= 7)) {
        conv_t64_plan(L, maxB, H, W);
        return;
    }
    if (kmode >= 6 && conv_t128_applies(cin, cout_padded, taps, mode) &&
        (kmode >= 8 || (taps == 9 && cin == 128 && (mode == EPI_F16 || mode == EPI_F16_POOL)) ||
         (kmode >= 7 && taps == 1 && mode == EPI_F32))) {
        conv_t128_plan(L, maxB, H, W, 148);
        return;
    }
    L.v2 = (kmode >= 2 && taps == 9 && cin == 64 && cout_padded <= 128 && cout_padded % 64 == 0) ? 1 : 0;
    if (L.v2) {
        L.halo_pitch = CONV2_TILE_W + 2;
        L.flags = 1;  // interleaved tile pairs
        L.box_w = L.halo_pitch;
        L.box_h = CONV2_TILE_H + 2;
        p.tiles_x = (W + CONV2_TILE_W - 1) / CONV2_TILE_W;
        p.tiles_y = (H + CONV2_TILE_H - 1) / CONV2_TILE_H;
        p.total_tiles = maxB * p.tiles_x * p.tiles_y;
        const int stage = (L.halo_pitch * L.box_h * 128 + 1023) / 1024 * 1024;
        const int wbytes = 9 * cout_padded * 128;
        int S = (222 * 1024 - wbytes - 4096) / stage;
        if (S > 6) S = 6;
        p.stages = S;
        L.smem_bytes = wbytes + S * stage + 1024 + 25 * 8 + 16 + 2048 + 64;
        return;
    }
    L.halo_pitch = 0;
    L.flags = 0;
    // Tile shape: 16 x 8 (the fused 2x2 pool needs both neighbours inside one warp), or 32 x 4 when that covers the
    // map with fewer tiles (60 x 94 coarse map: 45 instead of 48 tiles per frame, 10 instead of 11 tile rounds per SM
    // at batch 32).
    p.tile_w_log2 = 4;
    if (mode != EPI_F16_POOL) {
        const long t16 = (long)((W + 15) / 16) * ((H + 7) / 8), t32 = (long)((W + 31) / 32) * ((H + 3) / 4);
        if (t32 < t16) p.tile_w_log2 = 5;
    }
    L.box_w = 1 << p.tile_w_log2;
    L.box_h = 128 >> p.tile_w_log2;
    p.tiles_x = (W + L.box_w - 1) / L.box_w;
    p.tiles_y = (H + L.box_h - 1) / L.box_h;
    p.total_tiles = maxB * p.tiles_x * p.tiles_y;
    const int stage_bytes = CONV_A_BYTES + cout_padded * 128;
    int S = (200 * 1024) / stage_bytes;
    if (S > 8) S = 8;
    p.stages = S;
    L.smem_bytes = S * stage_bytes + 1024 /*align*/ + 20 * 8 + 16 + 256 * 4 + 64;
}

cudaError_t conv_tc_launch(const ConvLayer& L, int batch, int num_sms, cudaStream_t st) {
    if (L.v3 == 3) return conv_t128_launch(L, batch, num_sms, st);
    if (L.v3) return conv_t64_launch(L, L.wgt, batch, num_sms, st);
    // one-time kernel attributes; a function-local static initialiser is thread-safe (contexts on several host
    // threads launch through here concurrently)
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    ConvTcParams p = L.p;
    p.B = batch;
    p.total_tiles = batch * p.tiles_x * p.tiles_y;
    int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    if (L.v2) {
        const int stage = (L.halo_pitch * L.box_h * 128 + 1023) / 1024 * 1024;
        conv_tc2_kernel<<<grid, CONV_THREADS, L.smem_bytes, st>>>(L.mapA, L.mapB, p, L.hb, L.halo_pitch, stage, L.flags);
    } else {
        conv_tc_kernel<<<grid, CONV_THREADS, L.smem_bytes, st>>>(L.mapA, L.mapB, p, L.hb);
    }
    return cudaGetLastError();
}

}  // namespace ppg
