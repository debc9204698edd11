// Transposed tcgen05 convolution for the 3x3, Cin = Cout = 64 layers (conv1b, conv2a, conv2b: 1.4 of the 2.8 ms of
// convolution time per 32 EuRoC frames) -- replaces the cuDNN convolutions LibTorch runs for net/Backbone.pt
// (feature/src/PPGExtractor.cpp:152).
//
// Why: with pixels as the M operand (conv_tc2_kernel) every M128 N64 K16 instruction reads 4 KB of pixels + 2 KB of
// weights from shared memory for 32 cycles of math; measured on the B200 (tools/ts_probe.cu) such an instruction takes
// 57 cycles however it is issued, so those layers sat at 43 % of the tensor pipe.  Here the GEMM is transposed:
//     D[cout, pixel] += W_tap[cout, cin] * X[pixel + shift(tap), cin]
//   * A = weights, resident in TENSOR MEMORY for the whole kernel (tcgen05.mma with the A operand in TMEM): M = 128 rows
//     = 64 output channels x 2 taps;
//   * B = N = 160 consecutive pixels of the halo tile in shared memory (one TMA box per tile, 128-byte swizzled rows;
//     the window of a tap is just another start row of the same tile); 5 KB per instruction for 80 cycles of math:
//     measured 83 cycles per instruction = 96 % of the instruction's floor.
// Two taps share an instruction when their windows can share the B operand: the row halves are the taps (dy,-1) and
// (dy,0), whose accumulator columns then belong to output pixels that differ by one: out(n) = H0[n] + H1[n+1].  The three
// taps (dy,+1) have no partner (their second row half holds zero weights), so a tile takes 6 x 4 instead of 4.5 x 4
// instructions -- 75 % useful MMA work at 96 % of the floor instead of 100 % useful at 47 %.
//
// Tile = 38 x 4 output pixels, halo 40 x 6 (30 KB per stage + 1 KB of zero rows, 4 stages), accumulator column n = 40*oy + ox (158 used
// of 160).  TMEM: 192 columns of weights (6 groups x 4 K-steps x 8 columns) + two accumulators of 160 columns = 512.
// Row m of the A operand: lane 32q + 16h + r = tap half h, output channel 16q + r, so that the two partial sums of a
// channel sit 16 lanes apart in the same warp: the epilogue adds them with one shuffle per two outputs (each lane
// sends the half the other lane keeps), pools 2x2 in registers (a lane holds whole pool windows), adds the bias,
// applies ReLU and stores fp16 NHWC (16 lanes = 16 consecutive channels = one 32-byte sector per pixel).
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer (straight-line issue code: the probe showed that
// every instruction between two tcgen05.mma adds to the tile period), warps 2-5 / 6-9 epilogue of even / odd tiles.
#include "conv_tc.cuh"
#include "once.cuh"
#include "ptx.cuh"

namespace ppg {

namespace {

constexpr int T_TW = CONVT_TILE_W, T_TH = CONVT_TILE_H, T_HW = T_TW + 2, T_HH = T_TH + 2;
constexpr int T_N = 160;                                 // UMMA N: (T_TH - 1) * T_HW + T_TW + 1 = 159 columns used
constexpr int T_STAGES = 4;
constexpr int T_STAGE_BYTES = T_HW * T_HH * 128;         // 30720 bytes per TMA box
// The windows of the (dy,+1) taps start two pixels further and so read two rows past the box.  Their columns 158 / 159
// are junk for tap half 0, but the zero weights of half 1 turn a NaN / Inf bit pattern there into a NaN in H1[158], which
// IS used (0 * NaN): every stage therefore owns eight more rows, zeroed once, that TMA never writes.
constexpr int T_STAGE_STRIDE = T_STAGE_BYTES + 1024;     // a multiple of 1024 (swizzle pattern period)
constexpr uint32_t T_WCOLS = 6 * 4 * 8;                  // weights: 192 TMEM columns
constexpr uint32_t T_ACC0 = T_WCOLS;                     // accumulator a at column T_ACC0 + a * T_N
static_assert((T_TH - 1) * T_HW + T_TW + 1 <= T_N, "accumulator too narrow");
static_assert(T_WCOLS + 2 * T_N <= 512, "TMEM budget");
static_assert(T_STAGE_STRIDE % 1024 == 0, "halo stage must keep the swizzle phase");
static_assert(T_TW % 2 == 0 && T_TH % 2 == 0 && T_HW % 4 == 0, "pool windows / 4-column groups");

__device__ __forceinline__ uint16_t half_bits(float v) { return __half_as_ushort(__float2half_rn(v)); }

// The 24 MMAs of one tile; STAGE / ACC are compile-time so that every descriptor is `runtime base + constant`.
// group g = 2 * (dy + 1) + u: u = 0 the tap pair (dy,-1) | (dy,0), B window starts at halo row (dy+1) * T_HW;
//                             u = 1 the single tap (dy,+1), window starts two pixels further.
template <int STAGE, int ACC>
__device__ __forceinline__ void issue_tile(uint32_t tmem_base, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    const uint32_t d = tmem_base + T_ACC0 + ACC * T_N;
#pragma unroll
    for (int g = 0; g < 6; g++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t off = (uint32_t)(STAGE * T_STAGE_STRIDE + ((g >> 1) * T_HW + ((g & 1) ? 2 : 0)) * 128 + k * 32);
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (off >> 4));
            ptx::umma_ts_f16(d, tmem_base + (uint32_t)(g * 4 + k) * 8u, bdesc, idesc, (uint32_t)((g | k) != 0));
        }
    }
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_t64_kernel(const __grid_constant__ CUtensorMap mapA, const __half* __restrict__ wgt, const ConvTcParams p,
                const __grid_constant__ ConvBias cb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* shalo = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(shalo + T_STAGES * T_STAGE_STRIDE);
    uint64_t* empty = full + T_STAGES;
    uint64_t* tfull = empty + T_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < T_STAGES * 64) {  // the eight rows behind every box: zero, read by the tensor core (async proxy)
        const int s = threadIdx.x >> 6, i = threadIdx.x & 63;
        reinterpret_cast<uint4*>(shalo + (size_t)s * T_STAGE_STRIDE + T_STAGE_BYTES)[i] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < T_STAGES; i++) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 4);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&mapA);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int q = warp & 3, hi = lane >> 4, ch = 16 * q + (lane & 15);
    if (warp >= 2 && warp < 6) {
        // weights -> TMEM: this thread owns A row 32q + lane = (tap half hi, output channel ch)
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int g = 0; g < 6; g++) {
            const int tap = (g >> 1) * 3 + ((g & 1) ? 2 : hi);
            const bool zero = (g & 1) && hi;
            const uint4* src = reinterpret_cast<const uint4*>(wgt + ((size_t)tap * 64 + ch) * 64);
#pragma unroll
            for (int j = 0; j < 4; j++) {  // K-step j: 16 input channels = 8 columns
                uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
                if (!zero) {
                    a = src[2 * j];
                    b = src[2 * j + 1];
                }
                const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                ptx::tmem_st8(ta + (uint32_t)(g * 4 + j) * 8u, r);
            }
        }
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    const int per = p.tiles_x * p.tiles_y;
    auto decode = [&](int tile, int& n, int& y0, int& x0) {
        n = tile / per;
        const int r = tile - n * per, ty = r / p.tiles_x;
        y0 = ty * T_TH;
        x0 = (r - ty * p.tiles_x) * T_TW;
    };
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer: one halo box per tile =====================
        for (int it = 0; it < my_tiles; it++) {
            int n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, n, y0, x0);
            const uint32_t s = it & (T_STAGES - 1), ph = (it / T_STAGES) & 1;
            ptx::mbar_wait(&empty[s], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(&full[s], T_STAGE_BYTES);
                ptx::tma_load_4d(shalo + (size_t)s * T_STAGE_STRIDE, &mapA, &full[s], 0, x0 - 1, y0 - 1, n);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = ptx::make_idesc_f16(128, T_N, 0);
        const uint64_t b0 = ptx::make_sw128_desc(ptx::smem_u32(shalo));
        const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
        for (int it = 0; it < my_tiles; it++) {
            const uint32_t s = it & (T_STAGES - 1), acc = it & 1;
            ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            ptx::mbar_wait(&full[s], (it / T_STAGES) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                switch (s) {
                    case 0: issue_tile<0, 0>(tmem_base, b_lo, b_hi, idesc); break;
                    case 1: issue_tile<1, 1>(tmem_base, b_lo, b_hi, idesc); break;
                    case 2: issue_tile<2, 0>(tmem_base, b_lo, b_hi, idesc); break;
                    default: issue_tile<3, 1>(tmem_base, b_lo, b_hi, idesc); break;
                }
                ptx::umma_commit(&empty[s]);
                ptx::umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: group grp drains accumulator grp (every second tile) =====================
        const int grp = (warp - 2) >> 2;
        const float bias = cb.v[ch];
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + T_ACC0 + (uint32_t)grp * T_N;
        __half* const outp = reinterpret_cast<__half*>(p.out);
        for (int it = grp; it < my_tiles; it += 2) {
            int n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, n, y0, x0);
            ptx::mbar_wait(&tfull[grp], (it >> 1) & 1);
            ptx::tc_fence_after();
            if (p.mode == EPI_F16_POOL) {
                const int Ho = p.H >> 1, Wo = p.W >> 1;
#pragma unroll
                for (int rp = 0; rp < T_TH / 2; rp++) {
                    // two image rows = 80 consecutive accumulator columns
                    uint32_t ra[64], rb[16];
                    ptx::tmem_ld64(tq + rp * 2 * T_HW, ra);
                    ptx::tmem_ld16(tq + rp * 2 * T_HW + 64, rb);
                    ptx::tmem_ld_wait();
                    auto V = [&](int i) { return __uint_as_float(i < 64 ? ra[i] : rb[(i < 80 ? i : 79) - 64]); };
                    const int yo = (y0 >> 1) + rp;
                    __half* orow = outp + ((size_t)(n * Ho + yo) * Wo + (x0 >> 1)) * 64 + ch;
#pragma unroll
                    for (int i = 0; i < T_HW / 4; i++) {
                        // columns 4i .. 4i+4 of both rows: the low lane (tap half 0) keeps outputs 4i, 4i+1, the high
                        // lane (tap half 1, its columns shifted by one) outputs 4i+2, 4i+3; each sends the other its part
                        float o[2][2];
#pragma unroll
                        for (int rr = 0; rr < 2; rr++) {
                            const int b = rr * T_HW + 4 * i;
                            const float k0 = hi ? V(b + 3) : V(b), k1 = hi ? V(b + 4) : V(b + 1);
                            const float s0 = hi ? V(b + 1) : V(b + 2), s1 = hi ? V(b + 2) : V(b + 3);
                            o[rr][0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                            o[rr][1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                        }
                        float m = fmaxf(fmaxf(o[0][0], o[0][1]), fmaxf(o[1][0], o[1][1])) + bias;
                        if (p.relu) m = fmaxf(m, 0.f);
                        const int wdx = 2 * i + hi;  // pool window of this lane inside the tile row
                        if (wdx < T_TW / 2 && (x0 >> 1) + wdx < Wo && yo < Ho)
                            *reinterpret_cast<uint16_t*>(orow + (size_t)wdx * 64) = half_bits(m);
                    }
                }
            } else {  // EPI_F16
                const int odd = lane & 1;
#pragma unroll 1
                for (int rp = 0; rp < T_TH / 2; rp++) {
                    uint32_t ra[64], rb[16];
                    ptx::tmem_ld64(tq + rp * 2 * T_HW, ra);
                    ptx::tmem_ld16(tq + rp * 2 * T_HW + 64, rb);
                    ptx::tmem_ld_wait();
                    auto V = [&](int i) { return __uint_as_float(i < 64 ? ra[i] : rb[(i < 80 ? i : 79) - 64]); };
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        const int y = y0 + 2 * rp + rr;
                        // the lane pair (r, r ^ 1) holds channels (c, c + 1) of the same two pixels: one exchange lets the
                        // even lane store both channels of the first pixel and the odd lane both of the second (4-byte
                        // stores, 8 lanes = one 32-byte sector per pixel)
                        __half* orow = outp + ((size_t)(n * p.H + y) * p.W + x0) * 64 + (ch & ~1);
#pragma unroll
                        for (int i = 0; i < T_HW / 4; i++) {
                            const int b = rr * T_HW + 4 * i;
                            const float k0 = hi ? V(b + 3) : V(b), k1 = hi ? V(b + 4) : V(b + 1);
                            const float s0 = hi ? V(b + 1) : V(b + 2), s1 = hi ? V(b + 2) : V(b + 3);
                            float o0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16) + bias;
                            float o1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16) + bias;
                            if (p.relu) {
                                o0 = fmaxf(o0, 0.f);
                                o1 = fmaxf(o1, 0.f);
                            }
                            const float got = __shfl_xor_sync(0xffffffffu, odd ? o0 : o1, 1);
                            const __half2 h2 = odd ? __floats2half2_rn(got, o1) : __floats2half2_rn(o0, got);
                            const int ox = 4 * i + 2 * hi + odd;
                            if (ox < T_TW && x0 + ox < p.W && y < p.H)
                                *reinterpret_cast<__half2*>(orow + (size_t)ox * 64) = h2;
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[grp]);
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

bool conv_t64_applies(int cin, int cout_padded, int taps, int mode) {
    return taps == 9 && cin == 64 && cout_padded == 64 && (mode == EPI_F16 || mode == EPI_F16_POOL);
}

void conv_t64_plan(ConvLayer& L, int maxB, int H, int W) {
    ConvTcParams& p = L.p;
    L.v3 = 1;
    L.v2 = 0;
    L.flags = 0;
    L.halo_pitch = T_HW;
    L.box_w = T_HW;
    L.box_h = T_HH;
    p.tiles_x = (W + T_TW - 1) / T_TW;
    p.tiles_y = (H + T_TH - 1) / T_TH;
    p.total_tiles = maxB * p.tiles_x * p.tiles_y;
    p.stages = T_STAGES;
    L.smem_bytes = T_STAGES * T_STAGE_STRIDE + 1024 /*align*/ + (2 * T_STAGES + 4) * 8 + 64;
}

cudaError_t conv_t64_launch(const ConvLayer& L, const __half* wgt, int batch, int num_sms, cudaStream_t st) {
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        return cudaFuncSetAttribute(conv_t64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    ConvTcParams p = L.p;
    p.B = batch;
    p.total_tiles = batch * p.tiles_x * p.tiles_y;
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    conv_t64_kernel<<<grid, CONV_THREADS, L.smem_bytes, st>>>(L.mapA, wgt, p, L.hb);
    return cudaGetLastError();
}

}  // namespace ppg
