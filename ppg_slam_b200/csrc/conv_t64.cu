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
            ptx::mbar_init(&tempty[a], p.mode == EPI_F16_POOL ? 4 : 8);  // EPI_F16: both groups drain every tile
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
        // ===================== epilogue =====================
        // pooled layers: group grp drains accumulator grp (every second tile);
        // EPI_F16 (four times the values to add, convert and store): BOTH groups drain every tile, group grp the image rows
        // 2 grp, 2 grp + 1 -- with one group per tile the ~1150 instructions of a warp per tile took longer than the two
        // tile periods it has (ncu: tensor pipe 56 % active, 0.277 ms for conv2a)
        const int grp = (warp - 2) >> 2;
        const float bias = cb.v[ch];
        const uint32_t tq0 = tmem_base + ((uint32_t)(q * 32) << 16) + T_ACC0;
        __half* const outp = reinterpret_cast<__half*>(p.out);
        if (p.mode == EPI_F16_POOL) {
            const uint32_t tq = tq0 + (uint32_t)grp * T_N;
            const int Ho = p.H >> 1, Wo = p.W >> 1;
            for (int it = grp; it < my_tiles; it += 2) {
                int n, y0, x0;
                decode(blockIdx.x + it * gridDim.x, n, y0, x0);
                ptx::mbar_wait(&tfull[grp], (it >> 1) & 1);
                ptx::tc_fence_after();
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
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty[grp]);
            }
        } else {  // EPI_F16
            const int odd = lane & 1;
            for (int it = 0; it < my_tiles; it++) {
                int n, y0, x0;
                decode(blockIdx.x + it * gridDim.x, n, y0, x0);
                const int acc = it & 1;
                ptx::mbar_wait(&tfull[acc], (it >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t tq = tq0 + (uint32_t)acc * T_N + (uint32_t)(grp * 2 * T_HW);
                uint32_t ra[64], rb[16];
                ptx::tmem_ld64(tq, ra);
                ptx::tmem_ld16(tq + 64, rb);
                ptx::tmem_ld_wait();
                auto V = [&](int i) { return __uint_as_float(i < 64 ? ra[i] : rb[(i < 80 ? i : 79) - 64]); };
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int y = y0 + 2 * grp + rr;
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
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            }
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
// conv1a fused into conv1b: the 64-channel full-resolution map conv1a produces (46 MB per EuRoC frame, written and
// read back once = 2 x 1.48 GB per 32 frames, and a 0.32 ms kernel bound by exactly that write) never exists.
// Four producer warps replace the TMA producer: for the 240 halo pixels of a tile they run conv1a as conv1a_tc_kernel
// does -- A = the 9 taps of a pixel as exact u8 integers in fp16 plus a 1.0 for the bias, two accumulating M128 N64
// K16 instructions against the hi / lo halves of the weights, fp32 accumulator in TMEM -- and write ReLU(acc) as the
// pixel's 128-byte swizzled row of the halo stage, zeros for pixels outside the image (conv1b's padding).  Same
// instruction on the same operands: the stage holds bit for bit what TMA would have loaded from conv1a's output.
//
// Tensor memory: the conv1a accumulator needs 64 columns, and 192 (weights) + 320 (two accumulators) is everything.
// The three unpaired taps therefore take their weights from SHARED memory here (M128 with a zero half as before, 16 KB
// per tap, 128-byte swizzled): measured (tools/ts_probe.cu) an N = 160 instruction takes 83 cycles with A in either
// place.  TMEM: [0,96) paired weights, [96,160) conv1a accumulator, [192,512) the two conv1b accumulators.
// Warps: 0 MMA issuer of conv1b, 1-4 / 5-8 epilogue of even / odd tiles, 9-12 / 13-16 producers of halo pixels 0-127 /
// 128-239 (thread <-> TMEM lane; the first warp of a group issues its conv1a instructions behind a named barrier).
// A single group of four warps needed 4300 cycles per tile for its chain of dependent steps (fetch, build, barrier,
// MMA, tcgen05.ld, convert, store) against the 2600 cycles conv1b spends on a tile: 1.21 ms per 32 frames.
constexpr int F_THREADS = 544, F_PROD0 = 9;
constexpr int F_WU_BYTES = 3 * 16384;
constexpr int F_C1A_BYTES = 4096, F_C1B_BYTES = 2048;
constexpr uint32_t F_C1ACC = 96;
constexpr int F_HALO = T_HW * T_HH;  // 240 pixels
constexpr int F_SMEM = T_STAGES * T_STAGE_STRIDE + F_WU_BYTES + 2 * F_C1A_BYTES + 2 * F_C1B_BYTES + 256 + 1024;

__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// fp16 pair (lo = relu(a), hi = relu(b)) in one instruction
__device__ __forceinline__ uint32_t relu_pack(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    return d;
}

template <int STAGE, int ACC>
__device__ __forceinline__ void issue_tile_f(uint32_t tmem_base, uint32_t b_lo, uint32_t b_hi, uint32_t w_lo,
                                             uint32_t w_hi, uint32_t idesc) {
    const uint32_t d = tmem_base + T_ACC0 + ACC * T_N;
#pragma unroll
    for (int g = 0; g < 6; g++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t off = (uint32_t)(STAGE * T_STAGE_STRIDE + ((g >> 1) * T_HW + ((g & 1) ? 2 : 0)) * 128 + k * 32);
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + (off >> 4));
            if (g & 1) {
                const uint64_t adesc = ((uint64_t)w_hi << 32) | (uint64_t)(w_lo + (uint32_t)(((g >> 1) * 16384 + k * 32) >> 4));
                ptx::umma_f16(d, adesc, bdesc, idesc, 1u);
            } else {
                ptx::umma_ts_f16(d, tmem_base + (uint32_t)((g >> 1) * 4 + k) * 8u, bdesc, idesc, (uint32_t)((g | k) != 0));
            }
        }
    }
}

__global__ void __launch_bounds__(F_THREADS, 1)
conv_t64_fused_kernel(const uint8_t* __restrict__ gray, const float* __restrict__ w1a, const float* __restrict__ b1a,
                      const __half* __restrict__ wgt, const ConvTcParams p, const __grid_constant__ ConvBias cb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* shalo = smem;
    uint8_t* swu = shalo + T_STAGES * T_STAGE_STRIDE;  // 1024-aligned: the stride is a multiple of 1024
    uint8_t* sA1 = swu + F_WU_BYTES;                   // conv1a A operand, two buffers
    uint8_t* sBhi = sA1 + 2 * F_C1A_BYTES;
    uint8_t* sBlo = sBhi + F_C1B_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sBlo + F_C1B_BYTES);
    uint64_t* empty = full + T_STAGES;
    uint64_t* tfull = empty + T_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* c1full = tempty + 2;
    uint64_t* c1empty = c1full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c1empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < T_STAGES * 64) {  // the eight zero rows behind every stage (see T_STAGE_STRIDE)
        const int s = threadIdx.x >> 6, i = threadIdx.x & 63;
        reinterpret_cast<uint4*>(shalo + (size_t)s * T_STAGE_STRIDE + T_STAGE_BYTES)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // unpaired taps (dy,+1): A row m = (tap half (m >> 4) & 1, channel 16 * (m >> 5) + (m & 15)), half 1 = zeros
    for (int i = threadIdx.x; i < 3 * 128 * 8; i += F_THREADS) {
        const int gi = i >> 10, m = (i >> 3) & 127, c = i & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (!((m >> 4) & 1)) {
            const int chn = 16 * (m >> 5) + (m & 15);
            v = reinterpret_cast<const uint4*>(wgt + ((size_t)(gi * 3 + 2) * 64 + chn) * 64)[c];
        }
        *reinterpret_cast<uint4*>(swu + gi * 16384 + m * 128 + ((c ^ (m & 7)) << 4)) = v;
    }
    if (threadIdx.x < 64) {  // conv1a weights of channel t as in conv1a_tc_kernel: taps / 255, bias in column 9, hi + lo
        const int t = threadIdx.x;
        float v[16];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = w1a[t * 9 + k] * (1.0f / 255.0f);
        v[9] = b1a[t];
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
    }
    ptx::fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int i = 0; i < T_STAGES; i++) {
            ptx::mbar_init(&full[i], 8);
            ptx::mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; a++) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 4);
        }
        ptx::mbar_init(&c1full[0], 1);
        ptx::mbar_init(&c1full[1], 1);
        ptx::mbar_init(c1empty, 128);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int q = warp & 3, hi = lane >> 4, ch = 16 * q + (lane & 15);
    if (warp >= 1 && warp < 5) {
        // paired taps -> TMEM: this thread owns A row 32q + lane = (tap half hi, output channel ch)
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int gp = 0; gp < 3; gp++) {
            const uint4* src = reinterpret_cast<const uint4*>(wgt + ((size_t)(gp * 3 + hi) * 64 + ch) * 64);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 a = src[2 * j], b = src[2 * j + 1];
                const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                ptx::tmem_st8(ta + (uint32_t)(gp * 4 + j) * 8u, r);
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
        // ===================== MMA issuer of conv1b =====================
        const uint32_t idesc = ptx::make_idesc_f16(128, T_N, 0);
        const uint64_t b0 = ptx::make_sw128_desc(ptx::smem_u32(shalo));
        const uint64_t w0 = ptx::make_sw128_desc(ptx::smem_u32(swu));
        const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32), w_lo = (uint32_t)w0, w_hi = (uint32_t)(w0 >> 32);
        for (int it = 0; it < my_tiles; it++) {
            const uint32_t s = it & (T_STAGES - 1), acc = it & 1;
            ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            ptx::mbar_wait(&full[s], (it / T_STAGES) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                switch (s) {
                    case 0: issue_tile_f<0, 0>(tmem_base, b_lo, b_hi, w_lo, w_hi, idesc); break;
                    case 1: issue_tile_f<1, 1>(tmem_base, b_lo, b_hi, w_lo, w_hi, idesc); break;
                    case 2: issue_tile_f<2, 0>(tmem_base, b_lo, b_hi, w_lo, w_hi, idesc); break;
                    default: issue_tile_f<3, 1>(tmem_base, b_lo, b_hi, w_lo, w_hi, idesc); break;
                }
                ptx::umma_commit(&empty[s]);
                ptx::umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else if (warp >= F_PROD0) {
        // ===================== conv1a producers: group pg computes halo pixels [128 pg, 128 pg + 128) of every tile ====
        // One accumulator serves both groups: a group holds it from its two MMAs to the end of its tcgen05.ld (c1empty),
        // everything else -- fetching the taps, building A, ReLU / convert / store of the 128-byte rows -- overlaps with
        // the other group's turn.
        const int pg = (warp - F_PROD0) >> 2;
        const int m = 32 * q + lane;   // TMEM lane = A row
        const int row = pg * 128 + m;  // halo pixel
        const int hy = row / T_HW, hx = row - hy * T_HW;
        const bool live = row < F_HALO;
        const int H = p.H, W = p.W;
        const uint32_t idesc1 = ptx::make_idesc_f16(128, 64, 0);
        const uint64_t bhi = make_nosw_desc(ptx::smem_u32(sBhi), 1024u, 128u);
        const uint64_t blo = make_nosw_desc(ptx::smem_u32(sBlo), 1024u, 128u);
        uint8_t* const a = sA1 + pg * F_C1A_BYTES;
        const uint64_t adesc = make_nosw_desc(ptx::smem_u32(a), 2048u, 128u);
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + F_C1ACC;
        // tile coordinates advance by gridDim.x tiles per iteration: no divisions inside the loop
        const int step_ty = (int)gridDim.x / p.tiles_x, step_tx = (int)gridDim.x - step_ty * p.tiles_x;
        int fn, fty, ftx;  // tile of the NEXT fetch
        {
            int y0, x0;
            decode(blockIdx.x, fn, y0, x0);
            fty = y0 / T_TH;
            ftx = x0 / T_TW;
        }
        // the 9 taps of this thread's halo pixel of the next tile: raw bytes, consumed one iteration later so that the
        // loads stay in flight across the MMA round trip (the first version packed them at once: 7 % of all stall samples)
        auto fetch = [&](int it, uint32_t (&t)[9], bool& inside) {
#pragma unroll
            for (int k = 0; k < 9; k++) t[k] = 0u;
            inside = false;
            const int y = fty * T_TH - 1 + hy, x = ftx * T_TW - 1 + hx, n = fn;
            ftx += step_tx;
            fty += step_ty;
            if (ftx >= p.tiles_x) {
                ftx -= p.tiles_x;
                fty++;
            }
            while (fty >= p.tiles_y) {
                fty -= p.tiles_y;
                fn++;
            }
            if (it >= my_tiles || !live) return;
            if (y < 0 || y >= H || x < 0 || x >= W) return;
            inside = true;
            const uint8_t* g = gray + ((size_t)n * H + y) * W + x;
            if (y >= 1 && y < H - 1 && x >= 1 && x < W - 1) {
#pragma unroll
                for (int k = 0; k < 9; k++) t[k] = __ldg(g + (k / 3 - 1) * W + (k % 3 - 1));
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W) t[k] = __ldg(g + (k / 3 - 1) * W + (k % 3 - 1));
                }
            }
        };
        // fp16 pair (u0, u1) from two bytes: 0x6400 | u is the fp16 number 1024 + u, and (1024 + u) - 1024 is exact
        auto h2 = [](uint32_t u0, uint32_t u1) {
            const uint32_t v = 0x64006400u | u0 | (u1 << 16);
            const __half2 c = __halves2half2(__ushort_as_half((unsigned short)0x6400), __ushort_as_half((unsigned short)0x6400));
            const __half2 r2 = __hsub2(*reinterpret_cast<const __half2*>(&v), c);
            return *reinterpret_cast<const uint32_t*>(&r2);
        };
        uint32_t u[9];
        bool in_next;
        fetch(0, u, in_next);
        for (int it = 0; it < my_tiles; it++) {
            const int j = 2 * it + pg;  // turn on the shared accumulator
            const uint32_t s = it & (T_STAGES - 1);
            *reinterpret_cast<uint4*>(a + m * 16) = make_uint4(h2(u[0], u[1]), h2(u[2], u[3]), h2(u[4], u[5]), h2(u[6], u[7]));
            *reinterpret_cast<uint4*>(a + 2048 + m * 16) = make_uint4(h2(u[8], 0u) | 0x3C000000u, 0u, 0u, 0u);  // tap 8, 1.0
            ptx::fence_proxy_async();
            if (pg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (((warp - F_PROD0) & 3) == 0) {
                ptx::mbar_wait(c1empty, (j & 1) ^ 1);  // the other group has read the accumulator of turn j - 1
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    ptx::umma_f16(tmem_base + F_C1ACC, adesc, bhi, idesc1, 0u);
                    ptx::umma_f16(tmem_base + F_C1ACC, adesc, blo, idesc1, 1u);
                    ptx::umma_commit(&c1full[pg]);
                }
                __syncwarp();
            }
            const bool inside = in_next;
            fetch(it + 1, u, in_next);
            ptx::mbar_wait(&empty[s], ((it / T_STAGES) & 1) ^ 1);  // conv1b is done with the stage
            ptx::mbar_wait(&c1full[pg], it & 1);
            ptx::tc_fence_after();
            uint32_t r[64];
            ptx::tmem_ld64(tacc, r);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(c1empty);
            if (live) {
                uint8_t* dst = shalo + (size_t)s * T_STAGE_STRIDE + row * 128;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (inside) {
                        v.x = relu_pack(r[8 * c + 0], r[8 * c + 1]);
                        v.y = relu_pack(r[8 * c + 2], r[8 * c + 3]);
                        v.z = relu_pack(r[8 * c + 4], r[8 * c + 5]);
                        v.w = relu_pack(r[8 * c + 6], r[8 * c + 7]);
                    }
                    *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = v;
                }
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&full[s]);
        }
    } else {
        // ===================== epilogue: group grp drains accumulator grp (every second tile) =====================
        // one image row at a time (40 + 10 live values instead of 80: this kernel has 416 threads)
        const int grp = (warp - 1) >> 2;
        const float bias = cb.v[ch];
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + T_ACC0 + (uint32_t)grp * T_N;
        __half* const outp = reinterpret_cast<__half*>(p.out);
        const int Ho = p.H >> 1, Wo = p.W >> 1;
        for (int it = grp; it < my_tiles; it += 2) {
            int n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, n, y0, x0);
            ptx::mbar_wait(&tfull[grp], (it >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int rp = 0; rp < T_TH / 2; rp++) {
                float hm[T_HW / 4];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    uint32_t ra[32], rb[8];
                    ptx::tmem_ld32(tq + (2 * rp + rr) * T_HW, ra);
                    ptx::tmem_ld8(tq + (2 * rp + rr) * T_HW + 32, rb);
                    ptx::tmem_ld_wait();
                    auto V = [&](int i) { return __uint_as_float(i < 32 ? ra[i] : rb[(i < 40 ? i : 39) - 32]); };
#pragma unroll
                    for (int i = 0; i < T_HW / 4; i++) {
                        const int b = 4 * i;
                        const float k0 = hi ? V(b + 3) : V(b), k1 = hi ? V(b + 4) : V(b + 1);
                        const float s0 = hi ? V(b + 1) : V(b + 2), s1 = hi ? V(b + 2) : V(b + 3);
                        const float o0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                        const float o1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                        const float mx = fmaxf(o0, o1);
                        hm[i] = rr ? fmaxf(hm[i], mx) : mx;
                    }
                }
                const int yo = (y0 >> 1) + rp;
                __half* orow = outp + ((size_t)(n * Ho + yo) * Wo + (x0 >> 1)) * 64 + ch;
#pragma unroll
                for (int i = 0; i < T_HW / 4; i++) {
                    float mv = hm[i] + bias;
                    if (p.relu) mv = fmaxf(mv, 0.f);
                    const int wdx = 2 * i + hi;
                    if (wdx < T_TW / 2 && (x0 >> 1) + wdx < Wo && yo < Ho)
                        *reinterpret_cast<uint16_t*>(orow + (size_t)wdx * 64) = half_bits(mv);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[grp]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
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

void conv_t64_fuse_conv1a(ConvLayer& L, const uint8_t* gray, const float* w1a, const float* b1a) {
    L.v3 = 2;
    L.gray = gray;
    L.w1a = w1a;
    L.b1a = b1a;
    L.smem_bytes = F_SMEM;
}

static cudaError_t conv_t64_fused_launch(const ConvLayer& L, const __half* wgt, int batch, int num_sms, cudaStream_t st) {
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        return cudaFuncSetAttribute(conv_t64_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess) return attr_err;
    ConvTcParams p = L.p;
    p.B = batch;
    p.total_tiles = batch * p.tiles_x * p.tiles_y;
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    conv_t64_fused_kernel<<<grid, F_THREADS, L.smem_bytes, st>>>(L.gray, L.w1a, L.b1a, wgt, p, L.hb);
    return cudaGetLastError();
}

cudaError_t conv_t64_launch(const ConvLayer& L, const __half* wgt, int batch, int num_sms, cudaStream_t st) {
    if (L.v3 == 2) return conv_t64_fused_launch(L, wgt, batch, num_sms, st);
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
