// Transposed tcgen05 convolution for the 3x3 layers with Cin = 128 (conv3b, conv4a, conv4b, convPa, convDa) -- replaces
// the cuDNN convolutions LibTorch runs for net/Backbone.pt / PointHeatmap.pt / Descriptor.pt
// (feature/src/PPGExtractor.cpp:152-155).
//
// Why: the generic kernel (conv_tc.cu) fetches, for each of the 9 taps and 2 channel chunks, a shifted 16 KB pixel tile
// AND the 16-32 KB weight tile per 128 output pixels: 576 KB of L2 -> shared memory traffic per 72 tensor-core
// instructions.  Its ncu capture (profiles/r2_generic_conv_ncu_raw.csv) shows it bound by exactly that -- 12-14.6 TB/s
// arriving over the crossbar, tensor pipe 37-40 % active on conv3b / conv4a / conv4b, 61 % on the Cout = 256 layers.
// Here, as in conv_t64.cu,
//     D[cout, pixel] += W_tap[cout, cin] * X[pixel + shift(tap), cin]
//   * B = N consecutive pixels (up to 256) of a halo tile that is loaded ONCE per output tile (one TMA box per
//     64-channel chunk); the window of a tap is another start row of the same tile.  The chunks go through a ring of
//     nchunks + 1 buffers and the taps of a chunk are multiplied back to back, so a buffer is free for the next tile
//     while the other chunks of this tile are still in use -- a third less shared memory than two whole halo stages,
//     which the weight ring needs (below);
//   * A = the [128 cout x 64 cin] weight tile of (tap, chunk), streamed through a ring of 16 KB stages -- one tile per
//     four instructions of N columns, i.e. 16 KB per up to 240 output pixels instead of per 128.  As many stages as fit.
// Per 128 output pixels that is ~200 KB instead of 576 KB, and an N = 256 instruction runs at the full rate of the
// tensor pipe (128 cycles, tools/ts_probe.cu).  Cout = 256 layers run as two independent blocks of 128 output channels.
// The tile shape (TW x TH output pixels, N = (TH - 1) * (TW + 2) + TW columns rounded up to 16) is chosen per layer by
// conv_t128_plan so that the tiles cover the map with little waste and fill whole waves of SMs.
// Accumulator column n = (TW + 2) * oy + ox; TMEM lane = output channel, so the 2x2 max-pool happens inside a thread.
// Warps: 0 halo producer, 1 MMA issuer, 6 weight producer, 2-5 epilogue; for layers with fewer than 72 instructions per
// tile warps 7-10 are a second epilogue group (both drain every tile, group g takes the rows of its parity).
#include "conv_tc.cuh"
#include "once.cuh"
#include "ptx.cuh"

namespace ppg {

namespace {

constexpr int U_THREADS = 352;
constexpr int U_WBYTES = 16384;      // one weight tile: 128 output channels x 64 input channels
constexpr int U_MAX_WS = 10;
constexpr int U_MAX_HB = 5;       // halo chunk buffers: input chunks + 1
constexpr int U_SMEM_MAX = 227 * 1024;

struct T128 {
    int tw, th, hw;      // output tile, halo pitch (tw + 2 * pad)
    int ntaps, nchunks, pad;  // 9 taps (3x3, pad 1) or 1 (1x1, no halo); input channels / 64
    int hbox_bytes;      // bytes one halo box really holds: hw * (th + 2 * pad) * 128
    int hchunk_bytes;    // the same rounded up to 1024 (swizzle phase)
    int n;               // UMMA N
    int ws;              // weight ring stages
    int egroups;         // epilogue groups of four warps that share a tile (1 or 2)
    int nblk;            // blocks of 128 output channels
    int cout_total;      // rows per tap in the weight tensor
};

__device__ __forceinline__ uint16_t h_bits(float v) { return __half_as_ushort(__float2half_rn(v)); }

// MODE / EG (epilogue groups) are template parameters: with all four epilogues and a run-time row stride in one kernel
// ptxas settled on 74 registers and the plain fp16 epilogue lost its instruction-level parallelism (conv4a 0.051 ->
// 0.074 ms per 32 frames).
template <int MODE, int EG>
__global__ void __launch_bounds__(U_THREADS, 1)
conv_t128_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const ConvTcParams p,
                 const __grid_constant__ ConvBias cb, const T128 t) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* shalo = smem;                                   // [nchunks + 1][hchunk_bytes]: ring of halo chunks
    const int nbuf = t.nchunks + 1;
    uint8_t* swgt = shalo + (size_t)nbuf * t.hchunk_bytes;   // [ws][16 KB]
    uint64_t* hfull = reinterpret_cast<uint64_t*>(swgt + (size_t)t.ws * U_WBYTES);
    uint64_t* hempty = hfull + U_MAX_HB;
    uint64_t* wfull = hempty + U_MAX_HB;
    uint64_t* wempty = wfull + U_MAX_WS;
    uint64_t* tfull = wempty + U_MAX_WS;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            ptx::mbar_init(&tfull[i], 1);
            ptx::mbar_init(&tempty[i], 4 * EG);
        }
        for (int i = 0; i < U_MAX_HB; i++) {
            ptx::mbar_init(&hfull[i], 1);
            ptx::mbar_init(&hempty[i], 1);
        }
        for (int i = 0; i < U_MAX_WS; i++) {
            ptx::mbar_init(&wfull[i], 1);
            ptx::mbar_init(&wempty[i], 1);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&mapA);
        ptx::prefetch_tmap(&mapW);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int per = p.tiles_x * p.tiles_y;
    auto decode = [&](int tile, int& blk, int& n, int& y0, int& x0) {
        const int sp = tile / t.nblk;
        blk = tile - sp * t.nblk;
        n = sp / per;
        const int r = sp - n * per, ty = r / p.tiles_x;
        y0 = ty * t.th;
        x0 = (r - ty * p.tiles_x) * t.tw;
    };
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ===================== halo producer: one box per 64-channel chunk and tile =====================
        int hb = 0, hph = 0;
        for (int it = 0; it < my_tiles; it++) {
            int blk, n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, blk, n, y0, x0);
            for (int c = 0; c < t.nchunks; c++) {
                ptx::mbar_wait(&hempty[hb], hph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&hfull[hb], (uint32_t)t.hbox_bytes);
                    ptx::tma_load_4d(shalo + (size_t)hb * t.hchunk_bytes, &mapA, &hfull[hb], 64 * c, x0 - t.pad, y0 - t.pad, n);
                }
                __syncwarp();
                if (++hb == nbuf) {
                    hb = 0;
                    hph ^= 1;
                }
            }
        }
    } else if (warp == 6) {
        // ===================== weight producer: ntaps x nchunks tiles of [128 cout x 64 cin] per output tile ===========
        int wsi = 0, wph = 0;
        for (int it = 0; it < my_tiles; it++) {
            int blk, n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, blk, n, y0, x0);
            for (int c = 0; c < t.nchunks; c++)
            for (int tap = 0; tap < t.ntaps; tap++) {
                ptx::mbar_wait(&wempty[wsi], wph ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(&wfull[wsi], U_WBYTES);
                    ptx::tma_load_2d(swgt + (size_t)wsi * U_WBYTES, &mapW, &wfull[wsi], c * 64,
                                     tap * t.cout_total + blk * 128);
                }
                __syncwarp();
                if (++wsi == t.ws) {
                    wsi = 0;
                    wph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = ptx::make_idesc_f16(128, t.n, 0);
        const uint64_t h0 = ptx::make_sw128_desc(ptx::smem_u32(shalo));
        const uint64_t w0 = ptx::make_sw128_desc(ptx::smem_u32(swgt));
        const uint32_t h_lo = (uint32_t)h0, h_hi = (uint32_t)(h0 >> 32), w_lo = (uint32_t)w0, w_hi = (uint32_t)(w0 >> 32);
        int wsi = 0, wph = 0, hb = 0, hph = 0;
        for (int it = 0; it < my_tiles; it++) {
            const int acc = it & 1;
            ptx::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            const uint32_t d = tmem_base + (uint32_t)acc * 256u;
            for (int c = 0; c < t.nchunks; c++) {
            ptx::mbar_wait(&hfull[hb], hph);
            for (int tap = 0; tap < t.ntaps; tap++) {
                ptx::mbar_wait(&wfull[wsi], wph);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const int tc = tap + c;  // zero only for the first weight tile of the output tile
                    const uint32_t boff = (uint32_t)(hb * t.hchunk_bytes + ((tap / 3) * t.hw + tap % 3) * 128);
                    const uint32_t aoff = (uint32_t)(wsi * U_WBYTES);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint64_t adesc = ((uint64_t)w_hi << 32) | (uint64_t)(w_lo + ((aoff + k * 32) >> 4));
                        const uint64_t bdesc = ((uint64_t)h_hi << 32) | (uint64_t)(h_lo + ((boff + k * 32) >> 4));
                        ptx::umma_f16(d, adesc, bdesc, idesc, (uint32_t)((tc | k) != 0));
                    }
                    ptx::umma_commit(&wempty[wsi]);
                }
                __syncwarp();
                if (++wsi == t.ws) {
                    wsi = 0;
                    wph ^= 1;
                }
            }
            // the chunk's buffer is free as soon as its taps have been multiplied: the next tile's chunk lands there
            // while this tile's remaining chunks are still in use
            if (ptx::elect_one()) ptx::umma_commit(&hempty[hb]);
            __syncwarp();
            if (++hb == nbuf) {
                hb = 0;
                hph ^= 1;
            }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tfull[acc]);
            __syncwarp();
        }
    } else if (warp < 6 || EG == 2) {
        // ===================== epilogue: TMEM lane = output channel, columns = pixels =====================
        const int q = warp & 3, eg = warp > 6 ? 1 : 0;
        __half* const outp = reinterpret_cast<__half*>(p.out);
        for (int it = 0; it < my_tiles; it++) {
            int blk, n, y0, x0;
            decode(blockIdx.x + it * gridDim.x, blk, n, y0, x0);
            const int acc = it & 1;
            const int ch = blk * 128 + 32 * q + lane;
            const float bias = cb.v[ch];
            ptx::mbar_wait(&tfull[acc], (it >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            if (MODE == EPI_F16_POOL) {
                const int Ho = p.H >> 1, Wo = p.W >> 1;
                for (int oy = 2 * eg; oy < t.th; oy += 2 * EG) {
                    const int yo = (y0 + oy) >> 1;
                    for (int c0 = 0; c0 < t.tw; c0 += 16) {
                        uint32_t r0[16], r1[16];
                        ptx::tmem_ld16(tq + oy * t.hw + c0, r0);
                        ptx::tmem_ld16(tq + (oy + 1) * t.hw + c0, r1);
                        ptx::tmem_ld_wait();
                        __half* orow = outp + ((size_t)(n * Ho + yo) * Wo + ((x0 + c0) >> 1)) * p.out_ld + ch;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float m = fmaxf(fmaxf(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1])),
                                            fmaxf(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]))) + bias;
                            if (p.relu) m = fmaxf(m, 0.f);
                            if (c0 + 2 * j < t.tw && ((x0 + c0) >> 1) + j < Wo && yo < Ho)
                                *reinterpret_cast<uint16_t*>(orow + (size_t)j * p.out_ld) = h_bits(m);
                        }
                    }
                }
            } else if (MODE == EPI_F32) {  // lane = channel: a warp writes 128 contiguous bytes per pixel
                float* const outf = reinterpret_cast<float*>(p.out);
                for (int oy = eg; oy < t.th; oy += EG) {
                    const int y = y0 + oy;
                    for (int c0 = 0; c0 < t.tw; c0 += 16) {
                        uint32_t r0[16];
                        ptx::tmem_ld16(tq + oy * t.hw + c0, r0);
                        ptx::tmem_ld_wait();
                        float* orow = outf + ((size_t)(n * p.H + y) * p.W + x0 + c0) * p.out_ld + ch;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float m = __uint_as_float(r0[j]) + bias;
                            if (p.relu) m = fmaxf(m, 0.f);
                            if (c0 + j < t.tw && x0 + c0 + j < p.W && y < p.H) orow[(size_t)j * p.out_ld] = m;
                        }
                    }
                }
            } else if (MODE == EPI_F16_PS2) {  // pixel_shuffle(2): channel ch of (y, x) -> channel ch / 4 of (2y + i, 2x + j)
                const int oc = ch >> 2, pi = (ch >> 1) & 1, pj = ch & 1, Wo = 2 * p.W;
                for (int oy = eg; oy < t.th; oy += EG) {
                    const int y = y0 + oy;
                    for (int c0 = 0; c0 < t.tw; c0 += 16) {
                        uint32_t r0[16];
                        ptx::tmem_ld16(tq + oy * t.hw + c0, r0);
                        ptx::tmem_ld_wait();
                        __half* orow = outp + ((size_t)(n * 2 * p.H + 2 * y + pi) * Wo + 2 * (x0 + c0) + pj) * p.out_ld + oc;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float m = __uint_as_float(r0[j]) + bias;
                            if (p.relu) m = fmaxf(m, 0.f);
                            if (c0 + j < t.tw && x0 + c0 + j < p.W && y < p.H)
                                *reinterpret_cast<uint16_t*>(orow + (size_t)(2 * j) * p.out_ld) = h_bits(m);
                        }
                    }
                }
            } else {  // EPI_F16
                for (int oy = eg; oy < t.th; oy += EG) {
                    const int y = y0 + oy;
                    for (int c0 = 0; c0 < t.tw; c0 += 16) {
                        uint32_t r0[16];
                        ptx::tmem_ld16(tq + oy * t.hw + c0, r0);
                        ptx::tmem_ld_wait();
                        __half* orow = outp + ((size_t)(n * p.H + y) * p.W + x0 + c0) * p.out_ld + ch;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float m = __uint_as_float(r0[j]) + bias;
                            if (p.relu) m = fmaxf(m, 0.f);
                            if (c0 + j < t.tw && x0 + c0 + j < p.W && y < p.H)
                                *reinterpret_cast<uint16_t*>(orow + (size_t)j * p.out_ld) = h_bits(m);
                        }
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

int round_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

bool conv_t128_applies(int cin, int cout_padded, int taps, int mode) {
    return (taps == 9 || taps == 1) && (cin == 64 || cin == 128 || cin == 256) && cout_padded % 128 == 0 &&
           cout_padded <= 256 && (mode == EPI_F16 || mode == EPI_F16_POOL || mode == EPI_F32 || mode == EPI_F16_PS2);
}

// Tile shape: the cheapest cover of the map in whole waves of SMs.  Cost of a tile = max(its tensor-core instructions of N
// columns at the measured rate, its L2 -> shared memory bytes at the measured ~45 bytes per cycle and SM).
void conv_t128_plan(ConvLayer& L, int maxB, int H, int W, int num_sms) {
    ConvTcParams& p = L.p;
    const bool pool = p.mode == EPI_F16_POOL;
    const int nblk = L.cout / 128, nchunks = L.cin / 64, ntaps = p.taps, pad = ntaps == 9 ? 1 : 0;
    double best = 1e30;
    int btw = 0, bth = 0, bws = 0;
    for (int th = pool ? 2 : 1; th <= 8; th += pool ? 2 : 1) {
        for (int tw = 8; tw <= 254 && tw <= round_up(W, 2); tw += pool ? 2 : 1) {
            const int hw = tw + 2 * pad, hh = th + 2 * pad;
            const int n = round_up((th - 1) * hw + tw, 16);
            if (n > 256) break;
            if ((th - 1) * hw + round_up(tw, 16) > 256) continue;  // the epilogue reads 16 columns at a time
            const int hchunk = round_up(hw * hh * 128, 1024);
            int ws = (U_SMEM_MAX - 1024 - (nchunks + 1) * hchunk - 512) / U_WBYTES;
            if (ws < 3) continue;
            if (ws > U_MAX_WS) ws = U_MAX_WS;
            const long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * maxB * nblk;
            const long waves = (tiles + num_sms - 1) / num_sms;
            // 57 / 70.6 / 83 / 128 cycles per instruction at N = 64 / 128 / 160 / 256 (tools/ts_probe.cu)
            const double mma = 4.0 * ntaps * nchunks * (n > 128 ? n * 0.5 : 0.26 * n + 37.0) + 500.0;
            const double l2 = ((double)ntaps * nchunks * U_WBYTES + (double)nchunks * hw * hh * 128) / 45.0;
            const double cost = (double)waves * (mma > l2 ? mma : l2);
            if (cost < best) {
                best = cost;
                btw = tw;
                bth = th;
                bws = ws;
            }
        }
    }
    L.v3 = 3;
    L.v2 = 0;
    L.flags = 0;
    L.t_tw = btw;
    L.t_th = bth;
    L.t_ws = bws;
    L.halo_pitch = btw + 2 * pad;
    L.box_w = btw + 2 * pad;
    L.box_h = bth + 2 * pad;
    p.tiles_x = (W + btw - 1) / btw;
    p.tiles_y = (H + bth - 1) / bth;
    p.total_tiles = maxB * p.tiles_x * p.tiles_y * nblk;
    p.stages = bws;
    const int hchunk = round_up(L.box_w * L.box_h * 128, 1024);
    L.smem_bytes = 1024 + (nchunks + 1) * hchunk + bws * U_WBYTES + 512;
}

template <int MODE, int EG>
static cudaError_t launch_t128(const ConvLayer& L, const ConvTcParams& p, const T128& t, int grid, cudaStream_t st) {
    static bool attr_done[64];
    static std::mutex attr_mu;
    const cudaError_t attr_err = once_per_device(attr_done, attr_mu, [] {
        return cudaFuncSetAttribute(conv_t128_kernel<MODE, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, U_SMEM_MAX);
    });
    if (attr_err != cudaSuccess) return attr_err;
    conv_t128_kernel<MODE, EG><<<grid, EG == 2 ? U_THREADS : 224, L.smem_bytes, st>>>(L.mapA, L.mapB, p, L.hb, t);
    return cudaGetLastError();
}

cudaError_t conv_t128_launch(const ConvLayer& L, int batch, int num_sms, cudaStream_t st) {
    ConvTcParams p = L.p;
    T128 t;
    t.ntaps = p.taps;
    t.nchunks = L.cin / 64;
    t.pad = p.taps == 9 ? 1 : 0;
    t.tw = L.t_tw;
    t.th = L.t_th;
    t.hw = L.t_tw + 2 * t.pad;
    t.hbox_bytes = t.hw * (t.th + 2 * t.pad) * 128;
    t.hchunk_bytes = round_up(t.hbox_bytes, 1024);
    t.n = round_up((t.th - 1) * t.hw + t.tw, 16);
    t.ws = L.t_ws;
    t.nblk = L.cout / 128;
    t.cout_total = L.cout;
    // 72 instructions per tile (3x3, Cin = 128) hide one group's epilogue; the shorter tiles need two (measured: convDb
    // 0.119 -> 0.086 ms with two, conv4a 0.051 -> 0.056 ms)
    t.egroups = t.ntaps * t.nchunks >= 18 ? 1 : 2;
    p.B = batch;
    p.total_tiles = batch * p.tiles_x * p.tiles_y * t.nblk;
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (grid <= 0) return cudaSuccess;
    const bool two = t.egroups == 2;
    switch (p.mode) {
        case EPI_F16_POOL: return two ? launch_t128<EPI_F16_POOL, 2>(L, p, t, grid, st) : launch_t128<EPI_F16_POOL, 1>(L, p, t, grid, st);
        case EPI_F32: return two ? launch_t128<EPI_F32, 2>(L, p, t, grid, st) : launch_t128<EPI_F32, 1>(L, p, t, grid, st);
        case EPI_F16_PS2: return two ? launch_t128<EPI_F16_PS2, 2>(L, p, t, grid, st) : launch_t128<EPI_F16_PS2, 1>(L, p, t, grid, st);
        default: return two ? launch_t128<EPI_F16, 2>(L, p, t, grid, st) : launch_t128<EPI_F16, 1>(L, p, t, grid, st);
    }
}

}  // namespace ppg
