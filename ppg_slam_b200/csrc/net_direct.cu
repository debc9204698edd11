// CUDA-core pieces of the networks: the layers that are not GEMM-shaped (SURVEY.md s.7 K1/K4c/K3a tail).
//   conv1a_kernel      u8 -> x/255 -> conv3x3 1->64 + ReLU (fp32; |w| reaches 197) -> NHWC fp16
//   edge_tail_kernel   conv3x3 16->16 (BN folded) + ReLU + pixel_shuffle(2) + conv1x1 4->2 + softmax[:,1]
//   junction_d2s_kernel softmax over 65 channels, drop dustbin, depth-to-space(8) -> H x W prob map
//   conv_ref_kernel    plain fp32 direct convolution over NHWC fp16 (debug/validation of conv_tc only)
// Reference call sites: feature/src/PPGExtractor.cpp:151-154 (inference), :161-162, :242.
#include "net_direct.cuh"

namespace ppg {

// ------------------------------------------------------------------------------------------------
// conv1a: 8 threads per pixel, each produces 8 output channels (one 16-byte store); a warp writes 512
// contiguous bytes.  Input tile (+1 halo) staged in shared memory as fp32 already divided by 255.
__global__ void __launch_bounds__(256) conv1a_kernel(const uint8_t* __restrict__ gray, const float* __restrict__ w,
                                                     const float* __restrict__ bias, __half* __restrict__ out, int H,
                                                     int W) {
    constexpr int TX = 32, TY = 8;  // 256 pixels per block pass, 8 channel-groups -> loop
    __shared__ float tile[TY + 2][TX + 2];
    __shared__ float sw[64 * 9];
    __shared__ float sb[64];
    const int n = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const uint8_t* g = gray + (size_t)n * H * W;
    for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
    for (int i = threadIdx.x; i < (TY + 2) * (TX + 2); i += blockDim.x) {
        int ty = i / (TX + 2), tx = i - ty * (TX + 2);
        int y = y0 + ty - 1, x = x0 + tx - 1;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = __fdiv_rn((float)g[(size_t)y * W + x], 255.0f);  // :151
        tile[ty][tx] = v;
    }
    __syncthreads();
    // thread -> (pixel = t/8 + 32*pass, group = t%8)
    const int grp = threadIdx.x & 7;
    for (int pass = 0; pass < 8; pass++) {
        int pix = (threadIdx.x >> 3) + 32 * pass;
        int py = pix / TX, px = pix - py * TX;
        int y = y0 + py, x = x0 + px;
        float in[9];
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) in[ky * 3 + kx] = tile[py + ky][px + kx];
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const float* wc = &sw[(grp * 8 + c) * 9];
            float a = sb[grp * 8 + c];
#pragma unroll
            for (int k = 0; k < 9; k++) a = fmaf(in[k], wc[k], a);
            acc[c] = fmaxf(a, 0.f);
        }
        if (y < H && x < W) {
            __half2 h0 = __floats2half2_rn(acc[0], acc[1]), h1 = __floats2half2_rn(acc[2], acc[3]);
            __half2 h2 = __floats2half2_rn(acc[4], acc[5]), h3 = __floats2half2_rn(acc[6], acc[7]);
            uint4 u = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                 *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            *reinterpret_cast<uint4*>(out + (((size_t)n * H + y) * W + x) * 64 + grp * 8) = u;
        }
    }
}

cudaError_t conv1a_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                          cudaStream_t st) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    conv1a_kernel<<<grid, 256, 0, st>>>(gray, w, bias, out, H, W);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Edge decoder tail.  in: NHWC fp16, 16 channels at (Hh x Wh) = (H/2 x W/2).  One thread per low-res
// pixel: 16 conv outputs = 4 channels x 2x2 sub-pixels, then the 4->2 1x1 conv and the 2-way softmax
// for each of the 4 full-resolution pixels.  w3: [16 out][3][3][16 in] fp32, BN folded.
__global__ void __launch_bounds__(128) edge_tail_kernel(const __half* __restrict__ in, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, float* __restrict__ heat,
                                                        int Hh, int Wh) {
    constexpr int TX = 16, TY = 8;
    __shared__ float tile[TY + 2][TX + 2][17];  // +1 pad against bank conflicts
    __shared__ float sw[16 * 9 * 16];
    __shared__ float sb[16], s1[8], sb1[2];
    const int n = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    for (int i = threadIdx.x; i < 16 * 9 * 16; i += blockDim.x) sw[i] = w3[i];
    if (threadIdx.x < 16) sb[threadIdx.x] = b3[threadIdx.x];
    if (threadIdx.x < 8) s1[threadIdx.x] = w1[threadIdx.x];
    if (threadIdx.x < 2) sb1[threadIdx.x] = b1[threadIdx.x];
    const __half* src = in + (size_t)n * Hh * Wh * 16;
    for (int i = threadIdx.x; i < (TY + 2) * (TX + 2) * 2; i += blockDim.x) {
        int half8 = i & 1, pi = i >> 1;
        int ty = pi / (TX + 2), tx = pi - ty * (TX + 2);
        int y = y0 + ty - 1, x = x0 + tx - 1;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < Hh && x >= 0 && x < Wh)
            u = *reinterpret_cast<const uint4*>(src + ((size_t)y * Wh + x) * 16 + half8 * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float2 f = __half22float2(h2[k]);
            tile[ty][tx][half8 * 8 + 2 * k] = f.x;
            tile[ty][tx][half8 * 8 + 2 * k + 1] = f.y;
        }
    }
    __syncthreads();
    const int py = threadIdx.x / TX, px = threadIdx.x - py * TX;
    const int y = y0 + py, x = x0 + px;
    float o[16];
#pragma unroll
    for (int co = 0; co < 16; co++) o[co] = sb[co];
    for (int ky = 0; ky < 3; ky++)
        for (int kx = 0; kx < 3; kx++) {
            float iv[16];
#pragma unroll
            for (int ci = 0; ci < 16; ci++) iv[ci] = tile[py + ky][px + kx][ci];
#pragma unroll
            for (int co = 0; co < 16; co++) {
                const float* wc = &sw[((co * 3 + ky) * 3 + kx) * 16];
                float a = o[co];
#pragma unroll
                for (int ci = 0; ci < 16; ci++) a = fmaf(iv[ci], wc[ci], a);
                o[co] = a;
            }
        }
    if (y >= Hh || x >= Wh) return;
    const int H = Hh * 2, W = Wh * 2;
    float* dst = heat + (size_t)n * H * W;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float hv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            // pixel_shuffle(2): full-res channel c at (2y+i, 2x+j) = low-res channel 4c + 2i + j
            float l0 = sb1[0], l1 = sb1[1];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float v = fmaxf(o[4 * c + 2 * i + j], 0.f);
                l0 = fmaf(s1[c], v, l0);
                l1 = fmaf(s1[4 + c], v, l1);
            }
            float m = fmaxf(l0, l1);
            float e0 = expf(l0 - m), e1 = expf(l1 - m);
            hv[j] = e1 / (e0 + e1);  // softmax(dim=1)[:,1], PPGExtractor.cpp:242
        }
        *reinterpret_cast<float2*>(dst + (size_t)(2 * y + i) * W + 2 * x) = make_float2(hv[0], hv[1]);
    }
}

cudaError_t edge_tail_launch(const __half* in, const float* w3, const float* b3, const float* w1, const float* b1,
                             float* heat, int B, int Hh, int Wh, cudaStream_t st) {
    dim3 grid((Wh + 15) / 16, (Hh + 7) / 8, B);
    edge_tail_kernel<<<grid, 128, 0, st>>>(in, w3, b3, w1, b1, heat, Hh, Wh);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Junction head tail: logits NHWC fp32 [B][Hc][Wc][ld] (65 valid) -> prob [B][8Hc][8Wc].
// One warp per coarse cell.  P[8h+i][8w+j] = softmax(logits)[8i+j]  (PPGExtractor.cpp:161-162).
__global__ void __launch_bounds__(256) junction_d2s_kernel(const float* __restrict__ logits, float* __restrict__ prob,
                                                           int cells_total, int Hc, int Wc, int ld) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= cells_total) return;
    const int n = warp / (Hc * Wc), r = warp - n * Hc * Wc, hc = r / Wc, wc = r - hc * Wc;
    const float* l = logits + (size_t)warp * ld;
    float a = l[lane], b = l[lane + 32], c = (lane == 0) ? l[64] : -INFINITY;
    float m = fmaxf(fmaxf(a, b), c);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    float ea = expf(a - m), eb = expf(b - m), ec = (lane == 0) ? expf(c - m) : 0.f;
    float sum = ea + eb + ec;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const int W = Wc * 8, H = Hc * 8;
    float* dst = prob + (size_t)n * H * W;
    // channel ch -> (i = ch/8, j = ch%8)
    dst[(size_t)(8 * hc + (lane >> 3)) * W + 8 * wc + (lane & 7)] = ea / sum;
    dst[(size_t)(8 * hc + 4 + (lane >> 3)) * W + 8 * wc + (lane & 7)] = eb / sum;
}

cudaError_t junction_d2s_launch(const float* logits, float* prob, int B, int Hc, int Wc, int ld, cudaStream_t st) {
    int cells = B * Hc * Wc;
    int blocks = (cells * 32 + 255) / 256;
    junction_d2s_kernel<<<blocks, 256, 0, st>>>(logits, prob, cells, Hc, Wc, ld);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Validation-only direct convolution (fp32 accumulate over the same fp16 operands conv_tc consumes).
// w: [taps][N][Cin] fp16 (same tensor conv_tc reads).  out: fp32 NHWC [B][H][W][N], bias added, ReLU opt.
__global__ void conv_ref_kernel(const __half* __restrict__ in, const __half* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ out, int B, int H, int W, int Cin,
                                int N, int taps, int relu) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * H * W * N;
    if (idx >= total) return;
    int co = idx % N;
    size_t pix = idx / N;
    int x = pix % W;
    int y = (pix / W) % H;
    int n = pix / ((size_t)W * H);
    float acc = 0.f;
    for (int tap = 0; tap < taps; tap++) {
        int dy = taps == 9 ? tap / 3 - 1 : 0, dx = taps == 9 ? tap % 3 - 1 : 0;
        int yy = y + dy, xx = x + dx;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const __half* ip = in + (((size_t)n * H + yy) * W + xx) * Cin;
        const __half* wp = w + ((size_t)tap * N + co) * Cin;
        for (int ci = 0; ci < Cin; ci++) acc = fmaf(__half2float(ip[ci]), __half2float(wp[ci]), acc);
    }
    acc += bias[co];
    if (relu) acc = fmaxf(acc, 0.f);
    out[idx] = acc;
}

cudaError_t conv_ref_launch(const __half* in, const __half* w, const float* bias, float* out, int B, int H, int W,
                            int Cin, int N, int taps, int relu, cudaStream_t st) {
    size_t total = (size_t)B * H * W * N;
    conv_ref_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, w, bias, out, B, H, W, Cin, N, taps, relu);
    return cudaGetLastError();
}

}  // namespace ppg
