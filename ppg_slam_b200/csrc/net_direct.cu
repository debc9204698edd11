// Validation-only CUDA-core convolution (ppg_selftest_conv): plain fp32 direct convolution over the same NHWC fp16
// operands the tensor-core kernels consume.  The network itself runs on tcgen05 only (conv_tc.cu, conv_t64.cu,
// conv1a_tc.cu); the CUDA-core / mma.sync versions of conv1a, the edge-decoder tail and the junction softmax that round 1
// superseded are gone.
#include "net_direct.cuh"

namespace ppg {

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
