// Implicit-GEMM 3x3 / 1x1 convolution on tcgen05 tensor cores (sm_100a), NHWC fp16 activations.
//
// Replaces the cuDNN convolutions LibTorch runs for net/Backbone.pt, PointHeatmap.pt, EdgeHeatmap.pt
// and Descriptor.pt (reference call sites feature/src/PPGExtractor.cpp:152-155).
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} A[pixel shifted by tap, cin] * Wt[tap][cout][cin]
//   M tile  = 128 output pixels = an 8-row x 16-column image patch (TMEM lane = 16*row + col)
//   N       = Cout (padded to a multiple of 16, <= 256), one UMMA N
//   K step  = 64 input channels of one tap = one 128-byte swizzled smem row per pixel
// A tiles come straight from the NHWC activation tensor with one 4-D TMA box per (tap, 64-channel
// chunk); out-of-image coordinates are zero-filled by TMA, which is exactly the conv's zero padding.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-5 / 6-9 = epilogue of the
// even / odd tiles.
// Two TMEM accumulators (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppg {

enum ConvEpilogue {
    EPI_F16 = 0,       // bias (+ReLU) -> NHWC fp16
    EPI_F16_POOL = 1,  // bias + ReLU + 2x2 max-pool -> NHWC fp16 at half resolution
    EPI_F16_PS2 = 2,   // bias + ReLU + pixel_shuffle(2) -> NHWC fp16, Cout/4 channels at double resolution
    EPI_F32 = 3,       // bias (+ReLU) -> NHWC fp32 (all N padded channels stored)
    EPI_SOFTMAX_D2S = 4,  // junction head tail (N = 80, 65 valid): bias -> softmax over 65 channels -> drop the dustbin
                          // -> depth-to-space(8): P[8y+i][8x+j] = softmax[8i+j]  (PPGExtractor.cpp:161-162), fp32
};

struct ConvTcParams {
    int B, H, W;       // batch actually processed, conv (input == output) resolution
    int cin_chunks;    // Cin / 64
    int taps;          // 9 (3x3, pad 1) or 1 (1x1)
    int N;             // UMMA N (padded Cout)
    int mode, relu;
    int stages;
    int tiles_x, tiles_y, total_tiles;
    int tile_w_log2;   // generic kernel: output tile = (1 << tile_w_log2) x (128 >> tile_w_log2) pixels (16 x 8 or 32 x 4)
    const float* bias; // [N]
    void* out;
    int out_ld;        // channels per output pixel in the destination tensor
    // EPI_SOFTMAX_D2S only: the threshold scan of detectKeyPoint (PPGExtractor.cpp:168-176) fused into the epilogue --
    // every thread has the 64 probabilities of its coarse cell in registers, so the candidate list and the NMS state
    // map are written here and the H x W map is not read again for it (scan_cand == nullptr: not fused)
    uint32_t* scan_cand;    // [B][H*W*64] pixel indices of the in-border candidates (unordered)
    int* scan_counters;     // [B][8]: 0 in-border candidates, 1 pixels >= threshold
    uint8_t* scan_state2;   // [B][H*W*16] 2 bits per pixel (shared-memory NMS), or
    uint8_t* scan_state;    // [B][H*W*64] one byte per pixel (global-memory NMS)
    float scan_thresh;
    int scan_radius;
};

// Bias of one layer, passed BY VALUE as a __grid_constant__ kernel parameter: the epilogue then reads it from the
// constant bank (LDC / uniform operands) instead of shared memory.  The first ncu captures showed the shared-memory
// data pipe saturated (tensor-core operand reads 55 % + LDS/SHFL of the epilogue 31 %), so the epilogue must stay
// off that pipe.
struct ConvBias {
    float v[256];
};

struct ConvLayer {
    CUtensorMap mapA, mapB;
    ConvTcParams p;
    ConvBias hb;        // host copy of the (padded) bias, see above
    int smem_bytes;
    int cin, cout;
    // v2 ("halo") kernel: 3x3, Cin = 64, weights resident in shared memory, one TMA halo tile per output tile
    int v2;             // 0 = generic kernel, 1 = halo kernel
    int box_w, box_h;   // TMA box of the activation map (pixels): generic 16 x 8, halo (8+2 | 16) x 18
    int halo_pitch;     // pixels per halo row in shared memory (= box_w)
    int flags;          // conv_tc2_kernel flags: bit 0 = interleaved tile pairs
    // v3 ("transposed") kernel, conv_t64.cu: 3x3, Cin = Cout = 64, weights as the A operand in tensor memory
    int v3;             // 1 = transposed kernel, 2 = transposed kernel with conv1a computed by its producer warps,
                        // 3 = transposed kernel for Cin = 128 (conv_t128.cu)
    int t_tw, t_th, t_ws;  // v3 == 3: output tile and weight ring stages chosen by conv_t128_plan
    const __half* wgt;  // [tap][Cout][Cin] fp16 (the transposed kernel loads its A operand from here)
    const uint8_t* gray;            // v3 == 2: the u8 frames and conv1a's fp32 weights [64][9] / bias [64]
    const float *w1a, *b1a;
};

constexpr int CONV_TILE_W = 16, CONV_TILE_H = 8, CONV_A_BYTES = 16384, CONV_THREADS = 320;
constexpr int CONV2_TILE_W = 8, CONV2_TILE_H = 16;
constexpr int CONVT_TILE_W = 38, CONVT_TILE_H = 4;  // transposed kernel: output tile, halo 40 x 6

// Fills stages / tiles / smem size for the given shape. Tensor maps are encoded by the caller (api).
void conv_tc_plan(ConvLayer& L, int maxB, int H, int W, int cin, int cout_padded, int taps, int mode, int relu,
                  const float* bias, void* out, int out_ld);
cudaError_t conv_tc_launch(const ConvLayer& L, int batch, int num_sms, cudaStream_t st);

// Transposed kernel (conv_t64.cu).
bool conv_t64_applies(int cin, int cout_padded, int taps, int mode);
void conv_t64_plan(ConvLayer& L, int maxB, int H, int W);
cudaError_t conv_t64_launch(const ConvLayer& L, const __half* wgt, int batch, int num_sms, cudaStream_t st);
// Transposed kernel for Cin = 128 (conv_t128.cu); its weight tensor map has boxes of 128 rows.
bool conv_t128_applies(int cin, int cout_padded, int taps, int mode);
void conv_t128_plan(ConvLayer& L, int maxB, int H, int W, int num_sms);
cudaError_t conv_t128_launch(const ConvLayer& L, int batch, int num_sms, cudaStream_t st);
// conv1b only (the layer whose input is conv1a's output): compute conv1a inside the kernel, see conv_t64.cu
void conv_t64_fuse_conv1a(ConvLayer& L, const uint8_t* gray, const float* w1a, const float* b1a);

}  // namespace ppg
