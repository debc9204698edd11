// Thin network layers on tcgen05 (conv1a_tc.cu) and the validation convolution (net_direct.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppg {
// conv1a_tc.cu: conv1a (u8 -> x/255 -> conv3x3 1->64 + ReLU) and the edge-decoder tail on tcgen05
cudaError_t conv1a_tc_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                             cudaStream_t st);
cudaError_t edge_tail_tc_launch(const __half* in, const float* w3, const float* b3, const float* w1, const float* b1,
                                float* heat, int B, int Hh, int Wh, cudaStream_t st);  // conv1a_tc.cu
cudaError_t conv_ref_launch(const __half* in, const __half* w, const float* bias, float* out, int B, int H, int W,
                            int Cin, int N, int taps, int relu, cudaStream_t st);
}  // namespace ppg
