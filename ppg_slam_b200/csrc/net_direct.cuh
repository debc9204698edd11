// CUDA-core network pieces (see net_direct.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppg {
cudaError_t conv1a_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                          cudaStream_t st);
bool conv1a_tc_supported(int H, int W);  // conv1a_tc.cu: the tcgen05 version of conv1a
cudaError_t conv1a_tc_launch(const uint8_t* gray, const float* w, const float* bias, __half* out, int B, int H, int W,
                             cudaStream_t st);
cudaError_t edge_tail_tc_launch(const __half* in, const float* w3, const float* b3, const float* w1, const float* b1,
                                float* heat, int B, int Hh, int Wh, cudaStream_t st);  // conv1a_tc.cu
cudaError_t edge_tail_launch(const __half* in, const float* w3, const float* b3, const float* w1, const float* b1,
                             float* heat, int B, int Hh, int Wh, cudaStream_t st);
cudaError_t junction_d2s_launch(const float* logits, float* prob, int B, int Hc, int Wc, int ld, cudaStream_t st);
cudaError_t conv_ref_launch(const __half* in, const __half* w, const float* bias, float* out, int B, int H, int W,
                            int Cin, int N, int taps, int relu, cudaStream_t st);
}  // namespace ppg
