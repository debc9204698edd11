// One-time, per-device kernel attribute setup (cudaFuncSetAttribute is per device; contexts on several host threads
// and on several devices of one process may launch through the same function).
#pragma once
#include <cuda_runtime.h>

#include <mutex>

namespace ppg {

template <typename Fn>
inline cudaError_t once_per_device(bool (&done)[64], std::mutex& mu, Fn fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return cudaSuccess;
    e = fn();
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

}  // namespace ppg
