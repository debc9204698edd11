// Hardware probe for the transposed convolution kernel (weights as the A operand in TMEM, halo pixels as the B
// operand in shared memory).  Two parts, both printed as JSON lines:
//   check : D = A * B^T with A written to TMEM by tcgen05.st (32x32b), B a window of N rows starting at an arbitrary
//           128-byte row of a SWIZZLE_128B tile, compared with a host computation (confirms the A-in-TMEM layout,
//           the element order inside a 32-bit TMEM cell and the shifted-window B descriptor);
//   bench : cycles per tcgen05.mma for N in {64,128,160,240,256}, A from TMEM (TS) or shared memory (SS), alone and
//           with the other warps hammering the shared-memory pipe (SHFL / STS), on all SMs at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ppg_slam_b200/csrc
//        tools/ts_probe.cu -o tools/build/ts_probe
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "ptx.cuh"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

namespace {

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int B_ROWS = 352;  // rows of the B tile (128 B each)

// mode 0: TS (A in TMEM), 1: SS (A in shared memory)
__global__ void __launch_bounds__(160, 1)
check_kernel(const __half* __restrict__ A, const __half* __restrict__ Bm, float* __restrict__ D, int s, int N, int mode,
             int dcol) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sB = smem;
    uint8_t* sA = smem + B_ROWS * 128;  // 128 rows
    uint64_t* bar = reinterpret_cast<uint64_t*>(sA + 128 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 4) {
        ptx::tmem_alloc(slot, 512);
        ptx::tmem_relinquish();
    }
    for (int i = threadIdx.x; i < B_ROWS * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(Bm)[i];
    }
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(A)[i];
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tbase = *slot;
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)row * 64);
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = src[j];
        const uint32_t ta = tbase + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int g = 0; g < 4; g++) tmem_st8(ta + 8 * g, r + 8 * g);
        tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 4) {
        const uint32_t idesc = ptx::make_idesc_f16(128, N, 0);
        if (ptx::elect_one()) {
            const uint64_t bdesc = ptx::make_sw128_desc(ptx::smem_u32(sB) + (uint32_t)s * 128u);
            const uint64_t adesc = ptx::make_sw128_desc(ptx::smem_u32(sA));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (mode == 0)
                    umma_ts_f16(tbase + dcol, tbase + 8 * k, bdesc + 2 * k, idesc, (uint32_t)(k != 0));
                else
                    ptx::umma_f16(tbase + dcol, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)(k != 0));
            }
            ptx::umma_commit(bar);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    if (warp < 4) {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r[16];
            ptx::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + dcol + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) D[(size_t)row * 256 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tbase, 512);
    }
}

// noise bit 0: warps 0-3 run SHFL chains, bit 1: warps 0-3 write shared memory (STS.128), bit 2: warps 0-3 read TMEM
template <int N, int ilv, int mode>
__global__ void __launch_bounds__(192, 1)
bench_kernel(const __half* __restrict__ Bm, int iters, int noise, long long* __restrict__ cyc,
             long long* __restrict__ noise_ops) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sB = smem;                     // B_ROWS rows
    uint8_t* sA = smem + B_ROWS * 128;      // 6 x 128 rows (SS mode)
    uint8_t* sN = sA + 6 * 128 * 128;       // 16 KB scratch for the STS noise
    uint64_t* bar = reinterpret_cast<uint64_t*>(sN + 16384);  // [2]
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar[0], 1);
        ptx::mbar_init(&bar[1], 1);
        ptx::fence_barrier_init();
        *stop = 0;
    }
    if (warp == 4) {
        ptx::tmem_alloc(slot, 512);
        ptx::tmem_relinquish();
    }
    for (int i = threadIdx.x; i < (B_ROWS + 6 * 128) * 8; i += blockDim.x)
        reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bm)[i % (B_ROWS * 8)];
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tbase = *slot;
    if (warp < 4) {  // fill the A region of TMEM with the same data
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = reinterpret_cast<const uint32_t*>(sB)[(threadIdx.x * 8 + j) & 4095];
        for (int g = 0; g < 24; g++) tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 8 * g, r);
        tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 4) {
        const uint32_t idesc = ptx::make_idesc_f16(128, N, 0);
        // ilv > 1: consecutive instructions go to ilv different accumulators (column blocks of N), so that no
        // instruction depends on its predecessor; both tile buffers then share the same columns
        const uint32_t acc_col[2] = {192u, (ilv == 1 && N <= 160) ? 192u + (uint32_t)N : 192u};
        constexpr int srow[6] = {0, 2, 40, 42, 80, 82};
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const int acc = it & 1;
            if (it >= 2) ptx::mbar_wait(&bar[acc], ((it >> 1) - 1) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int g = 0; g < 6; g++) {
                    const uint64_t bdesc = ptx::make_sw128_desc(ptx::smem_u32(sB) + (uint32_t)srow[g] * 128u);
                    const uint64_t adesc = ptx::make_sw128_desc(ptx::smem_u32(sA) + (uint32_t)g * 16384u);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t dcol = acc_col[acc] + (uint32_t)(((g * 4 + k) % ilv) * N);
                        if (mode == 0)
                            umma_ts_f16(tbase + dcol, tbase + (uint32_t)(g * 4 + k) * 8u, bdesc + 2 * k, idesc,
                                        (uint32_t)((g * 4 + k) >= ilv));
                        else
                            ptx::umma_f16(tbase + dcol, adesc + 2 * k, bdesc + 2 * k, idesc,
                                          (uint32_t)((g * 4 + k) >= ilv));
                    }
                }
                ptx::umma_commit(&bar[acc]);
            }
            __syncwarp();
        }
        for (int a = 0; a < 2; a++) {
            const int n_a = (iters - a + 1) / 2;  // commits on bar[a]
            if (n_a > 0) ptx::mbar_wait(&bar[a], (n_a - 1) & 1);
        }
        const long long t1 = clock64();
        if (lane == 0) {
            cyc[blockIdx.x] = t1 - t0;
            *stop = 1;
        }
    } else if (warp < 4 && noise) {
        long long ops = 0;
        float v = (float)threadIdx.x;
        uint4 w = make_uint4(threadIdx.x, 1, 2, 3);
        uint32_t r[16];
        while (!*stop) {
            if (noise & 1) {
#pragma unroll
                for (int j = 0; j < 16; j++) v += __shfl_xor_sync(0xffffffffu, v, 16);
                ops += 16;
            }
            if (noise & 2) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    *reinterpret_cast<uint4*>(sN + ((threadIdx.x * 16 + j * 2048) & 16383)) = w;
                ops += 8;
            }
            if (noise & 4) {
                ptx::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + 192, r);
                ptx::tmem_ld_wait();
                v += __uint_as_float(r[3]);
                ops += 1;
            }
        }
        if (v == 1234.5f) sN[0] = 1;
        if (lane == 0) noise_ops[blockIdx.x * 4 + warp] = ops;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tbase, 512);
    }
}

}  // namespace

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": %d%d}\n", prop.name, prop.multiProcessorCount, prop.major,
           prop.minor);
    const int smem_check = B_ROWS * 128 + 128 * 128 + 2048;
    const int smem_bench = B_ROWS * 128 + 6 * 128 * 128 + 16384 + 2048;
    CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_check));

    srand(7);
    std::vector<__half> hA(128 * 64), hB((B_ROWS + 6 * 128) * 64);
    std::vector<float> fA(hA.size()), fB(hB.size());
    for (size_t i = 0; i < hA.size(); i++) {
        fA[i] = (float)((rand() % 2001) - 1000) / 1000.f;
        hA[i] = __float2half(fA[i]);
        fA[i] = __half2float(hA[i]);
    }
    for (size_t i = 0; i < hB.size(); i++) {
        fB[i] = (float)((rand() % 2001) - 1000) / 1000.f;
        hB[i] = __float2half(fB[i]);
        fB[i] = __half2float(hB[i]);
    }
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));

    struct Case {
        int s, N, mode, dcol;
    };
    const Case cases[] = {{0, 64, 1, 0},    {0, 64, 0, 64},    {3, 160, 0, 192}, {41, 160, 0, 352},
                          {82, 160, 0, 192}, {3, 240, 0, 192},  {5, 256, 1, 256}, {0, 16, 0, 32}};
    std::vector<float> hD(128 * 256);
    for (const Case& c : cases) {
        CK(cudaMemset(dD, 0xff, 128 * 256 * 4));
        check_kernel<<<1, 160, smem_check>>>(dA, dB, dD, c.s, c.N, c.mode, c.dcol);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("{\"check\": {\"s\": %d, \"N\": %d, \"mode\": %d, \"error\": \"%s\"}}\n", c.s, c.N, c.mode,
                   cudaGetErrorString(e));
            return 1;
        }
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        // hypotheses: 0 = as written; 1 = the two halves of every 32-bit A cell swapped; 2 = rows 0-63 / 64-127 swapped
        double err[3] = {0, 0, 0}, ref_max = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < c.N; n++) {
                double acc[3] = {0, 0, 0};
                for (int k = 0; k < 64; k++) {
                    const double b = fB[(size_t)(c.s + n) * 64 + k];
                    acc[0] += (double)fA[m * 64 + k] * b;
                    acc[1] += (double)fA[m * 64 + (k ^ 1)] * b;
                    acc[2] += (double)fA[(m ^ 64) * 64 + k] * b;
                }
                const double got = hD[(size_t)m * 256 + n];
                for (int h = 0; h < 3; h++) {
                    const double d = fabs(got - acc[h]);
                    if (!(d <= err[h])) err[h] = d;
                }
                ref_max = std::max(ref_max, fabs(acc[0]));
            }
        printf("{\"check\": {\"s\": %d, \"N\": %d, \"mode\": \"%s\", \"dcol\": %d, \"max_err\": %.3g, \"err_if_cell_halves_swapped\": %.3g, "
               "\"err_if_row_halves_swapped\": %.3g, \"ref_max\": %.3g}}\n",
               c.s, c.N, c.mode ? "SS" : "TS", c.dcol, err[0], err[1], err[2], ref_max);
    }

    // ---------------- throughput
    const int iters = argc > 1 ? atoi(argv[1]) : 400;
    const int nsm = prop.multiProcessorCount;
    long long *dcyc, *dops;
    CK(cudaMalloc(&dcyc, nsm * 8));
    CK(cudaMalloc(&dops, nsm * 4 * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](auto kern, int mode, int N, int ilv, int noise) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bench));
        CK(cudaMemset(dops, 0, nsm * 4 * 8));
        for (int rep = 0; rep < 2; rep++) {  // first run = warm-up
            CK(cudaEventRecord(e0));
            kern<<<nsm, 192, smem_bench>>>(dB, iters, noise, dcyc, dops);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> c(nsm), o(nsm * 4);
        CK(cudaMemcpy(c.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o.data(), dops, nsm * 4 * 8, cudaMemcpyDeviceToHost));
        std::sort(c.begin(), c.end());
        long long ops = 0;
        for (long long v : o) ops += v;
        const double per = (double)c[nsm / 2] / ((double)iters * 24);
        const double flops = 2.0 * 128 * N * 16 * 24.0 * iters * nsm;
        printf("{\"bench\": {\"mode\": \"%s\", \"N\": %d, \"ilv\": %d, \"noise\": %d, \"cycles_per_mma_median\": %.1f, "
               "\"min\": %.1f, \"max\": %.1f, \"floor\": %.1f, \"kernel_ms\": %.4f, \"tflops\": %.1f, "
               "\"noise_warp_ops_per_mma\": %.2f}}\n",
               mode ? "SS" : "TS", N, ilv, noise, per, (double)c[0] / (iters * 24.0), (double)c[nsm - 1] / (iters * 24.0),
               128.0 * N / 256.0, ms, flops / (ms * 1e-3) / 1e12, (double)ops / ((double)iters * 24 * nsm));
        fflush(stdout);
    };
#define RUN(N, ILV, MODE, NOISE) run(bench_kernel<N, ILV, MODE>, MODE, N, ILV, NOISE)
    RUN(160, 1, 0, 0); RUN(160, 2, 0, 0); RUN(160, 1, 0, 7); RUN(160, 2, 0, 7);
    RUN(80, 1, 0, 0);  RUN(80, 2, 0, 0);  RUN(80, 4, 0, 0);
    RUN(64, 1, 0, 0);  RUN(64, 2, 0, 0);  RUN(64, 4, 0, 0);
    RUN(128, 1, 0, 0); RUN(128, 2, 0, 0);
    RUN(240, 1, 0, 0); RUN(256, 1, 0, 0); RUN(32, 1, 0, 0); RUN(32, 4, 0, 0); RUN(16, 1, 0, 0); RUN(16, 4, 0, 0);
    RUN(160, 1, 1, 0); RUN(160, 2, 1, 0);
    RUN(64, 1, 1, 0);  RUN(64, 2, 1, 0);  RUN(64, 4, 1, 0);
    RUN(128, 1, 1, 0); RUN(128, 2, 1, 0); RUN(256, 1, 1, 0);
    return 0;
}
