// Post-processing kernels: the CPU stages of PPGExtractor (feature/src/PPGExtractor.cpp:158-589) on the GPU.
//
// Compiled with -fmad=false: the reference is built for baseline x86-64 (no FMA contraction,
// CMakeLists.txt:8-9), and keypoints / graph must come out bit-identical given the same dense maps.
// Everything order- or rounding-dependent in the reference is kept:
//   detectKeyPoint :158-234  threshold scan, (score desc, raster asc) order, greedy 9x9 NMS, top-500, undistort
//   refineHeatMap  :540-578  per 16x16 tile top-30% mean rescale
//   remap          :259-263  cv::remap INTER_LINEAR in 1/32 fixed point (pinhole only)
//   detectLines    :265-441  3-point heat test, greedy (i,j)-ordered overlap filter, line scoring, colinearity
//   genPointDescriptor :515-538  bilinear sampling + L2 normalisation
#include "post.cuh"

#include <math.h>

namespace ppg {

namespace {

#define PPG_PI 3.1415926535897932384626433832795
#define PPG_2PI 6.283185307179586476925286766559

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int* hdr_of(const PostParams& p, int b) {
    return reinterpret_cast<int*>(p.out + (size_t)b * p.lay.total + p.lay.hdr);
}
template <typename T>
__device__ __forceinline__ T* out_of(const PostParams& p, int b, size_t off) {
    return reinterpret_cast<T*>(p.out + (size_t)b * p.lay.total + off);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        int t = __shfl_up_sync(FULL, v, s);
        if (lane >= s) v += t;
    }
    return v;
}

// Exclusive prefix of one value per thread over the whole block; *total = block sum.  `ws` = 33 ints.
__device__ int block_excl_scan(int v, int* ws, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = warp_incl_scan(v, lane);
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? ws[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        ws[lane] = wi - w;
        if (lane == 31) ws[32] = wi;
    }
    __syncthreads();
    int r = ws[warp] + inc - v;
    *total = ws[32];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------------
// K5: threshold scan.  state[pix] = 1 for in-border pixels with score >= thresh (:173 skips "<", so a
// NaN would pass); their indices are appended (unordered) to cand[].  Pixels closer than R to the border
// can never be accepted and never suppress (:190-193), so they are dropped here.
// SCAN_Q quads of four pixels per thread, a block covers SCAN_Q * 256 consecutive quads: every thread has SCAN_Q
// independent 16-byte loads in flight (the first version, one quad per thread, ran at 0.16 of the HBM copy bandwidth:
// latency-bound), the row of a quad comes from a multiply-high with the precomputed reciprocal of W instead of a
// division, and a warp appends its candidates with one atomic for all its quads.
constexpr int SCAN_Q = 4;
__global__ void __launch_bounds__(256) scan_kernel(const PostParams p, const uint32_t w_magic) {
    const int b = blockIdx.y;
    const int HW = p.H * p.W, nq = HW >> 2;
    const int lane = threadIdx.x & 31;
    const int R = p.nms_radius;
    const int q0 = blockIdx.x * (256 * SCAN_Q) + threadIdx.x;
    const float* prob = p.prob + (size_t)b * HW;
    float4 v[SCAN_Q];
#pragma unroll
    for (int u = 0; u < SCAN_Q; u++) {
        const int q = q0 + u * 256;
        v[u] = q < nq ? __ldcs(reinterpret_cast<const float4*>(prob) + q) : make_float4(-1.f, -1.f, -1.f, -1.f);
    }
    int c = 0, call = 0;
    uint32_t stq[SCAN_Q];  // one byte per pixel of the quad: 1 = in-border candidate
#pragma unroll
    for (int u = 0; u < SCAN_Q; u++) {
        const int q = q0 + u * 256;
        stq[u] = 0;
        if (q >= nq) continue;
        const float s[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const int pix = q * 4;
        const int y = (int)__umulhi((uint32_t)pix, w_magic);  // pix / W (exact: pix < 2^32 / W, checked on the host)
        const int x0 = pix - y * p.W;
        const bool yin = (y >= R) && (y <= p.H - R - 1);
        uint32_t st = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool pass = !(s[k] < p.junction_thresh);
            const int x = x0 + k;
            const bool inb = yin && (x >= R) && (x <= p.W - R - 1);
            call += pass;
            if (pass && inb) st |= 1u << (8 * k);
        }
        stq[u] = st;
        c += __popc(st);
        if (p.nms_smem) {
            // 2 bits per pixel, 4 pixels per byte (pixel k of the quad in bits 2k..2k+1)
            const uint32_t packed = (st & 1u) | ((st >> 6) & 4u) | ((st >> 12) & 16u) | ((st >> 18) & 64u);
            p.state2[(size_t)b * nq + q] = (uint8_t)packed;
        } else {
            *reinterpret_cast<uint32_t*>(p.state + (size_t)b * HW + pix) = st;
        }
    }
    // block-aggregated append: ONE pair of atomics per block (per-frame counters: with one pair per warp the 1 400
    // same-address atomics of a frame serialised in L2 and set the kernel time)
    __shared__ int s_wtot[8], s_wcall[8], s_base;
    const int warp = threadIdx.x >> 5;
    const int inc = warp_incl_scan(c, lane);
    const int tot = __shfl_sync(FULL, inc, 31);
    const int call_w = __reduce_add_sync(FULL, call);
    if (lane == 0) {
        s_wtot[warp] = tot;
        s_wcall[warp] = call_w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0, cw = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int x = s_wtot[w];
            s_wtot[w] = t;  // exclusive prefix over the warps
            t += x;
            cw += s_wcall[w];
        }
        s_base = t ? atomicAdd(&p.counters[b * 8 + 0], t) : 0;
        if (cw) atomicAdd(&p.counters[b * 8 + 1], cw);
    }
    __syncthreads();
    const int base = s_base + s_wtot[warp];
    if (c) {
        uint32_t* cl = p.cand + (size_t)b * HW + base + inc - c;
#pragma unroll
        for (int u = 0; u < SCAN_Q; u++) {
            uint32_t st = stq[u];
            const uint32_t pix = (uint32_t)(q0 + u * 256) * 4u;
            while (st) {
                const int k = (__ffs(st) - 1) >> 3;
                *cl++ = pix + k;
                st &= st - 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K6+K7: exact greedy NMS without the sequential walk, then rank + top-k + undistort.
//
// The reference walks candidates in (score desc) order and accepts one iff no previously ACCEPTED
// candidate lies within Chebyshev distance R (:185-206).  Equivalent fixed point: a candidate is
//   suppressed  if some higher-priority candidate within R is accepted,
//   accepted    if every higher-priority candidate within R is suppressed.
// Rounds of that rule converge to the unique greedy result in any evaluation order (decisions are
// final and only read other final decisions), so stale reads are benign.  Priority = (score desc,
// raster index asc) -- the stable order the oracle uses for std::sort ties.  The 500 cap (:196-197) only
// truncates the walk, so it is applied afterwards on the ranked survivors.
// Rank the NMS survivors (bitonic sort of (~score, index) keys = score desc, raster asc), keep the first max_kp
// (:196-197), undistort them through the LUT and write the keypoint SoA + header of the frame record.
// If more survivors than key slots exist (degenerate maps: a uniform map leaves 14 400 survivors at 752x480), the
// max_kp best are selected exactly first: bitwise search for the max_kp-th smallest key over the accepted
// candidates (`accepted(idx)` reads the NMS state), then only the keys up to it are gathered -- keys are unique, so
// exactly max_kp of them.  The record is then NOT flagged: the result is the reference's.
template <typename AcceptedFn>
__device__ void nms_finish(const PostParams& p, int b, unsigned long long* keys, int nacc_all, int ovf, int rounds,
                           int ncand, AcceptedFn accepted) {
    const int tid = threadIdx.x, W = p.W;
    int nacc = nacc_all;
    int* hdr = hdr_of(p, b);
    if (nacc_all > p.acc_cap && p.max_kp <= p.acc_cap) {
        __shared__ int s_cnt;
        const float* prob = p.prob + (size_t)b * p.H * p.W;
        const uint32_t* cl = p.cand + (size_t)b * p.H * p.W;
        auto key_of = [&](int idx) {
            return ((unsigned long long)(~__float_as_uint(__ldg(prob + idx))) << 32) | (unsigned)idx;
        };
        unsigned long long kth = 0ull;  // smallest K with #(key <= K) >= max_kp, built from the top bit down
        for (int bit = 63; bit >= 0; bit--) {
            const unsigned long long trial = kth | ((1ull << bit) - 1ull);  // all keys with this bit clear (and prefix)
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            int c = 0;
            for (int t = tid; t < ncand; t += blockDim.x) {
                const int idx = cl[t];
                if (accepted(idx) && key_of(idx) <= trial) c++;
            }
            if (c) atomicAdd(&s_cnt, c);
            __syncthreads();
            if (s_cnt < p.max_kp) kth |= 1ull << bit;
            __syncthreads();
        }
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int t = tid; t < ncand; t += blockDim.x) {
            const int idx = cl[t];
            if (!accepted(idx)) continue;
            const unsigned long long k = key_of(idx);
            if (k <= kth) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < p.acc_cap) keys[slot] = k;
            }
        }
        __syncthreads();
        nacc = s_cnt < p.acc_cap ? s_cnt : p.acc_cap;
        ovf = 0;
        __syncthreads();
    }
    if (nacc > p.acc_cap) nacc = p.acc_cap;
    // bitonic sort of the survivors' keys (ascending = score desc, index asc)
    int P = 1;
    while (P < nacc) P <<= 1;
    for (int t = nacc + tid; t < P; t += blockDim.x) keys[t] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P; t += blockDim.x) {
                const int u = t ^ j;
                if (u > t) {
                    const bool asc = (t & k) == 0;
                    const unsigned long long a = keys[t], c = keys[u];
                    if ((a > c) == asc) {
                        keys[t] = c;
                        keys[u] = a;
                    }
                }
            }
            __syncthreads();
        }
    const int nkp = nacc < p.max_kp ? nacc : p.max_kp;
    float* kp_x = out_of<float>(p, b, p.lay.kp_x);
    float* kp_y = out_of<float>(p, b, p.lay.kp_y);
    int* opx = out_of<int>(p, b, p.lay.px);
    int* opy = out_of<int>(p, b, p.lay.py);
    float* osc = out_of<float>(p, b, p.lay.score);
    float* oxu = out_of<float>(p, b, p.lay.xun);
    float* oyu = out_of<float>(p, b, p.lay.yun);
    uint8_t* oout = out_of<uint8_t>(p, b, p.lay.kout);
    for (int t = tid; t < nkp; t += blockDim.x) {
        const unsigned long long key = keys[t];
        const int idx = (int)(unsigned)(key & 0xffffffffull);
        const float s = __uint_as_float(~(unsigned)(key >> 32));
        const int y = idx / W, x = idx - y * W;
        const float2 un = p.undist_lut[idx];  // cv::[fisheye::]undistortPoints of the integer pixel, :220-223
        uint8_t o = 1;                        // KeyPointEx ctor: mbOut(true)
        if (un.x >= 1.f && un.x < (float)(p.W - 1) && un.y >= 1.f && un.y < (float)(p.H - 1)) o = 0;  // :230
        opx[t] = x;
        opy[t] = y;
        osc[t] = s;
        oxu[t] = un.x;
        oyu[t] = un.y;
        oout[t] = o;
        if (p.fisheye) {  // run(): pinhole only mPos <- mPosUn (:141-145)
            kp_x[t] = (float)x;
            kp_y[t] = (float)y;
        } else {
            kp_x[t] = un.x;
            kp_y[t] = un.y;
        }
    }
    if (tid == 0) {
        hdr[HDR_NKP] = nkp;
        hdr[HDR_NEDGES] = 0;
        hdr[HDR_NCOL] = 0;
        hdr[HDR_STATUS] = ovf ? ST_OVF_ACCEPT : 0;
        hdr[HDR_NCAND] = p.counters[b * 8 + 1];
        hdr[HDR_NACC] = nacc_all;
        hdr[HDR_NPASS] = 0;
        hdr[HDR_NLINES] = 0;
        hdr[HDR_NMS_ROUNDS] = rounds;
    }
}


__global__ void __launch_bounds__(1024) nms_global_kernel(const PostParams p) {
    extern __shared__ unsigned long long keys[];  // [acc_cap]
    __shared__ int s_remaining, s_nacc, s_ovf;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = p.H * p.W, W = p.W, R = p.nms_radius;
    const float* prob = p.prob + (size_t)b * HW;
    volatile uint8_t* state = p.state + (size_t)b * HW;
    const uint32_t* cl = p.cand + (size_t)b * HW;
    int n = p.counters[b * 8 + 0];
    if (n > HW) n = HW;
    if (tid == 0) {
        s_nacc = 0;
        s_ovf = 0;
    }
    int rounds = 0;
    for (;;) {
        if (tid == 0) s_remaining = 0;
        __syncthreads();
        for (int c = tid; c < n; c += blockDim.x) {
            const int idx = cl[c];
            if (state[idx] != 1) continue;
            const float s = prob[idx];
            const int y = idx / W, x = idx - y * W;
            bool sup = false, wait = false;
            for (int dy = -R; dy <= R && !sup; dy++) {
                const int rowi = (y + dy) * W + x;
                for (int dx = -R; dx <= R; dx++) {
                    const int ni = rowi + dx;
                    const uint8_t st = state[ni];
                    if (st == 0 || st == 3 || ni == idx) continue;
                    const float sn = prob[ni];
                    const bool higher = (sn > s) || (sn == s && ni < idx);
                    if (!higher) continue;
                    if (st == 2) {
                        sup = true;
                        break;
                    }
                    wait = true;
                }
            }
            if (sup) {
                state[idx] = 3;
            } else if (!wait) {
                state[idx] = 2;
                const int k = atomicAdd(&s_nacc, 1);
                if (k < p.acc_cap)
                    keys[k] = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned)idx;
                else
                    s_ovf = 1;
            } else {
                atomicAdd(&s_remaining, 1);
            }
        }
        __syncthreads();
        rounds++;
        const int rem = s_remaining;
        __syncthreads();
        if (rem == 0 || rounds > n + 1) break;
    }
    nms_finish(p, b, keys, s_nacc, s_ovf, rounds, n, [&](int idx) { return state[idx] == 2; });
}

// Same fixed point with the per-pixel state held in SHARED memory as 2 bits per pixel (90 KB at 752x480): the
// global-memory version above spends its time in 81 dependent L2 reads per candidate and round (0.51 ms for 32
// frames in the first ncu capture).  A 9-pixel row of the window is one 18-bit field of a 64-bit load.
//   * an ACCEPTED pixel inside the window always has higher priority than an undecided one (it could not have been
//     accepted while this candidate was neither suppressed nor of lower priority), so it suppresses without a score
//     comparison;
//   * only UNDECIDED neighbours need their score (read-only global loads) to know whether to wait for them.
// Used whenever the bitmap and the ranking keys fit (all shipped shapes except 1024x1024, which takes the kernel above).
__global__ void __launch_bounds__(1024) nms_smem_kernel(const PostParams p) {
    extern __shared__ unsigned long long nms_dyn[];
    __shared__ int s_remaining, s_nacc, s_ovf;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int HW = p.H * p.W, W = p.W, R = p.nms_radius;
    unsigned long long* keys = nms_dyn;                                   // [acc_cap]
    uint32_t* st2 = reinterpret_cast<uint32_t*>(nms_dyn + p.acc_cap);     // [HW/16 + 2]
    const int nwords = HW / 16;
    const float* prob = p.prob + (size_t)b * HW;
    const uint32_t* cl = p.cand + (size_t)b * HW;
    const uint32_t* g2 = reinterpret_cast<const uint32_t*>(p.state2 + (size_t)b * (HW / 4));
    for (int t = tid; t < nwords + 2; t += blockDim.x) st2[t] = t < nwords ? g2[t] : 0u;
    int n = p.counters[b * 8 + 0];
    if (n > HW) n = HW;
    if (tid == 0) {
        s_nacc = 0;
        s_ovf = 0;
    }
    const int span = 2 * R + 1;
    const uint32_t field = (span >= 16) ? 0xffffffffu : ((1u << (2 * span)) - 1u);
    int rounds = 0;
    for (;;) {
        if (tid == 0) s_remaining = 0;
        __syncthreads();
        for (int c = tid; c < n; c += blockDim.x) {
            const int idx = cl[c];
            if (((st2[idx >> 4] >> ((idx & 15) * 2)) & 3u) != 1u) continue;
            const float s = __ldg(prob + idx);
            const int y = idx / W, x = idx - y * W;
            bool sup = false, wait = false;
            for (int dy = -R; dy <= R; dy++) {
                const int p0 = (y + dy) * W + x - R;
                const int wd = p0 >> 4, sh = (p0 & 15) * 2;
                const unsigned long long two = (unsigned long long)st2[wd] | ((unsigned long long)st2[wd + 1] << 32);
                const uint32_t bits = (uint32_t)(two >> sh) & field;
                const uint32_t lo = bits & 0x55555555u, hi = (bits >> 1) & 0x55555555u;
                if (hi & ~lo) {  // an accepted neighbour
                    sup = true;
                    break;
                }
                uint32_t und = lo & ~hi;
                if (dy == 0) und &= ~(1u << (2 * R));  // the candidate itself
                while (und && !wait) {
                    const int k = (__ffs(und) - 1) >> 1;
                    und &= und - 1;
                    const int ni = p0 + k;
                    const float sn = __ldg(prob + ni);
                    if ((sn > s) || (sn == s && ni < idx)) wait = true;
                }
            }
            if (sup) {
                atomicOr(&st2[idx >> 4], 2u << ((idx & 15) * 2));  // 01 -> 11 suppressed
            } else if (!wait) {
                atomicXor(&st2[idx >> 4], 3u << ((idx & 15) * 2));  // 01 -> 10 accepted
                const int k = atomicAdd(&s_nacc, 1);
                if (k < p.acc_cap)
                    keys[k] = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned)idx;
                else
                    s_ovf = 1;
            } else {
                atomicAdd(&s_remaining, 1);
            }
        }
        __syncthreads();
        rounds++;
        const int rem = s_remaining;
        __syncthreads();
        if (rem == 0 || rounds > n + 1) break;
    }
    nms_finish(p, b, keys, s_nacc, s_ovf, rounds, n,
               [&](int idx) { return ((st2[idx >> 4] >> ((idx & 15) * 2)) & 3u) == 2u; });
}

// ------------------------------------------------------------------------------------------------
// K8: refineHeatMap (:540-578) -- one warp per 16x16 tile, 8 raster-consecutive values per lane.
__global__ void __launch_bounds__(256) refine_kernel(const PostParams p, float* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int tiles_x = p.W >> 4, tiles_y = p.H >> 4;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int per = tiles_x * tiles_y;
    if (tile >= per * p.B) return;
    const int b = tile / per, r = tile - b * per, ty = r / tiles_x, tx = r - ty * tiles_x;
    const size_t base = (size_t)b * p.H * p.W + (size_t)(ty * 16 + (lane >> 1)) * p.W + tx * 16 + (lane & 1) * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(p.heat_raw + base);
    const float4 a1 = *reinterpret_cast<const float4*>(p.heat_raw + base + 4);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    uint32_t u[8];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const bool valid = v[k] > p.line_valid_thresh;  // :549
        u[k] = valid ? __float_as_uint(v[k]) : 0u;
        cnt += valid;
    }
    const int n = __reduce_add_sync(FULL, cnt);
    const int valCount = (int)(p.line_valid_ratio * (float)n);  // :554
    bool unchanged = valCount < 1, zero = false;                // :555 leaves the tile as it is
    if (!unchanged && (double)n >= 256.0 * 0.9) {               // :557
        const int k = (int)((double)n * 0.9);                   // index into the raster-ordered valid list
        const int inc = warp_incl_scan(cnt, lane);
        const int excl = inc - cnt;
        float cand = 0.f;
        const bool mine = (k >= excl) && (k < inc);
        if (mine) {
            int want = k - excl, seen = 0;
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (u[e]) {
                    if (seen == want) cand = v[e];
                    seen++;
                }
        }
        const unsigned who = __ballot_sync(FULL, mine);
        const float vk = who ? __shfl_sync(FULL, cand, __ffs(who) - 1) : 0.f;
        zero = who && ((double)vk > 0.1);
    }
    float o[8];
    if (unchanged) {
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = v[k];
    } else if (zero) {
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = 0.f;
    } else {
        // valCount-th largest valid value by bitwise binary search (positive floats order as uints): t ends as the largest
        // threshold with at least valCount values >= t.  Two shortcuts (the ncu capture showed the kernel issue-bound on
        // this loop, 31 rounds of 8 compares + a warp reduction): the bits above the highest bit in which the tile's
        // largest and smallest valid values differ are common to all of them, and as soon as EXACTLY valCount values are
        // >= t the answer is the smallest of those -- one more reduction instead of the remaining rounds.
        uint32_t umax = 0, umin = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            umax = max(umax, u[k]);
            if (u[k]) umin = min(umin, u[k]);
        }
        umax = __reduce_max_sync(FULL, umax);
        umin = __reduce_min_sync(FULL, umin);
        const uint32_t diff = umax ^ umin;
        uint32_t t = umax;  // all valid values equal
        if (diff) {
            const int top = 31 - __clz(diff);
            t = umax & ~((2u << top) - 1u);  // the common prefix: every valid value is >= it
            int ge_t = n;                    // values >= t
            for (int bit = top; bit >= 0 && ge_t != valCount; bit--) {
                const uint32_t c = t | (1u << bit);
                int ge = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) ge += (u[k] >= c);
                ge = __reduce_add_sync(FULL, ge);
                if (ge >= valCount) {
                    t = c;
                    ge_t = ge;
                }
            }
            if (ge_t == valCount) {  // the valCount-th largest is the smallest value >= t
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (u[k] >= t) m = min(m, u[k]);
                t = __reduce_min_sync(FULL, m);
            }
        }
        // sum of the top valCount values; double accumulation of <= 76 floats in (0.01, 1] is exact, so
        // the order is irrelevant (:563 std::accumulate(..., 0.0) over the sorted prefix)
        double sum = 0.0;
        int gt = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (u[k] > t) {
                sum += (double)v[k];
                gt++;
            }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            sum += __shfl_xor_sync(FULL, sum, s);
            gt += __shfl_xor_sync(FULL, gt, s);
        }
        sum += (double)(valCount - gt) * (double)__uint_as_float(t);
        const float ave = (float)(sum / (double)(float)valCount);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (u[k]) {
                const float ns = __fdiv_rn(v[k], ave);
                o[k] = (ns > 1.0f) ? 1.0f : ns;  // :570 compares as double: the same predicate
            } else {
                o[k] = 0.f;  // :573
            }
        }
    }
    *reinterpret_cast<float4*>(dst + base) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + base + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

// ------------------------------------------------------------------------------------------------
// K9: cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) as OpenCV evaluates it for CV_32F: coordinates in
// 1/32 fixed point, four f32 weights, sequential f32 sum.  4 pixels per thread.
// remap_lut (built once at ppg_create from rint(map * 32), api.cu) holds per pixel .x = iy * W + ix of the top-left
// source texel and .y = fx | fy << 5 | in-image bits of the four texels << 10: the first version kept (sx, sy) and
// spent two thirds of its instructions on shifts, bounds tests and address arithmetic -- it was issue-bound at 0.71 of
// the HBM rate.
// A thread handles pixels t, t + 256, t + 512, t + 768 of its block's 1024: every warp-wide access -- table, the four
// gathers of a pixel, the store -- then covers 32 neighbouring pixels, i.e. one or two 128-byte lines (with four
// CONSECUTIVE pixels per thread a gather touched four lines for a quarter of their bytes: 76 L1 wavefronts per 128
// pixels instead of 28).
__global__ void __launch_bounds__(256) remap_kernel(const PostParams p) {
    const int b = blockIdx.y, HW = p.H * p.W, W = p.W;
    const int q0 = blockIdx.x * 1024 + threadIdx.x;
    const float* src = p.heat_ref + (size_t)b * HW;
    int2 l[4];
#pragma unroll
    for (int k = 0; k < 4; k++) l[k] = (q0 + k * 256 < HW) ? p.remap_lut[q0 + k * 256] : make_int2(0, 0);
    float s[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {  // all sixteen loads first
        const float* t = src + l[k].x;
        const uint32_t pk = (uint32_t)l[k].y;
        s[k][0] = (pk & 0x400u) ? t[0] : 0.f;
        s[k][1] = (pk & 0x800u) ? t[1] : 0.f;
        s[k][2] = (pk & 0x1000u) ? t[W] : 0.f;
        s[k][3] = (pk & 0x2000u) ? t[W + 1] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t pk = (uint32_t)l[k].y;
        const float fx = (float)(pk & 31u) * (1.0f / 32.0f), fy = (float)((pk >> 5) & 31u) * (1.0f / 32.0f);
        const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;
        const float o = ((s[k][0] * w00 + s[k][1] * w01) + s[k][2] * w10) + s[k][3] * w11;
        if (q0 + k * 256 < HW) p.heat_final[(size_t)b * HW + q0 + k * 256] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// K10: 3-point heat test for every pair i<j of in-bounds keypoints (:293-313).  One warp per row i;
// the result is a bit matrix so that the (i asc, j asc) order the greedy filter needs is implicit:
// candidate id(i,j) = row_off[i] + row_prefix[i][j/32] + popc(bits[i][j/32] below j).
__global__ void __launch_bounds__(256) pair_test_kernel(const PostParams p) {
    __shared__ float sx[POST_MAX_KP], sy[POST_MAX_KP];
    __shared__ uint8_t so[POST_MAX_KP];
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n = hdr_of(p, b)[HDR_NKP];
    const float* xun = out_of<float>(p, b, p.lay.xun);
    const float* yun = out_of<float>(p, b, p.lay.yun);
    const uint8_t* ko = out_of<uint8_t>(p, b, p.lay.kout);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        sx[t] = xun[t];
        sy[t] = yun[t];
        so[t] = ko[t];
    }
    __syncthreads();
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* heat = p.heat_final + (size_t)b * p.H * p.W;
    uint32_t* bits = p.pair_bits + ((size_t)b * p.max_kp + i) * p.pair_words;
    uint16_t* prefix = p.row_prefix + ((size_t)b * p.max_kp + i) * p.pair_words;
    uint32_t* sym = p.sym_bits + ((size_t)b * p.max_kp + i) * p.pair_words;
    const float xi = sx[i], yi = sy[i];
    const bool iout = so[i] != 0;
    const float th = p.line_heatmap_thresh;
    int total = 0;
    const int words = (n + 31) >> 5;
    for (int w = 0; w < words; w++) {
        const int j = w * 32 + lane;
        bool pass = false;
        if (!iout && j > i && j < n && !so[j]) {
            const float xj = sx[j], yj = sy[j];
            const float c1x = xj * 0.2f + xi * 0.8f, c1y = yj * 0.2f + yi * 0.8f;  // :304
            const float c2x = xj * 0.8f + xi * 0.2f, c2y = yj * 0.8f + yi * 0.2f;  // :305
            const float c3x = xj * 0.5f + xi * 0.5f, c3y = yj * 0.5f + yi * 0.5f;  // :306
            pass = !(heat[(int)((double)c1y + 0.5) * p.W + (int)((double)c1x + 0.5)] < th) &&
                   !(heat[(int)((double)c2y + 0.5) * p.W + (int)((double)c2x + 0.5)] < th) &&
                   !(heat[(int)((double)c3y + 0.5) * p.W + (int)((double)c3x + 0.5)] < th);
        }
        const unsigned m = __ballot_sync(FULL, pass);
        if (lane == 0) {
            bits[w] = m;
            sym[w] = m;  // cand_build_kernel ORs the transposed bits in
            prefix[w] = (uint16_t)total;
        }
        total += __popc(m);
    }
    if (lane == 0) p.row_cnt[b * p.max_kp + i] = total;
}

// Candidate tables in global memory, per frame [pair_cap].
struct CandTables {
    uint32_t* se;  // s | e << 16
    float* dist;
    float* dirf;   // dir(s,e)  (:283)
    float* dirb;   // dir(e,s)  (:284-286)
};
__device__ __forceinline__ CandTables cand_of(const PostParams& p, int b) {
    CandTables t;
    const size_t o = (size_t)b * p.pair_cap;
    t.se = p.c_se + o;
    t.dist = p.c_dist + o;
    t.dirf = p.c_dirf + o;
    t.dirb = p.c_dirb + o;
    return t;
}
__device__ __forceinline__ int cand_id(const PostParams& p, int b, const int* row_off, int lo, int hi) {
    const size_t r = ((size_t)b * p.max_kp + lo) * p.pair_words + (hi >> 5);
    return row_off[lo] + p.row_prefix[r] + __popc(p.pair_bits[r] & ((1u << (hi & 31)) - 1u));
}

// K10b: row offsets (every block rescans the <= 1024 row counts) + candidate table (:265-288 evaluated only
// for the pairs that passed).  grid (ceil(max_kp/8), B), one warp per row.
__global__ void __launch_bounds__(256) cand_build_kernel(const PostParams p) {
    __shared__ int s_off[POST_MAX_KP + 1];
    __shared__ int ws[40];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int n = hdr_of(p, b)[HDR_NKP];
    int carry = 0;
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        const int t = t0 + tid;
        const int v = t < n ? p.row_cnt[b * p.max_kp + t] : 0;
        int tot;
        const int ex = block_excl_scan(v, ws, &tot);
        if (t < n) s_off[t] = carry + ex;
        carry += tot;
    }
    if (tid == 0) s_off[n] = carry;
    __syncthreads();
    int* row_off = p.row_off + (size_t)b * (p.max_kp + 1);
    if (blockIdx.x == 0)
        for (int t = tid; t <= n; t += blockDim.x) row_off[t] = s_off[t];
    const int i = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
    if (i >= n) return;
    const float* xun = out_of<float>(p, b, p.lay.xun);
    const float* yun = out_of<float>(p, b, p.lay.yun);
    const CandTables T = cand_of(p, b);
    const uint32_t* bits = p.pair_bits + ((size_t)b * p.max_kp + i) * p.pair_words;
    const float xi = xun[i], yi = yun[i];
    int pos = s_off[i];
    const int words = (n + 31) >> 5;
    for (int w = 0; w < words; w++) {
        const uint32_t mw = bits[w];
        if (mw & (1u << lane)) {
            const int id = pos + __popc(mw & ((1u << lane) - 1u));
            if (id < p.pair_cap) {
                const int j = w * 32 + lane;
                const float dx = xun[j] - xi, dy = yun[j] - yi;    // :278
                const float dist = sqrtf(dx * dx + dy * dy);       // :279
                const float ux = __fdiv_rn(dx, dist), uy = __fdiv_rn(dy, dist);
                const float d = (float)atan2((double)uy, (double)ux);  // :283 (see oracle divergence 2)
                float r = (float)((double)d - PPG_PI);                 // :284
                if ((double)r < -PPG_PI) r = (float)((double)r + PPG_2PI);
                T.dist[id] = dist;
                T.dirf[id] = d;
                T.dirb[id] = r;
                T.se[id] = (uint32_t)i | ((uint32_t)j << 16);
                atomicOr(p.sym_bits + ((size_t)b * p.max_kp + j) * p.pair_words + (i >> 5), 1u << (i & 31));
            }
        }
        pos += __popc(mw);
    }
}

// K11a: the pairwise part of the overlap filter (:316-335 / :338-357), evaluated for EVERY (new candidate,
// earlier candidate sharing an endpoint) in parallel.  Whether the earlier line is still alive when the new
// one is processed is the only sequential state; everything else is a pure function of the two candidates:
//   kill  : distNew <= distOld && distNew*sin(a) < thr   -> the old line becomes bad
//   block : distOld <  distNew && distOld*sin(a) < thr   -> the new line is not created
// One warp per candidate; entries = other endpoint q of the old line | type << 15, <= INTER_K per side.
constexpr int INTER_K = 16;  // lanes 0-15 <-> list at s, lanes 16-31 <-> list at e in the sequential pass

__device__ __forceinline__ int interact_one(const PostParams& p, const CandTables& T, int b, const int* row_off, int pt,
                                            int q, float dir_new, float dist_new) {
    const int lo = pt < q ? pt : q, hi = pt < q ? q : pt;
    const int id = cand_id(p, b, row_off, lo, hi);
    if (id >= p.pair_cap) return -1;
    const float dir_old = (pt == lo) ? T.dirf[id] : T.dirb[id];
    float a = dir_new - dir_old;
    if ((double)a < -PPG_PI) a = (float)((double)a + PPG_2PI);
    if ((double)a > PPG_PI) a = (float)((double)a - PPG_2PI);
    a = fabsf(a);
    if ((double)a > 0.2 * PPG_PI) return -1;
    const float dist_old = T.dist[id];
    const float sn = (float)sin((double)a);  // :330 unqualified sin -> double overload
    if (dist_new <= dist_old && dist_new * sn < p.line_dist_thresh) return 0;
    if (dist_old < dist_new && dist_old * sn < p.line_dist_thresh) return 1;
    return -1;
}

// One pass over the earlier candidates that share an endpoint with candidate (i,j):
//   side i: old lines (q,i) with q < i (rows above) and (i,q) with i < q < j (same row, earlier columns)
//           = the neighbours q < j of i in the symmetric pair matrix
//   side j: old lines (q,j) with q < i (rows above row i; later rows are not created yet) = the neighbours q < i of j
// Lane l owns word l of the two rows (max_kp <= 1024) and evaluates only the set bits; entries are written in
// ascending q (= the order of the reference's adjacency lists) while they fit.  (The first version tested all
// q < j with strided reads of the upper-triangular matrix: 0.11 ms per 32 frames.)
__device__ __forceinline__ uint32_t bits_below(uint32_t w, int lane, int limit) {  // keep bits whose index < limit
    const int lo = lane * 32;
    if (lo >= limit) return 0u;
    if (limit - lo >= 32) return w;
    return w & ((1u << (limit - lo)) - 1u);
}

__device__ __forceinline__ void interact_scan(const PostParams& p, const CandTables& T, int b, const int* row_off,
                                              const uint32_t* sym, int nwords, int i, int j, float dirf, float dirb,
                                              float dist_new, int lane, uint16_t* ent_i, int cap_i, uint16_t* ent_j,
                                              int cap_j, int* out_cnt_i, int* out_cnt_j) {
    uint32_t wi = 0u, wj = 0u;
    if (lane < nwords) {
        wi = bits_below(sym[(size_t)i * p.pair_words + lane], lane, j);
        wj = bits_below(sym[(size_t)j * p.pair_words + lane], lane, i);
    }
    uint32_t hit_i = 0u, typ_i = 0u, hit_j = 0u, typ_j = 0u;
    for (uint32_t m = wi; m; m &= m - 1) {
        const int bit = __ffs(m) - 1;
        const int ty = interact_one(p, T, b, row_off, i, lane * 32 + bit, dirf, dist_new);
        if (ty >= 0) {
            hit_i |= 1u << bit;
            typ_i |= (uint32_t)ty << bit;
        }
    }
    for (uint32_t m = wj; m; m &= m - 1) {
        const int bit = __ffs(m) - 1;
        const int ty = interact_one(p, T, b, row_off, j, lane * 32 + bit, dirb, dist_new);
        if (ty >= 0) {
            hit_j |= 1u << bit;
            typ_j |= (uint32_t)ty << bit;
        }
    }
    __syncwarp();
    const int ci = __popc(hit_i), cj = __popc(hit_j);
    const int inc_i = warp_incl_scan(ci, lane), inc_j = warp_incl_scan(cj, lane);
    int k = inc_i - ci;
    for (uint32_t m = hit_i; m; m &= m - 1, k++) {
        const int bit = __ffs(m) - 1;
        if (k < cap_i) ent_i[k] = (uint16_t)((lane * 32 + bit) | (((typ_i >> bit) & 1u) << 15));
    }
    k = inc_j - cj;
    for (uint32_t m = hit_j; m; m &= m - 1, k++) {
        const int bit = __ffs(m) - 1;
        if (k < cap_j) ent_j[k] = (uint16_t)((lane * 32 + bit) | (((typ_j >> bit) & 1u) << 15));
    }
    *out_cnt_i = __shfl_sync(FULL, inc_i, 31);
    *out_cnt_j = __shfl_sync(FULL, inc_j, 31);
}

// Lists longer than INTER_K per side (dense fans of near-parallel candidates) spill to a per-frame pool: the warp
// reserves cnt_i + cnt_j entries and repeats the scan into them.  inter_off = pool offset, or ~0 for inline.
__global__ void __launch_bounds__(256) interact_kernel(const PostParams p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n = hdr_of(p, b)[HDR_NKP];
    if (n == 0) return;
    const int* row_off = p.row_off + (size_t)b * (p.max_kp + 1);
    int npass = row_off[n];
    if (npass > p.pair_cap) npass = p.pair_cap;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= npass) return;
    const CandTables T = cand_of(p, b);
    const uint32_t se = T.se[c];
    const int i = se & 0xffff, j = se >> 16;
    const float dist_new = T.dist[c], dirf = T.dirf[c], dirb = T.dirb[c];
    uint16_t* ent = p.inter + ((size_t)b * p.pair_cap + c) * (2 * INTER_K);
    const uint32_t* sym = p.sym_bits + (size_t)b * p.max_kp * p.pair_words;
    const int nwords = (n + 31) >> 5;
    int cnt_i, cnt_j;
    interact_scan(p, T, b, row_off, sym, nwords, i, j, dirf, dirb, dist_new, lane, ent, INTER_K, ent + INTER_K, INTER_K,
                  &cnt_i, &cnt_j);
    uint32_t off = 0xffffffffu;
    if (cnt_i > INTER_K || cnt_j > INTER_K) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&p.counters[b * 8 + 2], cnt_i + cnt_j);
        base = __shfl_sync(FULL, base, 0);
        if (base + cnt_i + cnt_j <= p.pool_cap) {
            uint16_t* pool = p.inter_pool + (size_t)b * p.pool_cap + base;
            int ci2, cj2;
            interact_scan(p, T, b, row_off, sym, nwords, i, j, dirf, dirb, dist_new, lane, pool, cnt_i, pool + cnt_i,
                          cnt_j, &ci2, &cj2);
            off = (uint32_t)base;
        } else {
            off = 0xfffffffeu;  // pool exhausted: reported as ST_OVF_DEGREE by lines_kernel
        }
    }
    if (lane == 0) {
        p.inter_cnt[(size_t)b * p.pair_cap + c] = (uint32_t)cnt_i | ((uint32_t)cnt_j << 16);
        p.inter_off[(size_t)b * p.pair_cap + c] = off;
    }
}

// ------------------------------------------------------------------------------------------------
// K11b+K12: one CTA per frame.
//   A  alive[p] = bit set over the other endpoint q of every candidate line at p (symmetric matrix)
//   B  the sequential part of the greedy overlap filter: only candidates that HAVE an interaction entry are
//      visited (in candidate order, one thread); kills clear the old line's bits, a block clears the new one's
//   C  line scoring, one thread per surviving candidate (:366-389, :461-513)
//   D  final edges, mvConnected CSR (:433-441)
//   E  colinearity per keypoint, one thread per keypoint (:391-432), mvColine CSR
struct LineSmem {
    uint32_t* alive;   // [max_kp][pair_words]
    uint16_t* adj;     // [max_kp][deg_cap] candidate ids of the final lines at p, ascending
    int* adj_cnt;      // [max_kp]
    int* row_off;      // [max_kp + 1]
    float* kx;
    float* ky;
    int* ws;           // 48 ints scan scratch / flags
    uint32_t* seq_se;  // [SEQ_WIN] s | e << 16 of the candidates with interactions in the current window
    uint32_t* seq_own; // [SEQ_WIN] packed alive-bit address of the candidate's own line (see pack_bit)
    uint32_t* seq_cnt; // [SEQ_WIN] entries at s | entries at e << 16
    uint32_t* seq_off; // [SEQ_WIN] pool offset of spilled lists, ~0 = inline
    uint32_t* seq_ent; // [SEQ_WIN][2*INTER_K] inline lists, one packed alive-bit address per entry, ~0 = none
};
constexpr int SEQ_WIN = 512;  // >= blockDim.x of lines_kernel

// During the sequential pass a line (a,b) owns ONE bit: row min(a,b), column max(a,b) of the alive matrix (the mirror
// bits are rebuilt afterwards).  Packed address: bit index [0,5), word index [5,31), entry type (1 = block) bit 31.
__device__ __forceinline__ uint32_t pack_bit(int words, int a, int b2, uint32_t type) {
    const int lo = a < b2 ? a : b2, hi = a < b2 ? b2 : a;
    return (type << 31) | ((uint32_t)(lo * words + (hi >> 5)) << 5) | (uint32_t)(hi & 31);
}

__device__ __forceinline__ LineSmem carve_lines(const PostParams& p, uint8_t* s) {
    LineSmem m;
    m.alive = reinterpret_cast<uint32_t*>(s);
    m.adj_cnt = reinterpret_cast<int*>(m.alive + (size_t)p.max_kp * p.pair_words);
    m.row_off = m.adj_cnt + p.max_kp;
    m.kx = reinterpret_cast<float*>(m.row_off + p.max_kp + 1);
    m.ky = m.kx + p.max_kp;
    m.ws = reinterpret_cast<int*>(m.ky + p.max_kp);
    // plain pointer arithmetic only: an integer round trip would lose the shared address space and turn every
    // access of the sequential pass into a generic LD/ST
    m.seq_se = reinterpret_cast<uint32_t*>(m.ws + 48);
    m.seq_own = m.seq_se + SEQ_WIN;
    m.seq_cnt = m.seq_own + SEQ_WIN;
    m.seq_off = m.seq_cnt + SEQ_WIN;
    m.seq_ent = m.seq_off + SEQ_WIN;
    // the adjacency rows (lines_graph_kernel) share the region of the sequential-pass window (lines_filter_kernel)
    m.adj = reinterpret_cast<uint16_t*>(m.ws + 48);
    return m;
}

__constant__ float c_inv_gap[4] = {0.3333, 0.200, 0.1427, 0.1111};  // invSampleGapTable, PPGExtractor.cpp:19

__device__ __forceinline__ float bilinear_heat(const float* M, int W, float ptX, float ptY) {  // :580-589
    const int x1 = (int)ptX, x2 = x1 + 1, y1 = (int)ptY, y2 = y1 + 1;
    const float d1 = ((float)x2 - ptX) * M[y1 * W + x1] + (ptX - (float)x1) * M[y1 * W + x2];
    const float d2 = ((float)x2 - ptX) * M[y2 * W + x1] + (ptX - (float)x1) * M[y2 * W + x2];
    return ((float)y2 - ptY) * d1 + (ptY - (float)y1) * d2;
}

__device__ __forceinline__ bool alive_bit(const LineSmem& m, int words, int pt, int q) {
    return (m.alive[pt * words + (q >> 5)] >> (q & 31)) & 1u;
}
__device__ __forceinline__ void clear_line(const LineSmem& m, int words, int a, int b2) {
    atomicAnd(&m.alive[a * words + (b2 >> 5)], ~(1u << (b2 & 31)));
    atomicAnd(&m.alive[b2 * words + (a >> 5)], ~(1u << (a & 31)));
}

// K11b: one CTA per frame -- phases A and B; the alive matrix goes to global memory for the next two kernels.
__global__ void __launch_bounds__(512) lines_filter_kernel(const PostParams p) {
    extern __shared__ __align__(16) uint8_t smem_lines[];
    const LineSmem m = carve_lines(p, smem_lines);
    const int b = blockIdx.x, tid = threadIdx.x;
    int* hdr = hdr_of(p, b);
    const int n = hdr[HDR_NKP];
    if (n == 0) return;  // detectLines returns at :239-240
    const CandTables T = cand_of(p, b);
    const int words = p.pair_words, nwords = (n + 31) >> 5;
    unsigned status = 0;
    long long tphase = clock64();
    int nphase = 0;
    auto phase_mark = [&]() {  // thread 0 only; diagnostic phase durations in the record header
        if (tid == 0 && nphase < HDR_WORDS - HDR_DIAG) {
            const long long t = clock64();
            hdr[HDR_DIAG + nphase] = (int)((t - tphase) >> 4);
            tphase = t;
        }
        nphase++;
    };

    // ---- A: keypoints, row offsets, symmetric alive matrix
    const int* g_row_off = p.row_off + (size_t)b * (p.max_kp + 1);
    for (int t = tid; t <= n; t += blockDim.x) m.row_off[t] = g_row_off[t];
    const uint32_t* bits = p.pair_bits + (size_t)b * p.max_kp * words;
    for (int t = tid; t < n * words; t += blockDim.x) {
        const int w = t % words;
        m.alive[t] = (w < nwords) ? bits[t] : 0u;
    }
    __syncthreads();
    const int npass_all = m.row_off[n];
    int npass = npass_all;
    if (npass > p.pair_cap) {
        npass = p.pair_cap;
        status |= ST_OVF_PAIRS;
        // Pairs beyond the candidate table have no id: drop their bits, so that nothing downstream (scoring, edge
        // list, adjacency rows, colinearity) can look up a candidate that was never built.  The record is flagged
        // and the graph is that of the first pair_cap candidates.
        for (int t = tid; t < n; t += blockDim.x) {
            if (m.row_off[t + 1] <= p.pair_cap) continue;
            int keep = p.pair_cap - m.row_off[t];
            if (keep < 0) keep = 0;
            for (int w = 0; w < nwords; w++) {
                uint32_t v = m.alive[t * words + w];
                const int cnt = __popc(v);
                if (cnt <= keep) {
                    keep -= cnt;
                    continue;
                }
                uint32_t kept = 0u;
                for (; keep > 0; keep--) {
                    const uint32_t low = v & (0u - v);
                    kept |= low;
                    v ^= low;
                }
                m.alive[t * words + w] = kept;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        m.ws[40] = 0;  // blocked count
        m.ws[41] = 0;  // interaction-list overflow
        m.ws[42] = 0;  // diagnostic: candidates visited by the sequential pass
        m.ws[43] = 0;  // diagnostic: of those, lists read from the spill pool
        m.ws[44] = 0;  // diagnostic: visited candidates with at least one block entry (inline lists)
    }
    __syncthreads();

    phase_mark();
    // ---- B: sequential pass over the candidates that interact, window by window
    const uint32_t* g_cnt = p.inter_cnt + (size_t)b * p.pair_cap;
    const uint32_t* g_off = p.inter_off + (size_t)b * p.pair_cap;
    const uint16_t* g_ent = p.inter + (size_t)b * p.pair_cap * (2 * INTER_K);
    const uint16_t* g_pool = p.inter_pool + (size_t)b * p.pool_cap;
    for (int c0 = 0; c0 < npass;) {
        // compact up to SEQ_WIN interacting candidates starting at c0 (block-wide, order preserved)
        int filled = 0, c_next = c0;
        while (filled < SEQ_WIN && c_next < npass) {
            const int c = c_next + tid;
            uint32_t cnt = 0;
            if (c < npass) cnt = g_cnt[c];
            int tot;
            const int ex = block_excl_scan(cnt != 0 ? 1 : 0, m.ws, &tot);
            if (filled + tot > SEQ_WIN) break;  // this chunk does not fit: process what we have first
            if (cnt != 0) {
                const int k = filled + ex;
                const uint32_t off = g_off[c];
                const uint32_t se = T.se[c];
                const int i = se & 0xffff, j = se >> 16;
                m.seq_cnt[k] = cnt;
                m.seq_off[k] = off;
                m.seq_se[k] = se;
                m.seq_own[k] = pack_bit(words, i, j, 0u);
                if (off == 0xffffffffu) {
                    // expand the (other endpoint | type) entries into packed alive-bit addresses here, in parallel,
                    // so that the one-warp sequential pass below is load -> test -> vote -> clear and nothing else
                    const int ci = cnt & 0xffff, cj = cnt >> 16;
                    const uint4* src = reinterpret_cast<const uint4*>(g_ent + (size_t)c * 2 * INTER_K);
                    uint32_t* dst = m.seq_ent + (size_t)k * 2 * INTER_K;
                    bool any_block = false;
#pragma unroll
                    for (int v4 = 0; v4 < 2 * INTER_K * 2 / 16; v4++) {
                        const uint4 u = src[v4];
                        const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int h = 0; h < 8; h++) {
                            const int t = v4 * 8 + h;
                            const uint32_t e = (wv[h >> 1] >> ((h & 1) * 16)) & 0xffffu;
                            const int side = t / INTER_K, tt = t % INTER_K;
                            const bool valid = tt < (side ? cj : ci);
                            dst[t] = valid ? pack_bit(words, side ? j : i, (int)(e & 0x7fff), e >> 15) : 0xffffffffu;
                            any_block |= valid && (e >> 15);
                        }
                    }
                    if (any_block) atomicAdd(&m.ws[44], 1);  // diagnostic: candidates that can be blocked at all
                }
            }
            filled += tot;
            c_next += blockDim.x;
        }
        __syncthreads();
        if (tid < 32) {
            // One warp walks the window in candidate order.  Inside one candidate the entries are independent (each
            // looks only at its own old line), so lanes 0-15 take the list at s and lanes 16-31 the list at e in one
            // step: the kills at s always apply, the list at e is scanned only if no live line at s blocks the
            // candidate (:336-337), and a blocked candidate clears its own bit.
            // The next candidate's record is fetched while the current one is resolved: the only loop-carried
            // dependency is alive-bit load -> ballot -> clear.
            const int lane = tid;
            int blocked = 0, ovf = 0, nspill = 0;
            uint32_t n_own = 0, n_off = 0, n_e = 0;
            if (filled > 0) {
                n_own = m.seq_own[0];
                n_off = m.seq_off[0];
                n_e = m.seq_ent[lane];
            }
            for (int k = 0; k < filled; k++) {
                const uint32_t own = n_own, off = n_off, e = n_e;
                if (k + 1 < filled) {
                    n_own = m.seq_own[k + 1];
                    n_off = m.seq_off[k + 1];
                    n_e = m.seq_ent[(size_t)(k + 1) * 2 * INTER_K + lane];
                }
                if (off == 0xfffffffeu) {
                    ovf = 1;
                    continue;
                }
                bool blk;
                if (off == 0xffffffffu) {
                    const bool valid = e != 0xffffffffu;
                    const uint32_t w = (e >> 5) & 0x3ffffffu, bit = e & 31u;
                    const bool al = valid && ((m.alive[w] >> bit) & 1u);
                    const unsigned mb = __ballot_sync(FULL, al && (e >> 31));
                    const bool blk_i = (mb & ((1u << INTER_K) - 1u)) != 0;
                    if (al && !(e >> 31) && (lane < INTER_K || !blk_i)) atomicAnd(&m.alive[w], ~(1u << bit));
                    blk = mb != 0;  // a block at e only counts when s did not block, and then blk_i is set anyway
                } else {
                    nspill++;
                    const uint32_t se = m.seq_se[k], cnt2 = m.seq_cnt[k];
                    const int i = se & 0xffff, j = se >> 16;
                    const int ci = cnt2 & 0xffff, cj = cnt2 >> 16;
                    bool blk_i = false, blk_j = false;
                    for (int t0 = 0; t0 < ci; t0 += 32) {  // scan of adj[s] (:316-335)
                        const int t = t0 + lane;
                        const uint32_t g = t < ci ? g_pool[off + t] : 0u;
                        const uint32_t pk = pack_bit(words, i, (int)(g & 0x7fff), 0u);
                        const bool al = t < ci && ((m.alive[pk >> 5] >> (pk & 31u)) & 1u);
                        if (al && !(g >> 15)) atomicAnd(&m.alive[pk >> 5], ~(1u << (pk & 31u)));
                        blk_i |= __any_sync(FULL, al && (g >> 15));
                    }
                    if (!blk_i)
                        for (int t0 = 0; t0 < cj; t0 += 32) {  // scan of adj[e] (:338-357)
                            const int t = t0 + lane;
                            const uint32_t g = t < cj ? g_pool[off + ci + t] : 0u;
                            const uint32_t pk = pack_bit(words, j, (int)(g & 0x7fff), 0u);
                            const bool al = t < cj && ((m.alive[pk >> 5] >> (pk & 31u)) & 1u);
                            if (al && !(g >> 15)) atomicAnd(&m.alive[pk >> 5], ~(1u << (pk & 31u)));
                            blk_j |= __any_sync(FULL, al && (g >> 15));
                        }
                    blk = blk_i || blk_j;
                }
                if (blk) {
                    if (lane == 0) atomicAnd(&m.alive[own >> 5], ~(1u << (own & 31u)));
                    blocked++;
                }
                __syncwarp();
            }
            if (lane == 0) {
                m.ws[40] += blocked;
                m.ws[41] |= ovf;
                m.ws[42] += filled;
                m.ws[43] += nspill;
            }
        }
        __syncthreads();
        c0 = c_next;
    }
    if (m.ws[41]) status |= ST_OVF_DEGREE;
    const int ncreated = npass - m.ws[40];
    for (int c = tid; c < npass; c += blockDim.x) {  // mirror bits of the surviving lines: bit s of row e
        const uint32_t se = T.se[c];
        const int i = se & 0xffff, j = se >> 16;
        if (alive_bit(m, words, i, j)) atomicOr(&m.alive[j * words + (i >> 5)], 1u << (i & 31));
    }
    __syncthreads();
    phase_mark();
    uint32_t* alive_g = p.alive_g + (size_t)b * p.max_kp * words;
    for (int t = tid; t < n * words; t += blockDim.x) alive_g[t] = m.alive[t];
    if (tid == 0) {
        hdr[HDR_DIAG + 5] = m.ws[42];
        hdr[HDR_DIAG + 6] = m.ws[43];
        hdr[HDR_DIAG + 2] = m.ws[44];
        hdr[HDR_STATUS] |= (int)status;
        hdr[HDR_NPASS] = npass_all;
        hdr[HDR_NLINES] = ncreated;
    }
}

// K12a: line scoring (:367-389) for every surviving candidate of every frame -- one warp per candidate, grid-wide
// (inside the per-frame CTA this phase took 160 k cycles on 16 warps).  The lanes fetch the samples of the segment in
// parallel, then the values are added in sample order so that the f32 sum is the reference's sequential one.
__global__ void __launch_bounds__(256) lines_score_kernel(const PostParams p) {
    const int b = blockIdx.y, tid = threadIdx.x;
    const int n = hdr_of(p, b)[HDR_NKP];
    if (n == 0) return;
    const float* heat = p.heat_final + (size_t)b * p.H * p.W;
    const CandTables T = cand_of(p, b);
    const int words = p.pair_words;
    int npass = p.row_off[(size_t)b * (p.max_kp + 1) + n];
    if (npass > p.pair_cap) npass = p.pair_cap;
    const uint32_t* alive_g = p.alive_g + (size_t)b * p.max_kp * words;
    const float* kxg = out_of<float>(p, b, p.lay.xun);
    const float* kyg = out_of<float>(p, b, p.lay.yun);
    float* lscore = p.l_score + (size_t)b * p.pair_cap;
    int* ledge = p.l_edge + (size_t)b * p.pair_cap;
    {
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        for (int c = blockIdx.x * nwarps + warp; c < npass; c += gridDim.x * nwarps) {
            const uint32_t se = T.se[c];
            const int s = se & 0xffff, e = se >> 16;
            if (!((alive_g[s * words + (e >> 5)] >> (e & 31)) & 1u)) continue;
            const float psx = kxg[s], psy = kyg[s], pex = kxg[e], pey = kyg[e];
            const float dist = T.dist[c];
            int lenLevel = (int)((double)(dist * p.inv_scale) * 4.0);  // :485
            if (lenLevel > 3) lenLevel = 3;  // dist == diagonal cannot happen (points >= 1 px inside)
            if (lenLevel < 0) lenLevel = 0;
            const int segNum = (int)(dist * c_inv_gap[lenLevel]);     // :486
            const float step = (float)(1.0 / (double)(float)segNum);  // :487
            int cnt = 0;
            float sum = 0.f;
            // up to SC_CH x 32 samples are fetched before anything is added (one round of load latency per line)
            constexpr int SC_CH = 4;
            for (int i00 = 1; i00 < segNum; i00 += 32 * SC_CH) {
                float v[SC_CH];
                bool inl[SC_CH];
#pragma unroll
                for (int ch = 0; ch < SC_CH; ch++) {
                    const int i = i00 + ch * 32 + lane;
                    v[ch] = 0.f;
                    inl[ch] = false;
                    if (i < segNum) {
                        const float qx = (psx * step) * (float)i + (pex * step) * (float)(segNum - i);
                        const float qy = (psy * step) * (float)i + (pey * step) * (float)(segNum - i);
                        const int posx = (int)((double)qx + 0.5), posy = (int)((double)qy + 0.5);
                        inl[ch] = heat[posy * p.W + posx] > p.line_heatmap_thresh;
                        v[ch] = bilinear_heat(heat, p.W, qx, qy);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < SC_CH; ch++) {
                    const int i0 = i00 + ch * 32;
                    if (i0 < segNum) {
                        cnt += __popc(__ballot_sync(FULL, inl[ch]));
                        const int mm = min(32, segNum - i0);
                        for (int t = 0; t < mm; t++) sum += __shfl_sync(FULL, v[ch], t);
                    }
                }
            }
            if (lane == 0) {
                const float rate = __fdiv_rn((float)cnt, (float)(segNum - 1));  // segNum == 1 -> 0/0 = NaN, accepted (:376)
                bool bad = false;
                float sh = 0.f;
                if (rate < p.line_inlier_rate) {
                    bad = true;
                } else {
                    sh = __fdiv_rn(sum, (float)(segNum - 1));
                    if (sh < p.line_heatmap_thresh) bad = true;
                }
                // bad lines: marker only; lines_graph_kernel clears their bits
                lscore[c] = bad ? -INFINITY : rate * sh;
                ledge[c] = bad ? -2 : -1;
            }
        }
    }
}

// K12b: one CTA per frame -- final edges, mvConnected, colinearity.
__global__ void __launch_bounds__(512) lines_graph_kernel(const PostParams p) {
    extern __shared__ __align__(16) uint8_t smem_lines[];
    const LineSmem m = carve_lines(p, smem_lines);
    const int b = blockIdx.x, tid = threadIdx.x;
    int* hdr = hdr_of(p, b);
    const int n = hdr[HDR_NKP];
    int* conn_off = out_of<int>(p, b, p.lay.conn_off);
    int* col_off = out_of<int>(p, b, p.lay.col_off);
    if (n == 0) {  // detectLines returns at :239-240; the record carries no edges
        if (tid == 0) {
            conn_off[0] = 0;
            col_off[0] = 0;
        }
        return;
    }
    const CandTables T = cand_of(p, b);
    const int words = p.pair_words, nwords = (n + 31) >> 5;
    unsigned status = 0;
    long long tphase = clock64();
    int nphase = 3;
    auto phase_mark = [&]() {  // thread 0 only; diagnostic phase durations in the record header
        if (tid == 0 && nphase < HDR_WORDS - HDR_DIAG) {
            const long long t = clock64();
            hdr[HDR_DIAG + nphase] = (int)((t - tphase) >> 4);
            tphase = t;
        }
        nphase++;
    };
    const int* g_row_off = p.row_off + (size_t)b * (p.max_kp + 1);
    for (int t = tid; t < n; t += blockDim.x) m.adj_cnt[t] = 0;
    for (int t = tid; t <= n; t += blockDim.x) m.row_off[t] = g_row_off[t];
    const uint32_t* alive_g = p.alive_g + (size_t)b * p.max_kp * words;
    for (int t = tid; t < n * words; t += blockDim.x) m.alive[t] = alive_g[t];
    __syncthreads();
    int npass = m.row_off[n];
    if (npass > p.pair_cap) npass = p.pair_cap;
    float* lscore = p.l_score + (size_t)b * p.pair_cap;
    int* ledge = p.l_edge + (size_t)b * p.pair_cap;
    for (int c = tid; c < npass; c += blockDim.x) {
        const uint32_t se = T.se[c];
        const int s = se & 0xffff, e = se >> 16;
        if (alive_bit(m, words, s, e) && ledge[c] == -2) clear_line(m, words, s, e);
    }
    __syncthreads();

    // ---- D: final edge list = surviving candidates in candidate order (:433-441)
    int ecarry = 0;
    int* es = out_of<int>(p, b, p.lay.edge_s);
    int* ee = out_of<int>(p, b, p.lay.edge_e);
    float* esc = out_of<float>(p, b, p.lay.edge_score);
    for (int c0 = 0; c0 < npass; c0 += blockDim.x) {
        const int c = c0 + tid;
        uint32_t se = 0;
        int v = 0;
        if (c < npass) {
            se = T.se[c];
            v = alive_bit(m, words, se & 0xffff, se >> 16) ? 1 : 0;
        }
        int tot;
        const int ex = block_excl_scan(v, m.ws, &tot);
        if (v) {
            const int E = ecarry + ex;
            ledge[c] = E;
            if (E < p.lay.max_edges) {
                es[E] = se & 0xffff;
                ee[E] = se >> 16;
                esc[E] = lscore[c];
            }
        }
        ecarry += tot;
    }
    const int nedges = ecarry;
    if (nedges > p.lay.max_edges) status |= ST_OVF_EDGES;
    __syncthreads();

    // adjacency rows: final lines at p in ascending other-endpoint order = ascending candidate id (= the order
    // in which :366-388 re-pushes them)
    bool deg_ovf = false;
    for (int t = tid; t < n; t += blockDim.x) {
        uint16_t* row = m.adj + (size_t)t * p.deg_cap;
        int k = 0;
        for (int w = 0; w < nwords; w++) {
            uint32_t mw = m.alive[t * words + w];
            while (mw) {
                const int q = w * 32 + __ffs(mw) - 1;
                mw &= mw - 1;
                const int lo = t < q ? t : q, hi = t < q ? q : t;
                if (k < p.deg_cap)
                    row[k] = (uint16_t)cand_id(p, b, m.row_off, lo, hi);
                else
                    deg_ovf = true;
                k++;
            }
        }
        m.adj_cnt[t] = k < p.deg_cap ? k : p.deg_cap;
    }
    if (__syncthreads_or(deg_ovf)) status |= ST_OVF_DEGREE;
    // mvConnected CSR
    int* conn_idx = out_of<int>(p, b, p.lay.conn_idx);
    int ccarry = 0;
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        const int t = t0 + tid;
        const int v = t < n ? m.adj_cnt[t] : 0;
        int tot;
        const int ex = block_excl_scan(v, m.ws, &tot);
        if (t < n) {
            const int off = ccarry + ex;
            conn_off[t] = off;
            if (nedges <= p.lay.max_edges)
                for (int q = 0; q < v; q++) conn_idx[off + q] = ledge[m.adj[(size_t)t * p.deg_cap + q]];
        }
        ccarry += tot;
    }
    if (tid == 0) conn_off[n] = ccarry;

    phase_mark();
    // ---- E: colinearity (:392-432), in place on the adjacency row; pair k is parked at row[D-2(k+1)..]
    int ncol_mine = 0;
    int colcarry = 0;
    int* col_pairs = out_of<int>(p, b, p.lay.col_pairs);
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        const int pt = t0 + tid;
        int D = 0;
        ncol_mine = 0;
        if (pt < n) {
            uint16_t* row = m.adj + (size_t)pt * p.deg_cap;
            D = m.adj_cnt[pt];
            int mm = D;
            while (mm > 1) {
                const int l1 = row[mm - 1];
                const uint32_t se1 = T.se[l1];
                const int s1 = se1 & 0xffff, e1 = se1 >> 16;
                const int p1 = (pt == s1) ? e1 : s1;
                const float dir1 = (pt == s1) ? T.dirf[l1] : T.dirb[l1];
                const float dist1 = T.dist[l1];
                double minPD = 1e9;
                int best = -1, bp2 = -1;
                for (int i = 0; i < mm - 1; i++) {
                    const int l2 = row[i];
                    const uint32_t se2 = T.se[l2];
                    const int s2 = se2 & 0xffff, e2 = se2 >> 16;
                    const int p2 = (pt == s2) ? e2 : s2;
                    const float dir2 = (pt == s2) ? T.dirf[l2] : T.dirb[l2];
                    const float ad = dir1 - dir2;
                    const double pd = 0.5 * (double)(dist1 + T.dist[l2]) * (double)fabsf((float)sin((double)ad));  // :413
                    if (minPD > pd) {
                        minPD = pd;
                        best = i;
                        bp2 = p2;
                    }
                }
                if (minPD > (double)p.line_dist_thresh) {  // :421
                    mm--;
                    continue;
                }
                mm--;
                row[best] = row[mm - 1];  // :427-430
                mm--;
                const int slot = D - 2 * (ncol_mine + 1);
                row[slot] = (uint16_t)p1;
                row[slot + 1] = (uint16_t)bp2;
                ncol_mine++;
            }
        }
        int tot;
        const int ex = block_excl_scan(ncol_mine, m.ws, &tot);
        if (pt < n) {
            const int off = colcarry + ex;
            col_off[pt] = off;
            const uint16_t* row = m.adj + (size_t)pt * p.deg_cap;
            for (int k = 0; k < ncol_mine; k++) {
                const int slot = D - 2 * (k + 1);
                if (off + k < p.lay.max_col) {
                    col_pairs[2 * (off + k)] = row[slot];
                    col_pairs[2 * (off + k) + 1] = row[slot + 1];
                }
            }
        }
        colcarry += tot;
    }
    if (colcarry > p.lay.max_col) status |= ST_OVF_COLINE;
    phase_mark();
    if (tid == 0) {
        col_off[n] = colcarry;
        hdr[HDR_NEDGES] = nedges;
        hdr[HDR_NCOL] = colcarry;
        hdr[HDR_STATUS] |= (int)status;
    }
}

// ------------------------------------------------------------------------------------------------
// K13: genPointDescriptor (:515-538): grid_sampler(bilinear, zeros, align_corners=false) at the DISTORTED
// integer pixel + F.normalize.  One warp per keypoint, lane = 8 consecutive channels of the NHWC map.
__global__ void __launch_bounds__(256) desc_kernel(const PostParams p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n = hdr_of(p, b)[HDR_NKP];
    if (k >= n) return;
    float* o = out_of<float>(p, b, p.lay.desc) + (size_t)k * 256 + lane * 8;
    if (n < 10) {  // :520-524
        *reinterpret_cast<float4*>(o) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(o + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int px = out_of<int>(p, b, p.lay.px)[k], py = out_of<int>(p, b, p.lay.py)[k];
    const float gx = (float)((double)__fdiv_rn((float)px, (float)p.W) * 2. - 1.);  // :530
    const float gy = (float)((double)__fdiv_rn((float)py, (float)p.H) * 2. - 1.);  // :531
    const float ix = __fdiv_rn((gx + 1.f) * (float)p.Wc - 1.f, 2.f), iy = __fdiv_rn((gy + 1.f) * (float)p.Hc - 1.f, 2.f);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((float)x1 - ix) * ((float)y1 - iy), ne = (ix - (float)x0) * ((float)y1 - iy);
    const float sw = ((float)x1 - ix) * (iy - (float)y0), se = (ix - (float)x0) * (iy - (float)y0);
    const float* d = p.desc + (size_t)b * p.Hc * p.Wc * 256 + lane * 8;
    float r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = 0.f;
    auto tap = [&](int yy, int xx, float w) {
        if (yy >= 0 && yy < p.Hc && xx >= 0 && xx < p.Wc) {
            const float* q = d + ((size_t)yy * p.Wc + xx) * 256;
            const float4 a = *reinterpret_cast<const float4*>(q), c4 = *reinterpret_cast<const float4*>(q + 4);
            const float v[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int c = 0; c < 8; c++) r[c] += v[c] * w;
        }
    };
    tap(y0, x0, nw);
    tap(y0, x1, ne);
    tap(y1, x0, sw);
    tap(y1, x1, se);
    double ss = 0.0;
#pragma unroll
    for (int c = 0; c < 8; c++) ss += (double)r[c] * (double)r[c];
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) ss += __shfl_xor_sync(FULL, ss, s);
    float nrm = (float)sqrt(ss);
    if (nrm < 1e-12f) nrm = 1e-12f;  // F.normalize eps
#pragma unroll
    for (int c = 0; c < 8; c++) r[c] = __fdiv_rn(r[c], nrm);
    *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

OutLayout make_out_layout(int max_kp, int max_edges, int max_col) {
    OutLayout L;
    L.max_kp = max_kp;
    L.max_edges = max_edges;
    L.max_col = max_col;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o = align_up(o + bytes, 16);
        return r;
    };
    L.hdr = take(HDR_WORDS * 4);
    L.kp_x = take((size_t)max_kp * 4);
    L.kp_y = take((size_t)max_kp * 4);
    L.px = take((size_t)max_kp * 4);
    L.py = take((size_t)max_kp * 4);
    L.score = take((size_t)max_kp * 4);
    L.xun = take((size_t)max_kp * 4);
    L.yun = take((size_t)max_kp * 4);
    L.kout = take((size_t)max_kp);
    L.edge_s = take((size_t)max_edges * 4);
    L.edge_e = take((size_t)max_edges * 4);
    L.edge_score = take((size_t)max_edges * 4);
    L.conn_off = take((size_t)(max_kp + 1) * 4);
    L.conn_idx = take((size_t)max_edges * 8);
    L.col_off = take((size_t)(max_kp + 1) * 4);
    L.col_pairs = take((size_t)max_col * 8);
    o = align_up(o, 256);
    L.small_total = o;
    L.desc = take((size_t)max_kp * 256 * 4);
    L.total = align_up(o, 256);
    return L;
}

size_t post_nms_smem(const PostParams& p) {
    return (size_t)p.acc_cap * 8 + (p.nms_smem ? ((size_t)p.H * p.W / 16 + 2) * 4 : 0);
}

// Chooses the NMS variant for the frame shape: bitmap in shared memory when it fits next to >= 2048 ranking keys.
void post_plan_nms(PostParams& p) {
    const size_t budget = 226 * 1024;
    const size_t bitmap = ((size_t)p.H * p.W / 16 + 2) * 4;
    p.nms_smem = 0;
    p.acc_cap = 8192;
    if ((p.H * p.W) % 16 != 0) return;
    if (2 * p.nms_radius + 1 > 16) return;  // a window row must fit the 32-bit field the kernel extracts
    for (int cap = 8192; cap >= 2048; cap >>= 1)
        if (bitmap + (size_t)cap * 8 <= budget) {
            p.nms_smem = 1;
            p.acc_cap = cap;
            return;
        }
}

// shared memory both per-frame kernels carve: alive matrix, adj_cnt, row_off, kx, ky, ws
size_t post_lines_fixed_smem(int max_kp, int pair_words) {
    size_t s = (size_t)max_kp * pair_words * 4;
    s += (size_t)(max_kp * 4 + 1) * 4 + 48 * 4 + 16;
    return align_up(s, 16);
}

size_t post_lines_filter_smem(const PostParams& p) {  // + seq_se, seq_own, seq_cnt, seq_off, seq_ent
    return post_lines_fixed_smem(p.max_kp, p.pair_words) + (size_t)SEQ_WIN * (4 + 4 + 4 + 4 + 2 * INTER_K * 4);
}

size_t post_lines_smem(const PostParams& p) {  // + adjacency rows
    return align_up(post_lines_fixed_smem(p.max_kp, p.pair_words) + (size_t)p.max_kp * p.deg_cap * 2, 16);
}

cudaError_t post_init_attrs(const PostParams& p) {
    // The attribute belongs to the function, not to the ctx: contexts with different frame shapes share it, so it is
    // set to the opt-in maximum rather than to this ctx's size.
    (void)p;
    const int kMax = 226 * 1024;  // opt-in maximum minus room for the kernels' few static __shared__ words
    cudaError_t e = cudaFuncSetAttribute(nms_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nms_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lines_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(lines_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax);
}

cudaError_t post_keypoints_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark) {
    const int HW = p.H * p.W;
    if (!p.scan_fused) {
        cudaError_t e = cudaMemsetAsync(p.counters, 0, sizeof(int) * 8 * p.B, st);
        if (e != cudaSuccess) return e;
        dim3 g((HW / 4 + 256 * SCAN_Q - 1) / (256 * SCAN_Q), p.B);
        // ceil(2^32 / W): __umulhi(pix, magic) == pix / W for every pix < 2^32 / W (H * W is far below that)
        const uint32_t w_magic = (uint32_t)(((1ull << 32) + (uint64_t)p.W - 1) / (uint64_t)p.W);
        scan_kernel<<<g, 256, 0, st>>>(p, w_magic);
        *launches += 1;
        mark("post.scan");
    }
    if (p.nms_smem)
        nms_smem_kernel<<<p.B, 1024, post_nms_smem(p), st>>>(p);
    else
        nms_global_kernel<<<p.B, 1024, post_nms_smem(p), st>>>(p);
    *launches += 1;
    mark("post.nms+topk");
    return cudaGetLastError();
}

cudaError_t post_heat_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark) {
    const int tiles = (p.W >> 4) * (p.H >> 4) * p.B;
    refine_kernel<<<(tiles + 7) / 8, 256, 0, st>>>(p, p.do_remap ? p.heat_ref : p.heat_final);
    *launches += 1;
    mark("post.refine");
    if (p.do_remap) {
        dim3 g((p.H * p.W + 1023) / 1024, p.B);
        remap_kernel<<<g, 256, 0, st>>>(p);
        *launches += 1;
        mark("post.remap");
    }
    return cudaGetLastError();
}

cudaError_t post_lines_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark) {
    dim3 g((p.max_kp + 7) / 8, p.B);
    pair_test_kernel<<<g, 256, 0, st>>>(p);
    mark("post.pair_test");
    cand_build_kernel<<<g, 256, 0, st>>>(p);
    mark("post.cand_build");
    interact_kernel<<<dim3((p.pair_cap + 7) / 8, p.B), 256, 0, st>>>(p);
    mark("post.interact");
    lines_filter_kernel<<<p.B, 512, post_lines_filter_smem(p), st>>>(p);
    mark("post.lines_filter");
    lines_score_kernel<<<dim3(16, p.B), 256, 0, st>>>(p);
    mark("post.lines_score");
    lines_graph_kernel<<<p.B, 512, post_lines_smem(p), st>>>(p);
    mark("post.lines_graph");
    *launches += 6;
    return cudaGetLastError();
}

cudaError_t post_desc_launch(const PostParams& p, cudaStream_t st, long long* launches, PostMark mark) {
    dim3 g((p.max_kp + 7) / 8, p.B);
    desc_kernel<<<g, 256, 0, st>>>(p);
    *launches += 1;
    mark("post.descriptors");
    return cudaGetLastError();
}

}  // namespace ppg
