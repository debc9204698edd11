// Bag-of-words transform of a frame's descriptors: DBoW3::Vocabulary::transform(features, BowVector&, FeatureVector&,
// levelsup) as called by Frame::ComputeBoW (map/src/Frame.cpp:331-340) and at keyframe creation (:127-131).
// DBoW3 is a third-party dependency of the reference; include/ppg_b200.h (ppg_vocabulary) states the restated algorithm.
//
//   B1 bow_descend_kernel  one warp per feature: at every level lane c < k accumulates DescManip::distance to child c
//                          (float products summed in double, in index order -- the order is part of the result
//                          because ties and near-ties pick the branch), the warp takes the first minimum, until a leaf.
//   B2 bow_vector_kernel   one CTA per frame: (word, feature) keys of the features with a positive weight sorted in
//                          shared memory, one thread per word adds the weights of its run in feature order (what
//                          BowVector::addWeight does through the std::map), thread 0 takes the L1 / L2 norm over the
//                          words in ascending id (sequential double sum, as BowVector::normalize), all divide.
// Built with -fmad=false (the CPU sums are not fused).
#include <math.h>
#include <string.h>

#include <string>

#include "assoc.cuh"
#include "ctx.cuh"

namespace ppg {

struct BowState {
    int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, dim = 0;
    int* children = nullptr;
    int* word_id = nullptr;
    double* weight = nullptr;
    float* desc = nullptr;
    int ncap = 0, bcap = 0;
    float* fdesc = nullptr;  // staged descriptors of one frame (host path)
    int* f_word = nullptr;   // [bcap][ncap]
    double* f_weight = nullptr;
    int* f_node = nullptr;
    int* bow_word = nullptr;
    double* bow_value = nullptr;
    int* nb = nullptr;  // [bcap][2]: BowVector size, features
    uint8_t* h = nullptr;
    size_t h_bytes = 0;
};

namespace {

constexpr int BOW_NCAP = 1024;

__global__ void __launch_bounds__(256) bow_descend_kernel(FrameSrc src, int ncap, int k, int dim, int nid_level,
                                                          const int* __restrict__ children,
                                                          const int* __restrict__ word_id,
                                                          const double* __restrict__ weight,
                                                          const float* __restrict__ ndesc, int* __restrict__ f_word,
                                                          double* __restrict__ f_weight, int* __restrict__ f_node) {
    extern __shared__ float s_feat[];  // [8][dim]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, f = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;
    const int n = min(src.n_of(f), ncap);
    if (i >= n) return;
    const float* a = src.desc_of(f) + (size_t)i * dim;
    float* sa = s_feat + warp * dim;
    for (int t = lane; t < dim; t += 32) sa[t] = a[t];
    __syncwarp();
    int node = 0, level = 0, nid = 0;
    do {
        ++level;
        const int ch = lane < k ? children[(size_t)node * k + lane] : -1;
        double d = 1.7976931348623157e308;
        if (ch >= 0) {
            const float* b = ndesc + (size_t)ch * dim;
            double sqd = 0.;
            for (int t = 0; t < dim; t += 4) {  // dim is a multiple of 4
                const float4 bv = *reinterpret_cast<const float4*>(b + t);
                const float x0 = sa[t] - bv.x, x1 = sa[t + 1] - bv.y, x2 = sa[t + 2] - bv.z, x3 = sa[t + 3] - bv.w;
                sqd += (double)(x0 * x0);
                sqd += (double)(x1 * x1);
                sqd += (double)(x2 * x2);
                sqd += (double)(x3 * x3);
            }
            d = sqd;
        }
        int best = ch >= 0 ? lane : 0x7fffffff;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double od = __shfl_xor_sync(AFULL, d, m);
            const int ob = __shfl_xor_sync(AFULL, best, m);
            if (od < d || (od == d && ob < best)) {  // strict < in child order: the first minimum wins
                d = od;
                best = ob;
            }
        }
        // d < DBL_MAX is required for a child to be taken; a node whose children are all at +max keeps final_id
        if (best != 0x7fffffff && d < 1.7976931348623157e308) node = children[(size_t)node * k + best];
        if (level == nid_level) nid = node;
    } while (children[(size_t)node * k] >= 0);
    if (lane == 0) {
        const double w = weight[node];
        const size_t o = (size_t)f * ncap + i;
        f_word[o] = word_id[node];
        f_weight[o] = w;
        f_node[o] = w > 0 ? nid : -1;
    }
}

__global__ void __launch_bounds__(256) bow_vector_kernel(FrameSrc src, int ncap, int scoring,
                                                         const int* __restrict__ f_word,
                                                         const double* __restrict__ f_weight,
                                                         int* __restrict__ bow_word, double* __restrict__ bow_value,
                                                         int* __restrict__ nb_out) {
    __shared__ unsigned long long key[BOW_NCAP];
    __shared__ int s_start[BOW_NCAP + 1];
    __shared__ int s_nb;
    __shared__ double s_norm;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(src.n_of(f), ncap);
    const int* fw = f_word + (size_t)f * ncap;
    const double* fwt = f_weight + (size_t)f * ncap;
    for (int i = tid; i < BOW_NCAP; i += 256)
        key[i] = (i < n && fwt[i] > 0) ? (((unsigned long long)(unsigned)fw[i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= BOW_NCAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < BOW_NCAP / 2; t += 256) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = key[lo], y = key[hi];
                if ((x > y) == up) {
                    key[lo] = y;
                    key[hi] = x;
                }
            }
            __syncthreads();
        }
    if (tid == 0) {  // run starts, in order (<= 1024 steps)
        int cnt = 0, nbv = 0;
        for (int t = 0; t < BOW_NCAP && key[t] != ~0ull; t++) {
            if (t == 0 || (key[t] >> 32) != (key[t - 1] >> 32)) s_start[nbv++] = t;
            cnt++;
        }
        s_start[nbv] = cnt;
        s_nb = nbv;
    }
    __syncthreads();
    const int nbv = s_nb;
    int* bw = bow_word + (size_t)f * ncap;
    double* bv = bow_value + (size_t)f * ncap;
    for (int r = tid; r < nbv; r += 256) {  // v.addWeight(id, w) in feature order
        double v = 0.;
        for (int t = s_start[r]; t < s_start[r + 1]; t++) {
            const double w = fwt[(int)(key[t] & 0xffffffffu)];
            v = t == s_start[r] ? w : v + w;
        }
        bw[r] = (int)(key[s_start[r]] >> 32);
        bv[r] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double norm = 0.0;
        if (scoring == 5) {
            norm = (double)nbv;
        } else if (scoring == 1) {
            for (int r = 0; r < nbv; r++) norm += bv[r] * bv[r];
            norm = sqrt(norm);
        } else {
            for (int r = 0; r < nbv; r++) norm += fabs(bv[r]);
        }
        s_norm = norm;
        nb_out[2 * f] = nbv;
        nb_out[2 * f + 1] = n;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0)
        for (int r = tid; r < nbv; r += 256) bv[r] = bv[r] / norm;
}

void free_voc(BowState* b) {
    void* bufs[] = {b->children, b->word_id, b->weight, b->desc};
    for (void* p : bufs)
        if (p) cudaFree(p);
    b->children = b->word_id = nullptr;
    b->weight = nullptr;
    b->desc = nullptr;
}

int run_bow(ppg_ctx* c, const FrameSrc& src, int frames, int levelsup) {
    BowState* b = c->bow;
    const int smem = 8 * b->dim * 4;
    bow_descend_kernel<<<dim3(b->ncap / 8, frames), 256, smem, c->st>>>(src, b->ncap, b->k, b->dim, b->L - levelsup,
                                                                        b->children, b->word_id, b->weight, b->desc,
                                                                        b->f_word, b->f_weight, b->f_node);
    stage_mark(c, "bow.descend");
    bow_vector_kernel<<<frames, 256, 0, c->st>>>(src, b->ncap, b->scoring, b->f_word, b->f_weight, b->bow_word,
                                                 b->bow_value, b->nb);
    stage_mark(c, "bow.vector");
    c->launches += 2;
    PPG_CUDA(c, cudaGetLastError());
    return PPG_OK;
}

// results of `frames` slots -> pinned mirror -> caller arrays
int fetch_bow(ppg_ctx* c, int frames, ppg_bow_out* outs) {
    BowState* b = c->bow;
    const size_t N = b->ncap, F = frames;
    uint8_t* h = b->h;
    int* h_nb = reinterpret_cast<int*>(h);
    int* h_word = reinterpret_cast<int*>(h + 4096);
    int* h_node = h_word + b->bcap * N;
    int* h_bword = h_node + b->bcap * N;
    double* h_weight = reinterpret_cast<double*>(h_bword + b->bcap * N);
    double* h_bval = h_weight + b->bcap * N;
    PPG_CUDA(c, cudaMemcpyAsync(h_nb, b->nb, F * 8, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(h_word, b->f_word, F * N * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(h_node, b->f_node, F * N * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(h_bword, b->bow_word, F * N * 4, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(h_weight, b->f_weight, F * N * 8, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(h_bval, b->bow_value, F * N * 8, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    for (int f = 0; f < frames; f++) {
        ppg_bow_out* o = &outs[f];
        o->n_bow = h_nb[2 * f];
        o->n_features = h_nb[2 * f + 1];
        const size_t n = (size_t)o->n_features, nb = (size_t)o->n_bow;
        if (o->word_id) memcpy(o->word_id, h_word + f * N, n * 4);
        if (o->node_id) memcpy(o->node_id, h_node + f * N, n * 4);
        if (o->word_weight) memcpy(o->word_weight, h_weight + f * N, n * 8);
        if (o->bow_word) memcpy(o->bow_word, h_bword + f * N, nb * 4);
        if (o->bow_value) memcpy(o->bow_value, h_bval + f * N, nb * 8);
    }
    return PPG_OK;
}

}  // namespace

void bow_destroy(ppg_ctx* c) {
    BowState* b = c->bow;
    if (!b) return;
    free_voc(b);
    void* bufs[] = {b->fdesc, b->f_word, b->f_weight, b->f_node, b->bow_word, b->bow_value, b->nb};
    for (void* p : bufs)
        if (p) cudaFree(p);
    if (b->h) cudaFreeHost(b->h);
    delete b;
    c->bow = nullptr;
}

}  // namespace ppg

using namespace ppg;

extern "C" {

int ppg_upload_vocabulary(ppg_ctx* c, const ppg_vocabulary* v) {
    if (!c || !v || !v->children || !v->word_id || !v->weight || !v->desc)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: null argument");
    if (v->k < 1 || v->k > 32 || v->L < 1 || v->n_nodes < 2 || v->dim != PPG_DESC_DIM)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: need 1 <= k <= 32, L >= 1, 256-dimensional node descriptors");
    if (v->weighting != 0 && v->weighting != 1)
        return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: only TF_IDF (0) / TF (1) weighting is supported");
    if (v->scoring < 0 || v->scoring > 5) return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: unknown scoring type");
    for (int i = 0; i < v->n_nodes * v->k; i++)
        if (v->children[i] >= v->n_nodes || v->children[i] == 0)
            return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: child id out of range");
    if (v->children[0] < 0) return set_err(c, PPG_ERR_ARG, "ppg_upload_vocabulary: the root has no children");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    if (!c->bow) {
        BowState* b = new BowState();
        c->bow = b;
        b->ncap = BOW_NCAP;
        b->bcap = c->maxB;
        const size_t N = b->ncap, B = b->bcap;
        PPG_CUDA(c, dalloc(&b->fdesc, N * PPG_DESC_DIM));
        PPG_CUDA(c, dalloc(&b->f_word, B * N));
        PPG_CUDA(c, dalloc(&b->f_weight, B * N));
        PPG_CUDA(c, dalloc(&b->f_node, B * N));
        PPG_CUDA(c, dalloc(&b->bow_word, B * N));
        PPG_CUDA(c, dalloc(&b->bow_value, B * N));
        PPG_CUDA(c, dalloc(&b->nb, B * 2));
        b->h_bytes = 4096 + B * N * (3 * 4 + 2 * 8);
        PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&b->h), b->h_bytes));
    }
    BowState* b = c->bow;
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    free_voc(b);
    const size_t n = v->n_nodes;
    PPG_CUDA(c, dalloc(&b->children, n * v->k));
    PPG_CUDA(c, dalloc(&b->word_id, n));
    PPG_CUDA(c, dalloc(&b->weight, n));
    PPG_CUDA(c, dalloc(&b->desc, n * v->dim));
    PPG_CUDA(c, cudaMemcpyAsync(b->children, v->children, n * v->k * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(b->word_id, v->word_id, n * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(b->weight, v->weight, n * 8, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaMemcpyAsync(b->desc, v->desc, n * v->dim * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    b->k = v->k;
    b->L = v->L;
    b->scoring = v->scoring;
    b->weighting = v->weighting;
    b->n_nodes = v->n_nodes;
    b->dim = v->dim;
    return PPG_OK;
}

int ppg_bow_transform(ppg_ctx* c, const float* desc, int n_features, int levelsup, ppg_bow_out* out) {
    if (!c || !out || (n_features > 0 && !desc)) return set_err(c, PPG_ERR_ARG, "ppg_bow_transform: null argument");
    if (!c->bow || !c->bow->children) return set_err(c, PPG_ERR_ARG, "ppg_bow_transform: upload a vocabulary first");
    BowState* b = c->bow;
    if (n_features < 0 || n_features > b->ncap) return set_err(c, PPG_ERR_ARG, "ppg_bow_transform: too many features");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    if (n_features > 0)
        PPG_CUDA(c, cudaMemcpyAsync(b->fdesc, desc, (size_t)n_features * b->dim * 4, cudaMemcpyHostToDevice, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));  // pageable source
    FrameSrc src{};
    src.desc = reinterpret_cast<const uint8_t*>(b->fdesc);
    src.n = nullptr;
    src.n_val = n_features;
    int rc = run_bow(c, src, 1, levelsup);
    if (rc != PPG_OK) return rc;
    return fetch_bow(c, 1, out);
}

int ppg_bow_run_batch(ppg_ctx* c, int n_frames, int levelsup) {
    if (!c || !c->bow || !c->bow->children) return set_err(c, PPG_ERR_ARG, "ppg_bow_run_batch: upload a vocabulary first");
    if (n_frames < 1 || n_frames > c->maxB) return set_err(c, PPG_ERR_ARG, "ppg_bow_run_batch: bad frame count");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    const OutLayout& L = c->post.lay;
    FrameSrc src{};
    src.desc = c->d_out + L.desc;
    src.n = c->d_out + L.hdr + HDR_NKP * sizeof(int);
    src.stride = L.total;
    return run_bow(c, src, n_frames, levelsup);
}

int ppg_bow_fetch_batch(ppg_ctx* c, int n_frames, ppg_bow_out* outs) {
    if (!c || !c->bow || !outs || n_frames < 1 || n_frames > c->bow->bcap)
        return set_err(c, PPG_ERR_ARG, "ppg_bow_fetch_batch: bad arguments");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    return fetch_bow(c, n_frames, outs);
}

}  // extern "C"
