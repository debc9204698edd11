// Row-sharded association (BASELINE.json config 5): the map-descriptor table is split by rows over the
// GPUs, every rank scores its rows against the replicated frame (assoc.cu: tensor-core filter + exact re-score), and ONE
// ncclAllGather of the 20-byte per-row records {best_idx, second_idx, best_dist, second_dist, accept} rebuilds the whole
// answer on every rank.  Rows are independent (the frozen-state search core of Matcher.cpp:224-281), so the gather is a
// concatenation.  Pack kernel and collective are enqueued on the ctx stream right behind the re-score kernel: no host
// synchronisation between compute and exchange.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already holds -- torch's bundled one under
// bench.py / the tests -- or the system's), so that the library loads and runs without it on single-GPU hosts; without
// NCCL the ppg_comm_* calls fail with PPG_ERR_NCCL, there is no fallback.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "assoc.cuh"
#include "ctx.cuh"

namespace ppg {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.h, "ncclCommInitRank"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.h, "ncclAllGather"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.h, "ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.h, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
    });
    return api;
}

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int32_t* send = nullptr;  // [cap][5]
    int32_t* recv = nullptr;  // [world][cap][5]
    int32_t* h_recv = nullptr;
    int cap = 0;              // rows per rank the buffers hold
    int last_rows = 0;        // rows per rank of the last gather
    cudaEvent_t g0 = nullptr, g1 = nullptr;
};

static int nccl_fail(ppg_ctx* c, ncclResult_t r, const char* what) {
    return set_err(c, PPG_ERR_NCCL, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "?"));
}

__global__ void __launch_bounds__(256) pack_records_kernel(const int* __restrict__ best_idx,
                                                           const int* __restrict__ second_idx,
                                                           const float* __restrict__ best_d,
                                                           const float* __restrict__ second_d,
                                                           const uint8_t* __restrict__ accept, int32_t* __restrict__ out,
                                                           int n, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    int32_t r[5] = {-1, -1, 0, 0, 0};  // rows past this rank's shard (uneven split): "no candidate"
    if (i < n) {
        r[0] = best_idx[i];
        r[1] = second_idx[i];
        r[2] = __float_as_int(best_d[i]);
        r[3] = __float_as_int(second_d[i]);
        r[4] = accept[i];
    }
#pragma unroll
    for (int k = 0; k < 5; k++) out[(size_t)i * 5 + k] = r[k];
}

void comm_destroy(ppg_ctx* c) {
    CommState* s = c->comm;
    if (!s) return;
    if (s->comm && nccl().ok) nccl().CommDestroy(s->comm);
    if (s->send) cudaFree(s->send);
    if (s->recv) cudaFree(s->recv);
    if (s->h_recv) cudaFreeHost(s->h_recv);
    if (s->g0) cudaEventDestroy(s->g0);
    if (s->g1) cudaEventDestroy(s->g1);
    delete s;
    c->comm = nullptr;
}

}  // namespace ppg

using namespace ppg;

extern "C" {

int ppg_comm_unique_id(void* id128) {
    if (!id128) return PPG_ERR_ARG;
    if (!nccl().ok) return PPG_ERR_NCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return PPG_ERR_NCCL;
    memcpy(id128, &id, sizeof(id));
    return PPG_OK;
}

int ppg_comm_init(ppg_ctx* c, const void* id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return set_err(c, PPG_ERR_ARG, "ppg_comm_init: bad arguments");
    if (!nccl().ok) return set_err(c, PPG_ERR_NCCL, "libnccl.so.2 could not be loaded");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    comm_destroy(c);
    CommState* s = new CommState();
    c->comm = s;
    s->rank = rank;
    s->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const ncclResult_t r = nccl().CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) {
        s->comm = nullptr;
        return nccl_fail(c, r, "ncclCommInitRank");
    }
    PPG_CUDA(c, cudaEventCreate(&s->g0));
    PPG_CUDA(c, cudaEventCreate(&s->g1));
    return PPG_OK;
}

int ppg_comm_destroy(ppg_ctx* c) {
    if (!c) return PPG_ERR_ARG;
    cudaSetDevice(c->dev);
    if (c->st) cudaStreamSynchronize(c->st);
    comm_destroy(c);
    return PPG_OK;
}

// Enqueues pack + all-gather of the records of the n_local rows scored by the last ppg_assoc_run / _run_frame behind
// them on the ctx stream.  rows_per_rank >= n_local is the (common) send count of every rank.
int ppg_assoc_allgather(ppg_ctx* c, int n_local, int rows_per_rank) {
    if (!c || !c->comm || !c->comm->comm) return set_err(c, PPG_ERR_ARG, "ppg_assoc_allgather: ppg_comm_init first");
    // a rank whose shard is empty (fewer rows than ranks) has nothing staged and still takes part in the collective
    if (n_local < 0 || (n_local > 0 && (!c->assoc || n_local > c->assoc->staged_rows)) || rows_per_rank < n_local ||
        rows_per_rank < 1)
        return set_err(c, PPG_ERR_ARG, "ppg_assoc_allgather: bad row counts");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    CommState* s = c->comm;
    AssocState* a = c->assoc;
    if (rows_per_rank > s->cap) {
        PPG_CUDA(c, cudaStreamSynchronize(c->st));
        if (s->send) cudaFree(s->send);
        if (s->recv) cudaFree(s->recv);
        if (s->h_recv) cudaFreeHost(s->h_recv);
        s->send = s->recv = s->h_recv = nullptr;
        PPG_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&s->send), (size_t)rows_per_rank * 20));
        PPG_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&s->recv), (size_t)s->world * rows_per_rank * 20));
        PPG_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&s->h_recv), (size_t)s->world * rows_per_rank * 20));
        s->cap = rows_per_rank;
    }
    pack_records_kernel<<<(rows_per_rank + 255) / 256, 256, 0, c->st>>>(
        a ? a->best_idx : nullptr, a ? a->second_idx : nullptr, a ? a->best_d : nullptr, a ? a->second_d : nullptr,
        a ? a->accept : nullptr, s->send, n_local, rows_per_rank);
    c->launches++;
    PPG_CUDA(c, cudaGetLastError());
    stage_mark(c, "sharded.pack");
    PPG_CUDA(c, cudaEventRecord(s->g0, c->st));
    const ncclResult_t r = nccl().AllGather(s->send, s->recv, (size_t)rows_per_rank * 5, ncclInt32, s->comm, c->st);
    if (r != ncclSuccess) return nccl_fail(c, r, "ncclAllGather");
    PPG_CUDA(c, cudaEventRecord(s->g1, c->st));
    stage_mark(c, "sharded.allgather");
    s->last_rows = rows_per_rank;
    return PPG_OK;
}

// Waits for the stream and copies the gathered records ([world][rows_per_rank][5] int32) to `records` (may be NULL);
// gather_us: device time of the last all-gather (CUDA events around it on the ctx stream).
int ppg_assoc_allgather_fetch(ppg_ctx* c, int32_t* records, float* gather_us) {
    if (!c || !c->comm || c->comm->last_rows < 1) return set_err(c, PPG_ERR_ARG, "ppg_assoc_allgather_fetch: nothing gathered");
    PPG_CUDA(c, cudaSetDevice(c->dev));
    CommState* s = c->comm;
    const size_t bytes = (size_t)s->world * s->last_rows * 20;
    if (records) PPG_CUDA(c, cudaMemcpyAsync(s->h_recv, s->recv, bytes, cudaMemcpyDeviceToHost, c->st));
    PPG_CUDA(c, cudaStreamSynchronize(c->st));
    if (records) memcpy(records, s->h_recv, bytes);
    if (gather_us) {
        float ms = 0.f;
        PPG_CUDA(c, cudaEventElapsedTime(&ms, s->g0, s->g1));
        *gather_us = ms * 1e3f;
    }
    return PPG_OK;
}

}  // extern "C"
