// Multi-GPU exchange of the sharded query path (the reference is single-device, so this has no
// counterpart there): frames / cached embeddings are row-sharded over ranks, every rank runs K4 on its shard with
// GLOBAL row indices, the per-rank candidate lists meet in ONE in-place ncclAllGather of a packed message on the
// caller's stream, and every rank runs the same deterministic k-way merge (descending score, ties -> higher global
// index: identical to the single-GPU order).  No host synchronisation, no staging copies, no framework kernels.
//
// Message of one rank (B200CLIP_TOPK_MSG_BYTES(q, k) bytes): int64 idx[q][k] (global, -1 = empty) | float score[q][k]
// | padding to a multiple of 16 bytes.  K4's final kernel writes its results straight into this rank's slot of the
// gather buffer.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded in the process -- the one the caller's
// communicator came from, e.g. PyTorch's -- else the system one): libb200clip.so has no link-time NCCL dependency and
// single-GPU users never touch it.
#include <dlfcn.h>

#include <mutex>

#include "internal.h"

namespace {

typedef int (*PFN_ncclAllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, void* /*ncclComm_t*/, cudaStream_t);
typedef const char* (*PFN_ncclGetErrorString)(int);
constexpr int kNcclInt8 = 0;   // ncclInt8 / ncclChar

struct NcclApi {
    PFN_ncclAllGather all_gather = nullptr;
    PFN_ncclGetErrorString err = nullptr;
    const char* why = "not loaded";
};

const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the instance that is already mapped (same SONAME) owns the caller's communicator; load one only if none is
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { api.why = "libnccl.so.2 not found"; return; }
        api.all_gather = reinterpret_cast<PFN_ncclAllGather>(dlsym(lib, "ncclAllGather"));
        api.err = reinterpret_cast<PFN_ncclGetErrorString>(dlsym(lib, "ncclGetErrorString"));
        if (!api.all_gather) api.why = "ncclAllGather missing from libnccl.so.2";
    });
    return api;
}

size_t msg_bytes(int q, int k) { return (static_cast<size_t>(q) * k * 12 + 15) & ~size_t(15); }

int ensure_gather(b200clip_handle* h, size_t need, cudaStream_t st) {
    if (need <= h->ws_gather_bytes) return 0;
    B200_CUDA(h, cudaStreamSynchronize(st));
    cudaFree(h->ws_gather);
    h->ws_gather = nullptr; h->ws_gather_bytes = 0;
    B200_CUDA(h, cudaMalloc(&h->ws_gather, need));
    h->ws_gather_bytes = need;
    return 0;
}

int check_world(b200clip_handle* h, void* comm, int rank, int world, int q, int k, const char* what) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "%s: null handle", what);
    if (!comm || world <= 0 || rank < 0 || rank >= world || q <= 0 || k <= 0)
        return b200_fail(h, B200CLIP_E_ARG, "%s: bad communicator / rank %d of %d / q %d / k %d", what, rank, world, q, k);
    if (!nccl_api().all_gather) return b200_fail(h, B200CLIP_E_NCCL, "%s: NCCL unavailable (%s)", what, nccl_api().why);
    return 0;
}

// slot `rank` of the gather buffer holds this rank's message: exchange in place, then merge on every rank
int gather_and_merge(b200clip_handle* h, void* comm, int world, int q, int k, float thr, const double* ts, double clip_dur,
                     double vid_dur, float* top_scores, int64_t* top_idx, double* intervals, int32_t* counts, int rank,
                     cudaStream_t st) {
    const size_t mb = msg_bytes(q, k);
    const NcclApi& api = nccl_api();
    {
        ProfScope ps(h, PROF_COMM, static_cast<double>(world) * mb, st);
        const int r = api.all_gather(h->ws_gather + static_cast<size_t>(rank) * mb, h->ws_gather, mb, kNcclInt8, comm, st);
        if (r != 0) return b200_fail(h, B200CLIP_E_NCCL, "ncclAllGather failed: %s", api.err ? api.err(r) : "?");
    }
    const int64_t* ci = reinterpret_cast<const int64_t*>(h->ws_gather);
    const float* cs = reinterpret_cast<const float*>(h->ws_gather + static_cast<size_t>(q) * k * 8);
    return launch_topk_merge_strided(h, cs, ci, static_cast<int64_t>(mb / 4), static_cast<int64_t>(mb / 8), world, q, k, thr, ts,
                                     clip_dur, vid_dur, top_scores, top_idx, intervals, counts, st);
}

}  // namespace

extern "C" int64_t b200clip_topk_msg_bytes(int q, int k) { return q > 0 && k > 0 ? static_cast<int64_t>(msg_bytes(q, k)) : 0; }

extern "C" int b200clip_topk_merge_packed(b200clip_handle* h, const void* gathered_dev, int g, int q, int k, float threshold,
                                          const double* timestamps_dev, double clip_duration, double video_duration,
                                          float* top_scores_dev, int64_t* top_idx_dev, double* intervals_dev,
                                          int32_t* counts_dev, void* stream) {
    if (!h || !gathered_dev || q <= 0 || k <= 0) return b200_fail(h, B200CLIP_E_ARG, "topk_merge_packed: bad argument");
    if (reinterpret_cast<uintptr_t>(gathered_dev) & 7) return b200_fail(h, B200CLIP_E_ARG, "topk_merge_packed: buffer must be 8-byte aligned");
    B200_CUDA(h, cudaSetDevice(h->device));
    const size_t mb = msg_bytes(q, k);
    const uint8_t* base = static_cast<const uint8_t*>(gathered_dev);
    return launch_topk_merge_strided(h, reinterpret_cast<const float*>(base + static_cast<size_t>(q) * k * 8),
                                     reinterpret_cast<const int64_t*>(base), static_cast<int64_t>(mb / 4),
                                     static_cast<int64_t>(mb / 8), g, q, k, threshold, timestamps_dev, clip_duration,
                                     video_duration, top_scores_dev, top_idx_dev, intervals_dev, counts_dev,
                                     static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_topk_merge_nccl(b200clip_handle* h, void* nccl_comm, int rank, int world,
                                        const float* local_scores_dev, const int64_t* local_idx_dev, int q, int k,
                                        float threshold, const double* timestamps_dev, double clip_duration,
                                        double video_duration, float* top_scores_dev, int64_t* top_idx_dev,
                                        double* intervals_dev, int32_t* counts_dev, void* stream) {
    int rc = check_world(h, nccl_comm, rank, world, q, k, "topk_merge_nccl");
    if (rc) return rc;
    if (!local_scores_dev || !local_idx_dev || !top_scores_dev || !top_idx_dev)
        return b200_fail(h, B200CLIP_E_ARG, "topk_merge_nccl: null buffer");
    B200_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t mb = msg_bytes(q, k), qk = static_cast<size_t>(q) * k;
    if ((rc = ensure_gather(h, mb * world, st))) return rc;
    uint8_t* slot = h->ws_gather + static_cast<size_t>(rank) * mb;
    B200_CUDA(h, cudaMemcpyAsync(slot, local_idx_dev, qk * 8, cudaMemcpyDeviceToDevice, st));
    B200_CUDA(h, cudaMemcpyAsync(slot + qk * 8, local_scores_dev, qk * 4, cudaMemcpyDeviceToDevice, st));
    return gather_and_merge(h, nccl_comm, world, q, k, threshold, timestamps_dev, clip_duration, video_duration,
                            top_scores_dev, top_idx_dev, intervals_dev, counts_dev, rank, st);
}

extern "C" int b200clip_sim_topk_nccl(b200clip_handle* h, void* nccl_comm, int rank, int world, const void* img_emb_dev,
                                      int emb_dtype, int64_t n_local, int e, const float* txt_emb_dev, int q, int k,
                                      float threshold, const double* timestamps_dev, int64_t index_base,
                                      double clip_duration, double video_duration, float* top_scores_dev,
                                      int64_t* top_idx_dev, double* intervals_dev, int32_t* counts_dev, void* stream) {
    int rc = check_world(h, nccl_comm, rank, world, q, k, "sim_topk_nccl");
    if (rc) return rc;
    if (!top_scores_dev || !top_idx_dev) return b200_fail(h, B200CLIP_E_ARG, "sim_topk_nccl: null output");
    B200_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t mb = msg_bytes(q, k), qk = static_cast<size_t>(q) * k;
    if ((rc = ensure_gather(h, mb * world, st))) return rc;
    uint8_t* slot = h->ws_gather + static_cast<size_t>(rank) * mb;
    // local K4 (threshold and intervals are applied after the merge) writes straight into this rank's message slot
    if ((rc = launch_sim_topk(h, img_emb_dev, emb_dtype, n_local, e, txt_emb_dev, q, k, -3.0e38f, nullptr, index_base, 0.0, 0.0,
                              reinterpret_cast<float*>(slot + qk * 8), reinterpret_cast<int64_t*>(slot), nullptr, nullptr, nullptr,
                              st)))
        return rc;
    return gather_and_merge(h, nccl_comm, world, q, k, threshold, timestamps_dev, clip_duration, video_duration,
                            top_scores_dev, top_idx_dev, intervals_dev, counts_dev, rank, st);
}
