// Internal declarations shared by the translation units of libb200clip.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/b200clip.h"

typedef __nv_bfloat16 bf16;

struct b200clip_handle {
    b200clip_config cfg;
    int device = 0;
    int num_sms = 148;
    bool finalized = false;
    mutable std::string err;
    int64_t launches = 0;

    // ---- opt-in per-class kernel timing (CUDA events on the launching stream; bench.py's roofline numbers)
    struct ProfRec { cudaEvent_t a, b; int cls; double work; };
    bool prof_on = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;

    // ---- packed weights (device) ----
    struct Block {
        // ln_1 / ln_2 are folded into the QKV / fc GEMMs at finalize: w_qkv = W.diag(gamma1) (bf16),
        // c1 = row sums of the folded weights, c2 = W.beta + bias (see GemmEpilogue)
        bf16 *w_qkv = nullptr, *w_out = nullptr, *w_fc = nullptr, *w_proj = nullptr;
        float *c1_qkv = nullptr, *c2_qkv = nullptr, *c1_fc = nullptr, *c2_fc = nullptr;
        float *b_out = nullptr, *b_proj = nullptr;
        // host copies kept between set_weight and finalize
        std::vector<float> h_ln1_g, h_ln1_b, h_ln2_g, h_ln2_b, h_w_qkv, h_b_qkv, h_w_fc, h_b_fc;
    };
    struct Tower {
        int width = 0, layers = 0, heads = 0, mlp = 0;
        std::vector<Block> blocks;
    };
    Tower vis, txt;
    // vision stem / head
    bf16* w_patch = nullptr;      // [width, patch_k] (conv1.weight flattened, K padded to 64)
    float* pos_emb = nullptr;     // [tokens, width] fp32
    float* cls_pos0 = nullptr;    // [width] = class_embedding + positional_embedding[0]
    float *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr, *ln_post_b = nullptr;
    bf16* vis_proj = nullptr;     // [width, embed] row-major (x @ proj)
    // text stem / head
    float* tok_emb = nullptr;     // [vocab, text_width] fp32 (one rounding after + positional)
    float* txt_pos = nullptr;     // [ctx, text_width]
    float *ln_final_g = nullptr, *ln_final_b = nullptr;
    bf16* txt_proj = nullptr;     // [text_width, embed]
    std::map<std::string, bool> have;
    std::vector<float> host_cls, host_pos0;   // kept until finalize builds cls_pos0
    std::vector<void*> allocs;    // everything cudaMalloc'ed for weights

    // ---- derived geometry ----
    int grid = 0;       // patches per side
    int tokens = 0;     // grid^2 + 1
    int patch_k = 0;    // 3*P*P padded to a multiple of 64

    // ---- persistent workspace ----
    int ws_images = 0, ws_texts = 0;
    bf16 *ws_x = nullptr, *ws_y = nullptr, *ws_qkv = nullptr, *ws_h = nullptr, *ws_patches = nullptr;
    // the text tower has its own (small) activation buffers, so that a text call on one stream may overlap an image
    // call on another stream of the same handle
    bf16 *ws_tx = nullptr, *ws_ty = nullptr, *ws_tqkv = nullptr, *ws_th = nullptr;
    float* ws_tstats = nullptr;
    uint8_t* ws_stage_dev[2] = {nullptr, nullptr};   // device staging for host-frame calls
    uint8_t* ws_stage_host[2] = {nullptr, nullptr};  // pinned
    size_t ws_stage_bytes = 0, ws_stage_host_bytes = 0;
    uint64_t stage_seq = 0;                          // chunks uploaded so far: parity picks the staging buffer
    float* ws_stats = nullptr;     // [rows, LN_SLOTS, 2] per-row (sum, sum of squares) partials of the residual stream
    int32_t* ws_eot = nullptr;     // [ws_texts] row of the EOT token per text
    int64_t* ws_tokens = nullptr;  // [ws_texts * ctx]
    float* ws_emb = nullptr;       // device scratch for host-output calls
    size_t ws_emb_elems = 0;
    void* pre_plans = nullptr;     // preprocess.cu: cached resize tables per frame geometry
    float* pre_lut = nullptr;      // [3][256] ToTensor+Normalize lookup
    uint8_t* ws_pre = nullptr;     // preprocess intermediates
    size_t ws_pre_bytes = 0;
    uint8_t* ws_gather = nullptr;  // [world] candidate messages of the NCCL exchange (b200clip_*_nccl)
    size_t ws_gather_bytes = 0;
    uint8_t* ws_nv12 = nullptr;    // RGB scratch of the generic (unfused) NV12 path
    size_t ws_nv12_bytes = 0;
    void* ws_topk = nullptr;       // sim/top-k partial candidates
    size_t ws_topk_bytes = 0;
    bf16* ws_patches2 = nullptr;   // second patch buffer: K1 of chunk i+1 overlaps the tower of chunk i
    cudaStream_t pre_stream = nullptr;
    cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_tower[2] = {nullptr, nullptr}, ev_fork = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t gemm_side = nullptr;            // hybrid GEMM: plain CTA pairs on the SMs the 8-CTA clusters leave idle
    cudaEvent_t ev_gemm_fork = nullptr, ev_gemm_join = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    int64_t h2d_bytes = 0;         // bytes uploaded by the host-buffer entry points (b200clip_transfer_bytes)
    int64_t d2h_bytes = 0;
    int chunk_cap = 0;             // images per pass of the tower fixed by an explicit b200clip_reserve (0 = default)
    // cudaFuncSetAttribute is per DEVICE: which kernels of this handle's device already carry their opt-in (ATTR_*)
    uint32_t attr_done = 0;
    int k1_ctas_per_sm[2] = {0, 0};          // resident CTAs per SM of the persistent K1 area kernels (RGB, NV12)
    int g2_clusters[5] = {0, 0, 0, 0, 0};   // co-resident clusters of the 2-CTA GEMM on this device, per pairs
};
enum { ATTR_GEMM64 = 1u << 0, ATTR_GEMM128 = 1u << 1, ATTR_GEMM256 = 1u << 2, ATTR_GEMM_2CTA = 1u << 3, ATTR_SIM_TC = 1u << 4,
       ATTR_ATTN_PERSIST = 1u << 5, ATTR_SIM_STREAM = 1u << 6, ATTR_K1_VFIRST = 1u << 7, ATTR_K1_NV12 = 1u << 8, ATTR_K1_MMA = 1u << 9, ATTR_ATTN_TC2 = 1u << 10, ATTR_ATTN_TC64 = 1u << 11, ATTR_HEAD_MMA = 1u << 12 };

// Environment switches, read ONCE per process (first use).  They select between code paths that are all valid and
// parity-tested (fallback kernels that other geometries use anyway); measurement probes that invalidate results or
// launch forms that lost their A/B exist only in builds with -DB200CLIP_PROBES (make PROBES=1).
struct B200Knobs {
    bool k1_unfused, area_fp32, area_px1, area_hfirst, area_nostrip, vpass_generic;   // K1 fallbacks
    bool gemm_1cta, gemm_spin_wait;
    bool hpass_px1;            // B200CLIP_HPASS_PX1: generic horizontal pass with one output pixel per thread
    bool gemm_5stage;          // B200CLIP_GEMM_5STAGE: five-stage / two-box 2-CTA GEMM for every shape
    bool sim_simt, sim_stream_a;
    bool attn_oneshot, attn_tc, attn_tiled, attn_tc2, attn_tc64;
    bool head_simt;            // B200CLIP_HEAD_SIMT: CUDA-core head kernel instead of the mma.sync one
    bool overlap, full_upload;
    bool nv12_unfused, k1_persistent;
    bool k1_verbose;           // B200CLIP_K1_VERBOSE: report the K1 code path on stderr
    bool area_mma;             // K1 stages A + B on the integer tensor cores (preprocess_mma.cuh)
};
const B200Knobs& b200_knobs();

// kernel classes for the profiler
enum { PROF_GEMM = 0, PROF_ATTN = 1, PROF_LN = 2, PROF_PRE = 3, PROF_HEAD = 4, PROF_SIM = 5, PROF_MISC = 6, PROF_PRE_A = 7, PROF_PRE_B = 8, PROF_PRE_C = 9, PROF_COMM = 10, PROF_GEMM_SMALL = 11, PROF_NCLS = 12 };
// RAII bracket: records an event pair around the launches made in its scope when profiling is on
struct ProfScope {
    b200clip_handle* h; cudaStream_t st; int idx;
    ProfScope(b200clip_handle* h_, int cls, double work, cudaStream_t st_);
    ~ProfScope();
};

// error helpers (api.cu)
int b200_fail(const b200clip_handle* h, int code, const char* fmt, ...);
#define B200_CUDA(h, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return b200_fail(h, B200CLIP_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

// ---- kernel launchers (return cudaError_t-free int codes via handle) ----
namespace b200 {
struct GemmEpilogue;
}
int launch_gemm(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc, int M, int N,
                int K, const b200::GemmEpilogue& ep, cudaStream_t st);
int launch_layernorm(b200clip_handle* h, const bf16* x, const float* g, const float* b, bf16* y, int64_t rows,
                     int width, float eps, int t_per_img, const float* cls_row, float* stats_out, cudaStream_t st);
int launch_attention(b200clip_handle* h, const bf16* qkv, bf16* out, int n_seq, int t, int heads, int causal,
                     cudaStream_t st);
int launch_head(b200clip_handle* h, const bf16* x, int64_t row_stride, const int32_t* row_index, const float* g,
                const float* b, const bf16* proj, int n, int width, int embed, float eps, void* out, int out_dtype,
                int l2norm, cudaStream_t st);
int launch_patchify_chw(b200clip_handle* h, const float* chw, int n, bf16* patches, cudaStream_t st);
int launch_text_embed(b200clip_handle* h, const int64_t* tokens, int q, bf16* x, int32_t* eot_rows, float* stats_out,
                      cudaStream_t st);
int launch_preprocess(b200clip_handle* h, const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                      int64_t row_stride, int mode, bf16* patches, float* chw, cudaStream_t st);
int launch_sim_topk(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q, int k,
                    float thr, const double* ts, int64_t index_base, double clip_dur, double vid_dur,
                    float* top_scores, int64_t* top_idx, double* intervals, int32_t* counts, float* dense,
                    cudaStream_t st);
int launch_similarity(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q,
                      float* scores, cudaStream_t st);
int launch_topk_merge(b200clip_handle* h, const float* cs, const int64_t* ci, int g, int q, int k, float thr,
                      const double* ts, double clip_dur, double vid_dur, float* top_scores, int64_t* top_idx,
                      double* intervals, int32_t* counts, cudaStream_t st);
int launch_preprocess_nv12(b200clip_handle* h, const uint8_t* y, const uint8_t* uv, int n, int H, int W, int64_t y_fs,
                           int64_t uv_fs, int64_t rs, int mode, bf16* patches, float* chw, cudaStream_t st);
int preprocess_source_window_nv12(b200clip_handle* h, int H, int W, int mode, int* x0, int* x1, int* y0, int* y1);
int launch_topk_merge_strided(b200clip_handle* h, const float* cs, const int64_t* ci, int64_t ls_s, int64_t ls_i, int g, int q,
                              int k, float thr, const double* ts, double clip_dur, double vid_dur, float* top_scores,
                              int64_t* top_idx, double* intervals, int32_t* counts, cudaStream_t st);
void preprocess_free_plans(b200clip_handle* h);
int preprocess_source_window(b200clip_handle* h, int H, int W, int mode, int* x0, int* x1, int* y0, int* y1);
