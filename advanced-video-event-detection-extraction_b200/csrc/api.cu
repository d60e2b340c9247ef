// C ABI of libb200clip.so: handle life cycle, weight packing, workspace, tower orchestration.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "gemm_tcgen05.cuh"
#include "internal.h"

static std::string g_err;
static std::mutex g_err_mu;

int b200_fail(const b200clip_handle* h, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) {
        h->err = buf;
    } else {
        std::lock_guard<std::mutex> lk(g_err_mu);
        g_err = buf;
    }
    return code;
}

const B200Knobs& b200_knobs() {
    static const B200Knobs k = [] {
        auto on = [](const char* name) { const char* v = getenv(name); return v != nullptr && v[0] != '\0' && v[0] != '0'; };
        B200Knobs b{};
        b.k1_unfused = on("B200CLIP_K1_UNFUSED"); b.area_fp32 = on("B200CLIP_AREA_FP32"); b.area_px1 = on("B200CLIP_AREA_PX1");
        b.area_hfirst = on("B200CLIP_AREA_HFIRST"); b.area_nostrip = on("B200CLIP_AREA_NOSTRIP");
        b.vpass_generic = on("B200CLIP_VPASS_GENERIC");
        b.gemm_1cta = on("B200CLIP_GEMM_1CTA"); b.gemm_spin_wait = on("B200CLIP_GEMM_SPIN_WAIT"); b.gemm_5stage = on("B200CLIP_GEMM_5STAGE"); b.hpass_px1 = on("B200CLIP_HPASS_PX1");
        b.sim_simt = on("B200CLIP_SIM_SIMT"); b.sim_stream_a = on("B200CLIP_SIM_STREAM_A");
        b.attn_oneshot = on("B200CLIP_ATTN_ONESHOT"); b.attn_tc = on("B200CLIP_ATTN_TC"); b.attn_tiled = on("B200CLIP_ATTN_TILED"); b.attn_tc2 = !on("B200CLIP_ATTN_NOTC2"); b.attn_tc64 = !on("B200CLIP_ATTN_NOTC64"); b.head_simt = on("B200CLIP_HEAD_SIMT");
        b.overlap = on("B200CLIP_OVERLAP"); b.full_upload = on("B200CLIP_FULL_UPLOAD");
        b.nv12_unfused = on("B200CLIP_NV12_UNFUSED"); b.k1_persistent = on("B200CLIP_K1_PERSISTENT");
        b.area_mma = !on("B200CLIP_AREA_NOMMA"); b.k1_verbose = on("B200CLIP_K1_VERBOSE");
        return b;
    }();
    return k;
}

extern "C" const char* b200clip_version(void) { return "b200clip 0.1 (sm_100a)"; }

extern "C" const char* b200clip_last_error(const b200clip_handle* h) {
    if (h) return h->err.c_str();
    return g_err.c_str();
}

extern "C" int64_t b200clip_launch_count(const b200clip_handle* h) { return h ? h->launches : 0; }
extern "C" void b200clip_reset_launch_count(b200clip_handle* h) {
    if (h) h->launches = 0;
}
extern "C" int b200clip_transfer_bytes(b200clip_handle* h, int64_t* h2d_out, int64_t* d2h_out, int reset) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "transfer_bytes: null handle");
    if (h2d_out) *h2d_out = h->h2d_bytes;
    if (d2h_out) *d2h_out = h->d2h_bytes;
    if (reset) { h->h2d_bytes = 0; h->d2h_bytes = 0; }
    return 0;
}

// ------------------------------------------------------------------------------------------- profiler
static cudaEvent_t prof_get_event(b200clip_handle* h) {
    if (!h->prof_pool.empty()) {
        cudaEvent_t e = h->prof_pool.back();
        h->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
ProfScope::ProfScope(b200clip_handle* h_, int cls, double work, cudaStream_t st_) : h(h_), st(st_), idx(-1) {
    if (!h || !h->prof_on) return;
    b200clip_handle::ProfRec r{prof_get_event(h), prof_get_event(h), cls, work};
    if (!r.a || !r.b) return;
    cudaEventRecord(r.a, st);
    h->prof_recs.push_back(r);
    idx = static_cast<int>(h->prof_recs.size()) - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof_recs[idx].b, st);
}

extern "C" int b200clip_profile_enable(b200clip_handle* h, int on) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "profile_enable: null handle");
    h->prof_on = on != 0;
    return 0;
}

extern "C" int b200clip_profile_read(b200clip_handle* h, int cls, double* ms_out, double* work_out,
                                     int64_t* launches_out, int reset) {
    if (!h || cls < 0 || cls >= PROF_NCLS) return b200_fail(h, B200CLIP_E_ARG, "profile_read: bad argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    B200_CUDA(h, cudaDeviceSynchronize());
    double ms = 0, work = 0;
    int64_t n = 0;
    for (const auto& r : h->prof_recs) {
        if (r.cls != cls) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; work += r.work; ++n; }
    }
    if (ms_out) *ms_out = ms;
    if (work_out) *work_out = work;
    if (launches_out) *launches_out = n;
    if (reset) {
        for (const auto& r : h->prof_recs) { h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b); }
        h->prof_recs.clear();
    }
    return 0;
}

static int check_cfg(const b200clip_config* c) {
    if (c->image_size <= 0 || c->patch <= 0 || c->image_size % c->patch != 0)
        return b200_fail(nullptr, B200CLIP_E_SHAPE, "image_size %d must be a positive multiple of patch %d",
                         c->image_size, c->patch);
    if (c->width <= 0 || c->width % 64 != 0 || c->heads * 64 != c->width)
        return b200_fail(nullptr, B200CLIP_E_SHAPE, "vision width %d must equal heads %d * 64", c->width, c->heads);
    if (c->text_width <= 0 || c->text_heads * 64 != c->text_width)
        return b200_fail(nullptr, B200CLIP_E_SHAPE, "text width %d must equal text heads %d * 64", c->text_width,
                         c->text_heads);
    if (c->mlp_dim % 64 != 0 || c->text_mlp_dim % 64 != 0 || c->embed_dim % 32 != 0 || c->embed_dim > 1024)
        return b200_fail(nullptr, B200CLIP_E_SHAPE, "mlp dims must be multiples of 64, embed_dim of 32 (<= 1024)");
    if (c->layers <= 0 || c->text_layers <= 0 || c->text_ctx <= 0 || c->text_ctx > 128 || c->text_vocab <= 0)
        return b200_fail(nullptr, B200CLIP_E_SHAPE, "bad layer / context / vocab sizes");
    if (c->act != 0 && c->act != 1) return b200_fail(nullptr, B200CLIP_E_ARG, "act must be 0 or 1");
    return 0;
}

extern "C" int b200clip_create(const b200clip_config* cfg, int device, b200clip_handle** out) {
    if (!cfg || !out) return b200_fail(nullptr, B200CLIP_E_ARG, "create: null argument");
    *out = nullptr;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return b200_fail(nullptr, B200CLIP_E_CUDA, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= ndev) return b200_fail(nullptr, B200CLIP_E_ARG, "device %d out of range", device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return b200_fail(nullptr, B200CLIP_E_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return b200_fail(nullptr, B200CLIP_E_ARCH, "device %d is sm_%d%d; libb200clip is built for sm_100a only", device,
                         prop.major, prop.minor);
    if (cudaSetDevice(device) != cudaSuccess) return b200_fail(nullptr, B200CLIP_E_CUDA, "cudaSetDevice failed");
    b200clip_handle* h = new b200clip_handle();
    h->cfg = *cfg;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->grid = cfg->image_size / cfg->patch;
    h->tokens = h->grid * h->grid + 1;
    h->patch_k = ((3 * cfg->patch * cfg->patch + 63) / 64) * 64;
    h->vis.width = cfg->width; h->vis.layers = cfg->layers; h->vis.heads = cfg->heads; h->vis.mlp = cfg->mlp_dim;
    h->txt.width = cfg->text_width; h->txt.layers = cfg->text_layers; h->txt.heads = cfg->text_heads;
    h->txt.mlp = cfg->text_mlp_dim;
    h->vis.blocks.resize(cfg->layers);
    h->txt.blocks.resize(cfg->text_layers);
    *out = h;
    return 0;
}

extern "C" int b200clip_destroy(b200clip_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    preprocess_free_plans(h);
    for (const auto& r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
    for (void* p : h->allocs) cudaFree(p);
    cudaFree(h->ws_x); cudaFree(h->ws_y); cudaFree(h->ws_qkv); cudaFree(h->ws_h); cudaFree(h->ws_patches);
    cudaFree(h->ws_emb); cudaFree(h->ws_pre); cudaFree(h->ws_nv12); cudaFree(h->ws_gather); cudaFree(h->ws_topk); cudaFree(h->ws_eot); cudaFree(h->ws_tokens);
    cudaFree(h->ws_patches2); cudaFree(h->ws_stats);
    cudaFree(h->ws_tx); cudaFree(h->ws_ty); cudaFree(h->ws_tqkv); cudaFree(h->ws_th); cudaFree(h->ws_tstats);
    if (h->pre_stream) cudaStreamDestroy(h->pre_stream);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_pre[i]) cudaEventDestroy(h->ev_pre[i]);
        if (h->ev_tower[i]) cudaEventDestroy(h->ev_tower[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->ws_stage_dev[i]);
        if (h->ws_stage_host[i]) cudaFreeHost(h->ws_stage_host[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->gemm_side) cudaStreamDestroy(h->gemm_side);
    if (h->ev_gemm_fork) cudaEventDestroy(h->ev_gemm_fork);
    if (h->ev_gemm_join) cudaEventDestroy(h->ev_gemm_join);
    delete h;
    return 0;
}

// ------------------------------------------------------------------------------------------- weights
static int upload_f32(b200clip_handle* h, const float* src, size_t count, float** dst) {
    void* d = nullptr;
    B200_CUDA(h, cudaMalloc(&d, count * sizeof(float)));
    h->allocs.push_back(d);
    B200_CUDA(h, cudaMemcpy(d, src, count * sizeof(float), cudaMemcpyHostToDevice));
    *dst = static_cast<float*>(d);
    return 0;
}

// fp32 [rows, cols] -> bf16 [rows, ld] (ld >= cols, zero padded), round to nearest even
static int upload_bf16(b200clip_handle* h, const float* src, size_t rows, size_t cols, size_t ld, bf16** dst) {
    std::vector<bf16> tmp(rows * ld, __float2bfloat16(0.f));
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) tmp[r * ld + c] = __float2bfloat16(src[r * cols + c]);
    void* d = nullptr;
    B200_CUDA(h, cudaMalloc(&d, tmp.size() * sizeof(bf16)));
    h->allocs.push_back(d);
    B200_CUDA(h, cudaMemcpy(d, tmp.data(), tmp.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    *dst = static_cast<bf16*>(d);
    return 0;
}

static bool shape_is(const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
    if (ndim != static_cast<int>(want.size())) return false;
    int i = 0;
    for (int64_t w : want)
        if (shape[i++] != w) return false;
    return true;
}

#define WANT_SHAPE(...)                                                                        \
    if (!shape_is(shape, ndim, {__VA_ARGS__}))                                                 \
        return b200_fail(h, B200CLIP_E_SHAPE, "set_weight(%s): unexpected shape for this config", name)

static int set_block_weight(b200clip_handle* h, b200clip_handle::Tower& tw, const char* name, const char* rest,
                            const float* data, const int64_t* shape, int ndim) {
    // rest = "<i>.<param>"
    char* end = nullptr;
    const long li = strtol(rest, &end, 10);
    if (end == rest || *end != '.' || li < 0 || li >= tw.layers)
        return b200_fail(h, B200CLIP_E_ARG, "set_weight(%s): bad block index", name);
    b200clip_handle::Block& b = tw.blocks[li];
    const std::string p(end + 1);
    const int64_t W = tw.width, F = tw.mlp;
    // LayerNorm parameters and the matrices they fold into are kept on the host until finalize
    if (p == "ln_1.weight") { WANT_SHAPE(W); b.h_ln1_g.assign(data, data + W); return 0; }
    if (p == "ln_1.bias") { WANT_SHAPE(W); b.h_ln1_b.assign(data, data + W); return 0; }
    if (p == "ln_2.weight") { WANT_SHAPE(W); b.h_ln2_g.assign(data, data + W); return 0; }
    if (p == "ln_2.bias") { WANT_SHAPE(W); b.h_ln2_b.assign(data, data + W); return 0; }
    if (p == "attn.in_proj_weight") { WANT_SHAPE(3 * W, W); b.h_w_qkv.assign(data, data + 3 * W * W); return 0; }
    if (p == "attn.in_proj_bias") { WANT_SHAPE(3 * W); b.h_b_qkv.assign(data, data + 3 * W); return 0; }
    if (p == "attn.out_proj.weight") { WANT_SHAPE(W, W); return upload_bf16(h, data, W, W, W, &b.w_out); }
    if (p == "attn.out_proj.bias") { WANT_SHAPE(W); return upload_f32(h, data, W, &b.b_out); }
    if (p == "mlp.c_fc.weight") { WANT_SHAPE(F, W); b.h_w_fc.assign(data, data + F * W); return 0; }
    if (p == "mlp.c_fc.bias") { WANT_SHAPE(F); b.h_b_fc.assign(data, data + F); return 0; }
    if (p == "mlp.c_proj.weight") { WANT_SHAPE(W, F); return upload_bf16(h, data, W, F, F, &b.w_proj); }
    if (p == "mlp.c_proj.bias") { WANT_SHAPE(W); return upload_f32(h, data, W, &b.b_proj); }
    return b200_fail(h, B200CLIP_E_ARG, "set_weight(%s): unknown block parameter", name);
}

extern "C" int b200clip_set_weight(b200clip_handle* h, const char* name, const float* data, const int64_t* shape,
                                   int ndim) {
    if (!h || !name || !data || !shape) return b200_fail(h, B200CLIP_E_ARG, "set_weight: null argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    const b200clip_config& c = h->cfg;
    const std::string n(name);
    const int64_t W = c.width, E = c.embed_dim, TW = c.text_width;
    int rc = 0;
    static const char* kVis = "visual.transformer.resblocks.";
    static const char* kTxt = "transformer.resblocks.";
    if (n.rfind(kVis, 0) == 0) rc = set_block_weight(h, h->vis, name, name + strlen(kVis), data, shape, ndim);
    else if (n.rfind(kTxt, 0) == 0) rc = set_block_weight(h, h->txt, name, name + strlen(kTxt), data, shape, ndim);
    else if (n == "visual.conv1.weight") {
        WANT_SHAPE(W, 3, c.patch, c.patch);
        rc = upload_bf16(h, data, W, 3 * c.patch * c.patch, h->patch_k, &h->w_patch);
    } else if (n == "visual.class_embedding") {
        WANT_SHAPE(W);
        h->host_cls.assign(data, data + W);
    } else if (n == "visual.positional_embedding") {
        WANT_SHAPE(h->tokens, W);
        h->host_pos0.assign(data, data + W);
        rc = upload_f32(h, data, static_cast<size_t>(h->tokens) * W, &h->pos_emb);
    } else if (n == "visual.ln_pre.weight") { WANT_SHAPE(W); rc = upload_f32(h, data, W, &h->ln_pre_g); }
    else if (n == "visual.ln_pre.bias") { WANT_SHAPE(W); rc = upload_f32(h, data, W, &h->ln_pre_b); }
    else if (n == "visual.ln_post.weight") { WANT_SHAPE(W); rc = upload_f32(h, data, W, &h->ln_post_g); }
    else if (n == "visual.ln_post.bias") { WANT_SHAPE(W); rc = upload_f32(h, data, W, &h->ln_post_b); }
    else if (n == "visual.proj") { WANT_SHAPE(W, E); rc = upload_bf16(h, data, W, E, E, &h->vis_proj); }
    else if (n == "token_embedding.weight") {
        WANT_SHAPE(c.text_vocab, TW);
        rc = upload_f32(h, data, static_cast<size_t>(c.text_vocab) * TW, &h->tok_emb);
    } else if (n == "positional_embedding") {
        WANT_SHAPE(c.text_ctx, TW);
        rc = upload_f32(h, data, static_cast<size_t>(c.text_ctx) * TW, &h->txt_pos);
    } else if (n == "ln_final.weight") { WANT_SHAPE(TW); rc = upload_f32(h, data, TW, &h->ln_final_g); }
    else if (n == "ln_final.bias") { WANT_SHAPE(TW); rc = upload_f32(h, data, TW, &h->ln_final_b); }
    else if (n == "text_projection") { WANT_SHAPE(TW, E); rc = upload_bf16(h, data, TW, E, E, &h->txt_proj); }
    else if (n == "logit_scale") { return 0; /* unused by the reference (plain cosine) */ }
    else return b200_fail(h, B200CLIP_E_ARG, "set_weight: unknown tensor name '%s'", name);
    if (rc == 0) h->have[n] = true;
    return rc;
}

// LN(x) W^T + b = rstd * (x (W.diag(g))^T - mean * c1) + c2,  c1[n] = sum_k bf16(W[n,k] g[k]),  c2[n] = sum_k beta[k] W[n,k] + b[n]
static int fold_layernorm(b200clip_handle* h, std::vector<float>& w, const std::vector<float>& bias,
                          const std::vector<float>& g, const std::vector<float>& beta, size_t rows, size_t cols,
                          bf16** w_dev, float** c1_dev, float** c2_dev) {
    std::vector<float> c1(rows), c2(rows);
    for (size_t n = 0; n < rows; ++n) {
        double s1 = 0.0, s2 = 0.0;
        float* wr = &w[n * cols];
        for (size_t k = 0; k < cols; ++k) {
            s2 += static_cast<double>(beta[k]) * wr[k];
            const float folded = __bfloat162float(__float2bfloat16(wr[k] * g[k]));
            wr[k] = folded;
            s1 += folded;
        }
        c1[n] = static_cast<float>(s1);
        c2[n] = static_cast<float>(s2 + bias[n]);
    }
    int rc = upload_bf16(h, w.data(), rows, cols, cols, w_dev);
    if (rc) return rc;
    if ((rc = upload_f32(h, c1.data(), rows, c1_dev))) return rc;
    return upload_f32(h, c2.data(), rows, c2_dev);
}

static int fold_tower(b200clip_handle* h, b200clip_handle::Tower& tw) {
    const size_t W = tw.width, F = tw.mlp;
    for (auto& b : tw.blocks) {
        int rc = fold_layernorm(h, b.h_w_qkv, b.h_b_qkv, b.h_ln1_g, b.h_ln1_b, 3 * W, W, &b.w_qkv, &b.c1_qkv, &b.c2_qkv);
        if (rc) return rc;
        if ((rc = fold_layernorm(h, b.h_w_fc, b.h_b_fc, b.h_ln2_g, b.h_ln2_b, F, W, &b.w_fc, &b.c1_fc, &b.c2_fc))) return rc;
        for (auto* v : {&b.h_ln1_g, &b.h_ln1_b, &b.h_ln2_g, &b.h_ln2_b, &b.h_w_qkv, &b.h_b_qkv, &b.h_w_fc, &b.h_b_fc})
            std::vector<float>().swap(*v);
    }
    return 0;
}

extern "C" int b200clip_finalize(b200clip_handle* h) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "finalize: null handle");
    B200_CUDA(h, cudaSetDevice(h->device));
    std::vector<std::string> need = {"visual.conv1.weight", "visual.class_embedding", "visual.positional_embedding",
                                     "visual.ln_pre.weight", "visual.ln_pre.bias", "visual.ln_post.weight",
                                     "visual.ln_post.bias", "visual.proj", "token_embedding.weight",
                                     "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection"};
    static const char* per_block[] = {"ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.in_proj_weight",
                                      "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                                      "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias"};
    for (int i = 0; i < h->cfg.layers; ++i)
        for (const char* p : per_block) need.push_back("visual.transformer.resblocks." + std::to_string(i) + "." + p);
    for (int i = 0; i < h->cfg.text_layers; ++i)
        for (const char* p : per_block) need.push_back("transformer.resblocks." + std::to_string(i) + "." + p);
    for (const std::string& k : need)
        if (!h->have.count(k)) return b200_fail(h, B200CLIP_E_STATE, "finalize: weight '%s' was never set", k.c_str());
    std::vector<float> cp(h->cfg.width);
    for (int i = 0; i < h->cfg.width; ++i) cp[i] = h->host_cls[i] + h->host_pos0[i];
    int rc = upload_f32(h, cp.data(), cp.size(), &h->cls_pos0);
    if (rc) return rc;
    if (!h->finalized) {
        if ((rc = fold_tower(h, h->vis))) return rc;
        if ((rc = fold_tower(h, h->txt))) return rc;
    }
    h->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------- workspace
// The image and the text tower own disjoint buffers and grow independently (a text call on one stream may overlap an
// image call on another stream of the same handle).  Growing is rare and frees buffers other streams of this handle
// may still be using, so the whole device is drained first.
static int ensure_image_ws(b200clip_handle* h, int images) {
    if (images <= h->ws_images) return 0;
    B200_CUDA(h, cudaDeviceSynchronize());
    const b200clip_config& c = h->cfg;
    const size_t mv = static_cast<size_t>(images) * h->tokens;
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    const size_t x_el = mv * c.width, qkv_el = 3 * x_el, h_el = mv * c.mlp_dim;
    const size_t p_el = static_cast<size_t>(images) * h->grid * h->grid * h->patch_k;
    cudaFree(h->ws_x); cudaFree(h->ws_y); cudaFree(h->ws_qkv); cudaFree(h->ws_h); cudaFree(h->ws_patches);
    cudaFree(h->ws_patches2); cudaFree(h->ws_stats);
    h->ws_stats = nullptr;
    h->ws_x = h->ws_y = h->ws_qkv = h->ws_h = h->ws_patches = h->ws_patches2 = nullptr;
    h->ws_images = 0;
    B200_CUDA(h, cudaMalloc(&h->ws_x, mx(x_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_y, mx(x_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_qkv, mx(qkv_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_h, mx(h_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_patches, mx(p_el, 8) * 2));
    if (b200_knobs().overlap) B200_CUDA(h, cudaMalloc(&h->ws_patches2, mx(p_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_stats, mx(mv, 1) * 16 * sizeof(float)));
    h->ws_images = images;
    return 0;
}

static int ensure_text_ws(b200clip_handle* h, int texts) {
    if (texts <= h->ws_texts) return 0;
    B200_CUDA(h, cudaDeviceSynchronize());
    const b200clip_config& c = h->cfg;
    const size_t mt = static_cast<size_t>(texts) * c.text_ctx;
    auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
    const size_t tx_el = mt * c.text_width, th_el = mt * c.text_mlp_dim;
    cudaFree(h->ws_eot); cudaFree(h->ws_tokens);
    cudaFree(h->ws_tx); cudaFree(h->ws_ty); cudaFree(h->ws_tqkv); cudaFree(h->ws_th); cudaFree(h->ws_tstats);
    h->ws_tx = h->ws_ty = h->ws_tqkv = h->ws_th = nullptr; h->ws_tstats = nullptr;
    h->ws_eot = nullptr; h->ws_tokens = nullptr;
    h->ws_texts = 0;
    B200_CUDA(h, cudaMalloc(&h->ws_tx, mx(tx_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_ty, mx(tx_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_tqkv, mx(3 * tx_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_th, mx(th_el, 8) * 2));
    B200_CUDA(h, cudaMalloc(&h->ws_tstats, mx(mt, 1) * 16 * sizeof(float)));
    B200_CUDA(h, cudaMalloc(&h->ws_eot, mx(texts, 1) * sizeof(int32_t)));
    B200_CUDA(h, cudaMalloc(&h->ws_tokens, mx(mt, 1) * sizeof(int64_t)));
    h->ws_texts = texts;
    return 0;
}

static int ensure_workspace(b200clip_handle* h, int images, int texts, cudaStream_t) {
    int rc = ensure_image_ws(h, images);
    if (rc) return rc;
    return ensure_text_ws(h, texts);
}

static const int kDefaultChunk = 1024;  // images per pass of the tower when the caller reserved nothing

// An explicit reserve fixes the images per pass of the tower (the caller sized the workspace: bench --chunk,
// settings.B200_MAX_IMAGES_PER_PASS); without one the default chunk applies whatever earlier calls happened to need.
extern "C" int b200clip_reserve(b200clip_handle* h, int max_images, int max_texts) {
    if (!h || max_images < 0 || max_texts < 0) return b200_fail(h, B200CLIP_E_ARG, "reserve: bad argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    if (max_images > 0) h->chunk_cap = max_images;
    return ensure_workspace(h, max_images, max_texts, nullptr);
}

// ------------------------------------------------------------------------------------------- towers
// Residual blocks.  ws_x holds the residual stream and ws_stats the (sum, sum of squares) partials of its rows
// (written by ln_pre / the text embedding, then by every residual GEMM); ln_1 / ln_2 never run as kernels: they
// are folded into the QKV / fc GEMM epilogues.
struct TowerWs { bf16 *x, *y, *qkv, *hid; float* stats; };
static int run_blocks(b200clip_handle* h, b200clip_handle::Tower& tw, const TowerWs& ws, int n_seq, int T, int causal,
                      cudaStream_t st) {
    const int M = n_seq * T;
    const int W = tw.width, F = tw.mlp;
    const int act = h->cfg.act == 0 ? 1 : 2;
    int rc;
    for (int l = 0; l < tw.layers; ++l) {
        const b200clip_handle::Block& b = tw.blocks[l];
        b200::GemmEpilogue ep{};
        ep.bias = b.c2_qkv; ep.ln_stats = ws.stats; ep.ln_c1 = b.c1_qkv; ep.ln_eps = h->cfg.ln_eps; ep.ln_width = W;
        if ((rc = launch_gemm(h, ws.x, W, b.w_qkv, W, ws.qkv, 3 * W, M, 3 * W, W, ep, st))) return rc;
        if ((rc = launch_attention(h, ws.qkv, ws.y, n_seq, T, tw.heads, causal, st))) return rc;
        ep = {}; ep.bias = b.b_out; ep.resid = ws.x; ep.stats_out = ws.stats;
        if ((rc = launch_gemm(h, ws.y, W, b.w_out, W, ws.x, W, M, W, W, ep, st))) return rc;
        ep = {}; ep.bias = b.c2_fc; ep.ln_stats = ws.stats; ep.ln_c1 = b.c1_fc; ep.ln_eps = h->cfg.ln_eps;
        ep.ln_width = W; ep.act = act;
        if ((rc = launch_gemm(h, ws.x, W, b.w_fc, W, ws.hid, F, M, F, W, ep, st))) return rc;
        ep = {}; ep.bias = b.b_proj; ep.resid = ws.x; ep.stats_out = ws.stats;
        if ((rc = launch_gemm(h, ws.hid, F, b.w_proj, F, ws.x, W, M, W, F, ep, st))) return rc;
    }
    return 0;
}

// patches (device, [n*g*g, patch_k]) -> embeddings; n <= ws_images
static int encode_patches_chunk(b200clip_handle* h, const bf16* patches, int n, void* out, int out_dtype, int l2norm,
                                cudaStream_t st) {
    const b200clip_config& c = h->cfg;
    const int g2 = h->grid * h->grid;
    int rc;
    b200::GemmEpilogue ep{};
    ep.rowtab = h->pos_emb; ep.t_in = g2; ep.t_out = h->tokens; ep.row_off = 1;
    if ((rc = launch_gemm(h, patches, h->patch_k, h->w_patch, h->patch_k, h->ws_y, c.width, n * g2, c.width, h->patch_k,
                          ep, st)))
        return rc;
    B200_CUDA(h, cudaMemsetAsync(h->ws_stats, 0, static_cast<size_t>(n) * h->tokens * 16 * sizeof(float), st));
    if ((rc = launch_layernorm(h, h->ws_y, h->ln_pre_g, h->ln_pre_b, h->ws_x, static_cast<int64_t>(n) * h->tokens,
                               c.width, c.ln_eps, h->tokens, h->cls_pos0, h->ws_stats, st)))
        return rc;
    const TowerWs vws{h->ws_x, h->ws_y, h->ws_qkv, h->ws_h, h->ws_stats};
    if ((rc = run_blocks(h, h->vis, vws, n, h->tokens, 0, st))) return rc;
    return launch_head(h, h->ws_x, static_cast<int64_t>(h->tokens) * c.width, nullptr, h->ln_post_g, h->ln_post_b,
                       h->vis_proj, n, c.width, c.embed_dim, c.ln_eps, out, out_dtype, l2norm, st);
}

static int check_ready(b200clip_handle* h, const char* what) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "%s: null handle", what);
    if (!h->finalized) return b200_fail(h, B200CLIP_E_STATE, "%s: handle not finalized (weights missing)", what);
    if (cudaSetDevice(h->device) != cudaSuccess) return b200_fail(h, B200CLIP_E_CUDA, "%s: cudaSetDevice failed", what);
    return 0;
}

static size_t out_elem(int dt) { return dt == B200CLIP_BF16 ? 2 : 4; }

static int chunk_images(b200clip_handle* h, int n) {
    const int cap = h->chunk_cap > 0 ? h->chunk_cap : kDefaultChunk;
    return n < cap ? n : cap;
}

extern "C" int b200clip_encode_patches(b200clip_handle* h, const void* patches_dev, int n, void* emb_out_dev,
                                       int out_dtype, int l2norm, void* stream) {
    int rc = check_ready(h, "encode_patches");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!patches_dev || !emb_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "encode_patches: bad argument");
    if (out_dtype != B200CLIP_F32 && out_dtype != B200CLIP_BF16) return b200_fail(h, B200CLIP_E_ARG, "bad out_dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int chunk = chunk_images(h, n);
    if (n > 0 && (rc = ensure_workspace(h, chunk, 0, st))) return rc;
    const size_t prow = static_cast<size_t>(h->grid) * h->grid * h->patch_k;
    for (int i = 0; i < n; i += chunk) {
        const int nc = (n - i) < chunk ? (n - i) : chunk;
        rc = encode_patches_chunk(h, static_cast<const bf16*>(patches_dev) + i * prow, nc,
                                  static_cast<uint8_t*>(emb_out_dev) + static_cast<size_t>(i) * h->cfg.embed_dim * out_elem(out_dtype),
                                  out_dtype, l2norm, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int b200clip_encode_image_chw(b200clip_handle* h, const float* chw_dev, int n, void* emb_out_dev,
                                         int out_dtype, int l2norm, void* stream) {
    int rc = check_ready(h, "encode_image_chw");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!chw_dev || !emb_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "encode_image_chw: bad argument");
    if (out_dtype != B200CLIP_F32 && out_dtype != B200CLIP_BF16) return b200_fail(h, B200CLIP_E_ARG, "bad out_dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int chunk = chunk_images(h, n);
    if (n > 0 && (rc = ensure_workspace(h, chunk, 0, st))) return rc;
    const size_t S = h->cfg.image_size;
    for (int i = 0; i < n; i += chunk) {
        const int nc = (n - i) < chunk ? (n - i) : chunk;
        if ((rc = launch_patchify_chw(h, chw_dev + static_cast<size_t>(i) * 3 * S * S, nc, h->ws_patches, st))) return rc;
        rc = encode_patches_chunk(h, h->ws_patches, nc,
                                  static_cast<uint8_t*>(emb_out_dev) + static_cast<size_t>(i) * h->cfg.embed_dim * out_elem(out_dtype),
                                  out_dtype, l2norm, st);
        if (rc) return rc;
    }
    return 0;
}

// public device-frame entry points take whole frames: strides must cover them (the internal host-upload path hands
// K1 a compacted window instead and is checked against that window in launch_preprocess)
static int check_frame_strides(b200clip_handle* h, int n, int height, int width, int64_t frame_stride, int64_t row_stride) {
    if (n > 0 && (height <= 0 || width <= 0 || row_stride < static_cast<int64_t>(width) * 3 ||
                  frame_stride < row_stride * height))
        return b200_fail(h, B200CLIP_E_ARG, "bad frame geometry %dx%d strides %lld/%lld", width, height,
                         (long long)row_stride, (long long)frame_stride);
    return 0;
}

extern "C" int b200clip_preprocess_u8(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                                      int64_t frame_stride, int64_t row_stride, int resize_mode, void* patches_out_dev,
                                      void* stream) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "preprocess: null handle");
    if (n < 0 || (n > 0 && (!frames_dev || !patches_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "preprocess: bad argument");
    if (int rc = check_frame_strides(h, n, height, width, frame_stride, row_stride)) return rc;
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_preprocess(h, frames_dev, n, height, width, frame_stride, row_stride, resize_mode,
                             static_cast<bf16*>(patches_out_dev), nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_preprocess_u8_chw(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                                          int64_t frame_stride, int64_t row_stride, int resize_mode,
                                          float* chw_out_dev, void* stream) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "preprocess: null handle");
    if (n < 0 || (n > 0 && (!frames_dev || !chw_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "preprocess: bad argument");
    if (int rc = check_frame_strides(h, n, height, width, frame_stride, row_stride)) return rc;
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_preprocess(h, frames_dev, n, height, width, frame_stride, row_stride, resize_mode, nullptr,
                             chw_out_dev, static_cast<cudaStream_t>(stream));
}

// Device frames -> embeddings in `parts` pipelined chunks: K1 of chunk i+1 runs on a second, lower-priority stream
// while the tower of chunk i runs on the caller's stream (the GEMM kernels keep one register- and smem-heavy CTA per
// SM whose issue slots are mostly idle; the register-light K1 kernels co-reside with it).
static int ensure_pipeline(b200clip_handle* h) {
    if (h->pre_stream) return 0;
    int lo = 0, hi = 0;
    B200_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200_CUDA(h, cudaStreamCreateWithPriority(&h->pre_stream, cudaStreamNonBlocking, lo));
    for (int i = 0; i < 2; ++i) {
        B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_pre[i], cudaEventDisableTiming));
        B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_tower[i], cudaEventDisableTiming));
    }
    B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    return 0;
}

extern "C" int b200clip_encode_frames_u8(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                                         int64_t frame_stride, int64_t row_stride, int resize_mode, void* emb_out_dev,
                                         int out_dtype, int l2norm, void* stream) {
    int rc = check_ready(h, "encode_frames_u8");
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!frames_dev || !emb_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "encode_frames_u8: bad argument");
    if ((rc = check_frame_strides(h, n, height, width, frame_stride, row_stride))) return rc;
    if (out_dtype != B200CLIP_F32 && out_dtype != B200CLIP_BF16) return b200_fail(h, B200CLIP_E_ARG, "bad out_dtype");
    if (n == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int chunk = chunk_images(h, n);
    // large batches: at least 4 pipelined chunks so that only the first K1 pass is exposed
    const bool pipelined = b200_knobs().overlap && n >= 256;  // measured slower on B200 (K1 blocks crowd out the tower): opt-in
    if (pipelined) {
        int parts = (n + chunk - 1) / chunk;
        if (parts < 4) parts = 4;
        chunk = (n + parts - 1) / parts;
    }
    if ((rc = ensure_workspace(h, chunk, 0, st))) return rc;
    auto emb_at = [&](int i) {
        return static_cast<uint8_t*>(emb_out_dev) + static_cast<size_t>(i) * h->cfg.embed_dim * out_elem(out_dtype);
    };
    if (!pipelined) {
        for (int i = 0; i < n; i += chunk) {
            const int nc = (n - i) < chunk ? (n - i) : chunk;
            if ((rc = launch_preprocess(h, frames_dev + static_cast<int64_t>(i) * frame_stride, nc, height, width,
                                        frame_stride, row_stride, resize_mode, h->ws_patches, nullptr, st)))
                return rc;
            if ((rc = encode_patches_chunk(h, h->ws_patches, nc, emb_at(i), out_dtype, l2norm, st))) return rc;
        }
        return 0;
    }
    if ((rc = ensure_pipeline(h))) return rc;
    bf16* pbuf[2] = {h->ws_patches, h->ws_patches2};
    const int nchunks = (n + chunk - 1) / chunk;
    // fork: K1 may not start before the work already queued on the caller's stream (it produced the frames)
    B200_CUDA(h, cudaEventRecord(h->ev_fork, st));
    B200_CUDA(h, cudaStreamWaitEvent(h->pre_stream, h->ev_fork, 0));
    auto issue_pre = [&](int ci) -> int {
        const int b = ci & 1, i0 = ci * chunk;
        const int nc = (n - i0) < chunk ? (n - i0) : chunk;
        if (ci >= 2) B200_CUDA(h, cudaStreamWaitEvent(h->pre_stream, h->ev_tower[b], 0));  // patch buffer b is free
        int r = launch_preprocess(h, frames_dev + static_cast<int64_t>(i0) * frame_stride, nc, height, width,
                                  frame_stride, row_stride, resize_mode, pbuf[b], nullptr, h->pre_stream);
        if (r) return r;
        B200_CUDA(h, cudaEventRecord(h->ev_pre[b], h->pre_stream));
        return 0;
    };
    if ((rc = issue_pre(0))) return rc;
    for (int ci = 0; ci < nchunks; ++ci) {
        const int b = ci & 1, i0 = ci * chunk;
        const int nc = (n - i0) < chunk ? (n - i0) : chunk;
        if (ci + 1 < nchunks && (rc = issue_pre(ci + 1))) return rc;
        B200_CUDA(h, cudaStreamWaitEvent(st, h->ev_pre[b], 0));
        if ((rc = encode_patches_chunk(h, pbuf[b], nc, emb_at(i0), out_dtype, l2norm, st))) return rc;
        B200_CUDA(h, cudaEventRecord(h->ev_tower[b], st));
    }
    return 0;   // the caller's stream has joined every K1 pass through ev_pre
}

// HOST frames -> HOST embeddings.  Device staging is double buffered: the copy stream uploads chunk i+1 while the
// compute stream runs preprocess + tower on chunk i.  Pinned caller memory is copied directly; pageable memory
// goes through pinned bounce buffers.  nv12 = false: packed RGB frames (H*W*3 bytes each); nv12 = true: NV12 frames
// (H*W*3/2 bytes each: Y plane, then the interleaved UV plane) -- half the bytes over PCIe.
static int encode_frames_host(b200clip_handle* h, const uint8_t* frames_host, int n, int height, int width, bool nv12,
                              int resize_mode, float* emb_out_host, int l2norm, void* stream, const char* what) {
    int rc = check_ready(h, what);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!frames_host || !emb_out_host))) return b200_fail(h, B200CLIP_E_ARG, "%s: bad argument", what);
    if (n == 0) return 0;
    if (height <= 0 || width <= 0) return b200_fail(h, B200CLIP_E_ARG, "%s: bad frame size", what);
    if (nv12 && ((height | width) & 1)) return b200_fail(h, B200CLIP_E_SHAPE, "%s: NV12 needs even width and height", what);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bpp = nv12 ? 1 : 3;                                // bytes per pixel of a source row
    const size_t src_pitch = static_cast<size_t>(width) * bpp;
    const size_t src_rows = nv12 ? static_cast<size_t>(height) * 3 / 2 : static_cast<size_t>(height);   // rows per frame
    const size_t fbytes = src_pitch * src_rows;
    // Only the window of each frame that survives the transform's centre crop is uploaded (for 1080p: 1104 of the 1920
    // columns): strided 3-D copies pack rows [wy0, wy1) x columns [wx0, wx1) of every frame (NV12: of both planes)
    // into the staging buffer, and K1 is pointed at a virtual frame origin in front of it.
    int wx0, wx1, wy0, wy1;
    if (nv12) rc = preprocess_source_window_nv12(h, height, width, resize_mode, &wx0, &wx1, &wy0, &wy1);
    else rc = preprocess_source_window(h, height, width, resize_mode, &wx0, &wx1, &wy0, &wy1);
    if (rc) return rc;
    if (b200_knobs().full_upload) { wx0 = 0; wx1 = width; wy0 = 0; wy1 = height; }
    wx0 &= ~15;                                                     // 16-pixel (RGB: 48-byte) aligned window start
    const int wrows = wy1 - wy0;
    const int uv0 = wy0 >> 1, uvrows = nv12 ? ((wy1 + 1) >> 1) - uv0 : 0;       // chroma rows of the window
    const size_t wbytes = static_cast<size_t>(wx1 - wx0) * bpp;     // bytes copied per row
    const size_t pitch = (wbytes + 63) & ~size_t(63);               // staging row pitch
    const size_t fpitch = pitch * wrows, uvfpitch = pitch * uvrows; // staging frame pitches of the (luma) and chroma windows
    const size_t lead = 64;                                         // K1 may read the aligned word in front of a row
    int chunk = chunk_images(h, n);
    const size_t budget = size_t(1) << 30;  // ~1 GiB of frames per staging buffer
    if (static_cast<size_t>(chunk) * (fpitch + uvfpitch) > budget) chunk = static_cast<int>(budget / (fpitch + uvfpitch));
    if (chunk < 1) chunk = 1;
    if ((rc = ensure_workspace(h, chunk, 0, st))) return rc;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, frames_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const size_t sbytes = static_cast<size_t>(chunk) * (fpitch + uvfpitch) + 2 * lead;
    const size_t hbytes = static_cast<size_t>(chunk) * fbytes;      // pageable input: whole frames bounce through pinned memory
    if (sbytes > h->ws_stage_bytes || (!pinned && (!h->ws_stage_host[0] || hbytes > h->ws_stage_host_bytes))) {
        B200_CUDA(h, cudaStreamSynchronize(st));
        if (h->copy_stream) B200_CUDA(h, cudaStreamSynchronize(h->copy_stream));
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->ws_stage_dev[i]); h->ws_stage_dev[i] = nullptr;
            if (h->ws_stage_host[i]) { cudaFreeHost(h->ws_stage_host[i]); h->ws_stage_host[i] = nullptr; }
        }
        const size_t nb = sbytes > h->ws_stage_bytes ? sbytes : h->ws_stage_bytes;
        const size_t nh = hbytes > h->ws_stage_host_bytes ? hbytes : h->ws_stage_host_bytes;
        h->ws_stage_bytes = 0; h->ws_stage_host_bytes = 0;
        for (int i = 0; i < 2; ++i) {
            B200_CUDA(h, cudaMalloc(&h->ws_stage_dev[i], nb));
            if (!pinned) B200_CUDA(h, cudaMallocHost(&h->ws_stage_host[i], nh));
        }
        h->ws_stage_bytes = nb;
        if (!pinned) h->ws_stage_host_bytes = nh;
    }
    if (!h->copy_stream) {
        B200_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
    }
    // the result buffer may be host memory (reference behaviour: numpy out) or device memory (keeps the
    // embeddings resident for K4)
    cudaPointerAttributes oattr{};
    const bool out_on_device = cudaPointerGetAttributes(&oattr, emb_out_host) == cudaSuccess &&
                               (oattr.type == cudaMemoryTypeDevice || oattr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    const size_t emb_el = static_cast<size_t>(n) * h->cfg.embed_dim;
    if (!out_on_device && emb_el > h->ws_emb_elems) {
        B200_CUDA(h, cudaStreamSynchronize(st));
        cudaFree(h->ws_emb); h->ws_emb = nullptr; h->ws_emb_elems = 0;
        B200_CUDA(h, cudaMalloc(&h->ws_emb, emb_el * sizeof(float)));
        h->ws_emb_elems = emb_el;
    }
    float* emb_dev = out_on_device ? emb_out_host : h->ws_emb;
    // The two staging buffers alternate ACROSS calls too (stage_seq), and an upload into buffer b only waits for the
    // last compute that read buffer b (ev_done[b], possibly recorded by an earlier call): the first upload of a call
    // overlaps the tower of the previous call instead of waiting for the whole compute stream to drain.
    const int nchunks = (n + chunk - 1) / chunk;
    for (int ci = 0; ci < nchunks; ++ci) {
        const int b = static_cast<int>((h->stage_seq + ci) & 1);
        const int i0 = ci * chunk;
        const int nc = (n - i0) < chunk ? (n - i0) : chunk;
        const uint8_t* src = frames_host + static_cast<size_t>(i0) * fbytes;
        B200_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));   // (a never-recorded event is a no-op)
        if (!pinned) {
            B200_CUDA(h, cudaEventSynchronize(h->ev_h2d[b]));                   // bounce buffer b is free again
            memcpy(h->ws_stage_host[b], src, static_cast<size_t>(nc) * fbytes);
            src = h->ws_stage_host[b];
        }
        uint8_t* stage = h->ws_stage_dev[b] + lead;
        uint8_t* stage_uv = stage + static_cast<size_t>(nc) * fpitch;       // NV12: chroma windows behind the luma windows
        auto upload = [&](const uint8_t* from, uint8_t* to, int rows) -> int {
            cudaMemcpy3DParms cp{};
            cp.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(from), src_pitch, src_pitch, src_rows);
            cp.dstPtr = make_cudaPitchedPtr(to, pitch, pitch, rows);
            cp.extent = make_cudaExtent(wbytes, rows, nc);
            cp.kind = cudaMemcpyHostToDevice;
            B200_CUDA(h, cudaMemcpy3DAsync(&cp, h->copy_stream));
            h->h2d_bytes += static_cast<int64_t>(wbytes) * rows * nc;
            return 0;
        };
        if ((rc = upload(src + static_cast<size_t>(wy0) * src_pitch + static_cast<size_t>(wx0) * bpp, stage, wrows))) return rc;
        if (nv12 && (rc = upload(src + (static_cast<size_t>(height) + uv0) * src_pitch + wx0, stage_uv, uvrows))) return rc;
        B200_CUDA(h, cudaEventRecord(h->ev_h2d[b], h->copy_stream));
        B200_CUDA(h, cudaStreamWaitEvent(st, h->ev_h2d[b], 0));
        // virtual origin: staging (row 0, byte 0) is source (row wy0, column wx0)
        const int64_t ip = static_cast<int64_t>(pitch);
        if (nv12) {
            rc = launch_preprocess_nv12(h, stage - wy0 * ip - wx0, stage_uv - uv0 * ip - wx0, nc, height, width,
                                        static_cast<int64_t>(fpitch), static_cast<int64_t>(uvfpitch), ip, resize_mode,
                                        h->ws_patches, nullptr, st);
        } else {
            rc = launch_preprocess(h, stage - wy0 * ip - static_cast<int64_t>(wx0) * 3, nc, height, width,
                                   static_cast<int64_t>(fpitch), ip, resize_mode, h->ws_patches, nullptr, st);
        }
        if (rc) return rc;
        if ((rc = encode_patches_chunk(h, h->ws_patches, nc, emb_dev + static_cast<size_t>(i0) * h->cfg.embed_dim,
                                       B200CLIP_F32, l2norm, st)))
            return rc;
        B200_CUDA(h, cudaEventRecord(h->ev_done[b], st));
    }
    h->stage_seq += nchunks;
    if (out_on_device) {
        // frames_host may be reused by the caller once the uploads are done; compute stays asynchronous on `st`
        B200_CUDA(h, cudaStreamSynchronize(h->copy_stream));
        return 0;
    }
    B200_CUDA(h, cudaMemcpyAsync(emb_out_host, h->ws_emb, emb_el * sizeof(float), cudaMemcpyDeviceToHost, st));
    h->d2h_bytes += static_cast<int64_t>(emb_el * sizeof(float));
    B200_CUDA(h, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int b200clip_encode_frames_u8_host(b200clip_handle* h, const uint8_t* frames_host, int n, int height,
                                              int width, int resize_mode, float* emb_out_host, int l2norm,
                                              void* stream) {
    return encode_frames_host(h, frames_host, n, height, width, false, resize_mode, emb_out_host, l2norm, stream,
                              "encode_frames_u8_host");
}

extern "C" int b200clip_encode_frames_nv12_host(b200clip_handle* h, const uint8_t* nv12_host, int n, int height,
                                                int width, int resize_mode, float* emb_out_host, int l2norm,
                                                void* stream) {
    return encode_frames_host(h, nv12_host, n, height, width, true, resize_mode, emb_out_host, l2norm, stream,
                              "encode_frames_nv12_host");
}

// ---- NV12 frames resident on the device (what NVDEC writes): Y plane + interleaved UV plane, common row pitch
static int check_nv12_args(b200clip_handle* h, const void* y, const void* uv, int n, int height, int width, int64_t y_fs,
                           int64_t uv_fs, int64_t rs, const void* out, const char* what) {
    if (n < 0 || (n > 0 && (!y || !uv || !out))) return b200_fail(h, B200CLIP_E_ARG, "%s: bad argument", what);
    if (n > 0 && (height <= 0 || width <= 0 || ((height | width) & 1)))
        return b200_fail(h, B200CLIP_E_SHAPE, "%s: NV12 needs positive even width and height (%dx%d)", what, width, height);
    if (n > 0 && (rs < width || y_fs < rs * height || uv_fs < rs * (height / 2)))
        return b200_fail(h, B200CLIP_E_ARG, "%s: bad strides %lld/%lld/%lld for %dx%d", what, (long long)rs, (long long)y_fs,
                         (long long)uv_fs, width, height);
    return 0;
}

extern "C" int b200clip_preprocess_nv12(b200clip_handle* h, const uint8_t* y_dev, const uint8_t* uv_dev, int n, int height,
                                        int width, int64_t y_frame_stride, int64_t uv_frame_stride, int64_t row_stride,
                                        int resize_mode, void* patches_out_dev, float* chw_out_dev, void* stream) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "preprocess_nv12: null handle");
    if (!patches_out_dev && !chw_out_dev) return b200_fail(h, B200CLIP_E_ARG, "preprocess_nv12: no output buffer");
    if (int rc = check_nv12_args(h, y_dev, uv_dev, n, height, width, y_frame_stride, uv_frame_stride, row_stride,
                                 patches_out_dev ? patches_out_dev : chw_out_dev, "preprocess_nv12"))
        return rc;
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_preprocess_nv12(h, y_dev, uv_dev, n, height, width, y_frame_stride, uv_frame_stride, row_stride,
                                  resize_mode, static_cast<bf16*>(patches_out_dev), chw_out_dev,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_encode_frames_nv12(b200clip_handle* h, const uint8_t* y_dev, const uint8_t* uv_dev, int n,
                                           int height, int width, int64_t y_frame_stride, int64_t uv_frame_stride,
                                           int64_t row_stride, int resize_mode, void* emb_out_dev, int out_dtype,
                                           int l2norm, void* stream) {
    int rc = check_ready(h, "encode_frames_nv12");
    if (rc) return rc;
    if ((rc = check_nv12_args(h, y_dev, uv_dev, n, height, width, y_frame_stride, uv_frame_stride, row_stride, emb_out_dev,
                              "encode_frames_nv12")))
        return rc;
    if (out_dtype != B200CLIP_F32 && out_dtype != B200CLIP_BF16) return b200_fail(h, B200CLIP_E_ARG, "bad out_dtype");
    if (n == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int chunk = chunk_images(h, n);
    if ((rc = ensure_workspace(h, chunk, 0, st))) return rc;
    for (int i = 0; i < n; i += chunk) {
        const int nc = (n - i) < chunk ? (n - i) : chunk;
        if ((rc = launch_preprocess_nv12(h, y_dev + static_cast<int64_t>(i) * y_frame_stride,
                                         uv_dev + static_cast<int64_t>(i) * uv_frame_stride, nc, height, width, y_frame_stride,
                                         uv_frame_stride, row_stride, resize_mode, h->ws_patches, nullptr, st)))
            return rc;
        if ((rc = encode_patches_chunk(h, h->ws_patches, nc,
                                       static_cast<uint8_t*>(emb_out_dev) + static_cast<size_t>(i) * h->cfg.embed_dim * out_elem(out_dtype),
                                       out_dtype, l2norm, st)))
            return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------- text
static int encode_text_dev(b200clip_handle* h, const int64_t* tokens_dev, int q, float* out_dev, int l2norm,
                           cudaStream_t st) {
    const b200clip_config& c = h->cfg;
    int rc;
    const int chunk = q < 256 ? q : 256;      // texts per pass of the tower; the workspace grows to it
    if ((rc = ensure_text_ws(h, chunk))) return rc;
    for (int i = 0; i < q; i += chunk) {
        const int nc = (q - i) < chunk ? (q - i) : chunk;
        B200_CUDA(h, cudaMemsetAsync(h->ws_tstats, 0, static_cast<size_t>(nc) * c.text_ctx * 16 * sizeof(float), st));
        if ((rc = launch_text_embed(h, tokens_dev + static_cast<size_t>(i) * c.text_ctx, nc, h->ws_tx, h->ws_eot,
                                    h->ws_tstats, st)))
            return rc;
        const TowerWs tws{h->ws_tx, h->ws_ty, h->ws_tqkv, h->ws_th, h->ws_tstats};
        if ((rc = run_blocks(h, h->txt, tws, nc, c.text_ctx, 1, st))) return rc;
        if ((rc = launch_head(h, h->ws_tx, c.text_width, h->ws_eot, h->ln_final_g, h->ln_final_b, h->txt_proj, nc,
                              c.text_width, c.embed_dim, c.ln_eps, out_dev + static_cast<size_t>(i) * c.embed_dim,
                              B200CLIP_F32, l2norm, st)))
            return rc;
    }
    return 0;
}

extern "C" int b200clip_encode_text(b200clip_handle* h, const int64_t* tokens_dev, int q, float* emb_out_dev,
                                    int l2norm, void* stream) {
    int rc = check_ready(h, "encode_text");
    if (rc) return rc;
    if (q < 0 || (q > 0 && (!tokens_dev || !emb_out_dev))) return b200_fail(h, B200CLIP_E_ARG, "encode_text: bad argument");
    if (q == 0) return 0;
    return encode_text_dev(h, tokens_dev, q, emb_out_dev, l2norm, static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_encode_text_host(b200clip_handle* h, const int64_t* tokens_host, int q, float* emb_out_host,
                                         int l2norm, void* stream) {
    int rc = check_ready(h, "encode_text_host");
    if (rc) return rc;
    if (q < 0 || (q > 0 && (!tokens_host || !emb_out_host))) return b200_fail(h, B200CLIP_E_ARG, "encode_text_host: bad argument");
    if (q == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const b200clip_config& c = h->cfg;
    int64_t* tok_dev = nullptr;
    float* emb_dev = nullptr;
    B200_CUDA(h, cudaMalloc(&tok_dev, static_cast<size_t>(q) * c.text_ctx * sizeof(int64_t)));
    if (cudaMalloc(&emb_dev, static_cast<size_t>(q) * c.embed_dim * sizeof(float)) != cudaSuccess) {
        cudaFree(tok_dev);
        return b200_fail(h, B200CLIP_E_NOMEM, "encode_text_host: out of device memory");
    }
    cudaError_t e = cudaMemcpyAsync(tok_dev, tokens_host, static_cast<size_t>(q) * c.text_ctx * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, st);
    h->h2d_bytes += static_cast<int64_t>(q) * c.text_ctx * sizeof(int64_t);
    h->d2h_bytes += static_cast<int64_t>(q) * c.embed_dim * sizeof(float);
    if (e == cudaSuccess) {
        rc = encode_text_dev(h, tok_dev, q, emb_dev, l2norm, st);
        if (rc == 0)
            e = cudaMemcpyAsync(emb_out_host, emb_dev, static_cast<size_t>(q) * c.embed_dim * sizeof(float),
                                cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tok_dev);
    cudaFree(emb_dev);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess)
        return b200_fail(h, B200CLIP_E_CUDA, "encode_text_host: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return 0;
}

// ------------------------------------------------------------------------------------------- K4
extern "C" int b200clip_sim_topk(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                                 const float* txt_emb_dev, int q, int k, float threshold, const double* timestamps_dev,
                                 int64_t index_base, double clip_duration, double video_duration,
                                 float* top_scores_dev, int64_t* top_idx_dev, double* intervals_dev,
                                 int32_t* counts_dev, void* stream) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "sim_topk: null handle");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_sim_topk(h, img_emb_dev, emb_dtype, n, e, txt_emb_dev, q, k, threshold, timestamps_dev, index_base,
                           clip_duration, video_duration, top_scores_dev, top_idx_dev, intervals_dev, counts_dev, nullptr,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_sim_topk_dense(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                                       const float* txt_emb_dev, int q, int k, float threshold, float* top_scores_dev,
                                       int64_t* top_idx_dev, int32_t* counts_dev, float* dense_scores_dev, void* stream) {
    if (!h || !dense_scores_dev) return b200_fail(h, B200CLIP_E_ARG, "sim_topk_dense: null argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_sim_topk(h, img_emb_dev, emb_dtype, n, e, txt_emb_dev, q, k, threshold, nullptr, 0, 30.0, 0.0,
                           top_scores_dev, top_idx_dev, nullptr, counts_dev, dense_scores_dev,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_similarity(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                                   const float* txt_emb_dev, int q, float* scores_out_dev, void* stream) {
    if (!h || !scores_out_dev) return b200_fail(h, B200CLIP_E_ARG, "similarity: null argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_similarity(h, img_emb_dev, emb_dtype, n, e, txt_emb_dev, q, scores_out_dev,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_topk_merge(b200clip_handle* h, const float* cand_scores_dev, const int64_t* cand_idx_dev, int g,
                                   int q, int k, float threshold, const double* timestamps_dev, double clip_duration,
                                   double video_duration, float* top_scores_dev, int64_t* top_idx_dev,
                                   double* intervals_dev, int32_t* counts_dev, void* stream) {
    if (!h) return b200_fail(h, B200CLIP_E_ARG, "topk_merge: null handle");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_topk_merge(h, cand_scores_dev, cand_idx_dev, g, q, k, threshold, timestamps_dev, clip_duration,
                             video_duration, top_scores_dev, top_idx_dev, intervals_dev, counts_dev,
                             static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------- building blocks
extern "C" int b200clip_gemm_bf16(b200clip_handle* h, const void* a_dev, const void* w_dev, void* out_dev, int m,
                                  int n, int k, const float* bias_dev, const void* resid_dev, int act,
                                  void* stream) {
    if (!h || !a_dev || !w_dev || !out_dev) return b200_fail(h, B200CLIP_E_ARG, "gemm: null argument");
    if (act < 0 || act > 2) return b200_fail(h, B200CLIP_E_ARG, "gemm: act must be 0, 1 or 2");
    B200_CUDA(h, cudaSetDevice(h->device));
    b200::GemmEpilogue ep{};
    ep.bias = bias_dev;
    ep.resid = static_cast<const bf16*>(resid_dev);
    ep.act = act;
    return launch_gemm(h, static_cast<const bf16*>(a_dev), k, static_cast<const bf16*>(w_dev), k,
                       static_cast<bf16*>(out_dev), n, m, n, k, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_layernorm_bf16(b200clip_handle* h, const void* x_dev, const float* gamma_dev,
                                       const float* beta_dev, void* y_dev, int64_t rows, int width, float eps,
                                       void* stream) {
    if (!h || !x_dev || !gamma_dev || !beta_dev || !y_dev) return b200_fail(h, B200CLIP_E_ARG, "layernorm: null argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_layernorm(h, static_cast<const bf16*>(x_dev), gamma_dev, beta_dev, static_cast<bf16*>(y_dev), rows,
                            width, eps, 0, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int b200clip_attention_bf16(b200clip_handle* h, const void* qkv_dev, void* out_dev, int n_seq, int t,
                                       int heads, int causal, void* stream) {
    if (!h || !qkv_dev || !out_dev) return b200_fail(h, B200CLIP_E_ARG, "attention: null argument");
    B200_CUDA(h, cudaSetDevice(h->device));
    return launch_attention(h, static_cast<const bf16*>(qkv_dev), static_cast<bf16*>(out_dev), n_seq, t, heads, causal,
                            static_cast<cudaStream_t>(stream));
}
