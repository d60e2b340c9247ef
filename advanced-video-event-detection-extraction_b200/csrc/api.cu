// C ABI of libb200clip.so: handle life cycle, weight packing, workspace, tower orchestration.
#include <stdarg.h>
#include <stdio.h>
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

extern "C" const char* b200clip_version(void) { return "b200clip 0.1 (sm_100a)"; }

extern "C" const char* b200clip_last_error(const b200clip_handle* h) {
    if (h) return h->err.c_str();
    return g_err.c_str();
}

extern "C" int64_t b200clip_launch_count(const b200clip_handle* h) { return h ? h->launches : 0; }
extern "C" void b200clip_reset_launch_count(b200clip_handle* h) {
    if (h) h->launches = 0;
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
    for (void* p : h->allocs) cudaFree(p);
    cudaFree(h->ws_x); cudaFree(h->ws_y); cudaFree(h->ws_qkv); cudaFree(h->ws_h); cudaFree(h->ws_patches);
    cudaFree(h->ws_emb); cudaFree(h->ws_pre); cudaFree(h->ws_topk);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->ws_stage_dev[i]);
        if (h->ws_stage_host[i]) cudaFreeHost(h->ws_stage_host[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return 0;
}

extern "C" int b200clip_gemm_bf16(b200clip_handle* h, const void* a_dev, const void* w_dev, void* out_dev, int m,
                                  int n, int k, const float* bias_dev, const void* resid_dev, int act,
                                  void* stream) {
    if (!h || !a_dev || !w_dev || !out_dev) return b200_fail(h, B200CLIP_E_ARG, "gemm: null argument");
    if (act < 0 || act > 2) return b200_fail(h, B200CLIP_E_ARG, "gemm: act must be 0, 1 or 2");
    b200::GemmEpilogue ep{};
    ep.bias = bias_dev;
    ep.resid = static_cast<const bf16*>(resid_dev);
    ep.rowtab = nullptr;
    ep.act = act;
    ep.t_in = 0; ep.t_out = 0; ep.row_off = 0;
    return launch_gemm(h, static_cast<const bf16*>(a_dev), k, static_cast<const bf16*>(w_dev), k,
                       static_cast<bf16*>(out_dev), n, m, n, k, ep, static_cast<cudaStream_t>(stream));
}
