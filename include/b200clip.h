/*
 * b200clip.h -- C ABI of libb200clip.so: the B200 (sm_100a) implementation of the phase1_mvp
 * natural-language query hot path of nb-hmd/Advanced-Video-Event-Detection-Extraction.
 *
 * The reference is pure Python; what a maintainer would bind is the object protocol its wrapper uses
 * on the third-party `open_clip` model (see INTEGRATION.md for the ctypes stub).  Every entry point
 * below names the reference interface (file:line under the reference tree) it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; `*_dev` pointers are CUDA device pointers on the handle's device,
 *     `*_host` pointers are host memory (pageable or pinned).
 *   - every function returns 0 on success or a negative B200CLIP_E_* code; b200clip_last_error()
 *     returns a human-readable message for the last failure on that handle (or globally when h == NULL).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous
 *     on that stream unless they take/return host buffers, in which case they return after the result
 *     is in the host buffer.
 *   - a handle is not re-entrant: one in-flight call per handle.  One handle per GPU / rank.
 *   - there is no CPU fallback anywhere: on a device that is not sm_100 every call fails with
 *     B200CLIP_E_ARCH.
 */
#ifndef B200CLIP_H_
#define B200CLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CLIP_OK 0
#define B200CLIP_E_ARG (-1)    /* null pointer / bad enum / bad size */
#define B200CLIP_E_SHAPE (-2)  /* unsupported shape (e.g. width not a multiple of 64) */
#define B200CLIP_E_CUDA (-3)   /* a CUDA runtime / driver call failed */
#define B200CLIP_E_ARCH (-4)   /* device is not compute capability 10.x */
#define B200CLIP_E_STATE (-5)  /* weights missing / handle not finalized */
#define B200CLIP_E_NOMEM (-6)
#define B200CLIP_E_NCCL (-7)   /* NCCL could not be bound at run time, or a collective failed */

typedef struct b200clip_handle b200clip_handle;

/* Model geometry == open_clip model_configs/{ViT-B-32,ViT-L-14}.json. */
typedef struct b200clip_config {
    int32_t image_size;  /* 224 */
    int32_t patch;       /* 32 (B/32) or 14 (L/14) */
    int32_t width;       /* 768 / 1024 */
    int32_t layers;      /* 12 / 24 */
    int32_t heads;       /* width / 64 */
    int32_t mlp_dim;     /* 4 * width */
    int32_t embed_dim;   /* 512 / 768 */
    int32_t act;         /* 0 = QuickGELU (pretrained="openai", src/utils/config.py:25-26), 1 = erf GELU */
    float ln_eps;        /* 1e-5 */
    int32_t text_ctx;    /* 77 */
    int32_t text_vocab;  /* 49408 */
    int32_t text_width;  /* 512 / 768 */
    int32_t text_heads;  /* 8 / 12 */
    int32_t text_layers; /* 12 */
    int32_t text_mlp_dim;
} b200clip_config;

/* resize modes of the frame preprocess kernel */
#define B200CLIP_RESIZE_REFERENCE 0   /* decoded frame -> bit-exact reference chain: cv2 INTER_AREA shrink to fit
                                         512x512 (FrameExtractor) -> Pillow bicubic(aa) Resize(S) -> CenterCrop */
#define B200CLIP_RESIZE_BILINEAR_AA 1 /* Pillow antialiased BILINEAR Resize(S) -> CenterCrop straight from the frame
                                         (no 512 shrink, cheaper taps; not the reference's filter) */
#define B200CLIP_RESIZE_BICUBIC 2     /* open_clip's transform only (what `preprocess(PIL)` does): Pillow bicubic(aa)
                                         Resize(S) -> CenterCrop, no 512 shrink */

/* OR-ed into `resize_mode` of the uint8 RGB entry points (preprocess_u8, preprocess_u8_chw, encode_frames_u8,
   encode_frames_u8_host): the frames are in OpenCV's BGR order, as cv2.VideoCapture delivers them.  Replaces the
   per-frame cv2.cvtColor(frame, cv2.COLOR_BGR2RGB) of /root/reference/src/services/frame_extractor.py:191: every resize
   stage works per channel, so the swap is applied in K1's final store (normalisation constants and output plane of
   channel 2 - c); results are bit-identical to converting first.  Not accepted by the NV12 entry points. */
#define B200CLIP_INPUT_BGR 0x100

/* element types for embedding buffers */
#define B200CLIP_F32 0
#define B200CLIP_BF16 1

const char* b200clip_version(void);
const char* b200clip_last_error(const b200clip_handle* h);

/* ---- life cycle.  Replaces OpenCLIPModel.__init__/load_model (src/models/openclip_model.py:14-150),
 *      i.e. open_clip.create_model_and_transforms(...) + model.eval(). */
int b200clip_create(const b200clip_config* cfg, int device, b200clip_handle** out);
int b200clip_destroy(b200clip_handle* h);
/* One tensor of the open_clip state dict (key names of SURVEY.md Appendix A, e.g.
 * "visual.transformer.resblocks.3.attn.in_proj_weight"), fp32, host memory, row-major.  The library
 * converts to its own bf16 / fp32 device layouts; the caller keeps ownership of `data_host`. */
int b200clip_set_weight(b200clip_handle* h, const char* name, const float* data_host, const int64_t* shape,
                        int ndim);
/* Verifies that every tensor the config needs was supplied; after this the encode calls are legal. */
int b200clip_finalize(b200clip_handle* h);
/* Pre-sizes the persistent activation workspace (otherwise grown on first use). */
int b200clip_reserve(b200clip_handle* h, int max_images, int max_texts);

/* ---- K1: frame preprocess.  Replaces MemoryManager.resize_frame_for_memory
 *      (src/utils/memory_manager.py:299-322) + open_clip image_transform (PIL Resize(bicubic) /
 *      CenterCrop / ToTensor / Normalize; call sites src/models/openclip_model.py:165-174,188-193).
 *      frames_dev: uint8 RGB HWC, frame i at frames_dev + i*frame_stride, rows row_stride bytes apart.
 *      patches_out_dev: bf16 [n * grid^2, patch_k] patch-major rows, col = c*P*P + y*P + x
 *      (patch_k = 3*P*P rounded up to a multiple of 64, zero padded). */
int b200clip_preprocess_u8(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                           int64_t frame_stride, int64_t row_stride, int resize_mode, void* patches_out_dev,
                           void* stream);
/* The normalised image itself, fp32 [n,3,S,S] (what open_clip's `preprocess` returns, stacked). */
int b200clip_preprocess_u8_chw(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                               int64_t frame_stride, int64_t row_stride, int resize_mode, float* chw_out_dev,
                               void* stream);

/* ---- K2+K3: image tower.  Replaces model.encode_image(x[B,3,S,S]) (+ the L2 normalisation at
 *      src/models/openclip_model.py:177-178,196-197 when l2norm != 0).
 *      emb_out_dev: [n, embed_dim] of out_dtype (B200CLIP_F32 / B200CLIP_BF16). */
int b200clip_encode_patches(b200clip_handle* h, const void* patches_dev, int n, void* emb_out_dev, int out_dtype,
                            int l2norm, void* stream);
int b200clip_encode_image_chw(b200clip_handle* h, const float* chw_dev, int n, void* emb_out_dev, int out_dtype,
                              int l2norm, void* stream);
/* Fused K1 -> K3 on device-resident frames. */
int b200clip_encode_frames_u8(b200clip_handle* h, const uint8_t* frames_dev, int n, int height, int width,
                              int64_t frame_stride, int64_t row_stride, int resize_mode, void* emb_out_dev,
                              int out_dtype, int l2norm, void* stream);
/* Reference-facing call: OpenCLIPModel.encode_images(np.ndarray[N,H,W,3] uint8) -> float32[N,E]
 * (src/models/openclip_model.py:152-198).  HOST buffers in and out: frames are staged through pinned
 * double buffers (H2D overlapped with compute), embeddings are copied back; returns when emb_out_host is
 * complete.  emb_out_host may also be a DEVICE pointer: the embeddings then stay resident (no D2H), the call
 * returns once all uploads are done and the compute remains asynchronous on `stream`. */
int b200clip_encode_frames_u8_host(b200clip_handle* h, const uint8_t* frames_host, int n, int height, int width,
                                   int resize_mode, float* emb_out_host, int l2norm, void* stream);

/* ---- frame feed: 4:2:0 frames as a video decoder leaves them, converted inside K1.
 *      Replaces the CPU colour conversion the reference's decoders run before handing RGB to Python
 *      (src/services/frame_extractor.py:38-235: decord.VideoReader / cv2.VideoCapture + cvtColor :191) and, like the
 *      *_u8 calls, MemoryManager.resize_frame_for_memory + the open_clip transform.  Pixels are bit-identical to
 *      cv2.cvtColor(frame, cv2.COLOR_YUV2RGB_NV12) followed by the *_u8 chain; RGB never exists in HBM on the fused path
 *      (1080p / 720p class geometries with 16-byte aligned planes; everything else converts the crop window first).
 *      NV12: y_dev = luma plane [height rows], uv_dev = interleaved chroma plane [height/2 rows], both with
 *      row_stride bytes per row (NVDEC: uv_dev = y_dev + row_stride * aligned_height); frame i at
 *      y_dev + i*y_frame_stride / uv_dev + i*uv_frame_stride.  width and height must be even. */
int b200clip_preprocess_nv12(b200clip_handle* h, const uint8_t* y_dev, const uint8_t* uv_dev, int n, int height,
                             int width, int64_t y_frame_stride, int64_t uv_frame_stride, int64_t row_stride,
                             int resize_mode, void* patches_out_dev /* bf16 patch rows or NULL */,
                             float* chw_out_dev /* fp32 [n,3,S,S] or NULL */, void* stream);
int b200clip_encode_frames_nv12(b200clip_handle* h, const uint8_t* y_dev, const uint8_t* uv_dev, int n, int height,
                                int width, int64_t y_frame_stride, int64_t uv_frame_stride, int64_t row_stride,
                                int resize_mode, void* emb_out_dev, int out_dtype, int l2norm, void* stream);
/* HOST NV12 frames (frame i at nv12_host + i*height*width*3/2: Y plane then UV plane, the layout cv2 takes for
 * COLOR_YUV2RGB_NV12) -> embeddings; same staging, window upload and emb_out rules as
 * b200clip_encode_frames_u8_host, at half the PCIe bytes. */
int b200clip_encode_frames_nv12_host(b200clip_handle* h, const uint8_t* nv12_host, int n, int height, int width,
                                     int resize_mode, float* emb_out_host, int l2norm, void* stream);

/* ---- text tower.  Replaces model.encode_text(tokens[Q,77]) (+ L2 norm, openclip_model.py:200-210). */
int b200clip_encode_text(b200clip_handle* h, const int64_t* tokens_dev, int q, float* emb_out_dev, int l2norm,
                         void* stream);
int b200clip_encode_text_host(b200clip_handle* h, const int64_t* tokens_host, int q, float* emb_out_host,
                              int l2norm, void* stream);

/* ---- K4: similarity + top-k + threshold + clip intervals.  Replaces
 *      OpenCLIPModel.compute_similarity (np.dot, openclip_model.py:212-214),
 *      np.argsort(s)[::-1][:top_k] + threshold (src/pipeline/phase1_mvp.py:145-155; ties -> higher index
 *      first) and ClipExtractor.extract_clip_with_padding / extract_clip interval arithmetic
 *      (src/services/clip_extractor.py:175-183, 94-111).
 *      img_emb_dev [n,e] (emb_dtype), txt_emb_dev fp32 [q,e].
 *      timestamps_dev: float64 [index_base + n] indexed by GLOBAL row index, or NULL (then timestamp =
 *      global index).  index_base is added to every row index (global index of this shard's first row).
 *      video_duration <= 0 means unknown.  Timestamps and intervals are float64 because the reference
 *      computes them with Python floats.
 *      Outputs (device): top_scores fp32 [q,k] and top_idx int64 [q,k] = the k best rows in descending
 *      order BEFORE thresholding (-1 marks an empty slot when n < k); counts int32 [q] = length of the
 *      prefix with score >= threshold (the reference's result list); intervals float64 [q,k,2]
 *      (start,end seconds) for every non-empty slot.  intervals_dev / counts_dev may be NULL.
 *      Kernel selection: a bf16 cache with q >= 8, k <= 8, n >= 4096 and e % 64 == 0 runs the tcgen05 similarity GEMM
 *      with the top-k fused into its epilogue (text embedding rounded to bf16); everything else runs the HBM-streaming
 *      kernel (fp32 text embedding). */
int b200clip_sim_topk(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                      const float* txt_emb_dev, int q, int k, float threshold, const double* timestamps_dev,
                      int64_t index_base, double clip_duration, double video_duration, float* top_scores_dev,
                      int64_t* top_idx_dev, double* intervals_dev, int32_t* counts_dev, void* stream);
/* Test / debugging form of b200clip_sim_topk: additionally writes the fp32 score matrix [n,q] exactly as the selected
 * kernel computed it (the tensor-core path rounds the text embedding to bf16), so that the top-k order can be checked
 * bit-exactly against an argsort of the same scores. */
int b200clip_sim_topk_dense(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                            const float* txt_emb_dev, int q, int k, float threshold, float* top_scores_dev,
                            int64_t* top_idx_dev, int32_t* counts_dev, float* dense_scores_dev, void* stream);
/* Dense scores fp32 [n,q] (compute_similarity itself). */
int b200clip_similarity(b200clip_handle* h, const void* img_emb_dev, int emb_dtype, int64_t n, int e,
                        const float* txt_emb_dev, int q, float* scores_out_dev, void* stream);
/* k-way merge of g sorted candidate lists (e.g. after an all-gather over ranks): cand_scores fp32 [g,q,k],
 * cand_idx int64 [g,q,k] global indices (-1 = empty).  Same outputs as b200clip_sim_topk. */
int b200clip_topk_merge(b200clip_handle* h, const float* cand_scores_dev, const int64_t* cand_idx_dev, int g, int q,
                        int k, float threshold, const double* timestamps_dev, double clip_duration,
                        double video_duration, float* top_scores_dev, int64_t* top_idx_dev, double* intervals_dev,
                        int32_t* counts_dev, void* stream);

/* ---- multi-GPU (the reference is single-device).  Rows (frames or cached embeddings) are
 *      block-partitioned over ranks, weights replicated, no communication while embedding.  The only exchange is the
 *      per-rank candidate list: ONE in-place ncclAllGather of B200CLIP message bytes on `stream`, then the same
 *      deterministic merge on every rank (result == the single-GPU top-k, bit for bit).  `nccl_comm` is an ncclComm_t
 *      (void*) created by the host with the libnccl.so.2 mapped in the process (PyTorch:
 *      ProcessGroupNCCL._comm_ptr()); NCCL is bound with dlopen at first use, failures return B200CLIP_E_NCCL.
 *      Message of one rank, b200clip_topk_msg_bytes(q, k) bytes: int64 idx[q][k] (global, -1 = empty) |
 *      float score[q][k] | zero padding to 16 bytes. */
int64_t b200clip_topk_msg_bytes(int q, int k);
/* Shard-local K4 + exchange + merge in one call: img_emb_dev holds this rank's n_local rows, index_base = global index
 * of its first row; outputs as b200clip_sim_topk, identical on every rank. */
int b200clip_sim_topk_nccl(b200clip_handle* h, void* nccl_comm, int rank, int world, const void* img_emb_dev,
                           int emb_dtype, int64_t n_local, int e, const float* txt_emb_dev, int q, int k,
                           float threshold, const double* timestamps_dev, int64_t index_base, double clip_duration,
                           double video_duration, float* top_scores_dev, int64_t* top_idx_dev, double* intervals_dev,
                           int32_t* counts_dev, void* stream);
/* Exchange + merge of candidates the caller already holds (local_scores_dev fp32 [q,k], local_idx_dev int64 [q,k],
 * global indices). */
int b200clip_topk_merge_nccl(b200clip_handle* h, void* nccl_comm, int rank, int world, const float* local_scores_dev,
                             const int64_t* local_idx_dev, int q, int k, float threshold, const double* timestamps_dev,
                             double clip_duration, double video_duration, float* top_scores_dev, int64_t* top_idx_dev,
                             double* intervals_dev, int32_t* counts_dev, void* stream);
/* Merge of g gathered messages (layout above, message l at gathered_dev + l * msg_bytes) that some other transport
 * delivered (e.g. torch.distributed.all_gather_into_tensor on gloo in the CPU-side tests of the host logic). */
int b200clip_topk_merge_packed(b200clip_handle* h, const void* gathered_dev, int g, int q, int k, float threshold,
                               const double* timestamps_dev, double clip_duration, double video_duration,
                               float* top_scores_dev, int64_t* top_idx_dev, double* intervals_dev, int32_t* counts_dev,
                               void* stream);

/* ---- building blocks, exported for the parity tests and micro-benchmarks ---- */
/* out[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ resid); bf16 in/out, fp32 accumulate. act: 0 none,
 * 1 QuickGELU, 2 erf GELU.  bias fp32 [N] or NULL, resid bf16 [M,N] or NULL (may equal out). */
int b200clip_gemm_bf16(b200clip_handle* h, const void* a_dev, const void* w_dev, void* out_dev, int m, int n, int k,
                       const float* bias_dev, const void* resid_dev, int act, void* stream);
/* rows of `width` bf16 -> bf16, fp32 statistics. */
int b200clip_layernorm_bf16(b200clip_handle* h, const void* x_dev, const float* gamma_dev, const float* beta_dev,
                            void* y_dev, int64_t rows, int width, float eps, void* stream);
/* qkv bf16 [n_seq*t, 3*heads*64] -> out bf16 [n_seq*t, heads*64]; softmax(q k^T / 8) v per (seq, head). */
int b200clip_attention_bf16(b200clip_handle* h, const void* qkv_dev, void* out_dev, int n_seq, int t, int heads,
                            int causal, void* stream);

/* Opt-in per-kernel-class device timing: while enabled, every launch is bracketed by a CUDA event pair on the
 * launching stream.  profile_read synchronises the device and returns, for one class, the summed elapsed ms, the
 * summed algorithmic work (FLOP for class 0 = the cta_group::2 GEMM and 11 = the single-CTA GEMM, M < 2048; bytes for 1 = attention, 2 = LayerNorm, 3 = K1 preprocess chain,
 * 4 = head, 5 = K4 similarity/top-k, 6 = misc, 7/8/9 = the K1 stages area / horizontal / vertical, which are also
 * inside 3, 10 = the NCCL all-gather of the multi-GPU exchange: bytes gathered) and the number of timed launches. */
int b200clip_profile_enable(b200clip_handle* h, int on);
int b200clip_profile_read(b200clip_handle* h, int kernel_class, double* ms_out, double* work_out,
                          int64_t* launches_out, int reset);

/* Counts kernel launches issued through this handle since the last reset (bench.py's gpu_launches). */
int64_t b200clip_launch_count(const b200clip_handle* h);
void b200clip_reset_launch_count(b200clip_handle* h);

/* Bytes moved over PCIe by the *_host entry points since the last reset (bench.py's e2e h2d/d2h_bytes_per_step).
 * b200clip_encode_frames_u8_host uploads only the window of each frame that the transform's centre crop keeps
 * (reference: the whole frame travels through PIL, src/models/openclip_model.py:165-174), so this is less than
 * n*H*W*3. */
int b200clip_transfer_bytes(b200clip_handle* h, int64_t* h2d_out, int64_t* d2h_out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* B200CLIP_H_ */
