// Host side of the tcgen05 GEMM: TMA tensor-map encoding and launch.
#include "gemm_tcgen05_2cta.cuh"
#include <stdlib.h>

#include "internal.h"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows,
// 128-byte swizzle; out-of-bounds elements read as zero (that is how M/N/K tails are handled).
int make_tmap_bf16_2d(b200clip_handle* h, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return b200_fail(h, B200CLIP_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
        return b200_fail(h, B200CLIP_E_SHAPE, "TMA operand must be 16-byte aligned with a 16-byte multiple pitch");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return b200_fail(h, B200CLIP_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

template <int BLOCK_N>
static int launch_gemm_bn(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc,
                          int M, int N, int K, const b200::GemmEpilogue& ep, cudaStream_t st) {
    using Cfg = b200::GemmCfg<BLOCK_N>;
    CUtensorMap ta, tw;
    int rc;
    if ((rc = make_tmap_bf16_2d(h, &ta, a, M, K, lda, b200::GEMM_BLOCK_M, b200::GEMM_BLOCK_K,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if ((rc = make_tmap_bf16_2d(h, &tw, w, N, K, ldw, BLOCK_N, b200::GEMM_BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    static bool attr_set = false;
    auto kern = b200::gemm_bf16_tcgen05_kernel<BLOCK_N>;
    if (!attr_set) {
        B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int m_blocks = (M + b200::GEMM_BLOCK_M - 1) / b200::GEMM_BLOCK_M;
    const int n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = m_blocks * n_blocks;
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    {
        ProfScope ps(h, PROF_GEMM, 2.0 * M * static_cast<double>(N) * K, st);
        kern<<<grid, b200::GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tw, out, ldc, M, N, K, ep);
    }
    h->launches++;
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

static int launch_gemm_2cta(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc,
                            int M, int N, int K, const b200::GemmEpilogue& ep, cudaStream_t st) {
    CUtensorMap ta, tw;
    int rc;
    if ((rc = make_tmap_bf16_2d(h, &ta, a, M, K, lda, b200::GEMM_BLOCK_M, b200::GEMM_BLOCK_K,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if ((rc = make_tmap_bf16_2d(h, &tw, w, N, K, ldw, b200::G2_HALF_N, b200::GEMM_BLOCK_K,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    // staged (TMA store) epilogue whenever the output rows are an affine image of the accumulator rows
    const int use_tma_epi = ep.t_in == 0 && ep.rowtab == nullptr && (ldc % 8 == 0) ? 1 : 0;
    CUtensorMap to, tr;
    if (use_tma_epi) {
        if ((rc = make_tmap_bf16_2d(h, &to, out, M, N, ldc, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tmap_bf16_2d(h, &tr, ep.resid ? ep.resid : out, M, N, ldc, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
    } else {
        to = ta; tr = ta;   // unused
    }
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA(h, cudaFuncSetAttribute(b200::gemm_bf16_tcgen05_2cta_kernel<0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, b200::G2_SMEM_BYTES));
        attr_set = true;
    }
    const int m_blocks = (M + 2 * b200::GEMM_BLOCK_M - 1) / (2 * b200::GEMM_BLOCK_M);
    const int n_blocks = (N + b200::G2_BLOCK_N - 1) / b200::G2_BLOCK_N;
    const int tiles = m_blocks * n_blocks;
    int clusters = h->num_sms / 2;
    if (tiles < clusters) clusters = tiles;
    static const bool probe_on = getenv("B200CLIP_GEMM_PROBE") != nullptr;   // development aid, see GemmEpilogue::probe
    b200::GemmEpilogue epp = ep;
    // epilogue warps sleep between polls of the accumulator barrier (A/B on one box: 0.5-1 % faster steps under the
    // power cap); B200CLIP_GEMM_SPIN_WAIT=1 restores the tight spin
    static const int relaxed = getenv("B200CLIP_GEMM_SPIN_WAIT") ? 0 : 1;
    epp.relaxed_wait = relaxed;
    long long* probe = nullptr;
    if (probe_on) {
        B200_CUDA(h, cudaMallocManaged(&probe, sizeof(long long) * 4 * clusters));
        B200_CUDA(h, cudaMemset(probe, 0, sizeof(long long) * 4 * clusters));
        epp.probe = probe;
    }
    {
        ProfScope ps(h, PROF_GEMM, 2.0 * M * static_cast<double>(N) * K, st);
        b200::gemm_bf16_tcgen05_2cta_kernel<0><<<2 * clusters, b200::GEMM_THREADS, b200::G2_SMEM_BYTES, st>>>(
            ta, tw, to, tr, out, ldc, M, N, K, epp, use_tma_epi);
    }
    if (probe) {
        B200_CUDA(h, cudaStreamSynchronize(st));
        double t[4] = {0, 0, 0, 0};
        for (int c = 0; c < clusters; ++c)
            for (int j = 0; j < 4; ++j) t[j] += static_cast<double>(probe[c * 4 + j]) / clusters;
        fprintf(stderr, "[gemm probe] M=%d N=%d K=%d tiles/cluster=%.1f: total %.0f clk; MMA waits: TMA data %.1f%%, free accumulator "
                        "%.1f%%; epilogue waits for accumulator %.1f%%\n", M, N, K, static_cast<double>(tiles) / clusters, t[0],
                100.0 * t[1] / t[0], 100.0 * t[2] / t[0], 100.0 * t[3] / t[0]);
        cudaFree(probe);
    }
    h->launches++;
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

int launch_gemm(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc, int M, int N,
                int K, const b200::GemmEpilogue& ep, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return b200_fail(h, B200CLIP_E_ARG, "gemm: empty problem %dx%dx%d", M, N, K);
    if (N % 32 != 0 || K % 8 != 0 || ldc % 8 != 0)
        return b200_fail(h, B200CLIP_E_SHAPE, "gemm: N must be a multiple of 32, K and ldc of 8 (N=%d K=%d ldc=%d)", N,
                         K, ldc);
    // large-M problems: CTA pairs (256x256 tiles, half the B traffic per CTA); B200CLIP_GEMM_1CTA=1 forces the
    // single-CTA kernel (used by the parity tests to cover both)
    // Row statistics are emitted per (row, column segment of BLOCK_N/2); only LN_SLOTS segments are kept.
    auto stats_fit = [&](int block_n) { return !ep.stats_out || (N + block_n / 2 - 1) / (block_n / 2) <= b200::LN_SLOTS; };
    static const bool force_1cta = getenv("B200CLIP_GEMM_1CTA") != nullptr;
    if (!force_1cta && N % 256 == 0 && M >= 2048 && stats_fit(256)) return launch_gemm_2cta(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
    // tiny M (text tower: 77 rows per query): narrow tiles so that more CTAs share the latency-bound problem
    if (M <= 256 && N % 64 == 0 && stats_fit(64)) return launch_gemm_bn<64>(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
    // (statistics-emitting GEMMs keep the 256-wide tiles at every M: the width of the partial-sum segments is part of
    // the numerics, and a query's embedding must not depend on how many queries share the batch)
    if (N % 256 == 0 || N > 1024) {
        if (!stats_fit(256)) return b200_fail(h, B200CLIP_E_SHAPE, "gemm: width %d too large for the LayerNorm statistics slots", N);
        return launch_gemm_bn<256>(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
    }
    if (!stats_fit(128)) return launch_gemm_bn<256>(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
    return launch_gemm_bn<128>(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
}
