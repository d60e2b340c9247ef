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
    auto kern = b200::gemm_bf16_tcgen05_kernel<BLOCK_N>;
    const uint32_t abit = BLOCK_N == 64 ? ATTR_GEMM64 : BLOCK_N == 128 ? ATTR_GEMM128 : ATTR_GEMM256;
    if (!(h->attr_done & abit)) {
        B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        h->attr_done |= abit;
    }
    const int m_blocks = (M + b200::GEMM_BLOCK_M - 1) / b200::GEMM_BLOCK_M;
    const int n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = m_blocks * n_blocks;
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    {
        // (class "gemm_small": the single-CTA kernel serves M < 2048 -- the 77-row text-tower GEMMs, heads of tiny
        // batches; the roofline class "gemm" is the cta_group::2 kernel alone)
        ProfScope ps(h, PROF_GEMM_SMALL, 2.0 * M * static_cast<double>(N) * K, st);
        kern<<<grid, b200::GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tw, out, ldc, M, N, K, ep);
    }
    h->launches++;
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

// co-resident clusters of 2*pairs CTAs of the 2-CTA kernel on this handle's device (B200: GPCs of 16/18/20 SMs ->
// 74 / 33 / 15); also where the kernel gets its dynamic shared-memory opt-in (per device, hence per handle)
static int g2_max_clusters(b200clip_handle* h, int pairs) {
    if (!(h->attr_done & ATTR_GEMM_2CTA)) {
        cudaFuncSetAttribute(b200::gemm_bf16_tcgen05_2cta_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b200::G2_SMEM_BYTES);
        cudaFuncSetAttribute(b200::gemm_bf16_tcgen05_2cta_kernel<1, b200::G2_DEEP_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             b200::g2_smem_bytes(b200::G2_DEEP_STAGES));
        h->g2_clusters[1] = h->num_sms / 2;
#ifdef B200CLIP_PROBES
        cudaFuncSetAttribute(b200::gemm_bf16_tcgen05_2cta_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b200::G2_SMEM_BYTES);
        cudaFuncSetAttribute(b200::gemm_bf16_tcgen05_2cta_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, b200::G2_SMEM_BYTES);
        for (int pr = 2; pr <= 4; pr += 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * pr * (h->num_sms / (2 * pr)));
            cfg.blockDim = dim3(b200::GEMM_THREADS);
            cfg.dynamicSmemBytes = b200::G2_SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2 * pr; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = pr == 2 ? cudaOccupancyMaxActiveClusters(&h->g2_clusters[pr], b200::gemm_bf16_tcgen05_2cta_kernel<2>, &cfg)
                                    : cudaOccupancyMaxActiveClusters(&h->g2_clusters[pr], b200::gemm_bf16_tcgen05_2cta_kernel<4>, &cfg);
            if (e != cudaSuccess || h->g2_clusters[pr] <= 0) {
                cudaGetLastError();
                h->g2_clusters[pr] = h->num_sms / (2 * pr) - (pr == 2 ? 4 : 3);
            }
        }
        fprintf(stderr, "[gemm] co-resident clusters: 4-CTA %d, 8-CTA %d\n", h->g2_clusters[2], h->g2_clusters[4]);
#endif
        h->attr_done |= ATTR_GEMM_2CTA;
    }
    return h->g2_clusters[pairs];
}

// rows [r0, r0 + rows) of the problem on `pairs` CTA pairs per cluster, at most max_cl clusters, on stream s
static int launch_gemm_2cta_range(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc,
                                  int r0, int rows, int N, int K, const b200::GemmEpilogue& ep, int pairs, int max_cl,
                                  bool with_probe, cudaStream_t s) {
    CUtensorMap ta, tw;
    int rc;
    const int64_t o = r0;
    if ((rc = make_tmap_bf16_2d(h, &ta, a + o * lda, rows, K, lda, b200::GEMM_BLOCK_M, b200::GEMM_BLOCK_K,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if ((rc = make_tmap_bf16_2d(h, &tw, w, N, K, ldw, b200::G2_HALF_N / pairs, b200::GEMM_BLOCK_K,
                                CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    b200::GemmEpilogue epp = ep;
    if (r0) {   // only affine epilogues are split (checked by the caller): every per-row pointer moves with the rows
        if (epp.resid) epp.resid += o * ldc;
        if (epp.ln_stats) epp.ln_stats += o * (b200::LN_SLOTS * 2);
        if (epp.stats_out) epp.stats_out += o * (b200::LN_SLOTS * 2);
    }
    bf16* outp = out + o * ldc;
    // staged (TMA store) epilogue whenever the output rows are an affine image of the accumulator rows
    const int use_tma_epi = ep.t_in == 0 && ep.rowtab == nullptr && (ldc % 8 == 0) ? 1 : 0;
    CUtensorMap to, tr;
    if (use_tma_epi) {
        if ((rc = make_tmap_bf16_2d(h, &to, outp, rows, N, ldc, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tmap_bf16_2d(h, &tr, epp.resid ? epp.resid : outp, rows, N, ldc, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
    } else {
        to = ta; tr = ta;   // unused
    }
    const int m_blocks = (rows + 2 * pairs * b200::GEMM_BLOCK_M - 1) / (2 * pairs * b200::GEMM_BLOCK_M);
    const int n_blocks = (N + b200::G2_BLOCK_N - 1) / b200::G2_BLOCK_N;
    const int tiles = m_blocks * n_blocks;
    int clusters = max_cl;
    if (tiles < clusters) clusters = tiles;
    // epilogue warps sleep between polls of the accumulator barrier (A/B on one box: 0.5-1 % faster steps under the
    // power cap); B200CLIP_GEMM_SPIN_WAIT=1 restores the tight spin
    epp.relaxed_wait = b200_knobs().gemm_spin_wait ? 0 : 1;
    long long* probe = nullptr;
#ifdef B200CLIP_PROBES
    if (with_probe) {
        B200_CUDA(h, cudaMallocManaged(&probe, sizeof(long long) * 4 * clusters));
        B200_CUDA(h, cudaMemset(probe, 0, sizeof(long long) * 4 * clusters));
        epp.probe = probe;
    }
#else
    (void)with_probe;
#endif
    const unsigned grid = 2u * pairs * clusters;
#ifdef B200CLIP_PROBES
    if (pairs == 4)
        b200::gemm_bf16_tcgen05_2cta_kernel<4><<<grid, b200::GEMM_THREADS, b200::G2_SMEM_BYTES, s>>>(
            ta, tw, to, tr, outp, ldc, rows, N, K, epp, use_tma_epi);
    else if (pairs == 2)
        b200::gemm_bf16_tcgen05_2cta_kernel<2><<<grid, b200::GEMM_THREADS, b200::G2_SMEM_BYTES, s>>>(
            ta, tw, to, tr, outp, ldc, rows, N, K, epp, use_tma_epi);
    else
#endif
    // GEMMs without a residual (patch embedding, qkv, fc) run the deep-ring form: six operand stages, one staging box
    // per epilogue warp (B200CLIP_GEMM_5STAGE=1: the five-stage / two-box form for every shape)
    if (!epp.resid && !b200_knobs().gemm_5stage)
        b200::gemm_bf16_tcgen05_2cta_kernel<1, b200::G2_DEEP_STAGES><<<grid, b200::GEMM_THREADS, b200::g2_smem_bytes(b200::G2_DEEP_STAGES), s>>>(
            ta, tw, to, tr, outp, ldc, rows, N, K, epp, use_tma_epi);
    else
        b200::gemm_bf16_tcgen05_2cta_kernel<1><<<grid, b200::GEMM_THREADS, b200::G2_SMEM_BYTES, s>>>(
            ta, tw, to, tr, outp, ldc, rows, N, K, epp, use_tma_epi);
    if (probe) {
        B200_CUDA(h, cudaStreamSynchronize(s));
        double t[4] = {0, 0, 0, 0};
        for (int c = 0; c < clusters; ++c)
            for (int j = 0; j < 4; ++j) t[j] += static_cast<double>(probe[c * 4 + j]) / clusters;
        fprintf(stderr, "[gemm probe] M=%d N=%d K=%d tiles/cluster=%.1f: total %.0f clk; MMA waits: TMA data %.1f%%, free accumulator "
                        "%.1f%%; epilogue waits for accumulator %.1f%%\n", rows, N, K, static_cast<double>(tiles) / clusters, t[0],
                100.0 * t[1] / t[0], 100.0 * t[2] / t[0], 100.0 * t[3] / t[0]);
        cudaFree(probe);
    }
    h->launches++;
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

static int launch_gemm_2cta(b200clip_handle* h, const bf16* a, int lda, const bf16* w, int ldw, bf16* out, int ldc,
                            int M, int N, int K, const b200::GemmEpilogue& ep, cudaStream_t st) {
    int pairs = 1;
    bool probe_on = false;
#ifdef B200CLIP_PROBES
    // Launch forms that were measured and lost inside the power-capped step (profiles/r01aw_*); probe builds only.
    // B200CLIP_GEMM_PAIRS=2|4: clusters of two / four CTA pairs that share their B tile by TMA multicast (per SM 4 % /
    // 11-14 % faster, but 4- / 8-CTA clusters only fit 132 / 120 of the 148 SMs).
    static const int pairs_env = getenv("B200CLIP_GEMM_PAIRS") ? atoi(getenv("B200CLIP_GEMM_PAIRS")) : 1;
    // B200CLIP_GEMM_HYBRID=<per mille of the rows>: 8-CTA multicast clusters take that share of the rows and, on a
    // second low-priority stream, plain CTA pairs work on the rest on the SMs the big clusters cannot use.
    static const int hybrid = getenv("B200CLIP_GEMM_HYBRID") ? atoi(getenv("B200CLIP_GEMM_HYBRID")) : 0;
    static const bool probe_env = getenv("B200CLIP_GEMM_PROBE") != nullptr;   // clock64 waits, see GemmEpilogue::probe
    probe_on = probe_env;
    const bool affine = ep.t_in == 0 && ep.rowtab == nullptr;
    int rc;
    if (hybrid > 0 && hybrid < 1000 && affine && M >= 32768) {
        const int big = g2_max_clusters(h, 4);
        const int small = (h->num_sms - 8 * big) / 2;
        int m1 = static_cast<int>(static_cast<int64_t>(M) * hybrid / 1000) / 1024 * 1024;
        if (small > 0 && m1 > 0 && m1 < M) {
            if (!h->gemm_side) {
                int lo = 0, hi = 0;
                B200_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
                B200_CUDA(h, cudaStreamCreateWithPriority(&h->gemm_side, cudaStreamNonBlocking, lo));
                B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_gemm_fork, cudaEventDisableTiming));
                B200_CUDA(h, cudaEventCreateWithFlags(&h->ev_gemm_join, cudaEventDisableTiming));
            }
            ProfScope ps(h, PROF_GEMM, 2.0 * M * static_cast<double>(N) * K, st);
            B200_CUDA(h, cudaEventRecord(h->ev_gemm_fork, st));
            static const int only = getenv("B200CLIP_GEMM_HYBRID_ONLY") ? atoi(getenv("B200CLIP_GEMM_HYBRID_ONLY")) : 0;  // probe: 1 main, 2 side
            if (only != 2 && (rc = launch_gemm_2cta_range(h, a, lda, w, ldw, out, ldc, 0, m1, N, K, ep, 4, big, false, st))) return rc;
            B200_CUDA(h, cudaStreamWaitEvent(h->gemm_side, h->ev_gemm_fork, 0));
            if (only != 1 && (rc = launch_gemm_2cta_range(h, a, lda, w, ldw, out, ldc, m1, M - m1, N, K, ep, 1, small, false, h->gemm_side)))
                return rc;
            B200_CUDA(h, cudaEventRecord(h->ev_gemm_join, h->gemm_side));
            B200_CUDA(h, cudaStreamWaitEvent(st, h->ev_gemm_join, 0));
            h->launches--;   // one logical GEMM
            return 0;
        }
    }
    pairs = (pairs_env == 2 || pairs_env == 4) && M >= 2 * pairs_env * b200::GEMM_BLOCK_M * 8 ? pairs_env : 1;
#endif
    ProfScope ps(h, PROF_GEMM, 2.0 * M * static_cast<double>(N) * K, st);
    return launch_gemm_2cta_range(h, a, lda, w, ldw, out, ldc, 0, M, N, K, ep, pairs, g2_max_clusters(h, pairs), probe_on, st);
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
    if (!b200_knobs().gemm_1cta && N % 256 == 0 && M >= 2048 && stats_fit(256)) return launch_gemm_2cta(h, a, lda, w, ldw, out, ldc, M, N, K, ep, st);
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
