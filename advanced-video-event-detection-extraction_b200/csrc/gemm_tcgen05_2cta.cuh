// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on one TPC computes a
// 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 weight
// rows), so per output FLOP only 2/3 of the shared-memory fill traffic of the 1-CTA 128x256 kernel crosses
// L2 -> SM (128 vs 85 FLOP per byte) -- the 1-CTA kernel is L2-bandwidth bound on the K=768 ViT GEMMs.
//
//   rank 0 (leader): TMA producer, the single MMA-issuing thread (UMMA 256x256x16, cta_group::2), epilogue
//   rank 1         : TMA producer (signals the leader's full barrier), epilogue for its own 128 rows
// Barrier plumbing: full[s] lives in the leader (expects both CTAs' bytes); empty[s] and tmem_full[a] exist in
// both CTAs and are signalled by multicast tcgen05.commit; tmem_empty[a] lives in the leader and collects the
// epilogue warps of BOTH CTAs (remote mbarrier.arrive through mapa).
#pragma once
#include "gemm_tcgen05.cuh"

namespace b200 {

constexpr int G2_BLOCK_N = 256;
constexpr int G2_HALF_N = 128;
constexpr int G2_STAGES = 5;
constexpr int G2_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;   // 16 KB: this CTA's 128 rows of A
constexpr int G2_B_BYTES = G2_HALF_N * GEMM_BLOCK_K * 2;      // 16 KB: this CTA's half of B
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
// epilogue staging: every epilogue warp owns two [32 rows x 64 cols] bf16 boxes (128-byte rows, SWIZZLE_128B)
constexpr int G2_BOX_BYTES = 32 * 128;
constexpr int G2_STAGING_BYTES = GEMM_EPI_WARPS * 2 * G2_BOX_BYTES;   // 64 KB
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_STAGING_BYTES + 512 + 1024;
// Deep-ring form (STAGES = 6): the sixth 32 KB operand stage is paid for by halving the epilogue staging -- ONE box per
// epilogue warp, reused for the warp's second 64 columns once the bulk store of the first has read it (the wait sits
// behind the second half's epilogue math, so it is normally free).  No residual in this form (the residual tile of the
// second half could only be requested after that wait, and a measured residual variant gained nothing on proj while
// its extra code slowed the GELU epilogue of fc): the launcher uses it for GEMMs without a residual.
constexpr int G2_DEEP_STAGES = 6;
constexpr int g2_boxes(int stages) { return stages >= G2_DEEP_STAGES ? 1 : 2; }
constexpr int g2_smem_bytes(int stages) { return stages * G2_STAGE_BYTES + GEMM_EPI_WARPS * g2_boxes(stages) * G2_BOX_BYTES + 512 + 1024; }
static_assert(g2_smem_bytes(G2_STAGES) == G2_SMEM_BYTES && g2_smem_bytes(G2_DEEP_STAGES) <= 227 * 1024, "shared-memory budget");
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster barrier address

// 32 accumulator columns of one row -> bf16 into a [32 rows][64 cols] SWIZZLE_128B box (row = lane).
// chunk0 = index of the first 16-byte chunk of the row these 32 columns occupy (0 or 4).  If has_res, the box
// already holds the residual tile (TMA-loaded) and it is added in place.
__device__ __forceinline__ void epilogue_chunk_smem(uint32_t (&acc)[32], int n0, int N, const GemmEpilogue& ep,
                                                    uint8_t* box, int lane, int chunk0, bool has_res, float mean,
                                                    float rstd, f32x2_t& ssum, f32x2_t& ssq, bool wait_read = false) {
    if (n0 >= N) return;
    f32x2_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = f2_pack(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
    epilogue_math(v, n0, ep, nullptr, mean, rstd);
    if (wait_read) {   // the box is being reused: the bulk store issued from it must have read it
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
    }
    uint8_t* rowp = box + lane * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4* slot = reinterpret_cast<uint4*>(rowp + (((chunk0 + j) ^ (lane & 7)) << 4));
        if (has_res) {
            const uint4 r = *slot;
            v[j * 4] = f2_add(v[j * 4], bf16x2_to_f2(r.x));
            v[j * 4 + 1] = f2_add(v[j * 4 + 1], bf16x2_to_f2(r.y));
            v[j * 4 + 2] = f2_add(v[j * 4 + 2], bf16x2_to_f2(r.z));
            v[j * 4 + 3] = f2_add(v[j * 4 + 3], bf16x2_to_f2(r.w));
        }
        if (ep.stats_out) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { ssum = f2_add(ssum, v[j * 4 + e]); ssq = f2_fma(v[j * 4 + e], v[j * 4 + e], ssq); }
        }
        uint4 o;
        o.x = f2_to_bf16x2(v[j * 4]);
        o.y = f2_to_bf16x2(v[j * 4 + 1]);
        o.z = f2_to_bf16x2(v[j * 4 + 2]);
        o.w = f2_to_bf16x2(v[j * 4 + 3]);
        *slot = o;
    }
}

// PAIRS = 1: a cluster is one CTA pair (the kernel described above).
// PAIRS = 2 / 4: a cluster is two / four pairs working on M-adjacent 256 x 256 tiles that share their B tile.  Each
// CTA loads its own 128 rows of A and only 1/PAIRS of its B half (64 / 32 weight rows), TMA-multicast to the CTA
// holding the same B half in every pair: 24 / 20 KB instead of 32 KB of L2 reads per CTA and k-block.  Extra plumbing:
// a stage is written by PAIRS CTAs, so empty[s] collects the MMA commits of ALL pairs (multicast to every CTA);
// tmem_full / tmem_empty stay inside a pair.  Multicast completions land on the full barrier of each destination
// pair's leader (cta_group::2 address with the peer bit cleared, as in CUTLASS' SM100_TMA_2SM_LOAD_MULTICAST).
template <int PAIRS, int STAGES = G2_STAGES>
__global__ void __cluster_dims__(2 * PAIRS, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                              const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                              __nv_bfloat16* out, int ldc, int M, int N, int K, GemmEpilogue ep, int use_tma_epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic keeps the shared address space: LDS/STS, not generic LD/ST */;
    constexpr int BOXES = g2_boxes(STAGES);              // staging boxes per epilogue warp
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * G2_A_BYTES;
    uint8_t* staging = smem + STAGES * G2_STAGE_BYTES;   // 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + GEMM_EPI_WARPS * BOXES * G2_BOX_BYTES);
    uint64_t* full_bar = bars;                           // [STAGES]  (used in the leader)
    uint64_t* empty_bar = bars + STAGES;                 // [STAGES]  (both CTAs)
    uint64_t* tmem_full_bar = bars + 2 * STAGES;         // [2]       (both CTAs)
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;    // [2]       (used in the leader)
    uint64_t* res_bar = bars + 2 * STAGES + 4;            // [GEMM_EPI_WARPS] residual-tile loads, one per warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + GEMM_EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;             // position inside the CTA pair
    const uint32_t pair = crank >> 1;             // which pair of the cluster (0 when PAIRS == 1)
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / (2 * PAIRS);
    const int num_clusters = gridDim.x / (2 * PAIRS);

    // a cluster walks "super tiles" of PAIRS M-adjacent 256-row blocks x one 256-column block
    const int m_blocks = (M + 2 * PAIRS * GEMM_BLOCK_M - 1) / (2 * PAIRS * GEMM_BLOCK_M);
    const int n_blocks = (N + G2_BLOCK_N - 1) / G2_BLOCK_N;
    const int num_tiles = m_blocks * n_blocks;
    const int k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIRS);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 2 * GEMM_EPI_WARPS);  // epilogue warps of both CTAs
        }
        for (int s = 0; s < GEMM_EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
        tma_prefetch_desc(&tmap_out);
        tma_prefetch_desc(&tmap_res);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();   // barriers of BOTH CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_sup = tile / n_blocks;
                const int n_blk = tile - m_sup * n_blocks;
                const int m_blk = m_sup * PAIRS + static_cast<int>(pair);
                const int m0 = m_blk * 2 * GEMM_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M;
                const int n0 = n_blk * G2_BLOCK_N + static_cast<int>(rank) * G2_HALF_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 11);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
                    const uint32_t bar = smem_u32(&full_bar[stage]) & kPeerBitMask;  // always the leader's barrier
                    tma_load_2d_cg2(smem_a + stage * G2_A_BYTES, &tmap_a, bar, kb * GEMM_BLOCK_K, m0);
                    if (PAIRS > 1) {
                        // this CTA's 1/PAIRS of the shared B half (64 or 32 weight rows), delivered to the CTA holding
                        // the same half in every pair of the cluster
                        constexpr uint32_t kAllPairs = PAIRS == 2 ? 0x5u : 0x55u;    // rank 0 of every pair
                        tma_load_2d_cg2_mc(smem_b + stage * G2_B_BYTES + pair * (G2_B_BYTES / PAIRS), &tmap_w, bar,
                                           kb * GEMM_BLOCK_K, n0 + static_cast<int>(pair) * (G2_HALF_N / PAIRS),
                                           static_cast<uint16_t>(kAllPairs << rank));
                    } else
                    tma_load_2d_cg2(smem_b + stage * G2_B_BYTES, &tmap_w, bar, kb * GEMM_BLOCK_K, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, G2_BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            long long p_full = 0, p_tmem = 0;
            const long long p_t0 = clock64();
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                long long c0 = ep.probe ? clock64() : 0;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 12);
                if (ep.probe) p_tmem += clock64() - c0;
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    c0 = ep.probe ? clock64() : 0;
                    mbar_wait(&full_bar[stage], phase, 13);
                    if (ep.probe) p_full += clock64() - c0;
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * G2_A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * G2_B_BYTES));
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
                        umma_bf16<2>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_cg2(&empty_bar[stage], static_cast<uint16_t>((1u << (2 * PAIRS)) - 1u));   // frees the stage in every CTA that feeds it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_cg2(&tmem_full_bar[as], static_cast<uint16_t>(0b11u << (2 * pair)));   // accumulator ready in both CTAs of the pair
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            if (ep.probe && pair == 0) {
                ep.probe[cluster_id * 4 + 0] = clock64() - p_t0;
                ep.probe[cluster_id * 4 + 1] = p_full;
                ep.probe[cluster_id * 4 + 2] = p_tmem;
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs, own 128 rows) =====================
        long long p_epi = 0;
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = G2_BLOCK_N / (GEMM_EPI_WARPS / 4);
        constexpr int NCH = COLS_PER_WARP / 32;
        int as = 0;
        uint32_t aphase = 0;
        uint32_t res_phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int m_sup = tile / n_blocks;
            const int n_blk = tile - m_sup * n_blocks;
            const int m_blk = m_sup * PAIRS + static_cast<int>(pair);
            const int row = m_blk * 2 * GEMM_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M + q * 32 + lane;
            const bool row_ok = row < M;
            long out_row = row;
            int tpos = 0;
            if (ep.t_in > 0) {
                const int img = row / ep.t_in;
                tpos = row - img * ep.t_in + ep.row_off;
                out_row = static_cast<long>(img) * ep.t_out + tpos;
            }
            __nv_bfloat16* out_ptr = out + out_row * ldc;
            const __nv_bfloat16* res_ptr = ep.resid ? ep.resid + out_row * ldc : nullptr;
            const float* tab_ptr = ep.rowtab ? ep.rowtab + static_cast<long>(tpos) * N : nullptr;
            const int col0 = n_blk * G2_BLOCK_N + half * COLS_PER_WARP;
            float mean, rstd;
            f32x2_t ssum = f2_pack(0.f, 0.f), ssq = f2_pack(0.f, 0.f);
            ln_row_stats(ep, out_row, row_ok, mean, rstd);

            if (use_tma_epi) {
                // ---- staged epilogue: TMEM -> registers -> swizzled smem box -> TMA bulk store (full 128-byte
                // lines); the residual tile is TMA-loaded into the same box while the MMAs of this tile still run.
                uint8_t* my_stage = staging + (warp - 2) * BOXES * G2_BOX_BYTES;
                uint64_t* my_bar = &res_bar[warp - 2];
                const int box_row0 = m_blk * 2 * GEMM_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M + q * 32;
                if (lane == 0) tma_store_wait_read<0>();   // previous tile's stores no longer read the boxes
                __syncwarp();
                if (BOXES == 2 && res_ptr && lane == 0) {
                    mbar_arrive_expect_tx(my_bar, 2 * G2_BOX_BYTES);
                    tma_load_2d(my_stage, &tmap_res, my_bar, col0, box_row0);
                    tma_load_2d(my_stage + G2_BOX_BYTES, &tmap_res, my_bar, col0 + 64, box_row0);
                }
                { const long long c0 = ep.probe ? clock64() : 0; if (ep.relaxed_wait) mbar_wait_relaxed(&tmem_full_bar[as], aphase, 14); else mbar_wait(&tmem_full_bar[as], aphase, 14); if (ep.probe) p_epi += clock64() - c0; }
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G2_BLOCK_N + half * COLS_PER_WARP;
                uint32_t acc_a[32], acc_b[32];
                tmem_ld_32x32(taddr, acc_a);
                const bool has_res = BOXES == 2 && res_ptr != nullptr;   // (the deep-ring form takes no residual)
                if (has_res) mbar_wait(my_bar, res_phase, 15);
#pragma unroll
                for (int c = 0; c < NCH; c += 2) {   // one 64-column box per iteration (a rolled loop measured 1-2 % slower)
                    uint8_t* box = my_stage + (BOXES == 2 ? (c >> 1) : 0) * G2_BOX_BYTES;
                    tmem_ld_wait_regs(acc_a);
                    tmem_ld_32x32(taddr + (c + 1) * 32, acc_b);
                    epilogue_chunk_smem(acc_a, col0 + c * 32, N, ep, box, lane, 0, has_res, mean, rstd, ssum, ssq, BOXES == 1 && c > 0);
                    tmem_ld_wait_regs(acc_b);
                    if (c + 2 < NCH) tmem_ld_32x32(taddr + (c + 2) * 32, acc_a);
                    epilogue_chunk_smem(acc_b, col0 + (c + 1) * 32, N, ep, box, lane, 4, has_res, mean, rstd, ssum, ssq);
                    if (BOXES == 1) {   // the single box leaves right away; it is rewritten after the store has read it
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (col0 + (c >> 1) * 64 < N) tma_store_2d(&tmap_out, box, col0 + (c >> 1) * 64, box_row0);
                            tma_store_commit();
                        }
                    }
                }
                if (BOXES == 2) {
                    // one generic->async proxy fence for both boxes, then the bulk stores
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int bx = 0; bx < NCH / 2; ++bx) {
                            if (col0 + bx * 64 < N) tma_store_2d(&tmap_out, my_stage + bx * G2_BOX_BYTES, col0 + bx * 64, box_row0);
                        }
                        tma_store_commit();
                    }
                }
                if (has_res) res_phase ^= 1;
            } else {
                { const long long c0 = ep.probe ? clock64() : 0; if (ep.relaxed_wait) mbar_wait_relaxed(&tmem_full_bar[as], aphase, 14); else mbar_wait(&tmem_full_bar[as], aphase, 14); if (ep.probe) p_epi += clock64() - c0; }
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G2_BLOCK_N + half * COLS_PER_WARP;
                uint32_t acc_a[32], acc_b[32];
                tmem_ld_32x32(taddr, acc_a);
    #pragma unroll
                for (int c = 0; c < NCH; c += 2) {
                    tmem_ld_wait_regs(acc_a);
                    if (c + 1 < NCH) tmem_ld_32x32(taddr + (c + 1) * 32, acc_b);
                    epilogue_chunk(acc_a, col0 + c * 32, N, row_ok, ep, out_ptr, res_ptr, tab_ptr, mean, rstd, ssum, ssq);
                    if (c + 1 < NCH) {
                        tmem_ld_wait_regs(acc_b);
                        if (c + 2 < NCH) tmem_ld_32x32(taddr + (c + 2) * 32, acc_a);
                        epilogue_chunk(acc_b, col0 + (c + 1) * 32, N, row_ok, ep, out_ptr, res_ptr, tab_ptr, mean, rstd, ssum, ssq);
                    }
                }
            }
            if (ep.stats_out && row_ok && col0 < N) store_row_stats(ep, out_row, col0 / COLS_PER_WARP, ssum, ssq);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], crank & ~1u);  // the pair leader's barrier
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if (ep.probe && leader && pair == 0 && warp == 2 && lane == 0) ep.probe[cluster_id * 4 + 3] = p_epi;
    }

    if (warp >= 2 && lane == 0) tma_store_wait<0>();   // bulk stores of the last tile are complete
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal / read
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, 512);
    }
}

}  // namespace b200
