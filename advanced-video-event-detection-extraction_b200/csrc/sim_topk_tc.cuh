// K4, tensor-core form (many queries x a large cached embedding matrix; BASELINE config 4):
// scores = E_img[N,E] . E_txt[Q,E]^T on tcgen05 with the 2-CTA mainloop of gemm_tcgen05_2cta.cuh (embedding rows are
// the M operand, up to 256 queries the N operand), fused with the top-k: the score tile never leaves the SM -- each
// epilogue warp reads its 32 rows x 128 queries from TMEM and maintains a sorted top-k list per query in shared
// memory (ballot + shuffle insert; a cheap ">= current k-th" filter keeps the common case at ~5 instructions per
// (32 rows, query)).  Per-warp lists go to global memory once at the end and are merged by topk_final_kernel.
//
// Reference semantics: np.dot (src/models/openclip_model.py:212-214) + np.argsort(s)[::-1][:k]
// (src/pipeline/phase1_mvp.py:145), ties -> higher index first.  bf16 x bf16 -> fp32 scores.
#pragma once
#include "gemm_tcgen05_2cta.cuh"

namespace b200 {

constexpr int STC_MAXK = 8;   // per-warp lists: 8 warps x 128 queries x k x 8 B must fit the 64 KB staging area

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
sim_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int N_rows,
                   int Q, int q0, int q_total, int K, int k, float* __restrict__ dense_out,
                   float* __restrict__ part_s, int* __restrict__ part_i) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
    uint8_t* lists = smem + G2_STAGES * G2_STAGE_BYTES;          // [8 warps][128 queries][k] scores, then indices
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + G2_STAGING_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + G2_STAGES;
    uint64_t* tmem_full_bar = bars + 2 * G2_STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * G2_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4 + GEMM_EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = (N_rows + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
    const int k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < G2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * GEMM_EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<2>(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = tile * 2 * GEMM_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M;
                const int n0 = q0 + static_cast<int>(rank) * G2_HALF_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 21);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
                    const uint32_t bar = smem_u32(&full_bar[stage]) & kPeerBitMask;
                    tma_load_2d_cg2(smem_a + stage * G2_A_BYTES, &tmap_a, bar, kb * GEMM_BLOCK_K, m0);
                    tma_load_2d_cg2(smem_b + stage * G2_B_BYTES, &tmap_w, bar, kb * GEMM_BLOCK_K, n0);
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, G2_BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 22);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 23);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * G2_A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * G2_B_BYTES));
#pragma unroll
                    for (int kk = 0; kk < GEMM_BLOCK_K / GEMM_UMMA_K; ++kk)
                        umma_bf16<2>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                    umma_commit_cg2(&empty_bar[stage], 0b11);
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_cg2(&tmem_full_bar[as], 0b11);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: per-warp top-k lists =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int w = warp - 2;
        float* ls = reinterpret_cast<float*>(lists) + static_cast<size_t>(w) * 128 * k;            // [128][k]
        int* li = reinterpret_cast<int*>(lists + GEMM_EPI_WARPS * 128 * k * 4) + static_cast<size_t>(w) * 128 * k;
        for (int i = lane; i < 128 * k; i += 32) { ls[i] = -INFINITY; li[i] = -1; }
        __syncwarp();
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int row = tile * 2 * GEMM_BLOCK_M + static_cast<int>(rank) * GEMM_BLOCK_M + q * 32 + lane;
            const bool row_ok = row < N_rows;
            mbar_wait(&tmem_full_bar[as], aphase, 24);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G2_BLOCK_N + half * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = half * 128 + c * 32;        // query column inside this 256-query pass
                if (q0 + col0 >= Q) break;                   // warp-uniform
                uint32_t acc[32];
                tmem_ld_32x32(taddr + c * 32, acc);
                tmem_ld_wait_regs(acc);
                if (dense_out && row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (q0 + col0 + j < Q) dense_out[static_cast<size_t>(row) * q_total + q0 + col0 + j] = __uint_as_float(acc[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (q0 + col0 + j >= Q) break;           // warp-uniform
                    const float s = row_ok ? __uint_as_float(acc[j]) : -INFINITY;
                    float* l_s = ls + (c * 32 + j) * k;
                    int* l_i = li + (c * 32 + j) * k;
                    // rows only ever grow within a warp, so ">=" is the exact "beats the k-th entry" test for ties
                    unsigned m = __ballot_sync(0xffffffffu, row_ok && s >= l_s[k - 1]);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const float cs = __shfl_sync(0xffffffffu, s, src);
                        const int ci = row - lane + src;
                        if (cs >= l_s[k - 1]) {
                            // sorted insert: lane e owns entry e; entries that stay ahead are strictly better or
                            // equal with a higher index (never here: ci is the highest index seen so far)
                            const float es = lane < k ? l_s[lane] : 0.f;
                            const int ei = lane < k ? l_i[lane] : 0;
                            const bool ahead = lane < k && es > cs;
                            const int p = __popc(__ballot_sync(0xffffffffu, ahead));
                            const float ps = __shfl_up_sync(0xffffffffu, es, 1);
                            const int pi = __shfl_up_sync(0xffffffffu, ei, 1);
                            if (lane < k) {
                                if (lane == p) { l_s[lane] = cs; l_i[lane] = ci; }
                                else if (lane > p) { l_s[lane] = ps; l_i[lane] = pi; }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        // publish: list id = ((cluster * 2 + rank) * 4 + lane quarter); each warp owns 128 query columns
        const int list_id = (cluster_id * 2 + static_cast<int>(rank)) * 4 + q;
        for (int i = lane; i < 128 * k; i += 32) {
            const int col = half * 128 + i / k;
            if (q0 + col < Q) {
                const size_t o = (static_cast<size_t>(list_id) * q_total + q0 + col) * k + (i % k);
                part_s[o] = ls[i];
                part_i[o] = li[i];
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, 512);
    }
}

}  // namespace b200
