// K4, tensor-core form (many queries x a large cached embedding matrix; BASELINE config 4):
// scores^T = E_txt[Q,E] . E_img[N,E]^T on tcgen05 with the 2-CTA mainloop of gemm_tcgen05_2cta.cuh.  The QUERIES are
// the M operand (256 per pass: 128 TMEM lanes in each CTA of the pair) and 256 embedding rows per tile are the N
// operand, so that in the epilogue every thread owns ONE query and sees the scores of consecutive embedding rows as
// the columns of its TMEM lane: the top-k is a register-resident sorted list per thread -- one compare per score on
// the common path, a branch-free compare/swap chain on insertion, no cross-lane traffic and no shared memory.  The
// score tile never leaves the SM.  Each (cluster, column half) publishes one list per query at the end and
// topk_final_kernel merges them.
// Filter: a thread only sees N / SMs rows, so its own k-th score is a weak bound (~k*ln(rows/k) inserts per thread).
// Every thread therefore also publishes its k-th score to gthr[query] with an atomic max -- the k-th best of ANY
// subset of the rows is a lower bound of the k-th best of all rows -- and re-reads that global bound once per tile.
// Scores below the bound are dropped by a branch-free compare that builds a 32-bit hit mask per 32-column chunk; the
// insertion code runs only for chunks with a hit, which after the first few tiles is rare.  Scores EQUAL to the bound
// still pass (ties resolve by index in the final merge), so the result does not depend on the publication race.
//
// Reference semantics: np.dot (src/models/openclip_model.py:212-214) + np.argsort(s)[::-1][:k]
// (src/pipeline/phase1_mvp.py:145), ties -> higher index first.  bf16 x bf16 -> fp32 scores.
#pragma once
#include "gemm_tcgen05_2cta.cuh"

namespace b200 {

constexpr int STC_MAXK = 8;   // length of the per-thread register list (the first k entries are published)

// order-preserving float <-> int map (an involution), so that atomicMax on ints orders floats of either sign
__device__ __forceinline__ int ord_encode(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord_decode(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ARES: the query operand (this CTA's 128 queries x E <= 512, 128 KB) is loaded ONCE and stays resident in shared
// memory; only the embedding rows stream through the (6-stage) ring.  Halves the L2 -> SM fill traffic, which is what
// bounds this kernel once the epilogue is cheap.  Without ARES (E > 512) both operands stream as in the GEMM.
constexpr int STC_MAX_STAGES = 6;
template <bool ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
sim_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int N_rows,
                   int row_offset /* index of row 0 of this launch's slice of the cache */, int Q, int q0, int q_total, int K, int k, float* __restrict__ dense_out,
                   float* __restrict__ part_s, int* __restrict__ part_i, int* __restrict__ gthr) {
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on smem_raw (not an integer round trip) keeps the shared address space visible: LDS / STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int NS = ARES ? STC_MAX_STAGES : G2_STAGES;
    uint8_t* smem_a = smem;                                              // ARES: [k_blocks <= 8][128 x 64] resident
    uint8_t* smem_b = smem + (ARES ? 8 : G2_STAGES) * G2_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + G2_STAGING_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STC_MAX_STAGES;
    uint64_t* tmem_full_bar = bars + 2 * STC_MAX_STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * STC_MAX_STAGES + 2;
    uint64_t* a_full_bar = bars + 2 * STC_MAX_STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STC_MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = (N_rows + G2_BLOCK_N - 1) / G2_BLOCK_N;
    const int k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * GEMM_EPI_WARPS); }
        mbar_init(a_full_bar, 1);
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
            const int m0 = q0 + static_cast<int>(rank) * GEMM_BLOCK_M;                         // this CTA's 128 queries
            if (ARES) {
                if (leader) mbar_arrive_expect_tx(a_full_bar, 2 * k_blocks * G2_A_BYTES);
                const uint32_t abar = smem_u32(a_full_bar) & kPeerBitMask;
                for (int kb = 0; kb < k_blocks; ++kb)
                    tma_load_2d_cg2(smem_a + kb * G2_A_BYTES, &tmap_a, abar, kb * GEMM_BLOCK_K, m0);
            }
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int n0 = tile * G2_BLOCK_N + static_cast<int>(rank) * G2_HALF_N;         // its half of the tile's rows
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 21);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], ARES ? 2 * G2_B_BYTES : 2 * G2_STAGE_BYTES);
                    const uint32_t bar = smem_u32(&full_bar[stage]) & kPeerBitMask;
                    if (!ARES) tma_load_2d_cg2(smem_a + stage * G2_A_BYTES, &tmap_a, bar, kb * GEMM_BLOCK_K, m0);
                    tma_load_2d_cg2(smem_b + stage * G2_B_BYTES, &tmap_w, bar, kb * GEMM_BLOCK_K, n0);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
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
            if (ARES) mbar_wait(a_full_bar, 0, 25);
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 22);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 23);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + (ARES ? kb : stage) * G2_A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * G2_B_BYTES));
#pragma unroll
                    for (int kk = 0; kk < GEMM_BLOCK_K / GEMM_UMMA_K; ++kk)
                        umma_bf16<2>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (kb | kk) != 0);
                    umma_commit_cg2(&empty_bar[stage], 0b11);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                umma_commit_cg2(&tmem_full_bar[as], 0b11);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: one query per thread, register top-k =====================
        const int q = warp & 3;                    // TMEM lane quarter of this warp
        const int half = (warp - 2) >> 2;          // which 128 of the tile's 256 embedding rows (columns)
        const int query = q0 + static_cast<int>(rank) * GEMM_BLOCK_M + q * 32 + lane;
        float sc[STC_MAXK];
        int id[STC_MAXK];
#pragma unroll
        for (int e = 0; e < STC_MAXK; ++e) { sc[e] = -INFINITY; id[e] = -1; }
        float thr = -INFINITY;                     // max(own k-th score, global bound of the query)
        const bool warp_has_query = query - lane < Q;                 // warp-uniform
        int* my_gthr = gthr + (query < Q ? query : 0);
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int row0 = tile * G2_BLOCK_N + half * 128;          // embedding row of this warp's first column
            if (warp_has_query) thr = fmaxf(thr, ord_decode(__ldcg(my_gthr)));   // a stale value is still a valid bound
            mbar_wait(&tmem_full_bar[as], aphase, 24);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G2_BLOCK_N + half * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int rbase = row0 + c * 32;
                if (rbase >= N_rows || !warp_has_query) break;         // warp-uniform: columns of zero-filled rows
                uint32_t acc[32];
                tmem_ld_32x32(taddr + c * 32, acc);
                tmem_ld_wait_regs(acc);
                if (dense_out && query < Q) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (rbase + j < N_rows) dense_out[static_cast<size_t>(rbase + j) * q_total + query] = __uint_as_float(acc[j]);
                }
                // branch-free filter: bit j = score of embedding row rbase + j reaches this query's bound
                uint32_t hit = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(acc[j]) >= thr) ? (1u << j) : 0u;
                const int nvalid = N_rows - rbase;                     // < 32 only in the last tile
                if (nvalid < 32) hit &= (1u << nvalid) - 1u;
                if (hit == 0) continue;                                // per thread; whole warps skip after warm-up
                // rare path: walk the set bits.  acc[] is indexed dynamically here, so the compiler parks this copy in
                // (L1-resident) local memory -- 32 stores, paid only by chunks with a candidate, instead of running the
                // insertion chain under predicate for all 32 columns
                float cand[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) cand[j] = __uint_as_float(acc[j]);
                bool changed = false;
                while (hit) {
                    const int j = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const float s = cand[j];
                    // rows arrive in increasing order, so ">=" is the exact "beats the k-th entry" test under the
                    // reference's tie rule (equal scores: higher index first)
                    if (s >= sc[STC_MAXK - 1]) {
                        changed = true;
                        sc[STC_MAXK - 1] = s; id[STC_MAXK - 1] = row_offset + rbase + j;
#pragma unroll
                        for (int e = STC_MAXK - 1; e > 0; --e) {       // bubble up; the newcomer passes equal scores
                            const bool sw = sc[e] >= sc[e - 1];
                            const float ts = sc[e]; const int ti = id[e];
                            sc[e] = sw ? sc[e - 1] : ts; id[e] = sw ? id[e - 1] : ti;
                            sc[e - 1] = sw ? ts : sc[e - 1]; id[e - 1] = sw ? ti : id[e - 1];
                        }
                    }
                }
                if (changed) {
                    // own k-th score (runtime k <= STC_MAXK): the new local bound, shared when it improves the global one
                    float kth = sc[0];
#pragma unroll
                    for (int e = 1; e < STC_MAXK; ++e) kth = (e < k) ? sc[e] : kth;
                    if (kth > thr) {
                        thr = kth;
                        if (query < Q) atomicMax(my_gthr, ord_encode(kth));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        // publish: one list per (cluster, column half) and query
        if (query < Q) {
            const int list_id = cluster_id * 2 + half;
            const size_t o = (static_cast<size_t>(list_id) * q_total + query) * k;
#pragma unroll
            for (int e = 0; e < STC_MAXK; ++e)
                if (e < k) { part_s[o + e] = sc[e]; part_i[o + e] = id[e]; }
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
