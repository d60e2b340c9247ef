// Attention for T = 257 (ViT-L/14, H/14, g/14 at 224 px: 256 patches + class token), unmasked, head dim 64, on the
// 5th-generation tensor cores with the probabilities handed over IN TENSOR MEMORY.  Included by vit_kernels.cu only.
//
// One persistent CTA per SM walks (sequence, head) items; Q, K and V of an item (100 KB) arrive by TMA into one of two
// shared-memory buffers, so the loads of the next item run under the arithmetic of the current one.  An item is two
// "units" = 128-row query tiles (rows 0..127 and 128..255); unit u lives in TMEM slot u & 1 (256 columns each) and is
// served by softmax warp group u & 1: 8 warps, TWO per TMEM lane quadrant, each taking 128 of the row's 256 key columns
// (partial maxima and sums are exchanged through shared memory) -- one warp per quadrant cannot overlap its own MUFU
// and TMEM phases.  The 257th row is scored on the CUDA cores by four more warps at the same time (attention_odd_row,
// as in attention_tc_kernel); a store warp writes the output tiles.
//
//   S   = Q_tile[128 x 64] . K[256 keys x 64]^T      4 x tcgen05.mma M=128 N=256 K=16, fp32 into slot columns [0, 256)
//   s'  = Q_tile . k_256                             the 257th key cannot join the N = 256 tile (and TMEM has no room for
//                                                    a second one): two mma.sync m16n8k16 tiles per quadrant, one shuffle set
//   P   = exp2(scale (S - max))                      FULL-ROW softmax (all of a row's scores are resident, so there is no
//                                                    running maximum and no rescaling of O): pass 1 = maximum, pass 2 =
//                                                    exponentials, both over double-buffered 16-column tcgen05.ld chunks;
//                                                    P (bf16, two keys per column) is written back with tcgen05.st over
//                                                    columns the SAME warp has already consumed: keys 0..127 -> [0, 64),
//                                                    128..255 -> [128, 192), key 256 and 15 zero partners -> [192, 200)
//                                                    -- it never touches shared memory
//   O   = P[128 x 272] . V[272 keys x 64]            17 x tcgen05.mma with A FROM TENSOR MEMORY (pinned by
//                                                    tools/probes/umma_tmem_a_probe.cu) and V as an MN-major B operand,
//                                                    fp32 into slot columns [64, 128) (scores of the first half, consumed)
//   out = O / sum                                    tcgen05.ld, normalise, bf16 rows staged in the unit's Q tile (the
//                                                    tensor core is done with it), one TMA store per tile
//
// A single thread issues all MMAs; it polls (mbarrier.test_wait) which of S(next unit) -- needs the item's data and a
// drained slot -- and PV(previous unit) -- needs its P -- is ready and sends S first, so the tensor core produces S for
// one group while the other one exponentiates.  What bounds the kernel is the tensor-memory READ port, 64 B/clk/SM:
// a full-row softmax reads S twice, 2 x 128 KB + 32 KB of O per unit = 9.2 k clocks per item, the measured period
// (MUFU.EX2 would allow 4.4 k, HBM 5.7 k; profiles/r02r_attention_tc2_steps.md).
#pragma once

constexpr int A2_T = 257, A2_SM_WARPS = 16, A2_ODD_WARPS = 4;
constexpr int A2_TMA_WARP = A2_SM_WARPS + A2_ODD_WARPS, A2_MMA_WARP = A2_TMA_WARP + 1, A2_ST_WARP = A2_MMA_WARP + 1;
constexpr int A2_THREADS = (A2_ST_WARP + 1) * 32;                  // 736
constexpr int A2_Q_BYTES = 2 * 128 * 128;                          // two 128-row tiles of 128-byte rows
constexpr int A2_KV_ROWS = 272;                                    // 257 keys, padded to the 16-key MMA step
constexpr int A2_KV_BYTES = A2_KV_ROWS * 128;
constexpr int A2_BUF_BYTES = A2_Q_BYTES + 2 * A2_KV_BYTES;         // 102400
constexpr int A2_SCRATCH_BYTES = 8192;                             // odd-row scratch (648 floats) + row max / sum exchange
constexpr int A2_SMEM_BYTES = 1024 + 2 * A2_BUF_BYTES + A2_SCRATCH_BYTES + 256;

__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_regs16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 : : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: A is read from tensor memory (128 lanes x K/2 columns, two bf16 per column)
__device__ __forceinline__ void umma_bf16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

#ifdef B200CLIP_PROBES
// development probe: clock64 stamps of CTA 0 (softmax warps 0 and 4, MMA issuer) for items 8 .. 15, read back with
// b200clip_debug_a2_probe (tools/attn_tc2_check.py prints the timeline)
__device__ long long g_a2_probe[3 * 8 * 8];
#define A2_STAMP(who, j, k) do { if (blockIdx.x == 0 && (j) >= 8 && (j) < 16) g_a2_probe[((who) * 8 + ((j) - 8)) * 8 + (k)] = clock64(); } while (0)
#else
#define A2_STAMP(who, j, k) do { } while (0)
#endif

__global__ void __launch_bounds__(A2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q128, const __grid_constant__ CUtensorMap tmap_kv64,
                     const __grid_constant__ CUtensorMap tmap_kv16, const __grid_constant__ CUtensorMap tmap_o,
                     const bf16* __restrict__ qkv, bf16* __restrict__ out, int heads, int n_items) {
    extern __shared__ uint8_t a2_raw[];
    uint8_t* base = a2_raw + ((1024u - (smem_u32(a2_raw) & 1023u)) & 1023u);
    float* scratch = reinterpret_cast<float*>(base + 2 * A2_BUF_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * A2_BUF_BYTES + A2_SCRATCH_BYTES);
    uint64_t* full = bars;              // [2] Q, K, V of an item landed in buffer b
    uint64_t* empty = bars + 2;         // [2] buffer b no longer read (tensor core + softmax warps + odd-row warps)
    uint64_t* s_full = bars + 4;        // [2] S of the unit in slot s is in TMEM
    uint64_t* p_full = bars + 6;        // [2] P of the unit in slot s is in TMEM
    uint64_t* o_full = bars + 8;        // [2] O of the unit in slot s is in TMEM (and P consumed)
    uint64_t* slot_free = bars + 10;    // [2] O read out: the slot may take the next S
    uint64_t* stage_full = bars + 12;   // [2] the unit's output rows are staged in its Q tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = A2_T, D = heads * ATT_D;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + 1 + A2_ODD_WARPS); mbar_init(&stage_full[i], 8);
            mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&o_full[i], 1); mbar_init(&slot_free[i], 8);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_q128); tma_prefetch_desc(&tmap_kv64); tma_prefetch_desc(&tmap_kv16); tma_prefetch_desc(&tmap_o);
    }
    if (warp == A2_MMA_WARP) tmem_alloc<1>(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == A2_TMA_WARP) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int j = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
                const int b = j & 1;
                mbar_wait_relaxed(&empty[b], ((j >> 1) & 1) ^ 1, 61);
                const int seq = item / heads, head = item - seq * heads;
                const int row0 = seq * T;
                uint8_t* sQ = base + b * A2_BUF_BYTES;
                uint8_t* sK = sQ + A2_Q_BYTES;
                uint8_t* sV = sK + A2_KV_BYTES;
                mbar_arrive_expect_tx(&full[b], A2_BUF_BYTES);
                tma_load_2d(sQ, &tmap_q128, &full[b], head * ATT_D, row0);
                tma_load_2d(sQ + 128 * 128, &tmap_q128, &full[b], head * ATT_D, row0 + 128);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    tma_load_2d(sK + kb * 64 * 128, &tmap_kv64, &full[b], D + head * ATT_D, row0 + kb * 64);
                    tma_load_2d(sV + kb * 64 * 128, &tmap_kv64, &full[b], 2 * D + head * ATT_D, row0 + kb * 64);
                }
                // rows 256 .. 271: key 256 and 15 rows of the next sequence (or zero fill), which P multiplies by zero
                tma_load_2d(sK + 256 * 128, &tmap_kv16, &full[b], D + head * ATT_D, row0 + 256);
                tma_load_2d(sV + 256 * 128, &tmap_kv16, &full[b], 2 * D + head * ATT_D, row0 + 256);
            }
        }
    } else if (warp == A2_MMA_WARP) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(128, 256);
            const uint32_t idesc_pv = make_idesc_bf16(128, ATT_D) | (1u << 16);     // B (= V) MN-major
            auto test = [](uint64_t* bar, uint32_t parity) {       // non-blocking: has the phase of that parity completed?
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
                return ok != 0;
            };
            int my_items = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) ++my_items;
            const int U = 2 * my_items;
            // S(us) needs the item's data and a free slot, PV(up) the unit's probabilities: whichever is ready goes to the
            // tensor core, S first (it releases a whole warpgroup, PV only an epilogue); a fixed S(u), PV(u - 1) order
            // would park PV behind the load of the next item
            int us = 0, up = 0;
            uint32_t idle = 0;
            while (up < U) {
                bool did = false;
                if (us < U) {
                    const int tile = us & 1, j = us >> 1, b = j & 1;
                    if (test(&full[b], (j >> 1) & 1) && test(&slot_free[tile], (j & 1) ^ 1)) {
                        A2_STAMP(2, j, tile * 2);
                        tc_fence_after();
                        uint8_t* sQ = base + b * A2_BUF_BYTES;
                        const uint64_t kdesc = make_sw128_kmajor_desc(smem_u32(sQ + A2_Q_BYTES));
                        const uint64_t qdesc = make_sw128_kmajor_desc(smem_u32(sQ + tile * 128 * 128));
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k)
                            umma_bf16<1>(tmem + tile * 256, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
                        umma_commit(&s_full[tile]);
                        ++us; did = true;
                    }
                }
                if (up < us) {
                    const int slot = up & 1, j = up >> 1, b = j & 1;
                    if (test(&p_full[slot], j & 1)) {
                        A2_STAMP(2, j, 4 + slot * 2);
                        tc_fence_after();
                        const uint32_t tbase = tmem + slot * 256;
                        const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(base + b * A2_BUF_BYTES + A2_Q_BYTES + A2_KV_BYTES));
#pragma unroll
                        for (int k = 0; k < A2_KV_ROWS / 16; ++k)     // P: keys 0..127 at columns [0, 64), 128..255 at [128, 192), 256.. at [192, 200)
                            umma_bf16_tmem_a(tbase + 64, tbase + (k < 8 ? 8 * k : 64 + 8 * k), bdesc + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, k != 0);
                        umma_commit(&o_full[slot]);
                        if (slot == 1) umma_commit(&empty[b]);     // every MMA that reads buffer b has been issued
                        ++up; did = true;
                    }
                }
                if (did) idle = 0;
                else if (++idle > (1u << 26)) { printf("b200clip: attention_tc2 MMA issuer stuck (block %d, us %d, up %d)\n", (int)blockIdx.x, us, up); __trap(); }
            }
        }
    } else if (warp == A2_ST_WARP) {
        // ===================== output store: one TMA store per unit out of its (recycled) Q tile =====================
        if (lane == 0) {
            int j = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
                const int b = j & 1;
                const int seq = item / heads, head = item - seq * heads;
                for (int tile = 0; tile < 2; ++tile) {
                    mbar_wait_relaxed(&stage_full[tile], j & 1, 69);
                    tma_store_2d(&tmap_o, base + b * A2_BUF_BYTES + tile * 128 * 128, head * ATT_D, seq * T + tile * 128);
                    tma_store_commit();
                }
                tma_store_wait_read<0>();
                mbar_arrive(&empty[b]);          // both staging tiles of buffer b have been read
            }
            tma_store_wait<0>();
        }
    } else if (warp >= A2_SM_WARPS) {
        // ===================== odd-row warps: query row 256 on the CUDA cores =====================
        const int ow = warp - A2_SM_WARPS;
        int j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int b = j & 1;
            mbar_wait_relaxed(&full[b], (j >> 1) & 1, 65);
            const int seq = item / heads, head = item - seq * heads;
            const uint8_t* sK = base + b * A2_BUF_BYTES + A2_Q_BYTES;
            attention_odd_row<A2_ODD_WARPS, 320>(sK, sK + A2_KV_BYTES, scratch,
                                                 qkv + (static_cast<int64_t>(seq) * T + T - 1) * 3 * D + head * ATT_D,
                                                 out + (static_cast<int64_t>(seq) * T + T - 1) * D + head * ATT_D, T, ow, lane, 2);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);
        }
    } else {
        // ===================== softmax warps: group of 8 warps = query tile; a row is shared by the two warps of a TMEM
        // lane quadrant, each taking 128 of its 256 key columns (half 1 also the 257th key) =====================
        const int tile = warp >> 3, half = (warp >> 2) & 1, wq = warp & 3;
        const int row_in_tile = wq * 32 + lane;
        const uint32_t tbase = tmem + (static_cast<uint32_t>(wq * 32) << 16) + tile * 256;
        const float scale_log2 = 0.125f * 1.4426950408889634f;
        const int g = lane >> 2, t4 = lane & 3;
        float* xmax = scratch + 1024 + tile * 512;           // [2 halves][128 rows] partial row maxima of the unit
        float* xsum = xmax + 256;                            // [2 halves][128 rows] partial row sums
        int j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int b = j & 1;
            const bool stamp = lane == 0 && wq == 0 && half == 0;
            if (stamp) A2_STAMP(tile, j, 0);
            mbar_wait(&full[b], (j >> 1) & 1, 66);
            if (stamp) A2_STAMP(tile, j, 1);
            // ---- s' = q_row . k_256 for the 32 rows of the quadrant (half 1): two m16n8k16 tiles, only column 0 of B
            //      is non-zero
            float s_last = -INFINITY;
            if (half == 1) {
                const uint32_t qa = smem_u32(base + b * A2_BUF_BYTES + tile * 128 * 128);
                const uint32_t ka = smem_u32(base + b * A2_BUF_BYTES + A2_Q_BYTES + 256 * 128);    // row 256: chunks unswizzled
                float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t b0 = 0u, b1 = 0u;
                    if (g == 0) {
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(ka + 32u * ks + 4u * t4));
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b1) : "r"(ka + 32u * ks + 16u + 4u * t4));
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        uint32_t a[4];
                        const int r = wq * 32 + mt * 16 + (lane & 15);
                        ldmatrix_x4(qa + sw_off(r, ks * 2 + (lane >> 4)), a[0], a[1], a[2], a[3]);
                        mma_bf16_16816(c[mt], a, b0, b1);
                    }
                }
                // row L of the warp = m-tile L >> 4, fragment row L & 15: column 0 sits in lane 4 (L & 7), c[.][0] / c[.][2]
                const int src = 4 * (lane & 7);
                const float v00 = __shfl_sync(0xffffffffu, c[0][0], src), v02 = __shfl_sync(0xffffffffu, c[0][2], src);
                const float v10 = __shfl_sync(0xffffffffu, c[1][0], src), v12 = __shfl_sync(0xffffffffu, c[1][2], src);
                s_last = (lane & 16) ? ((lane & 8) ? v12 : v10) : ((lane & 8) ? v02 : v00);
            }
            // ---- pass 1: maximum over this warp's 128 of the row's scores in TMEM (+ s'), exchanged with the other half
            if (stamp) A2_STAMP(tile, j, 2);
            mbar_wait(&s_full[tile], j & 1, 67);
            if (stamp) A2_STAMP(tile, j, 3);
            tc_fence_after();
            const uint32_t scol = tbase + half * 128;
            float mx4[4] = {s_last, s_last, s_last, s_last};     // independent chains: FMNMX3 has ~5 clocks of latency
            {
                uint32_t sa[2][16];          // 16-column chunks, the next one in flight while this one is reduced
                tmem_ld_32x32_x16(scol, sa[0]);
                tmem_ld_wait_regs16(sa[0]);
#pragma unroll
                for (int cchunk = 0; cchunk < 8; ++cchunk) {
                    if (cchunk + 1 < 8) tmem_ld_32x32_x16(scol + (cchunk + 1) * 16, sa[(cchunk + 1) & 1]);
#pragma unroll
                    for (int i = 0; i < 16; i += 2)
                        mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(__uint_as_float(sa[cchunk & 1][i]), __uint_as_float(sa[cchunk & 1][i + 1])));
                    if (cchunk + 1 < 8) tmem_ld_wait_regs16(sa[(cchunk + 1) & 1]);
                }
            }
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            xmax[half * 128 + row_in_tile] = mx;
            named_bar_sync(5 + tile, 256);
            mx = fmaxf(mx, xmax[(half ^ 1) * 128 + row_in_tile]);
            // ---- pass 2: exponentials, partial row sum, P (bf16 pairs) back over THIS warp's already consumed columns:
            //      half 0 -> columns [0, 64), half 1 -> [128, 192) and key 256 (+ 15 zero partners) -> [192, 200)
            if (stamp) A2_STAMP(tile, j, 4);
            const float nm = -mx * scale_log2;
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
            {
                uint32_t sa[2][16];
                tmem_ld_32x32_x16(scol, sa[0]);
                tmem_ld_wait_regs16(sa[0]);
#pragma unroll
                for (int cchunk = 0; cchunk < 8; ++cchunk) {
                    if (cchunk + 1 < 8) tmem_ld_32x32_x16(scol + (cchunk + 1) * 16, sa[(cchunk + 1) & 1]);
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sa[cchunk & 1][2 * i]), scale_log2, nm));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sa[cchunk & 1][2 * i + 1]), scale_log2, nm));
                        rs4[i & 3] += p0 + p1;
                        pk[i] = pack_bf16x2(p0, p1);
                    }
                    // columns [8 c, 8 c + 8) of this half held scores of chunks <= c / 2: all in registers or consumed
                    if (cchunk + 1 < 8) tmem_ld_wait_regs16(sa[(cchunk + 1) & 1]);
                    tmem_st_32x32_x8(scol + cchunk * 8, pk);
                }
            }
            if (half == 1) {
                const float pl = ex2_approx(fmaf(s_last, scale_log2, nm));
                rs4[0] += pl;
                uint32_t pk[8] = {pack_bf16x2(pl, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tmem_st_32x32_x8(tbase + 192, pk);
            }
            const float rs = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            xsum[half * 128 + row_in_tile] = rs;             // read by the other half behind the group barrier below
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[tile]);
            if (stamp) A2_STAMP(tile, j, 5);
            // ---- O of the unit (slot columns [64, 128)): this warp normalises and stages 32 of the 64 output columns
            mbar_wait(&o_full[tile], j & 1, 68);
            if (stamp) A2_STAMP(tile, j, 6);
            tc_fence_after();
            uint8_t* stage = base + b * A2_BUF_BYTES + tile * 128 * 128;     // the unit's Q tile: S is complete, and the
            uint32_t o[32];                                                  // ldmatrix reads end before the barrier below
            tmem_ld_32x32(tbase + 64 + half * 32, o);
            tmem_ld_wait_regs(o);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_free[tile]);
            named_bar_sync(3 + tile, 256);
            const float inv = 1.f / (rs + xsum[(half ^ 1) * 128 + row_in_tile]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint4 wv;
                wv.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * inv, __uint_as_float(o[cc * 8 + 1]) * inv);
                wv.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * inv, __uint_as_float(o[cc * 8 + 3]) * inv);
                wv.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * inv, __uint_as_float(o[cc * 8 + 5]) * inv);
                wv.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * inv, __uint_as_float(o[cc * 8 + 7]) * inv);
                *reinterpret_cast<uint4*>(stage + sw_off(row_in_tile, half * 4 + cc)) = wv;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stage_full[tile]);      // the store warp takes it from here
            if (stamp) A2_STAMP(tile, j, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == A2_MMA_WARP) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}
