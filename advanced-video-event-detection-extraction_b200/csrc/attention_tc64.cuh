// Attention for sequences of at most 64 tokens without a mask (the ViT-B/32 image tower: T = 50, 12 heads -> 43 200
// (sequence, head) items per 3600-frame layer) on the 5th-generation tensor cores.  Included by vit_kernels.cu only.
//
// Two items share one 128-row tensor-core tile: item A in rows / keys 0..63, item B in rows / keys 64..127 (T of the
// 64 are real).  Q, K and V of a tile arrive by TMA as six [T x 64] boxes into one of four shared-memory slots
// (3 x 16 KB each); slot s owns 128 TMEM columns and softmax warpgroup s, so four tiles are in flight per SM and the
// loads, the two products and the exponentials of different tiles overlap without any role waiting on its own tile.
//
//   S  = Q[128 x 64] . K[128 keys x 64]^T     4 x tcgen05.mma M=128 N=128 K=16 -> slot columns [0, 128).  Only the two
//                                             diagonal 64 x 64 blocks mean anything; the products across items are never read.
//   P  = softmax of the row's own block       one row per thread, its 64 scores in registers after ONE tcgen05.ld pass
//                                             (TMEM reads are 64 B/clk/SM: the ViT-L/14 kernel is bound by reading S twice);
//                                             P (bf16) goes back to columns [0, 64) as a 128-key row: the row's own block
//                                             plus ZEROS for the other item's keys, so that
//   O  = P[128 x 128] . [V_A; V_B][128 x 64]  8 x tcgen05.mma with A from tensor memory, V MN-major, is exact per item
//                                             (columns [64, 128): the part of S this row never needed, or already holds)
//   out = O / sum                             staged in the tile's Q slot, one TMA store of [T x 64] per item
//
// Rows >= T of the K tiles produce scores that are masked by selection, rows >= T of the V tiles are zeroed once (TMA
// never writes them) because P is exactly 0 there and 0 * NaN is not; rows >= T of Q only feed rows nobody stores.
#pragma once

constexpr int A6_SLOTS = 4, A6_SM_WARPS = 4 * A6_SLOTS;
constexpr int A6_TMA_WARP = A6_SM_WARPS, A6_MMA_WARP = A6_SM_WARPS + 1, A6_ST_WARP = A6_SM_WARPS + 2;
constexpr int A6_THREADS = (A6_ST_WARP + 1) * 32;                  // 608
constexpr int A6_TILE_BYTES = 128 * 128;                           // 128 rows of 128 bytes
constexpr int A6_SLOT_BYTES = 3 * A6_TILE_BYTES;                   // Q | K | V
constexpr int A6_SMEM_BYTES = 1024 + A6_SLOTS * A6_SLOT_BYTES + 256;

__global__ void __launch_bounds__(A6_THREADS, 1)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_o, int T, int heads,
                      int n_items) {
    extern __shared__ uint8_t a6_raw[];
    uint8_t* base = a6_raw + ((1024u - (smem_u32(a6_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + A6_SLOTS * A6_SLOT_BYTES);
    uint64_t* full = bars;                      // [slot] Q, K, V of the tile landed
    uint64_t* empty = bars + A6_SLOTS;          // [slot] output stored (shared memory read), TMEM slot drained
    uint64_t* s_full = bars + 2 * A6_SLOTS;     // [slot] S in TMEM
    uint64_t* p_full = bars + 3 * A6_SLOTS;     // [slot] P in TMEM
    uint64_t* o_full = bars + 4 * A6_SLOTS;     // [slot] O in TMEM
    uint64_t* stage_full = bars + 5 * A6_SLOTS; // [slot] output rows staged in the Q tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 * A6_SLOTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = heads * ATT_D;
    const int n_tiles = (n_items + 1) >> 1;
    const uint32_t box_bytes = static_cast<uint32_t>(T) * 128u;

    // V rows the TMA never writes must be finite (P is exactly 0 there); zero every slot once
    for (int i = threadIdx.x; i < A6_SLOTS * A6_SLOT_BYTES / 16; i += A6_THREADS) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < A6_SLOTS; ++i) {
            mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&stage_full[i], 4);
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_qkv); tma_prefetch_desc(&tmap_o);
    }
    if (warp == A6_MMA_WARP) tmem_alloc<1>(tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // tile k of this CTA = global tile blockIdx.x + k * gridDim.x, in slot k % 4, phase (k / 4) & 1 of the slot's barriers
    if (warp == A6_TMA_WARP) {
        if (lane == 0) {
            int k = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
                const int slot = k & (A6_SLOTS - 1);
                mbar_wait_relaxed(&empty[slot], ((k / A6_SLOTS) & 1) ^ 1, 71);
                uint8_t* sl = base + slot * A6_SLOT_BYTES;
                const int items_here = (2 * tile + 1 < n_items) ? 2 : 1;
                mbar_arrive_expect_tx(&full[slot], 3u * box_bytes * items_here);
                for (int it = 0; it < items_here; ++it) {
                    const int item = 2 * tile + it;
                    const int seq = item / heads, head = item - seq * heads;
                    uint8_t* dst = sl + it * 64 * 128;
                    tma_load_2d(dst, &tmap_qkv, &full[slot], head * ATT_D, seq * T);
                    tma_load_2d(dst + A6_TILE_BYTES, &tmap_qkv, &full[slot], D + head * ATT_D, seq * T);
                    tma_load_2d(dst + 2 * A6_TILE_BYTES, &tmap_qkv, &full[slot], 2 * D + head * ATT_D, seq * T);
                }
            }
        }
    } else if (warp == A6_MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc_bf16(128, 128);
            const uint32_t idesc_pv = make_idesc_bf16(128, ATT_D) | (1u << 16);     // B (= V) MN-major
            auto test = [](uint64_t* bar, uint32_t parity) {
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
                return ok != 0;
            };
            int my_tiles = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my_tiles;
            int us = 0, up = 0;             // next tile whose S / whose PV goes to the tensor core
            uint32_t idle = 0;
            while (up < my_tiles) {
                bool did = false;
                if (us < my_tiles) {
                    const int slot = us & (A6_SLOTS - 1);
                    if (test(&full[slot], (us / A6_SLOTS) & 1)) {
                        tc_fence_after();
                        uint8_t* sl = base + slot * A6_SLOT_BYTES;
                        const uint64_t qdesc = make_sw128_kmajor_desc(smem_u32(sl));
                        const uint64_t kdesc = make_sw128_kmajor_desc(smem_u32(sl + A6_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k)
                            umma_bf16<1>(tmem + slot * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
                        umma_commit(&s_full[slot]);
                        ++us; did = true;
                    }
                }
                if (up < us) {
                    const int slot = up & (A6_SLOTS - 1);
                    if (test(&p_full[slot], (up / A6_SLOTS) & 1)) {
                        tc_fence_after();
                        const uint32_t tb = tmem + slot * 128;
                        const uint64_t vdesc = make_sw128_kmajor_desc(smem_u32(base + slot * A6_SLOT_BYTES + 2 * A6_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16_tmem_a(tb + 64, tb + 8 * k, vdesc + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, k != 0);
                        umma_commit(&o_full[slot]);
                        ++up; did = true;
                    }
                }
                if (did) idle = 0;
                else if (++idle > (1u << 26)) { printf("b200clip: attention_tc64 MMA issuer stuck (block %d, us %d, up %d)\n", (int)blockIdx.x, us, up); __trap(); }
            }
        }
    } else if (warp == A6_ST_WARP) {
        if (lane == 0) {
            int k = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
                const int slot = k & (A6_SLOTS - 1);
                mbar_wait_relaxed(&stage_full[slot], (k / A6_SLOTS) & 1, 72);
                uint8_t* sl = base + slot * A6_SLOT_BYTES;
                const int items_here = (2 * tile + 1 < n_items) ? 2 : 1;
                for (int it = 0; it < items_here; ++it) {
                    const int item = 2 * tile + it;
                    const int seq = item / heads, head = item - seq * heads;
                    tma_store_2d(&tmap_o, sl + it * 64 * 128, head * ATT_D, seq * T);
                }
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(&empty[slot]);
            }
            tma_store_wait<0>();
        }
    } else {
        // ===================== softmax warps: group = slot, one row per thread =====================
        const int slot = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;                    // row of the 128-row tile: item row >> 6, token row & 63
        const int blk = wq >> 1;                           // which item (and which diagonal block of S) this warp serves
        const uint32_t tb = tmem + (static_cast<uint32_t>(wq * 32) << 16) + slot * 128;
        const float scale_log2 = 0.125f * 1.4426950408889634f;
        uint8_t* sl = base + slot * A6_SLOT_BYTES;
        int k = slot;        // tiles slot, slot + 4, ... of this CTA
        int my_tiles = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my_tiles;
        for (; k < my_tiles; k += A6_SLOTS) {
            const uint32_t ph = (k / A6_SLOTS) & 1;
            mbar_wait(&s_full[slot], ph, 73);
            tc_fence_after();
            uint32_t s[64];
            {
                uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[0]);
                uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&s[32]);
                tmem_ld_32x32(tb + blk * 64, s0);
                tmem_ld_32x32(tb + blk * 64 + 32, s1);
                tmem_ld_wait_regs(s0);
                tmem_ld_wait_regs(s1);
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                if (i >= T) s[i] = 0xff800000u;            // -inf: keys past the sequence
                mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[i]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            const float nm = -mx * scale_log2;
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), scale_log2, nm));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2, nm));
                rs4[i & 3] += p0 + p1;
                pk[i] = pack_bf16x2(p0, p1);
            }
            // the 128-key P row: own block at columns [32 blk, +32), zeros for the other item's keys
            {
                uint32_t zero[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) zero[i] = 0u;
                tmem_st_32x32_x32(tb + blk * 32, pk);
                tmem_st_32x32_x32(tb + (blk ^ 1) * 32, zero);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[slot]);
            const float inv = 1.f / ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
            mbar_wait(&o_full[slot], ph, 74);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t o[32];
                tmem_ld_32x32(tb + 64 + hh * 32, o);
                tmem_ld_wait_regs(o);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    uint4 wv;
                    wv.x = pack_bf16x2(__uint_as_float(o[cc * 8 + 0]) * inv, __uint_as_float(o[cc * 8 + 1]) * inv);
                    wv.y = pack_bf16x2(__uint_as_float(o[cc * 8 + 2]) * inv, __uint_as_float(o[cc * 8 + 3]) * inv);
                    wv.z = pack_bf16x2(__uint_as_float(o[cc * 8 + 4]) * inv, __uint_as_float(o[cc * 8 + 5]) * inv);
                    wv.w = pack_bf16x2(__uint_as_float(o[cc * 8 + 6]) * inv, __uint_as_float(o[cc * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(sl + sw_off(row, hh * 4 + cc)) = wv;     // the Q tile: S is complete
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&stage_full[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == A6_MMA_WARP) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}
