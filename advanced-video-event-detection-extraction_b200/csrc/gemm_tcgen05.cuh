// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
//   * A (activations, token rows) and W (nn.Linear weight [out,in]) are both K-major, so both
//     operands are fed to tcgen05.mma straight from 128B-swizzled shared memory tiles that TMA
//     fills (box 64 x rows, bf16) -- no transposes anywhere.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
//     warps 2..9 = epilogue (each may read the 32 TMEM lanes of its warp-id % 4; two warps split the
//     columns of a lane quarter), TMEM loads software-pipelined against the epilogue math.
//   * the fp32 accumulator lives in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of
//     tile i overlaps the main loop of tile i+1; one CTA per SM loops over tiles (grid = #SMs).
//   * fused epilogues: +bias, +bias+QuickGELU / erf-GELU, +bias+residual (in place), and the
//     patch-embed variant (row remap 49->50 tokens per image, + positional-embedding row).
//
// Replaces the cuBLAS/cuDNN calls eager PyTorch makes inside open_clip's VisionTransformer.forward
// (reference call site: src/models/openclip_model.py:177,196 -> model.encode_image).
#pragma once
#include "ptx.cuh"

namespace b200 {

struct GemmEpilogue {
    const float* bias;          // [N] or nullptr
    const __nv_bfloat16* resid; // [rows_out, ldc] added to the result, or nullptr (may alias out)
    const float* rowtab;        // [t_out, N] fp32 row table added by token position, or nullptr
    int act;                    // 0 none, 1 QuickGELU x*sigmoid(1.702x), 2 GELU(erf)
    int t_in, t_out, row_off;   // out_row = (r / t_in) * t_out + r % t_in + row_off   (t_in == 0: identity)
    // LayerNorm folded into this GEMM:  LN(x) W^T + b  ==  rstd * (x (W.gamma)^T - mean * c1) + c2  with
    // c1[n] = sum_k (W.gamma)[n,k], c2[n] = sum_k beta[k] W[n,k] + b[n] (passed as `bias`).  The per-row sums
    // (sum x, sum x^2) come from `ln_stats` [rows, LN_SLOTS, 2], written by whoever produced x.
    const float* ln_stats;      // nullptr: no fold
    const float* ln_c1;         // [N]
    float ln_eps;
    int ln_width;               // row length of x (= K)
    // this GEMM produces the residual stream: emit (sum, sum of squares) of each output row segment
    float* stats_out;           // [rows, LN_SLOTS, 2] or nullptr
    // development probe (B200CLIP_GEMM_PROBE): per cluster {total, MMA wait on TMA data, MMA wait on a free
    // accumulator, epilogue wait on the accumulator} in SM clocks; nullptr in production
    long long* probe;
    int relaxed_wait;           // nanosleep back-off in the epilogue's accumulator wait (set by the launcher)
};
constexpr int LN_SLOTS = 8;     // row segments (128 or 64 columns wide) whose partial sums are kept separately

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + GEMM_EPI_WARPS * 32;  // TMA warp + MMA warp + epilogue warps

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*align slack*/;
};

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == 1) {
        // QuickGELU: x * sigmoid(1.702 x), sigmoid(y) = 0.5 + 0.5 tanh(y/2): one MUFU (tanh.approx, rel. err ~2^-11)
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
        return x * fmaf(0.5f, t, 0.5f);
    } else if (act == 2) {
        return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    }
    return x;
}

// per-row LayerNorm statistics from the partial sums
__device__ __forceinline__ void ln_row_stats(const GemmEpilogue& ep, long row, bool row_ok, float& mean, float& rstd) {
    mean = 0.f; rstd = 1.f;
    if (!ep.ln_stats || !row_ok) return;
    const float4* p = reinterpret_cast<const float4*>(ep.ln_stats + row * (LN_SLOTS * 2));
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_SLOTS / 2; ++i) {
        const float4 v = __ldg(p + i);
        s += v.x + v.z; s2 += v.y + v.w;
    }
    const float inv = 1.0f / static_cast<float>(ep.ln_width);
    mean = s * inv;
    const float var = fmaxf(s2 * inv - mean * mean, 0.f);
    rstd = rsqrtf(var + ep.ln_eps);
}

// Epilogue math runs on packed fp32x2 registers (FFMA2 / FADD2 / FMUL2: two columns per instruction); the 32
// accumulator columns of a thread are 16 pairs.
__device__ __forceinline__ f32x2_t act2(f32x2_t x, int act) {
    // scalar on purpose: the packed formulation (FMUL2 + 2 MUFU + FFMA2 + FMUL2) measured 35% slower on the fc GEMM
    float a, b;
    f2_unpack(x, a, b);
    return f2_pack(apply_act(a, act), apply_act(b, act));
}

// bias / LN-fold / row table / activation (everything before the residual add)
__device__ __forceinline__ void epilogue_math(f32x2_t (&v)[16], int n0, const GemmEpilogue& ep, const float* tab_ptr,
                                              float mean, float rstd) {
    if (ep.ln_stats) {
        const f32x2_t nm = f2_pack(-mean, -mean), rs = f2_pack(rstd, rstd);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(ep.ln_c1 + n0 + j));
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
            v[j / 2] = f2_fma(rs, f2_fma(nm, f2_pack(c1.x, c1.y), v[j / 2]), f2_pack(c2.x, c2.y));
            v[j / 2 + 1] = f2_fma(rs, f2_fma(nm, f2_pack(c1.z, c1.w), v[j / 2 + 1]), f2_pack(c2.z, c2.w));
        }
    } else if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
            v[j / 2] = f2_add(v[j / 2], f2_pack(b.x, b.y));
            v[j / 2 + 1] = f2_add(v[j / 2 + 1], f2_pack(b.z, b.w));
        }
    }
    if (tab_ptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(tab_ptr + n0 + j));
            v[j / 2] = f2_add(v[j / 2], f2_pack(b.x, b.y));
            v[j / 2 + 1] = f2_add(v[j / 2 + 1], f2_pack(b.z, b.w));
        }
    }
    if (ep.act) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act2(v[j], ep.act);
    }
}

__device__ __forceinline__ f32x2_t bf16x2_to_f2(uint32_t u) {
    const float2 f = unpack_bf16x2(u);
    return f2_pack(f.x, f.y);
}
__device__ __forceinline__ uint32_t f2_to_bf16x2(f32x2_t v) {
    float a, b;
    f2_unpack(v, a, b);
    return pack_bf16x2(a, b);
}

// one 32-column chunk, direct global stores (row-remapped outputs, small problems)
__device__ __forceinline__ void epilogue_chunk(uint32_t (&acc)[32], int n0, int N, bool row_ok, const GemmEpilogue& ep,
                                               __nv_bfloat16* out_ptr, const __nv_bfloat16* res_ptr,
                                               const float* tab_ptr, float mean, float rstd, f32x2_t& ssum,
                                               f32x2_t& ssq) {
    if (!row_ok || n0 >= N) return;
    f32x2_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = f2_pack(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
    epilogue_math(v, n0, ep, tab_ptr, mean, rstd);
    if (res_ptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const uint4 r = *reinterpret_cast<const uint4*>(res_ptr + n0 + j);
            v[j / 2] = f2_add(v[j / 2], bf16x2_to_f2(r.x));
            v[j / 2 + 1] = f2_add(v[j / 2 + 1], bf16x2_to_f2(r.y));
            v[j / 2 + 2] = f2_add(v[j / 2 + 2], bf16x2_to_f2(r.z));
            v[j / 2 + 3] = f2_add(v[j / 2 + 3], bf16x2_to_f2(r.w));
        }
    }
    if (ep.stats_out) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { ssum = f2_add(ssum, v[j]); ssq = f2_fma(v[j], v[j], ssq); }
    }
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = f2_to_bf16x2(v[j / 2]);
        o.y = f2_to_bf16x2(v[j / 2 + 1]);
        o.z = f2_to_bf16x2(v[j / 2 + 2]);
        o.w = f2_to_bf16x2(v[j / 2 + 3]);
        *reinterpret_cast<uint4*>(out_ptr + n0 + j) = o;
    }
}

// (sum, sum of squares) of a row segment -> its statistics slot
__device__ __forceinline__ void store_row_stats(const GemmEpilogue& ep, long out_row, int seg, f32x2_t ssum, f32x2_t ssq) {
    float a, b, c, d;
    f2_unpack(ssum, a, b);
    f2_unpack(ssq, c, d);
    if (seg < LN_SLOTS) *reinterpret_cast<float2*>(ep.stats_out + (out_row * LN_SLOTS + seg) * 2) = make_float2(a + b, c + d);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         __nv_bfloat16* out, int ldc, int M, int N, int K, GemmEpilogue ep) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)   /* pointer arithmetic keeps the shared address space: LDS/STS, not generic LD/ST */;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                      // [STAGES]
    uint64_t* empty_bar = bars + STAGES;            // [STAGES]
    uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_blocks = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    const int n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = m_blocks * n_blocks;
    const int k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], GEMM_EPI_WARPS);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc<1>(tmem_slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_blocks;
                const int n_blk = tile - m_blk * n_blocks;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K,
                                m_blk * GEMM_BLOCK_M);
                    tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_w, &full_bar[stage], kb * GEMM_BLOCK_K,
                                n_blk * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (>>4) address field
                        umma_bf16<1>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full_bar[as]);     // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (2..2+GEMM_EPI_WARPS) =====================
        // warp w may read TMEM lanes [32*(w%4), +32); two warps share each lane quarter and split the columns.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BLOCK_N / (GEMM_EPI_WARPS / 4);
        constexpr int NCH = COLS_PER_WARP / 32;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / n_blocks;
            const int n_blk = tile - m_blk * n_blocks;
            const int row = m_blk * GEMM_BLOCK_M + q * 32 + lane;
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
            const int col0 = n_blk * BLOCK_N + half * COLS_PER_WARP;

            float mean, rstd;
            f32x2_t ssum = f2_pack(0.f, 0.f), ssq = f2_pack(0.f, 0.f);
            ln_row_stats(ep, out_row, row_ok, mean, rstd);

            mbar_wait(&tmem_full_bar[as], aphase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + half * COLS_PER_WARP;
            // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is processed
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
            if (ep.stats_out && row_ok && col0 < N) store_row_stats(ep, out_row, col0 / COLS_PER_WARP, ssum, ssq);
            // this warp is done reading the accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace b200
