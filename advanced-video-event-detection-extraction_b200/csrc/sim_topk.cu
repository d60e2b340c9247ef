// K4: image x text cosine scores fused with top-k, threshold and clip-interval emission.
//
// Replaces (reference file:line)
//   OpenCLIPModel.compute_similarity          src/models/openclip_model.py:212-214   (np.dot)
//   np.argsort(s)[::-1][:top_k] + threshold   src/pipeline/phase1_mvp.py:145-155     (ties -> higher index first)
//   extract_clip_with_padding / extract_clip  src/services/clip_extractor.py:175-183, 94-111 (interval clamps)
// restated in oracle/phase1_ref.py.
//
// This is the HBM-bound form (few queries): every embedding row is streamed once with 128-bit loads, each
// warp keeps a sorted top-k list per query (entries spread over lanes, inserted with ballot + shuffle), lists
// are merged per CTA and then by a final kernel that also applies the threshold and writes the intervals.
// The same final kernel merges candidate lists gathered from other ranks (multi-GPU path).
#include <math.h>

#include <stdlib.h>

#include "internal.h"
#include "ptx.cuh"
#include "sim_topk_tc.cuh"

using namespace b200;

constexpr int SIM_WARPS = 8;
constexpr int SIM_QTILE = 16;   // queries per pass of the streaming kernel
constexpr int SIM_MAXK = 32;          // result slots per pass (one lane each)
constexpr int SIM_MAXK_TOTAL = 4096;  // top_k served by ceil(k / 32) passes

__device__ __forceinline__ bool beats(float sa, int ia, float sb, int ib) {
    // descending score; equal scores -> higher index first (np.argsort(...)[::-1])
    return sa > sb || (sa == sb && ia > ib);
}

// Insert (s, i) into a sorted list of k entries held in shared memory; lane j owns entry j.
__device__ __forceinline__ void warp_insert(float* ls, int* li, int k, float s, int i, int lane) {
    const float es = lane < k ? ls[lane] : 0.f;
    const int ei = lane < k ? li[lane] : 0;
    const bool ahead = lane < k && beats(es, ei, s, i);
    const int p = __popc(__ballot_sync(0xffffffffu, ahead));  // entries that stay in front
    const float ps = __shfl_up_sync(0xffffffffu, es, 1);
    const int pi = __shfl_up_sync(0xffffffffu, ei, 1);
    if (lane < k) {
        if (lane == p) { ls[lane] = s; li[lane] = i; }
        else if (lane > p) { ls[lane] = ps; li[lane] = pi; }
    }
    __syncwarp();
}

// MAXC: 16-byte chunks of a row per lane (ceil(e / VEC / 32)); a template parameter so that the row registers are
// sized for the actual embedding width (E = 512: 2 for bf16, 4 for fp32) and two CTAs fit an SM
template <bool BF16, int MAXC>
__global__ void __launch_bounds__(SIM_WARPS * 32, 2)
sim_stream_kernel(const void* __restrict__ img, int64_t n, int e, const float* __restrict__ txt, int q0, int qn,
                  int q_total, int k, float* __restrict__ scores_out, float* __restrict__ part_s,
                  int* __restrict__ part_i, const float* __restrict__ ceil_s, const int64_t* __restrict__ ceil_i,
                  int ceil_stride, int64_t index_base) {
    extern __shared__ __align__(16) float sm[];
    float* sq = sm;                                                  // [qn][e]
    float* ls = sq + static_cast<size_t>(qn) * e;                    // [SIM_WARPS][qn][k]
    int* li = reinterpret_cast<int*>(ls + static_cast<size_t>(SIM_WARPS) * qn * k);
    // ceiling of a continuation pass (top_k > SIM_MAXK runs ceil(k / 32) passes): only rows that come strictly AFTER
    // the last entry the previous pass emitted are eligible.  First pass: (+inf, INT_MAX) -> every row is eligible.
    float* cl_s = reinterpret_cast<float*>(li + static_cast<size_t>(SIM_WARPS) * qn * k);   // [qn]
    int* cl_i = reinterpret_cast<int*>(cl_s + qn);                                          // [qn]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < qn * e; i += blockDim.x) sq[i] = txt[static_cast<size_t>(q0) * e + i];
    for (int i = threadIdx.x; i < SIM_WARPS * qn * k; i += blockDim.x) { ls[i] = -INFINITY; li[i] = -1; }
    for (int i = threadIdx.x; i < qn; i += blockDim.x) {
        float cs = INFINITY;
        int ci = 0x7fffffff;
        if (ceil_i) {
            const int64_t g = ceil_i[static_cast<size_t>(q0 + i) * ceil_stride];
            cs = g < 0 ? -INFINITY : ceil_s[static_cast<size_t>(q0 + i) * ceil_stride];
            ci = g < 0 ? -1 : static_cast<int>(g - index_base);
        }
        cl_s[i] = cs; cl_i[i] = ci;
    }
    __syncthreads();
    float* wls = ls + static_cast<size_t>(warp) * qn * k;
    int* wli = li + static_cast<size_t>(warp) * qn * k;

    constexpr int VEC = BF16 ? 8 : 4;
    const int chunks = e / VEC;                  // 16-byte chunks per row
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * SIM_WARPS;
    const int64_t wid = static_cast<int64_t>(blockIdx.x) * SIM_WARPS + warp;
    // rows in flight per warp: as many as 64 data registers hold (E = 512: 8 bf16 rows = 8 KB, 2 fp32 rows = 4 KB).
    // With two rows the bf16 stream was latency bound at 48 % of the HBM peak while fp32 reached 84 %.
    constexpr int ROWS = MAXC <= 1 ? 8 : (BF16 ? (MAXC <= 2 ? 8 : 4) : (MAXC <= 2 ? 4 : 2));   // <= 64 data registers
    for (int64_t row = wid * ROWS; row < n; row += warps_total * ROWS) {
        uint4 d[ROWS][MAXC];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int64_t rr = row + r < n ? row + r : row;      // tail: re-read row 0 of the group, result unused
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(img) +
                                                             rr * static_cast<int64_t>(e) * (BF16 ? 2 : 4));
#pragma unroll
            for (int j = 0; j < MAXC; ++j) {
                const int c = lane + j * 32;
                if (c < chunks) d[r][j] = __ldg(rp + c);
            }
        }
        for (int qi = 0; qi < qn; ++qi) {
            const float* qv = sq + static_cast<size_t>(qi) * e;
            float acc[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
#pragma unroll
            for (int j = 0; j < MAXC; ++j) {
                const int c = lane + j * 32;
                if (c < chunks) {
                    const float4 t0 = *reinterpret_cast<const float4*>(qv + c * VEC);
                    if (BF16) {
                        const float4 t1 = *reinterpret_cast<const float4*>(qv + c * VEC + 4);
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            float2 f;
                            float a0 = acc[r];
                            f = unpack_bf16x2(d[r][j].x); a0 = fmaf(f.x, t0.x, a0); a0 = fmaf(f.y, t0.y, a0);
                            f = unpack_bf16x2(d[r][j].y); a0 = fmaf(f.x, t0.z, a0); a0 = fmaf(f.y, t0.w, a0);
                            f = unpack_bf16x2(d[r][j].z); a0 = fmaf(f.x, t1.x, a0); a0 = fmaf(f.y, t1.y, a0);
                            f = unpack_bf16x2(d[r][j].w); a0 = fmaf(f.x, t1.z, a0); a0 = fmaf(f.y, t1.w, a0);
                            acc[r] = a0;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            float a0 = acc[r];
                            a0 = fmaf(__uint_as_float(d[r][j].x), t0.x, a0); a0 = fmaf(__uint_as_float(d[r][j].y), t0.y, a0);
                            a0 = fmaf(__uint_as_float(d[r][j].z), t0.z, a0); a0 = fmaf(__uint_as_float(d[r][j].w), t0.w, a0);
                            acc[r] = a0;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (row + r >= n) break;                         // warp-uniform
                if (scores_out && lane == 0) scores_out[(row + r) * q_total + q0 + qi] = acc[r];
                if (k > 0) {
                    float* l_s = wls + qi * k;
                    int* l_i = wli + qi * k;
                    if (beats(acc[r], static_cast<int>(row + r), l_s[k - 1], l_i[k - 1]) &&
                        beats(cl_s[qi], cl_i[qi], acc[r], static_cast<int>(row + r)))
                        warp_insert(l_s, l_i, k, acc[r], static_cast<int>(row + r), lane);
                }
            }
        }
    }
    if (k == 0) return;
    __syncthreads();
    // merge the SIM_WARPS lists of each query into warp 0's slot, queries round-robin over warps
    for (int qi = warp; qi < qn; qi += SIM_WARPS) {
        float* dst_s = ls + qi * k;   // warp 0's list for query qi
        int* dst_i = li + qi * k;
        for (int w = 1; w < SIM_WARPS; ++w) {
            const float* src_s = ls + (static_cast<size_t>(w) * qn + qi) * k;
            const int* src_i = li + (static_cast<size_t>(w) * qn + qi) * k;
            for (int j = 0; j < k; ++j) {
                const float s = src_s[j];
                const int i = src_i[j];
                if (i < 0) break;
                if (!beats(s, i, dst_s[k - 1], dst_i[k - 1])) break;  // sorted: nothing further can enter
                warp_insert(dst_s, dst_i, k, s, i, lane);
            }
        }
        if (lane < k) {
            part_s[(static_cast<size_t>(blockIdx.x) * q_total + q0 + qi) * k + lane] = dst_s[lane];
            part_i[(static_cast<size_t>(blockIdx.x) * q_total + q0 + qi) * k + lane] = dst_i[lane];
        }
    }
}

// Final merge: one CTA per query merges g sorted candidate lists, applies the threshold and emits intervals.
// IDX64: candidates carry int64 global indices (multi-GPU merge) instead of int32 local ones.
// One launch emits kp <= SIM_MAXK entries into output slots [off, off + kp) of the ko-wide result rows; top_k > 32 is
// served by consecutive launches (off = 0, 32, 64, ...): a continuation pass only admits candidates that come
// strictly after the entry in slot off - 1 (descending score, ties -> higher index first), so the concatenation is
// exactly np.argsort(s)[::-1][:k] (src/pipeline/phase1_mvp.py:145).  lk = length of every candidate list.
template <bool IDX64>
__global__ void __launch_bounds__(SIM_WARPS * 32)
topk_final_kernel(const float* __restrict__ cand_s, const void* __restrict__ cand_i, int64_t ls_s, int64_t ls_i, int g,
                  int q_total, int lk, int kp, int off, int ko, float thr, const double* __restrict__ ts, int64_t index_base, double clip_dur,
                  double vid_dur, float* __restrict__ top_s, int64_t* __restrict__ top_i, double* __restrict__ intervals,
                  int32_t* __restrict__ counts) {
    __shared__ float ls[SIM_WARPS][SIM_MAXK];
    __shared__ long long gi[SIM_WARPS][SIM_MAXK];  // candidate index (global), the tie-break key
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = kp;
    if (lane < k) { ls[warp][lane] = -INFINITY; gi[warp][lane] = -1; }
    __syncwarp();
    float ceil_s = INFINITY;
    long long ceil_i = 0x7fffffffffffffffLL;
    if (off > 0) {
        ceil_i = top_i[static_cast<size_t>(q) * ko + off - 1];
        ceil_s = ceil_i < 0 ? -INFINITY : top_s[static_cast<size_t>(q) * ko + off - 1];
    }
    // list `l` starts ls_s floats / ls_i indices behind list l - 1 (partial lists: q_total * lk; gathered messages: the
    // message pitch)
    (void)q_total;
    auto cand_index = [&](int list, int j) -> long long {
        const size_t o = static_cast<size_t>(list) * ls_i + static_cast<size_t>(q) * lk + j;
        return IDX64 ? static_cast<const long long*>(cand_i)[o] : static_cast<long long>(static_cast<const int*>(cand_i)[o]);
    };
    // 64-bit aware insert (tie-break on the global index)
    auto insert = [&](int w, float s, long long idx) {
        const float es = lane < k ? ls[w][lane] : 0.f;
        const long long ei = lane < k ? gi[w][lane] : 0;
        const bool ahead = lane < k && (es > s || (es == s && ei > idx));
        const int p = __popc(__ballot_sync(0xffffffffu, ahead));
        const float ps = __shfl_up_sync(0xffffffffu, es, 1);
        const long long pi = __shfl_up_sync(0xffffffffu, ei, 1);
        if (lane < k) {
            if (lane == p) { ls[w][lane] = s; gi[w][lane] = idx; }
            else if (lane > p) { ls[w][lane] = ps; gi[w][lane] = pi; }
        }
        __syncwarp();
    };
    for (int list = warp; list < g; list += SIM_WARPS) {
        for (int j = 0; j < lk; ++j) {
            const float s = cand_s[static_cast<size_t>(list) * ls_s + static_cast<size_t>(q) * lk + j];
            long long idx = cand_index(list, j);
            if (idx < 0) break;
            idx += index_base;
            if (!(ceil_s > s || (ceil_s == s && ceil_i > idx))) continue;   // emitted by an earlier pass
            const float ws = ls[warp][k - 1];
            const long long wi = gi[warp][k - 1];
            if (!(s > ws || (s == ws && idx > wi))) break;
            insert(warp, s, idx);
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < SIM_WARPS; ++w) {
            for (int j = 0; j < k; ++j) {
                const float s = ls[w][j];
                const long long idx = gi[w][j];
                if (idx < 0) break;
                const float ws = ls[0][k - 1];
                const long long wi = gi[0][k - 1];
                if (!(s > ws || (s == ws && idx > wi))) break;
                insert(0, s, idx);
            }
        }
        if (lane < k) {
            const float s = ls[0][lane];
            const long long gidx = gi[0][lane];
            const size_t o = static_cast<size_t>(q) * ko + off + lane;
            top_s[o] = s;
            top_i[o] = gidx;
            double start = 0.0, end = 0.0;
            if (gidx >= 0) {
                // clip_extractor.py:175-183 then :94-111
                const double t = ts ? ts[gidx] : static_cast<double>(gidx);
                start = fmax(0.0, t - clip_dur / 2);
                end = t + clip_dur / 2;
                if (end <= start) end = start + 5.0;
                if (vid_dur > 0.0) {
                    if (start >= vid_dur) { start = fmax(0.0, vid_dur - 5.0); end = vid_dur; }
                    else if (end > vid_dur) end = vid_dur;
                }
            }
            if (intervals) { intervals[o * 2] = start; intervals[o * 2 + 1] = end; }
        }
        const bool pass = lane < k && gi[0][lane] >= 0 && ls[0][lane] >= thr;
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (lane == 0 && counts) counts[q] = (off > 0 ? counts[q] : 0) + __popc(m);
    }
}

// Tensor-core path epilogue: the tcgen05 kernel SELECTS candidates on bf16-rounded queries; this kernel re-scores the
// k survivors of each query with the fp32 query in exactly sim_stream_kernel's arithmetic (same lane/chunk layout, same
// fma order, same butterfly reduction), re-sorts them and re-applies the threshold, so that the confidences a caller
// sees -- and every threshold decision -- do not depend on how many queries shared the batch.  One warp per query.
template <bool BF16>
__global__ void __launch_bounds__(32)
rescore_sort_kernel(const void* __restrict__ img, int e, const float* __restrict__ txt, int k, float thr,
                    const double* __restrict__ ts, int64_t index_base, double clip_dur, double vid_dur,
                    float* __restrict__ top_s, int64_t* __restrict__ top_i, double* __restrict__ intervals,
                    int32_t* __restrict__ counts) {
    const int q = blockIdx.x, lane = threadIdx.x;
    constexpr int VEC = BF16 ? 8 : 4;
    const int chunks = e / VEC;
    const float* qv = txt + static_cast<size_t>(q) * e;
    float my_s = -INFINITY;
    long long my_i = -1;
    for (int j = 0; j < k; ++j) {
        const long long gidx = top_i[static_cast<size_t>(q) * k + j];
        if (gidx < 0) continue;                              // warp-uniform
        const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(img) +
                                                         (gidx - index_base) * static_cast<int64_t>(e) * (BF16 ? 2 : 4));
        float acc = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            const uint4 d = __ldg(rp + c);
            const float4 t0 = *reinterpret_cast<const float4*>(qv + c * VEC);
            if (BF16) {
                const float4 t1 = *reinterpret_cast<const float4*>(qv + c * VEC + 4);
                float2 f;
                f = unpack_bf16x2(d.x); acc = fmaf(f.x, t0.x, acc); acc = fmaf(f.y, t0.y, acc);
                f = unpack_bf16x2(d.y); acc = fmaf(f.x, t0.z, acc); acc = fmaf(f.y, t0.w, acc);
                f = unpack_bf16x2(d.z); acc = fmaf(f.x, t1.x, acc); acc = fmaf(f.y, t1.y, acc);
                f = unpack_bf16x2(d.w); acc = fmaf(f.x, t1.z, acc); acc = fmaf(f.y, t1.w, acc);
            } else {
                acc = fmaf(__uint_as_float(d.x), t0.x, acc); acc = fmaf(__uint_as_float(d.y), t0.y, acc);
                acc = fmaf(__uint_as_float(d.z), t0.z, acc); acc = fmaf(__uint_as_float(d.w), t0.w, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == j) { my_s = acc; my_i = gidx; }
    }
    // rank of my entry among the k (descending score, ties -> higher index first); empty slots sort last
    int rank = 0;
    for (int j = 0; j < k; ++j) {
        const float s = __shfl_sync(0xffffffffu, my_s, j);
        const long long i = __shfl_sync(0xffffffffu, my_i, j);
        if (lane < k && j != lane) {
            const bool ahead = my_i < 0 ? (i >= 0 || j < lane) : (i >= 0 && (s > my_s || (s == my_s && i > my_i)));
            rank += ahead ? 1 : 0;
        }
    }
    if (lane < k) {
        const size_t o = static_cast<size_t>(q) * k + rank;
        top_s[o] = my_s;
        top_i[o] = my_i;
        double start = 0.0, end = 0.0;
        if (my_i >= 0) {                                     // clip_extractor.py:175-183 then :94-111
            const double t = ts ? ts[my_i] : static_cast<double>(my_i);
            start = fmax(0.0, t - clip_dur / 2);
            end = t + clip_dur / 2;
            if (end <= start) end = start + 5.0;
            if (vid_dur > 0.0) {
                if (start >= vid_dur) { start = fmax(0.0, vid_dur - 5.0); end = vid_dur; }
                else if (end > vid_dur) end = vid_dur;
            }
        }
        if (intervals) { intervals[o * 2] = start; intervals[o * 2 + 1] = end; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, lane < k && my_i >= 0 && my_s >= thr);
    if (lane == 0 && counts) counts[q] = __popc(m);
}

static int sim_grid(b200clip_handle* h, int64_t n) {
    int64_t g = (n + SIM_WARPS * 2 - 1) / (SIM_WARPS * 2);
    const int64_t cap = static_cast<int64_t>(h->num_sms) * 4;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

static int run_stream(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q,
                      int k, float* scores, float* part_s, int* part_i, int grid, cudaStream_t st,
                      const float* ceil_s = nullptr, const int64_t* ceil_i = nullptr, int ceil_stride = 0,
                      int64_t index_base = 0) {
    for (int q0 = 0; q0 < q; q0 += SIM_QTILE) {
        const int qn = (q - q0) < SIM_QTILE ? (q - q0) : SIM_QTILE;
        const size_t smem = static_cast<size_t>(qn) * e * 4 + static_cast<size_t>(SIM_WARPS) * qn * (k > 0 ? k : 0) * 8 +
                            static_cast<size_t>(qn) * 8;
        const int vec = dtype == B200CLIP_BF16 ? 8 : 4;
        const int cpl = (e / vec + 31) / 32;                     // chunks per lane
        void (*kern)(const void*, int64_t, int, const float*, int, int, int, int, float*, float*, int*, const float*,
                     const int64_t*, int, int64_t) = nullptr;
        if (dtype == B200CLIP_BF16) kern = cpl <= 1 ? sim_stream_kernel<true, 1> : cpl <= 2 ? sim_stream_kernel<true, 2> : sim_stream_kernel<true, 4>;
        else kern = cpl <= 2 ? sim_stream_kernel<false, 2> : cpl <= 4 ? sim_stream_kernel<false, 4> : sim_stream_kernel<false, 8>;
        if (smem > 48 * 1024) B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        kern<<<grid, SIM_WARPS * 32, smem, st>>>(img, n, e, txt, q0, qn, q, k, scores, part_s, part_i, ceil_s, ceil_i,
                                                 ceil_stride, index_base);
        h->launches++;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

static int check_sim_args(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q) {
    if ((!img && n > 0) || !txt) return b200_fail(h, B200CLIP_E_ARG, "similarity: null embedding pointer");
    if (dtype != B200CLIP_F32 && dtype != B200CLIP_BF16) return b200_fail(h, B200CLIP_E_ARG, "similarity: bad dtype");
    if (n < 0 || n >= (int64_t(1) << 31)) return b200_fail(h, B200CLIP_E_SHAPE, "similarity: n out of range");
    if (e <= 0 || e % 8 != 0 || e > 1024) return b200_fail(h, B200CLIP_E_SHAPE, "similarity: e must be a multiple of 8, <= 1024");
    if (q <= 0) return b200_fail(h, B200CLIP_E_ARG, "similarity: q must be positive");
    return 0;
}

int launch_similarity(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q,
                      float* scores, cudaStream_t st) {
    int rc = check_sim_args(h, img, dtype, n, e, txt, q);
    if (rc) return rc;
    if (n == 0) return 0;
    ProfScope ps(h, PROF_SIM, static_cast<double>(n) * e * (dtype == B200CLIP_BF16 ? 2.0 : 4.0) + static_cast<double>(n) * q * 4.0, st);
    return run_stream(h, img, dtype, n, e, txt, q, 0, scores, nullptr, nullptr, sim_grid(h, n), st);
}

int make_tmap_bf16_2d(b200clip_handle* h, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz);

// queries fp32 -> bf16 (the tensor-core operand) and the per-query global k-th-score bounds reset to -inf
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n, int* gthr, int q) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        dst[i] = __float2bfloat16(src[i]);
        if (i < q) gthr[i] = b200::ord_encode(-INFINITY);
    }
}

// fp32 -> bf16 cast of a cache slice, 8 elements per thread (two 16-byte loads, one 16-byte store); n % 8 == 0
__global__ void __launch_bounds__(256)
cast_f32_bf16_x8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int64_t n8) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 a = __ldcs(src + 2 * i), b = __ldcs(src + 2 * i + 1);
        uint4 o;
        o.x = b200::pack_bf16x2(a.x, a.y); o.y = b200::pack_bf16x2(a.z, a.w);
        o.z = b200::pack_bf16x2(b.x, b.y); o.w = b200::pack_bf16x2(b.z, b.w);
        dst[i] = o;
    }
}

static int ensure_topk_ws(b200clip_handle* h, size_t need, cudaStream_t st) {
    if (need <= h->ws_topk_bytes) return 0;
    B200_CUDA(h, cudaStreamSynchronize(st));
    if (h->ws_topk) cudaFree(h->ws_topk);
    h->ws_topk = nullptr; h->ws_topk_bytes = 0;
    B200_CUDA(h, cudaMalloc(&h->ws_topk, need));
    h->ws_topk_bytes = need;
    return 0;
}

// Tensor-core path (sim_topk_tc.cuh): many queries, small k.  A bf16 cache is consumed in place; an fp32 cache (what
// the reference holds) is converted to bf16 slice by slice into a scratch buffer and each slice is scored right away
// (one extra pass over the data instead of Q/8 passes of the streaming kernel); the global k-th-score bounds carry
// over from slice to slice, every slice appends its candidate lists and one final merge sees them all.
constexpr int64_t STC_SLICE_ROWS = int64_t(1) << 18;

static int launch_sim_topk_tc(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q,
                              int k, float thr, const double* ts, int64_t index_base, double clip_dur, double vid_dur,
                              float* top_scores, int64_t* top_idx, double* intervals, int32_t* counts, float* dense,
                              cudaStream_t st) {
    const bool f32 = dtype != B200CLIP_BF16;
    const int64_t slice = f32 ? (n < STC_SLICE_ROWS ? n : STC_SLICE_ROWS) : n;
    const int n_slices = static_cast<int>((n + slice - 1) / slice);
    const int tiles = static_cast<int>((slice + b200::G2_BLOCK_N - 1) / b200::G2_BLOCK_N);
    int clusters = h->num_sms / 2;
    if (tiles < clusters) clusters = tiles;
    const int g = clusters * 2;      // one candidate list per (cluster, column half) and query, per slice
    const size_t txt_bytes = (static_cast<size_t>(q) * e * 2 + 255) & ~size_t(255);
    const size_t thr_bytes = (static_cast<size_t>(q) * 4 + 255) & ~size_t(255);
    const size_t cast_bytes = f32 ? ((static_cast<size_t>(slice) * e * 2 + 255) & ~size_t(255)) : 0;
    const size_t list_elems = static_cast<size_t>(g) * n_slices * q * k;
    const size_t need = txt_bytes + thr_bytes + cast_bytes + list_elems * 8;
    int rc = ensure_topk_ws(h, need, st);
    if (rc) return rc;
    uint8_t* wsb = static_cast<uint8_t*>(h->ws_topk);
    bf16* txt16 = reinterpret_cast<bf16*>(wsb);
    int* gthr = reinterpret_cast<int*>(wsb + txt_bytes);
    bf16* cast = reinterpret_cast<bf16*>(wsb + txt_bytes + thr_bytes);
    float* part_s = reinterpret_cast<float*>(wsb + txt_bytes + thr_bytes + cast_bytes);
    int* part_i = reinterpret_cast<int*>(part_s + list_elems);
    ProfScope ps(h, PROF_SIM, static_cast<double>(n) * e * (f32 ? 4.0 : 2.0) + static_cast<double>(q) * (e * 4.0 + k * 12.0), st);
    f32_to_bf16_kernel<<<(static_cast<int64_t>(q) * e + 255) / 256, 256, 0, st>>>(txt, txt16, static_cast<int64_t>(q) * e, gthr, q);
    h->launches++;
    CUtensorMap ta, tw;
    // A (M side, 128 rows per CTA) = the queries, B (N side, 128 rows per CTA per tile) = the embedding rows
    if ((rc = make_tmap_bf16_2d(h, &ta, txt16, q, e, e, b200::GEMM_BLOCK_M, b200::GEMM_BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if (!(h->attr_done & ATTR_SIM_TC)) {
        B200_CUDA(h, cudaFuncSetAttribute(b200::sim_topk_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          b200::G2_SMEM_BYTES));
        B200_CUDA(h, cudaFuncSetAttribute(b200::sim_topk_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          b200::G2_SMEM_BYTES));
        h->attr_done |= ATTR_SIM_TC;
    }
    const bool no_ares = b200_knobs().sim_stream_a;
    auto kern = (e <= 512 && !no_ares) ? b200::sim_topk_tc_kernel<true> : b200::sim_topk_tc_kernel<false>;
    for (int si = 0; si < n_slices; ++si) {
        const int64_t r0 = si * slice;
        const int64_t nr = (n - r0) < slice ? (n - r0) : slice;
        const void* rows = img;
        if (f32) {
            const int64_t n8 = nr * e / 8;
            const int64_t blocks = (n8 + 255) / 256;
            cast_f32_bf16_x8_kernel<<<static_cast<unsigned>(blocks < h->num_sms * 16 ? blocks : h->num_sms * 16), 256, 0, st>>>(
                reinterpret_cast<const float4*>(static_cast<const float*>(img) + r0 * e), reinterpret_cast<uint4*>(cast), n8);
            h->launches++;
            rows = cast;
        }
        if ((rc = make_tmap_bf16_2d(h, &tw, rows, nr, e, e, b200::G2_HALF_N, b200::GEMM_BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
        const size_t loff = static_cast<size_t>(si) * g * q * k;
        for (int q0 = 0; q0 < q; q0 += b200::G2_BLOCK_N) {
            kern<<<2 * clusters, b200::GEMM_THREADS, b200::G2_SMEM_BYTES, st>>>(
                ta, tw, static_cast<int>(nr), static_cast<int>(r0), q, q0, q, e, k,
                dense ? dense + static_cast<size_t>(r0) * q : nullptr, part_s + loff, part_i + loff, gthr);
            h->launches++;
        }
    }
    topk_final_kernel<false><<<q, SIM_WARPS * 32, 0, st>>>(part_s, part_i, static_cast<int64_t>(q) * k, static_cast<int64_t>(q) * k,
                                                          g * n_slices, q, k, k, 0, k, thr, ts, index_base,
                                                          clip_dur, vid_dur, top_scores, top_idx, intervals, counts);
    h->launches++;
    if (!dense) {
        // (the dense test hook keeps the scores exactly as the tensor-core kernel computed them)
        if (f32)
            rescore_sort_kernel<false><<<q, 32, 0, st>>>(img, e, txt, k, thr, ts, index_base, clip_dur, vid_dur, top_scores,
                                                        top_idx, intervals, counts);
        else
            rescore_sort_kernel<true><<<q, 32, 0, st>>>(img, e, txt, k, thr, ts, index_base, clip_dur, vid_dur, top_scores,
                                                       top_idx, intervals, counts);
        h->launches++;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

int launch_sim_topk(b200clip_handle* h, const void* img, int dtype, int64_t n, int e, const float* txt, int q, int k,
                    float thr, const double* ts, int64_t index_base, double clip_dur, double vid_dur,
                    float* top_scores, int64_t* top_idx, double* intervals, int32_t* counts, float* dense,
                    cudaStream_t st) {
    int rc = check_sim_args(h, img, dtype, n, e, txt, q);
    if (rc) return rc;
    // many queries over a large bf16 cache: tensor-core similarity with the top-k fused into the epilogue
    const bool force_simt = b200_knobs().sim_simt;
    // (an fp32 cache takes this path only from 16 queries on: below that the streaming kernel reads it once anyway
    // and keeps full fp32 scores)
    if (!force_simt && e % 64 == 0 && k >= 1 && k <= b200::STC_MAXK && n >= 4096 && n < (int64_t(1) << 31) &&
        (dtype == B200CLIP_BF16 ? q >= 8 : q >= 16) && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && top_scores && top_idx)
        return launch_sim_topk_tc(h, img, dtype, n, e, txt, q, k, thr, ts, index_base, clip_dur, vid_dur, top_scores,
                                  top_idx, intervals, counts, dense, st);
    if (k <= 0 || k > SIM_MAXK_TOTAL)
        return b200_fail(h, B200CLIP_E_SHAPE, "sim_topk: k must be in [1, %d]", SIM_MAXK_TOTAL);
    if (!top_scores || !top_idx) return b200_fail(h, B200CLIP_E_ARG, "sim_topk: null output");
    const int grid = sim_grid(h, n > 0 ? n : 1);
    const int kp_max = k < SIM_MAXK ? k : SIM_MAXK;
    const int passes = (k + SIM_MAXK - 1) / SIM_MAXK;      // the embeddings are streamed once per 32 result slots
    ProfScope ps(h, PROF_SIM, passes * static_cast<double>(n) * e * (dtype == B200CLIP_BF16 ? 2.0 : 4.0) + static_cast<double>(q) * (e * 4.0 + k * 12.0), st);
    const size_t need = static_cast<size_t>(grid) * q * kp_max * 8;
    int rc2 = ensure_topk_ws(h, need, st);
    if (rc2) return rc2;
    for (int off = 0; off < k; off += SIM_MAXK) {
        const int kp = (k - off) < SIM_MAXK ? (k - off) : SIM_MAXK;
        float* part_s = static_cast<float*>(h->ws_topk);
        int* part_i = reinterpret_cast<int*>(part_s + static_cast<size_t>(grid) * q * kp);
        int g = grid;
        if (n == 0 || off >= n) {
            g = 0;      // no rows (left): the final kernel writes empty slots (-inf, -1)
        } else {
            rc = run_stream(h, img, dtype, n, e, txt, q, kp, off == 0 ? dense : nullptr, part_s, part_i, grid, st,
                            off > 0 ? top_scores + off - 1 : nullptr, off > 0 ? top_idx + off - 1 : nullptr, k, index_base);
            if (rc) return rc;
        }
        topk_final_kernel<false><<<q, SIM_WARPS * 32, 0, st>>>(part_s, part_i, static_cast<int64_t>(q) * kp,
                                                              static_cast<int64_t>(q) * kp, g, q, kp, kp, off, k, thr, ts, index_base,
                                                              clip_dur, vid_dur, top_scores, top_idx, intervals, counts);
        h->launches++;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

// cs / ci: g candidate lists of [q, k] entries each, list l at cs + l * ls_s (floats) / ci + l * ls_i (int64)
int launch_topk_merge_strided(b200clip_handle* h, const float* cs, const int64_t* ci, int64_t ls_s, int64_t ls_i, int g, int q,
                              int k, float thr, const double* ts, double clip_dur, double vid_dur, float* top_scores,
                              int64_t* top_idx, double* intervals, int32_t* counts, cudaStream_t st) {
    if (!cs || !ci || !top_scores || !top_idx) return b200_fail(h, B200CLIP_E_ARG, "topk_merge: null argument");
    if (g <= 0 || q <= 0 || k <= 0 || k > SIM_MAXK_TOTAL) return b200_fail(h, B200CLIP_E_SHAPE, "topk_merge: bad g/q/k");
    ProfScope ps(h, PROF_SIM, static_cast<double>(g) * q * k * 12.0, st);
    for (int off = 0; off < k; off += SIM_MAXK) {
        const int kp = (k - off) < SIM_MAXK ? (k - off) : SIM_MAXK;
        topk_final_kernel<true><<<q, SIM_WARPS * 32, 0, st>>>(cs, ci, ls_s, ls_i, g, q, k, kp, off, k, thr, ts, 0, clip_dur, vid_dur,
                                                             top_scores, top_idx, intervals, counts);
        h->launches++;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

int launch_topk_merge(b200clip_handle* h, const float* cs, const int64_t* ci, int g, int q, int k, float thr,
                      const double* ts, double clip_dur, double vid_dur, float* top_scores, int64_t* top_idx,
                      double* intervals, int32_t* counts, cudaStream_t st) {
    return launch_topk_merge_strided(h, cs, ci, static_cast<int64_t>(q) * k, static_cast<int64_t>(q) * k, g, q, k, thr, ts,
                                     clip_dur, vid_dur, top_scores, top_idx, intervals, counts, st);
}
