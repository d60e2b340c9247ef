// K1 stages A + B on the (legacy-path) integer tensor cores: cv2 INTER_AREA shrink and Pillow's horizontal fixed-point
// pass as chains of IMMA.16832 (mma.sync m16n8k32, u8/s8 operands, s32 accumulators -- exact integer arithmetic, so the
// results are the same bits as the CUDA-core kernels in preprocess.cu; B200 issues 2034 int8 MAC/clk/SM on this path,
// tools/probes/imma_probe.cu).  Included by preprocess.cu only.
//
// Why: the CUDA-core form (area_hpass_vfirst_kernel) is purely issue bound -- ~754 k warp instructions per 1080p frame,
// 58 % issue utilisation, the bulk-copy feed completely hidden.  Written as matrix products the
// three separable passes need half the instructions (375 M instead of 754 M per 1024 frames), and the weights (small
// integers) become constant A / B fragments:
//
//   stage 1  H^T [area col (x, c) | source row]  = Wx^T [15 area cols x 64 row bytes] . rows^T      (u8 x u8)
//            M tile = 5 area pixels x 3 channels (15 of the 16 MMA rows), K = the <= 64 bytes of the RGB source row under
//            them, N = 8 source rows.  The B fragment is 4 x 4 bytes of source row g per lane, straight out of the bulk-copy
//            ring (LDS.32, conflict free, no conversion); the weight fragments live in registers for the whole kernel.
//   stage 2  N2^T [area col | area row]          = H^T [.. x 32 source rows] . (2 Wy)^T + D          (u8 x u8, twice)
//            H <= 255 * Dx needs 16 bits: the accumulators of stage 1 are split into a low- and a high-byte plane with
//            three PRMT per 8 values and fed back as A fragments (the C layout of one MMA pairs up with the A layout of
//            the next when K, the source rows, is taken in the order the fragments hold them -- the Wy table is permuted
//            on the host to match).  The product runs as K = 16 steps (m16n8k16), one per finished pair of 8-row
//            blocks, with the two planes folded into one accumulator after each step, so a tile carries 6 live registers
//            through a group instead of 10.  N2 = 2 N + D, so the area sample is umulhi(N2, magic) >> s exactly like
//            the CUDA-core kernel (rint(N / D) for odd D); it is parked PLANAR (R, G, B planes) in shared memory.
//   stage 3  Pillow^T [output col | area row]    = Wp^T [16 output cols x 64 area cols] . plane^T     (3 byte planes)
//            the 22-bit signed fixed-point coefficients are split into three byte planes (u8, u8, s8); one set of weight
//            fragments (resident in shared memory, 3 KB per tile) serves the three colour planes.  (acc + 2^21) >> 22,
//            clipped; a tile's 8 rows x 48 bytes go through a per-warp scratch and leave as 16-byte stores to mid2.
//
// Work decomposition: a CTA is persistent and walks (strip, frame) items; a strip is `gps` groups of 8 area rows; a group
// is NB blocks of 8 source rows = one ring stage (8 bulk copies, one per row, issued by 8 lanes of the producer warp).
// Each of the NCW consumer warps owns TPW stage-1 tiles for the whole kernel, reads ITS 64-byte column window of every
// block and releases the block as soon as the bytes are in registers; the only cross-warp hand-off is the parked area
// rows of a group (NBUF = 4 buffers with an mbarrier pair each, the Pillow pass of group i deferred until after stage 2 of
// group i + PDEF, so no warp waits for the others in the steady state).  Ring row pitch = 16 mod 128 and parked-row
// pitch = 32 mod 128 bytes put the 8 rows a fragment load touches on distinct shared-memory banks.
#pragma once

struct MmaParams {
    const uint8_t* src; int64_t frame_stride, row_stride;
    uint8_t* mid2; int64_t mid2_frame_stride;
    int S, ny, nx;                 // Pillow output columns, area rows and area columns of the window
    int seg, pitch1, pitchA;       // bytes per bulk-copied row, ring row pitch (16 mod 128), parked-row pitch (32 mod 128)
    int ntx, npt;                  // stage-1 tiles (5 area pixels), Pillow tiles (16 output pixels)
    int gps, nstrips, ngroups, nitems, nst;
    uint32_t d, div_mul; int div_shift;
    const uint4* a1;               // [ntx][2 K steps][32 lanes]   stage-1 weight fragments
    const int* kb1;                // [ntx]                        first byte of the tile's K window inside a ring row
    const uint2* b2;               // [ngroups][32 lanes]          stage-2 (vertical) weight fragments, doubled
    const int2* gmeta;             // [ngroups]                    (first source row, source rows)
    const uint4* ap;               // [npt][3 planes][2 K steps][32 lanes]  Pillow coefficient fragments
    const int* kbp;                // [npt]                        first area column of the tile's K window
};

template <bool A_SIGNED>
__device__ __forceinline__ void imma16832(int (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    if constexpr (A_SIGNED)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity, int tag) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    uint32_t spins = 0;          // bounded: a protocol bug must trap, not hang the GPU
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) {
            printf("b200clip: K1 mma kernel barrier timeout tag=%d block=%d thread=%d\n", tag, (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    } while (!ok);
}

__device__ __forceinline__ void imma16816(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

constexpr int MMA_MAX_NST = 8;
constexpr int NBUF = 4;      // parked-row buffers (groups of 8 area rows)
constexpr int PDEF = 2;      // the Pillow pass of a group runs PDEF groups after its rows were parked
template <int NCW, int TPW, int NB>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) area_hpass_mma_kernel(const __grid_constant__ MmaParams P) {
    extern __shared__ __align__(128) uint8_t mma_smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(mma_smem);
    uint64_t* empty_bar = full_bar + MMA_MAX_NST;
    uint64_t* afull_bar = empty_bar + MMA_MAX_NST;
    uint64_t* aempty_bar = afull_bar + NBUF;
    uint8_t* ring = mma_smem + 256;
    const uint32_t blockbytes = 8u * static_cast<uint32_t>(P.pitch1);
    const uint32_t planebytes = 8u * static_cast<uint32_t>(P.pitchA) + 8u;    // + 8: the three colour planes of a pixel on distinct banks
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < P.nst; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NCW); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(&afull_bar[s], NCW); mbar_init(&aempty_bar[s], NCW); }
        fence_mbar_init();
    }
    {   // the Pillow coefficient fragments stay in shared memory for the whole kernel
        uint4* apd = reinterpret_cast<uint4*>(ring + P.nst * blockbytes + NBUF * 3u * planebytes);
        for (int i = tid; i < P.npt * 3 * 2 * 32; i += blockDim.x) apd[i] = __ldg(P.ap + i);
    }
    __syncthreads();

    if (w == NCW) {
        // ------------------------------------------------------------------ producer warp: lane r copies row r of a block
        const uint8_t* gbase = P.src;
        uint32_t s = 0, ph = 0;
        for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
            const int strip = item % P.nstrips;
            const int64_t f = item / P.nstrips;
            const uint8_t* gf = gbase + f * P.frame_stride;
            const int g0 = strip * P.gps, g1 = min(g0 + P.gps, P.ngroups);
            for (int gi = g0; gi < g1; ++gi) {
                const int2 meta = __ldg(P.gmeta + gi);
#pragma unroll 1
                for (int b = 0; b < NB; ++b) {
                    mbar_wait(&empty_bar[s], ph ^ 1u, 21);
                    const int rows_here = min(max(meta.y - 8 * b, 0), 8);
                    if (lane == 0) {
                        if (rows_here > 0) mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(rows_here) * static_cast<uint32_t>(P.seg));
                        else mbar_arrive(&full_bar[s]);
                    }
                    __syncwarp();
                    if (lane < rows_here)
                        bulk_load_1d(ring + s * blockbytes + static_cast<uint32_t>(lane) * static_cast<uint32_t>(P.pitch1),
                                     gf + static_cast<int64_t>(meta.x + 8 * b + lane) * P.row_stride, static_cast<uint32_t>(P.seg), &full_bar[s]);
                    if (++s == static_cast<uint32_t>(P.nst)) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    const int g = lane >> 2, t4 = lane & 3;
    // stage-1 tiles w, w + NCW, ...: weight fragments and the lane's byte offset inside a ring block.  A slot past the last
    // tile recomputes the last tile (no divergent code around the MMAs) and is masked where it would be parked.
    uint4 a1[TPW][2];
    uint32_t toff[TPW];
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
        const int tii = min(w + k * NCW, P.ntx - 1);
        a1[k][0] = __ldg(P.a1 + (tii * 2 + 0) * 32 + lane);
        a1[k][1] = __ldg(P.a1 + (tii * 2 + 1) * 32 + lane);
        toff[k] = static_cast<uint32_t>(g) * static_cast<uint32_t>(P.pitch1) + 4u * t4 + static_cast<uint32_t>(__ldg(P.kb1 + tii));
    }
    // where the accumulator rows of this lane (MMA rows g and g + 8 = area column (pixel, channel)) are parked: plane,
    // pixel inside the tile and the two area rows 2 t4, 2 t4 + 1 (columns of the accumulator)
    const uint32_t prow = 2u * t4 * static_cast<uint32_t>(P.pitchA);
    const uint32_t park0 = static_cast<uint32_t>(g % 3) * planebytes + static_cast<uint32_t>(g / 3) + prow;
    const uint32_t park1 = static_cast<uint32_t>((g + 8) % 3) * planebytes + static_cast<uint32_t>((g + 8) / 3) + prow;
    const int xlim0 = P.nx - g / 3, xlim1 = g == 7 ? -1 : P.nx - (g + 8) / 3;     // tile start x0 is parked iff x0 < xlim

    uint32_t full0 = smem_u32(full_bar), ring0 = smem_u32(ring), bbytes = blockbytes, pitchA = static_cast<uint32_t>(P.pitchA);
    uint32_t arows0 = ring0 + static_cast<uint32_t>(P.nst) * blockbytes;     // [NBUF][3 planes][8 rows][pitchA]
    uint32_t ap0 = arows0 + NBUF * 3u * planebytes;                         // Pillow fragments, then the per-warp scratch
    uint32_t afull0 = smem_u32(afull_bar), aempty0 = smem_u32(aempty_bar);
    asm volatile("" : "+r"(full0), "+r"(ring0), "+r"(bbytes), "+r"(arows0), "+r"(pitchA), "+r"(ap0));
    const uint32_t lane_offA = static_cast<uint32_t>(g) * pitchA + 8u * t4;
    const int S3 = P.S * 3;
    const uint32_t dinit = P.d, dmul = P.div_mul;
    const int dsh = P.div_shift;

    // Pillow pass over the parked rows of one group: buffer pp, `nvalid` rows, to `out` (row 0 of the group in mid2).
    // A tile's 8 rows x 16 pixels go through a per-warp scratch so that they leave as 16-byte stores.
    const uint32_t scr0 = ap0 + static_cast<uint32_t>(P.npt) * (3u * 2u * 32u * 16u) + static_cast<uint32_t>(w) * 384u;
    const uint32_t scr_w = scr0 + (2u * t4) * 48u + static_cast<uint32_t>(g) * 3u;     // (area row 2 t4, pixel g, channel 0)
    const uint32_t scr_r = scr0 + static_cast<uint32_t>(lane) * 16u;                   // row lane / 3, 16-byte part lane % 3
    auto pillow = [&](uint32_t pp, uint32_t wait_parity, uint8_t* out, int nvalid) {
        for (int pt = w; pt < P.npt; pt += NCW) {
            const uint32_t kbp = static_cast<uint32_t>(__ldg(P.kbp + pt));
            if (pt == w) mbar_wait_u32(afull0 + 8u * pp, wait_parity, 22);
            const uint32_t abase = arows0 + pp * 3u * planebytes + lane_offA + kbp;
            const uint32_t apl = ap0 + static_cast<uint32_t>(pt) * (3u * 2u * 32u * 16u) + static_cast<uint32_t>(lane) * 16u;
            uint4 ap[3][2];
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) ap[pl][ks] = lds128(apl + static_cast<uint32_t>(pl * 2 + ks) * 512u);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const uint2 b0 = lds64(abase + ch * planebytes);
                const uint2 b1 = lds64(abase + ch * planebytes + 32u);
                int c0[4] = {1 << 21, 1 << 21, 1 << 21, 1 << 21}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
                imma16832<false>(c0, ap[0][0], b0.x, b0.y);
                imma16832<false>(c1, ap[1][0], b0.x, b0.y);
                imma16832<true>(c2, ap[2][0], b0.x, b0.y);
                imma16832<false>(c0, ap[0][1], b1.x, b1.y);
                imma16832<false>(c1, ap[1][1], b1.x, b1.y);
                imma16832<true>(c2, ap[2][1], b1.x, b1.y);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int acc = c0[e] + (c1[e] << 8) + (c2[e] << 16);      // wraps like the 32-bit sum it stands for
                    const int v = min(max(acc >> 22, 0), 255);
                    sts8(scr_w + (e & 1) * 48u + (e >> 1) * 24u + ch, static_cast<uint32_t>(v));
                }
            }
            __syncwarp();
            if (lane < 24 && lane / 3 < nvalid) {
                const uint4 v = lds128(scr_r);
                *reinterpret_cast<uint4*>(out + (lane / 3) * S3 + pt * 48 + (lane % 3) * 16) = v;
            }
            __syncwarp();
        }
        if (lane == 0) mbar_arrive_u32(aempty0 + 8u * pp);
    };

    uint32_t s = 0, ph = 0;
    uint32_t gcount = 0;                  // groups finished by this CTA: parked-row buffer = gcount % NBUF
    uint8_t* out_q[PDEF];                 // mid2 rows of the groups whose Pillow pass is still owed (oldest first)
    int valid_q[PDEF];
#pragma unroll
    for (int i = 0; i < PDEF; ++i) { out_q[i] = nullptr; valid_q[i] = 0; }
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const int strip = item % P.nstrips;
        const int64_t f = item / P.nstrips;
        const int g0 = strip * P.gps, g1 = min(g0 + P.gps, P.ngroups);
        for (int gi = g0; gi < g1; ++gi) {
            const uint2 b2 = __ldg(P.b2 + gi * 32 + lane);
            uint32_t pA[TPW], pB[TPW];
            int acc[TPW][4];
#pragma unroll
            for (int k = 0; k < TPW; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = static_cast<int>(dinit);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                mbar_wait_u32(full0 + 8u * s, ph, 23);
                const uint32_t base = ring0 + s * bbytes;
                uint32_t bb[TPW][4];
#pragma unroll
                for (int k = 0; k < TPW; ++k)
#pragma unroll
                    for (int q = 0; q < 4; ++q) bb[k][q] = lds32(base + toff[k] + 16u * q);
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(full0 + 8u * (MMA_MAX_NST + s));     // empty_bar[s]: bytes are in registers
                if (++s == static_cast<uint32_t>(P.nst)) { s = 0; ph ^= 1u; }
                int c[TPW][4];
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0;
                    imma16832<false>(c[k], a1[k][0], bb[k][0], bb[k][1]);
                }
#pragma unroll
                for (int k = 0; k < TPW; ++k) imma16832<false>(c[k], a1[k][1], bb[k][2], bb[k][3]);
#pragma unroll
                for (int k = 0; k < TPW; ++k) {
                    // (row g | rows 2 t4, 2 t4 + 1 of the block) and (row g + 8 | same): two 16-bit sums per word
                    const uint32_t wA = __byte_perm(static_cast<uint32_t>(c[k][0]), static_cast<uint32_t>(c[k][1]), 0x5410);
                    const uint32_t wB = __byte_perm(static_cast<uint32_t>(c[k][2]), static_cast<uint32_t>(c[k][3]), 0x5410);
                    if ((b & 1) == 0 && b != NB - 1) { pA[k] = wA; pB[k] = wB; continue; }
                    // a pair of blocks (or the last, unpaired one) is complete: its low / high byte planes are the A
                    // fragment of a K = 16 step of the vertical product; the two planes are folded after each step
                    const uint32_t eA = (b & 1) ? pA[k] : wA, oA = (b & 1) ? wA : 0u;
                    const uint32_t eB = (b & 1) ? pB[k] : wB, oB = (b & 1) ? wB : 0u;
                    int ch[4] = {0, 0, 0, 0};
                    const uint32_t bw = (b >> 1) ? b2.y : b2.x;
                    imma16816(acc[k], __byte_perm(eA, oA, 0x6420), __byte_perm(eB, oB, 0x6420), bw);
                    imma16816(ch, __byte_perm(eA, oA, 0x7531), __byte_perm(eB, oB, 0x7531), bw);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[k][e] += ch[e] << 8;
                }
            }
            // ---- stage 2 epilogue: acc = 2 N + D for the 8 area rows of the group -> parked planar in buffer `par`
            const uint32_t par = gcount % NBUF;
            mbar_wait_u32(aempty0 + 8u * par, (((gcount / NBUF) & 1u) ^ 1u), 24);     // Pillow pass of group gcount - NBUF is done
            const uint32_t pbase = arows0 + par * 3u * planebytes;
#pragma unroll
            for (int k = 0; k < TPW; ++k) {
                const int ti = w + k * NCW;
                const int x0 = ti * 5;
                const bool slot = ti < P.ntx;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t q = __umulhi(static_cast<uint32_t>(acc[k][e]), dmul) >> dsh;
                    if (slot && x0 < ((e >> 1) ? xlim1 : xlim0))
                        sts8(pbase + ((e >> 1) ? park1 : park0) + (e & 1) * pitchA + static_cast<uint32_t>(x0), q);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(afull0 + 8u * par);
            // ---- stage 3 of the group finished PDEF groups ago (every warp completed its rows long since)
            if (gcount >= PDEF) pillow((gcount - PDEF) % NBUF, ((gcount - PDEF) / NBUF) & 1u, out_q[0], valid_q[0]);
#pragma unroll
            for (int i = 0; i + 1 < PDEF; ++i) { out_q[i] = out_q[i + 1]; valid_q[i] = valid_q[i + 1]; }
            out_q[PDEF - 1] = P.mid2 + f * P.mid2_frame_stride + static_cast<int64_t>(gi) * 8 * S3;
            valid_q[PDEF - 1] = min(8, P.ny - gi * 8);
            ++gcount;
        }
    }
    // drain: the Pillow passes still owed
#pragma unroll
    for (int i = 0; i < PDEF; ++i) {
        const uint32_t gq = gcount + i;       // the group at queue position i is gq - PDEF
        if (gq >= PDEF && gq - PDEF < gcount) pillow((gq - PDEF) % NBUF, ((gq - PDEF) / NBUF) & 1u, out_q[i], valid_q[i]);
    }
}
