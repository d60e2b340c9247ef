// Development probe (TMEM A operand): the same product as umma_mn_probe.cu, but P is written to TENSOR MEMORY by the
// threads (one row per thread, two bf16 per 32-bit column, tcgen05.st) and consumed from there as the A operand
// (tcgen05.mma ... [d], [a_tmem], b_desc ...; 8 columns per K = 16 step) -- the hand-off the attention kernel uses
// for the softmax probabilities.  argv[4] = 1 swaps the two bf16 of a column (to pin the packing order).
// Development probe: D[128 x 64] = P[128 x 64 keys] * V[64 keys x 64 d] on tcgen05 with V consumed as an MN-MAJOR B
// operand (rows of V = keys = the K dimension; the 64 d-values of a key are contiguous: exactly what a TMA
// SWIZZLE_128B box of the qkv matrix delivers).  Tries a (LBO, SBO) pair given on the command line and reports the
// max error against a CPU reference, so the descriptor encoding can be pinned before building attention on it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I ../../advanced-video-event-detection-extraction_b200/csrc -o umma_mn_probe umma_mn_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ptx.cuh"
using namespace b200;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tp, const __grid_constant__ CUtensorMap tv, float* out, uint32_t lbo,
             uint32_t sbo, uint32_t kstep_bytes, const __nv_bfloat16* __restrict__ Pg, int swap) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sP = smem;               // 128 rows x 128 B
    uint8_t* sV = smem + 16384;       // 64 keys x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
    uint64_t* mbar = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<1>(slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    {   // P row of this thread -> TMEM columns [64, 96): column c holds keys 2c (low half) and 2c + 1 (high half)
        uint32_t pr[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Pg + (warp * 32 + lane) * 64);
        for (int j = 0; j < 32; ++j) { uint32_t v = src[j]; pr[j] = swap ? ((v >> 16) | (v << 16)) : v; }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 64),
              "r"(pr[0]), "r"(pr[1]), "r"(pr[2]), "r"(pr[3]), "r"(pr[4]), "r"(pr[5]), "r"(pr[6]), "r"(pr[7]),
              "r"(pr[8]), "r"(pr[9]), "r"(pr[10]), "r"(pr[11]), "r"(pr[12]), "r"(pr[13]), "r"(pr[14]), "r"(pr[15]),
              "r"(pr[16]), "r"(pr[17]), "r"(pr[18]), "r"(pr[19]), "r"(pr[20]), "r"(pr[21]), "r"(pr[22]), "r"(pr[23]),
              "r"(pr[24]), "r"(pr[25]), "r"(pr[26]), "r"(pr[27]), "r"(pr[28]), "r"(pr[29]), "r"(pr[30]), "r"(pr[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 16384 + 8192);
        tma_load_2d(sP, &tp, bar, 0, 0);
        tma_load_2d(sV, &tv, bar, 0, 0);
        mbar_wait(bar, 0, 1);
        tc_fence_after();
        // idesc: f32 accum, bf16 A/B, A K-major, B MN-major (bit 16), N = 64, M = 128
        uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 16);
        for (int k = 0; k < 4; ++k) {
            uint64_t bdesc = 0;
            const uint32_t addr = smem_u32(sV) + k * kstep_bytes;
            bdesc |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
            bdesc |= static_cast<uint64_t>(lbo >> 4) << 16;
            bdesc |= static_cast<uint64_t>(sbo >> 4) << 32;
            bdesc |= static_cast<uint64_t>(1) << 46;
            bdesc |= static_cast<uint64_t>(2) << 61;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmem), "r"(tmem + 64 + 8 * k), "l"(bdesc), "r"(idesc), "r"(k != 0 ? 1u : 0u) : "memory");
        }
        umma_commit(mbar);
    }
    mbar_wait(mbar, 0, 2);
    tc_fence_after();
    uint32_t r[32];
    for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait_regs(r);
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tmem, 128); }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(PFN_encodeTiled enc, CUtensorMap* m, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main(int argc, char** argv) {
    const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 16, sbo = argc > 2 ? atoi(argv[2]) : 1024;
    const uint32_t kstep = argc > 3 ? atoi(argv[3]) : 2048;
    const int swap = argc > 4 ? atoi(argv[4]) : 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
    std::vector<__nv_bfloat16> P(128 * 64), V(64 * 64);
    std::vector<float> Pf(128 * 64), Vf(64 * 64), ref(128 * 64, 0.f), got(128 * 64);
    srand(1);
    for (size_t i = 0; i < P.size(); ++i) { P[i] = __float2bfloat16((rand() % 200 - 100) / 100.0f); Pf[i] = __bfloat162float(P[i]); }
    for (size_t i = 0; i < V.size(); ++i) { V[i] = __float2bfloat16((rand() % 200 - 100) / 100.0f); Vf[i] = __bfloat162float(V[i]); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            float a = 0.f;
            for (int k = 0; k < 64; ++k) a += Pf[m * 64 + k] * Vf[k * 64 + n];
            ref[m * 64 + n] = a;
        }
    __nv_bfloat16 *dP, *dV;
    float* dO;
    cudaMalloc(&dP, P.size() * 2); cudaMalloc(&dV, V.size() * 2); cudaMalloc(&dO, got.size() * 4);
    cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tp, tv;
    if (make_map(enc, &tp, dP, 128, 64, 128) || make_map(enc, &tv, dV, 64, 64, 64)) { printf("tensor map failed\n"); return 1; }
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    probe_kernel<<<1, 128, 40000>>>(tp, tv, dO, lbo, sbo, kstep, dP, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo=%u sbo=%u kstep=%u: CUDA error %s\n", lbo, sbo, kstep, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost);
    float mx = 0.f;
    for (size_t i = 0; i < got.size(); ++i) mx = fmaxf(mx, fabsf(got[i] - ref[i]));
    printf("TMEM-A swap=%d lbo=%u sbo=%u kstep=%u: max |err| = %g  (%s)\n", swap, lbo, sbo, kstep, mx, mx < 1e-2f ? "MATCH" : "mismatch");
    return 0;
}
