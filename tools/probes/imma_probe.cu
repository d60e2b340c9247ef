#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE, int ILP>
__global__ void k(int iters, int* out, long long* clk) {
    uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u}, b[2] = {threadIdx.x * 3u, 7u};
    int c[ILP][4] = {}; float f[ILP][4] = {};
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) { if (MODE == 0) imma(c[j], a, b); else hmma(f[j], a, b); }
    }
    long long t1 = clock64();
    int s = 0; for (int j = 0; j < ILP; ++j) for (int q = 0; q < 4; ++q) s += c[j][q] + (int)f[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP> void run(const char* name, int warps) {
    int* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    int iters = 4096;
    k<MODE, ILP><<<148, warps * 32>>>(iters, out, clk); cudaDeviceSynchronize();
    k<MODE, ILP><<<148, warps * 32>>>(iters, out, clk); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double macs = (MODE == 0 ? 16.0 * 8 * 32 : 16.0 * 8 * 16) * ILP * iters * warps;
    printf("%s warps/SM=%d ILP=%d: %.1f MAC/clk/SM (%lld clk)\n", name, warps, ILP, macs / h[0], h[0]);
    cudaFree(out); cudaFree(clk);
}
int main() {
    run<0, 1>("IMMA.16832.u8.s8", 4); run<0, 4>("IMMA.16832.u8.s8", 4); run<0, 4>("IMMA.16832.u8.s8", 8); run<0, 4>("IMMA.16832.u8.s8", 16);
    run<1, 1>("HMMA.16816.f16", 4); run<1, 4>("HMMA.16816.f16", 4); run<1, 4>("HMMA.16816.f16", 8); run<1, 4>("HMMA.16816.f16", 16);
    return 0;
}
