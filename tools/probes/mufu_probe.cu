// Development probe: MUFU.EX2 issue rate per SM (independent chains) next to FFMA, and a Cody-Waite + cubic exp2 on the
// FMA pipe (scalar and packed fp32x2), for the softmax of the tcgen05 attention kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly_ex2(float x) {
    const float t = x + 12582912.f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (t - 12582912.f);           // in [-0.5, 0.5]
    float p = fmaf(0.0555041087f, f, 0.2402265070f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int MODE>
__global__ void k(int iters, float* out, long long* clk) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2(v[i]) - 1.0f;
            else if (MODE == 1) v[i] = fmaf(v[i], 0.999f, -0.001f);
            else v[i] = poly_ex2(v[i]) - 1.0f;
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps) {
    float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 2048;
    k<MODE><<<148, warps * 32>>>(iters, out, clk); cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(iters, out, clk); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s warps/SM=%d: %.2f results/clk/SM (%lld clk)\n", name, warps, 8.0 * iters * warps * 32 / h[0], h[0]);
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int w : {4, 8, 16}) run<0>("MUFU.EX2 (+FADD)", w);
    for (int w : {4, 8, 16}) run<1>("FFMA", w);
    for (int w : {4, 8, 16}) run<2>("poly exp2 (+FADD)", w);
    float h[4]; float* d; cudaMalloc(&d, 16);
    return 0;
}
