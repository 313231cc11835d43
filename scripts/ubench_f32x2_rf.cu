// Development aid: issue cost and latency of packed f32x2 instructions on sm_100a.
//   mode 0  8 independent FFMA2 per iteration, three distinct 64-bit register operands each
//   mode 1  16 independent scalar FFMA (the same lane operations)
//   mode 2  one dependent FFMA2 chain (latency), 1 warp per SM sub-partition
//   mode 3  one dependent scalar FFMA chain (latency)
//   mode 4  8 FFMA2 whose multiplicand is a scalar broadcast constant (2 register operands of 64 bit)
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float s)
{
    u64 a[8], b[8], c[8];
    float f[16];
    for (int i = 0; i < 8; ++i) { a[i] = pk(threadIdx.x * 1e-3f + i, s); b[i] = pk(s + i * 1e-4f, s - i * 1e-4f); c[i] = pk(1e-3f * i, 2e-3f); }
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3f + i;
    u64 const bc = pk(s, s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma2(a[i], b[i], c[i]);
            if (MODE == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i]) : "f"(f[(2 * i + 5) & 15] * 0 + s), "f"(s)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i + 1]) : "f"(s), "f"(s)); }
            if (MODE == 2) a[0] = fma2(a[0], b[0], c[0]);
            if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[0]) : "f"(s), "f"(s));
            if (MODE == 4) a[i] = fma2(a[i], bc, c[i]);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    for (int i = 0; i < 16; ++i) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 5; ++mode) for (int rep = 0; rep < 2; ++rep) {
        bool lat = mode == 2 || mode == 3;
        dim3 grid(lat ? 148 : 148 * 8), block(lat ? 128 : 256);
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, block>>>(out, iters, 0.999f);
        if (mode == 1) k<1><<<grid, block>>>(out, iters, 0.999f);
        if (mode == 2) k<2><<<grid, block>>>(out, iters, 0.999f);
        if (mode == 3) k<3><<<grid, block>>>(out, iters, 0.999f);
        if (mode == 4) k<4><<<grid, block>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double warps_per_smsp = lat ? 1 : 16;
        double inst = (mode == 1 ? 16.0 : 8.0) * iters * warps_per_smsp;      // warp instructions per SM sub-partition
        if (rep) printf("mode %d: %.3f ms, %.2f cycles per warp instruction per sub-partition (1.965 GHz)\n", mode, ms, ms * 1e-3 * 1.965e9 / inst);
    }
    return 0;
}
