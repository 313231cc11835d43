#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b); rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a); rb = *reinterpret_cast<unsigned long long*>(&b);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    float2 b = make_float2(s, s * 0.5f), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, b.x, c.x); a[i].y = __fmaf_rn(a[i].y, b.y, c.y); }
            if (MODE == 1) a[i] = fma2(a[i], b, c);
            if (MODE == 2) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }
            if (MODE == 3) a[i] = add2(a[i], c);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 0.999f);
            if (mode == 3) k<3><<<148 * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = 148.0 * 8 * 256 * iters * 16;
            if (rep) printf("mode %d: %.3f ms, %.2f T lane-ops/s\n", mode, ms, ops / ms / 1e9);
        }
    }
    return 0;
}
