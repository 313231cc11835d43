// Does an FFMA2 (packed f32x2) occupy the warp scheduler's issue port for one cycle or two?
// mode 0: 8 FFMA2 per iteration; mode 1: 8 FFMA2 + 8 independent LOP3/IADD3; mode 2: 16 FFMA (scalar) ; mode 3: 16 FFMA + 8 int
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    u64 a[8]; float f[16]; unsigned q[8];
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i); a[i] = *reinterpret_cast<u64*>(&t); q[i] = threadIdx.x * 7 + i; }
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3f + i;
    float2 bb = make_float2(s, s * 0.5f), cc = make_float2(1e-3f, 2e-3f);
    u64 b = *reinterpret_cast<u64*>(&bb), c = *reinterpret_cast<u64*>(&cc);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 1) a[i] = fma2(a[i], b, c);
            if (MODE == 2 || MODE == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2*i]) : "f"(s), "f"(cc.x)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2*i+1]) : "f"(s), "f"(cc.y)); }
            if (MODE == 1 || MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(it), "r"(q[(i + 1) & 7]));
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); r += t.x + t.y + (float)q[i]; }
    for (int i = 0; i < 16; ++i) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f);
        if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
        if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 0.999f);
        if (mode == 3) k<3><<<148 * 8, 256>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("mode %d: %.3f ms\n", mode, ms);
    }
    return 0;
}
