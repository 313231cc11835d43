// refgpu_util.hpp — helpers shared by the host-side harness files of oracle/_ref/libhalmd_ref_gpu.so.
// TEST INFRASTRUCTURE ONLY (see refgpu.cpp).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <exception>

using namespace halmd;
using namespace halmd::mdsim::gpu;

#define RG_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "refgpu: %s: %s\n", #x, cudaGetErrorString(e_)); return static_cast<int>(e_); } } while (0)
// every launch block is bracketed by events: refgpu_last_kernel_ms() returns the device time of the last one (the
// reference's kernels timed on this GPU, scripts/ref_kernel_times.py); uploads and downloads lie outside
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static float g_last_ms = -1.f;
#define RG_TRY(stmt) try { \
        if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); } \
        cudaEventRecord(g_ev0); \
        stmt; \
        cudaEventRecord(g_ev1); cudaEventSynchronize(g_ev1); cudaEventElapsedTime(&g_last_ms, g_ev0, g_ev1); \
    } catch (std::exception const& e) { fprintf(stderr, "refgpu: %s\n", e.what()); return -1; }

namespace {

template <typename T>
struct dbuf {
    T* p = nullptr;
    size_t n = 0;
    explicit dbuf(size_t n_) : n(n_) { if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) p = nullptr; }
    dbuf(const T* host, size_t n_) : dbuf(n_) { if (p && n && host) cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice); }
    ~dbuf() { if (p) cudaFree(p); }
    int down(T* host) const { return static_cast<int>(cudaMemcpy(host, p, n * sizeof(T), cudaMemcpyDeviceToHost)); }
    dbuf(dbuf const&) = delete;
    dbuf& operator=(dbuf const&) = delete;
};

// linear-memory texture object, built as cuda::texture<T>(vector) builds it (cuda_wrapper/texture.hpp:33-46)
template <typename T>
struct lintex {
    cudaTextureObject_t t = 0;
    lintex(const T* dev, size_t n)
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = const_cast<T*>(dev);
        rd.res.linear.desc = cudaCreateChannelDesc<T>();
        rd.res.linear.sizeInBytes = n * sizeof(T);
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaCreateTextureObject(&t, &rd, &td, nullptr);
    }
    ~lintex() { if (t) cudaDestroyTextureObject(t); }
};

template <typename V>
V vec3(const float* v) { V r; for (int d = 0; d < 3; ++d) r[d] = v[d]; return r; }

} // namespace

