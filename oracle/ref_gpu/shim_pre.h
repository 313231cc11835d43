// Pre-included (-include) into every translation unit that compiles a reference .cu file in place.
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile): never linked into the product.
#pragma once
#include <cuda_runtime.h>
// CUDA 12 no longer declares the global ::saturate(float) device function that the reference calls in
// halmd/numeric/blas/detail/rounding.hpp:77 (it was __saturatef's alias in the CUDA 8-10 headers HALMD targets)
static __host__ __device__ inline float saturate(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
