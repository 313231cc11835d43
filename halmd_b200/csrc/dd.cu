// Slab domain decomposition: device-side helpers for particle migration and ghost (halo) layers.
//
// The reference is single-GPU (doc/usage/multi-gpu.rst.in: independent replicas only); the decomposition is new
// (SURVEY §8e).  Each rank owns the particles whose wrapped x coordinate lies in its slab and keeps ghost copies of
// the neighbouring ranks' particles within r_c + skin of its two faces.  Everything here is deterministic: selections
// are expressed as small integer keys that the caller feeds to the stable radix sort (hx_sort_pairs_stable_u32), so
// that migrants and ghosts keep their relative memory order.  Host orchestration: halmd_b200/decomposition.py.
#include "common.cuh"

namespace hx {

// owner slab of a wrapped coordinate: floor((x - x0) / width) folded into [0, nranks)
__device__ __forceinline__ int slab_of(float x, float x0, float inv_width, int nranks)
{
    int s = __float2int_rd(__fmul_rn(__fsub_rn(x, x0), inv_width));
    s %= nranks;
    return s < 0 ? s + nranks : s;
}

// key 0: stays, 1: migrates to the left neighbour (rank - 1), 2: to the right neighbour (rank + 1),
// 3: left the union of the three slabs (moved more than a slab width between two list updates: fatal)
__global__ void k_dd_owner_keys(const float4* __restrict__ pos_hi, unsigned n, float x0, float inv_width, int nranks,
                                int rank, unsigned* __restrict__ keys, unsigned* __restrict__ counts)
{
    __shared__ unsigned s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int s = slab_of(pos_hi[i].x, x0, inv_width, nranks);
        unsigned key = 3;
        if (s == rank) key = 0;
        else if (s == (rank + nranks - 1) % nranks) key = 1;
        else if (s == (rank + 1) % nranks) key = 2;
        // two ranks: both neighbours are the same peer, which takes everything as "left" traffic
        if (nranks == 2 && key == 2) key = 1;
        keys[i] = key;
        atomicAdd(&s_cnt[key], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

// key 0: within `halo` of the slab face(s) on the given side (0: low face, 1: high face, 2: either), else 1
__global__ void k_dd_ghost_keys(const float4* __restrict__ pos_hi, unsigned n, float lo, float hi, float halo, int side,
                                unsigned* __restrict__ keys, unsigned* __restrict__ count)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (i < n) {
        float x = pos_hi[i].x;
        bool near_lo = __fsub_rn(x, lo) < halo, near_hi = __fsub_rn(hi, x) < halo;
        sel = side == 0 ? near_lo : (side == 1 ? near_hi : (near_lo || near_hi));
        keys[i] = sel ? 0u : 1u;
    }
    unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(count, __popc(m));
}

// out[i] = in[index[i]] for the first n slots of each particle array (double-single planes at offset A)
__global__ void k_dd_permute(const unsigned* __restrict__ index, unsigned n, unsigned A,
                             const float4* __restrict__ pos_in, float4* __restrict__ pos_out,
                             const float4* __restrict__ vel_in, float4* __restrict__ vel_out,
                             const float4* __restrict__ img_in, float4* __restrict__ img_out,
                             const unsigned* __restrict__ id_in, unsigned* __restrict__ id_out)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned j = index[i];
    pos_out[i] = __ldg(&pos_in[j]);
    pos_out[i + A] = __ldg(&pos_in[j + A]);
    vel_out[i] = __ldg(&vel_in[j]);
    vel_out[i + A] = __ldg(&vel_in[j + A]);
    img_out[i] = __ldg(&img_in[j]);
    id_out[i] = __ldg(&id_in[j]);
}

__global__ void k_gather_f4(const float4* __restrict__ src, const unsigned* __restrict__ index, unsigned n,
                            float4* __restrict__ dst, const unsigned* __restrict__ ctl)
{
    if (ctl && ctl[CTL_HALTED]) return;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __ldg(&src[index[i]]);
}

__global__ void k_gather_u32(const unsigned* __restrict__ src, const unsigned* __restrict__ index, unsigned n,
                             unsigned* __restrict__ dst)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __ldg(&src[index[i]]);
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_dd_owner_keys(const hx_float4* position_hi_dev, size_t nparticle, float x0, float width, int nranks, int rank,
                     uint32_t* keys_dev, uint32_t* counts_dev /*4, zeroed here*/, void* stream)
{
    if (!position_hi_dev || !keys_dev || !counts_dev || nranks < 1 || rank < 0 || rank >= nranks || !(width > 0)
        || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_dd_owner_keys: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(counts_dev, 0, 4 * sizeof(uint32_t), st);
    if (err != cudaSuccess) { set_error("hx_dd_owner_keys: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    if (nparticle == 0) return HX_OK;
    k_dd_owner_keys<<<div_up(nparticle, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(position_hi_dev),
                                                          static_cast<unsigned>(nparticle), x0, 1.0f / width, nranks, rank,
                                                          keys_dev, counts_dev);
    return check_launch("k_dd_owner_keys");
}

int hx_dd_ghost_keys(const hx_float4* position_hi_dev, size_t nparticle, float lo, float hi, float halo, int side,
                     uint32_t* keys_dev, uint32_t* count_dev /*1, zeroed here*/, void* stream)
{
    if (!position_hi_dev || !keys_dev || !count_dev || side < 0 || side > 2 || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_dd_ghost_keys: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st);
    if (err != cudaSuccess) { set_error("hx_dd_ghost_keys: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    if (nparticle == 0) return HX_OK;
    k_dd_ghost_keys<<<div_up(nparticle, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(position_hi_dev),
                                                          static_cast<unsigned>(nparticle), lo, hi, halo, side, keys_dev, count_dev);
    return check_launch("k_dd_ghost_keys");
}

int hx_dd_permute(const uint32_t* index_dev, size_t nparticle, size_t array_size, const hx_float4* position_in_dev,
                  hx_float4* position_out_dev, const hx_float4* velocity_in_dev, hx_float4* velocity_out_dev,
                  const hx_float4* image_in_dev, hx_float4* image_out_dev, const uint32_t* id_in_dev, uint32_t* id_out_dev,
                  void* stream)
{
    if (!index_dev || !position_in_dev || !position_out_dev || !velocity_in_dev || !velocity_out_dev || !image_in_dev
        || !image_out_dev || !id_in_dev || !id_out_dev || nparticle > array_size || array_size >= 0x7FFFFFFFull
        || position_in_dev == position_out_dev || velocity_in_dev == velocity_out_dev) {
        set_error("hx_dd_permute: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    k_dd_permute<<<div_up(nparticle, 256), 256, 0, as_stream(stream)>>>(
        index_dev, static_cast<unsigned>(nparticle), static_cast<unsigned>(array_size),
        reinterpret_cast<const float4*>(position_in_dev), reinterpret_cast<float4*>(position_out_dev),
        reinterpret_cast<const float4*>(velocity_in_dev), reinterpret_cast<float4*>(velocity_out_dev),
        reinterpret_cast<const float4*>(image_in_dev), reinterpret_cast<float4*>(image_out_dev), id_in_dev, id_out_dev);
    return check_launch("k_dd_permute");
}

int hx_gather_float4(const hx_float4* src_dev, const uint32_t* index_dev, size_t n, hx_float4* dst_dev,
                     const uint32_t* ctl_dev, void* stream)
{
    if (!src_dev || !index_dev || !dst_dev || n >= 0x7FFFFFFFull) { set_error("hx_gather_float4: invalid argument"); return HX_E_INVALID; }
    if (n == 0) return HX_OK;
    k_gather_f4<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src_dev), index_dev,
                                                             static_cast<unsigned>(n), reinterpret_cast<float4*>(dst_dev), ctl_dev);
    return check_launch("k_gather_f4");
}

int hx_gather_u32(const uint32_t* src_dev, const uint32_t* index_dev, size_t n, uint32_t* dst_dev, void* stream)
{
    if (!src_dev || !index_dev || !dst_dev || n >= 0x7FFFFFFFull) { set_error("hx_gather_u32: invalid argument"); return HX_E_INVALID; }
    if (n == 0) return HX_OK;
    k_gather_u32<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(src_dev, index_dev, static_cast<unsigned>(n), dst_dev);
    return check_launch("k_gather_u32");
}

} // extern "C"
