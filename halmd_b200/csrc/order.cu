// Hilbert-curve keys, stable LSD radix sort of (key,value) pairs, particle permutation.
//
// Reference: halmd/mdsim/gpu/sorts/hilbert.cpp:71-117, hilbert_kernel.cu:38-69,
// halmd/mdsim/sorts/hilbert_kernel.hpp:70-142,205-230 (curve), halmd/algorithm/gpu/radix_sort*.{hpp,cu}
// (Zagha–Blelloch sort), halmd/mdsim/gpu/particle.cpp:173-218 + particle_kernel.cu:45-68 (rearrange).
#include "common.cuh"

namespace hx {

// ---- Hilbert key ----------------------------------------------------------------------------------
// The reference snaps r/L to the centre of a 2^depth lattice cell and then walks `depth` levels,
// reading the octant from the sign bits and scaling by two.  All of that arithmetic is exact, so the
// octant sequence is the binary expansion of the integer lattice coordinate (MSB first, a set bit
// meaning the non-negative half); the coordinate n (r/L - floor(r/L) rounding up to 1) behaves as n-1.
// The vertex/orientation update (Wang & Shan 2005) is kept table driven.
// The float arithmetic that decides the lattice cell is the reference build's, read off the SASS of its map kernel
// (oracle/_ref/obj/hilbert_kernel.o; -prec-div=false -ftz=true): element_div(r, L) is r * MUFU.RCP(L) (FMUL.FTZ, not
// fused), then t - floor(t) and the product with float(n), each rounded separately.
__device__ __forceinline__ unsigned lattice_coord(float r, float L, unsigned n)
{
    float const t = mul_ftz(rcp_approx(L), r);
    float const f = sub_ftz(t, floor_ftz(t));
    unsigned const c = __float2uint_rz(floor_ftz(mul_ftz(static_cast<float>(n), f)));
    return min(c, n - 1);
}

__device__ __forceinline__ unsigned hilbert_code(unsigned ix, unsigned iy, unsigned iz, unsigned depth)
{
    // bit offsets (within the packed table vc) of the eight curve vertices a..h
    unsigned pos[8] = {21, 18, 12, 15, 3, 0, 6, 9};
    // vertex → code table: 3 bits per vertex
    unsigned vc = (1u << 18) ^ (2u << 12) ^ (3u << 15) ^ (4u << 3) ^ (5u << 0) ^ (6u << 6) ^ (7u << 9);
    // permutation rules: two transpositions of table slots per code value
    const unsigned char rule[8][4] = {
        {1, 7, 2, 4},  // v = 0: (b h)(c e)
        {2, 6, 3, 7},  // v = 1: (c g)(d h)
        {2, 6, 3, 7},  // v = 2
        {0, 2, 5, 7},  // v = 3: (a c)(f h)
        {0, 2, 5, 7},  // v = 4
        {0, 4, 1, 5},  // v = 5: (a e)(b f)
        {0, 4, 1, 5},  // v = 6
        {0, 6, 3, 5},  // v = 7: (a g)(d f)
    };
    unsigned code = 0;
#pragma unroll 1
    for (unsigned level = 0; level < depth; ++level) {
        unsigned sh = depth - 1 - level;
        // sign bit set (negative half) ⇔ coordinate bit clear
        unsigned x = ((ix >> sh) & 1u) ^ 1u;
        unsigned y = ((iy >> sh) & 1u) ^ 1u;
        unsigned z = ((iz >> sh) & 1u) ^ 1u;
        unsigned v = (vc >> (3 * (x + (y << 1) + (z << 2)))) & 7u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            unsigned p = rule[v][2 * s], q = rule[v][2 * s + 1];
            unsigned a = pos[p], b = pos[q];
            unsigned va = (vc >> a) & 7u, vb = (vc >> b) & 7u;
            vc = vc ^ (va << a) ^ (vb << b) ^ (va << b) ^ (vb << a);
            pos[p] = b;
            pos[q] = a;
        }
        code = (code << 3) + v;
    }
    return code;
}

__global__ void k_hilbert_keys(const float4* __restrict__ pos_hi, unsigned nslots, vec3 L, unsigned depth,
                               unsigned* __restrict__ keys)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    float4 p = pos_hi[i];
    if (__float_as_uint(p.w) == HX_PH) {
        keys[i] = HX_PH;
        return;
    }
    unsigned n = 1u << depth;
    keys[i] = hilbert_code(lattice_coord(p.x, L.x, n), lattice_coord(p.y, L.y, n),
                           lattice_coord(p.z, L.z, n), depth);
}

__global__ void k_iota(unsigned* __restrict__ out, unsigned n)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

// ---- stable LSD radix sort, 8 bits per pass --------------------------------------------------------
constexpr unsigned RS_THREADS = 256;
constexpr unsigned RS_ITEMS = 8;
constexpr unsigned RS_TILE = RS_THREADS * RS_ITEMS;
constexpr unsigned RS_RADIX = 256;
constexpr unsigned RS_WARPS = RS_THREADS / 32;

// per-tile digit histogram, stored digit-major: hist[d * ntiles + tile]
__global__ void k_rs_hist(const unsigned* __restrict__ keys, unsigned n, unsigned shift, unsigned ntiles,
                          unsigned* __restrict__ hist)
{
    __shared__ unsigned s_h[RS_RADIX];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    unsigned base = blockIdx.x * RS_TILE;
#pragma unroll
    for (unsigned r = 0; r < RS_ITEMS; ++r) {
        unsigned i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x];
}

// block d: exclusive scan of row d (ntiles entries) in place; row total → totals[d]
__global__ void k_rs_scan_rows(unsigned* __restrict__ hist, unsigned ntiles, unsigned* __restrict__ totals)
{
    __shared__ unsigned s_warp[RS_WARPS];
    __shared__ unsigned s_carry;
    unsigned* row = hist + static_cast<size_t>(blockIdx.x) * ntiles;
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < ntiles; base += RS_THREADS) {
        unsigned i = base + threadIdx.x;
        unsigned v = (i < ntiles) ? row[i] : 0;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (unsigned w = 0; w < warp; ++w) woff += s_warp[w];
        unsigned carry = s_carry;
        if (i < ntiles) row[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == RS_THREADS - 1) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

// exclusive scan of the 256 digit totals (one block of 256 threads)
__global__ void k_rs_scan_totals(unsigned* __restrict__ totals)
{
    __shared__ unsigned s_warp[RS_WARPS];
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned v = totals[threadIdx.x];
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (unsigned w = 0; w < warp; ++w) woff += s_warp[w];
    totals[threadIdx.x] = woff + incl - v;
}

// stable scatter: element order inside a tile is its index order
__global__ void k_rs_scatter(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                             unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out, unsigned n,
                             unsigned shift, unsigned ntiles, const unsigned* __restrict__ hist,
                             const unsigned* __restrict__ totals)
{
    __shared__ unsigned s_off[RS_RADIX];              // next output slot per digit for this tile
    __shared__ unsigned s_wc[RS_WARPS][RS_RADIX];     // per-warp digit counts of the current round
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    s_off[threadIdx.x] = totals[threadIdx.x] + hist[threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (unsigned w = 0; w < RS_WARPS; ++w) s_wc[w][threadIdx.x] = 0;
    __syncthreads();
    unsigned base = blockIdx.x * RS_TILE;
    for (unsigned r = 0; r < RS_ITEMS; ++r) {
        unsigned i = base + r * RS_THREADS + threadIdx.x;
        bool active = i < n;
        unsigned key = active ? keys_in[i] : 0;
        unsigned val = active ? vals_in[i] : 0;
        // inactive lanes get a unique pseudo-digit so that they never match an active lane
        unsigned digit = active ? ((key >> shift) & 0xFFu) : (RS_RADIX + lane);
        unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        bool leader = (rank == 0);
        if (active && leader) s_wc[warp][digit] = __popc(peers);
        __syncthreads();
        if (active) {
            unsigned o = s_off[digit];
            for (unsigned w = 0; w < warp; ++w) o += s_wc[w][digit];
            o += rank;
            keys_out[o] = key;
            vals_out[o] = val;
        }
        __syncthreads();
        // thread d owns digit d: advance the tile offset, clear the per-warp counts
        unsigned add = 0;
#pragma unroll
        for (unsigned w = 0; w < RS_WARPS; ++w) {
            add += s_wc[w][threadIdx.x];
            s_wc[w][threadIdx.x] = 0;
        }
        s_off[threadIdx.x] += add;
        __syncthreads();
    }
}

// ---- particle::rearrange ----------------------------------------------------------------------------
__global__ void k_permute(const unsigned* __restrict__ index, unsigned n, unsigned A,
                          const float4* __restrict__ pos_in, float4* __restrict__ pos_out,
                          const float4* __restrict__ vel_in, float4* __restrict__ vel_out,
                          const float4* __restrict__ img_in, float4* __restrict__ img_out,
                          const unsigned* __restrict__ id_in, unsigned* __restrict__ id_out,
                          unsigned* __restrict__ rid_out)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    unsigned j = (i < n) ? index[i] : i;
    pos_out[i] = __ldg(&pos_in[j]);
    pos_out[i + A] = __ldg(&pos_in[j + A]);
    vel_out[i] = __ldg(&vel_in[j]);
    vel_out[i + A] = __ldg(&vel_in[j + A]);
    img_out[i] = __ldg(&img_in[j]);
    unsigned id = __ldg(&id_in[j]);
    id_out[i] = id;
    // reverse_id = values of stable sort (id, iota) over the first n slots (particle.cpp:212-217);
    // ids are a permutation of 0..n-1, so that is the inverse permutation.  Slots >= n keep iota.
    if (i < n) {
        if (id < n) rid_out[id] = i;
    } else {
        rid_out[i] = i;
    }
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_hilbert_keys(const hx_float4* position_hi_dev, size_t nslots, const float box_length[3],
                    unsigned depth, uint32_t* keys_dev, void* stream)
{
    if (!position_hi_dev || !box_length || !keys_dev || depth == 0 || depth > 10) {
        set_error("hx_hilbert_keys: invalid argument (depth must be in 1..10)");
        return HX_E_INVALID;
    }
    if (nslots == 0) return HX_OK;
    k_hilbert_keys<<<div_up(nslots, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(position_hi_dev), static_cast<unsigned>(nslots), make_vec3(box_length),
        depth, keys_dev);
    return check_launch("k_hilbert_keys");
}

int hx_iota_u32(uint32_t* out_dev, size_t n, void* stream)
{
    if (!out_dev) { set_error("hx_iota_u32: null pointer"); return HX_E_INVALID; }
    if (n == 0) return HX_OK;
    k_iota<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(out_dev, static_cast<unsigned>(n));
    return check_launch("k_iota");
}

size_t hx_sort_scratch_bytes(size_t n)
{
    size_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    if (ntiles == 0) ntiles = 1;
    // keys_tmp[n] + vals_tmp[n] + hist[256*ntiles] + totals[256]
    return (2 * n + RS_RADIX * ntiles + RS_RADIX) * sizeof(uint32_t);
}

int hx_sort_pairs_stable_u32(uint32_t* keys_dev, uint32_t* values_dev, size_t n, unsigned key_bits,
                             void* scratch_dev, size_t scratch_bytes, void* stream)
{
    if (n == 0) return HX_OK;
    if (!keys_dev || !values_dev || !scratch_dev || key_bits == 0 || key_bits > 32
        || scratch_bytes < hx_sort_scratch_bytes(n) || n >= 0xFFFFFFFFull) {
        set_error("hx_sort_pairs_stable_u32: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    unsigned ntiles = div_up(n, RS_TILE);
    uint32_t* keys_tmp = static_cast<uint32_t*>(scratch_dev);
    uint32_t* vals_tmp = keys_tmp + n;
    uint32_t* hist = vals_tmp + n;
    uint32_t* totals = hist + static_cast<size_t>(RS_RADIX) * ntiles;
    unsigned passes = (key_bits + 7) / 8;
    uint32_t *kin = keys_dev, *vin = values_dev, *kout = keys_tmp, *vout = vals_tmp;
    int rc;
    for (unsigned p = 0; p < passes; ++p) {
        unsigned shift = 8 * p;
        k_rs_hist<<<ntiles, RS_THREADS, 0, st>>>(kin, static_cast<unsigned>(n), shift, ntiles, hist);
        if ((rc = check_launch("k_rs_hist"))) return rc;
        k_rs_scan_rows<<<RS_RADIX, RS_THREADS, 0, st>>>(hist, ntiles, totals);
        if ((rc = check_launch("k_rs_scan_rows"))) return rc;
        k_rs_scan_totals<<<1, RS_RADIX, 0, st>>>(totals);
        if ((rc = check_launch("k_rs_scan_totals"))) return rc;
        k_rs_scatter<<<ntiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, static_cast<unsigned>(n), shift,
                                                   ntiles, hist, totals);
        if ((rc = check_launch("k_rs_scatter"))) return rc;
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (kin != keys_dev) {
        cudaError_t err;
        if ((err = cudaMemcpyAsync(keys_dev, kin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess
            || (err = cudaMemcpyAsync(values_dev, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) {
            set_error("hx_sort_pairs_stable_u32: copy back: %s", cudaGetErrorString(err));
            return static_cast<int>(err);
        }
    }
    return HX_OK;
}

int hx_state_permute(const uint32_t* index_dev, size_t nparticle, size_t array_size,
                     const hx_float4* position_in_dev, hx_float4* position_out_dev,
                     const hx_float4* velocity_in_dev, hx_float4* velocity_out_dev,
                     const hx_float4* image_in_dev, hx_float4* image_out_dev,
                     const uint32_t* id_in_dev, uint32_t* id_out_dev, uint32_t* reverse_id_out_dev,
                     void* stream)
{
    if (!index_dev || !position_in_dev || !position_out_dev || !velocity_in_dev || !velocity_out_dev
        || !image_in_dev || !image_out_dev || !id_in_dev || !id_out_dev || !reverse_id_out_dev
        || nparticle > array_size || array_size >= 0x7FFFFFFFull
        || position_in_dev == position_out_dev || velocity_in_dev == velocity_out_dev
        || image_in_dev == image_out_dev || id_in_dev == id_out_dev) {
        set_error("hx_state_permute: invalid argument (in/out arrays must be distinct)");
        return HX_E_INVALID;
    }
    k_permute<<<div_up(array_size, 256), 256, 0, as_stream(stream)>>>(
        index_dev, static_cast<unsigned>(nparticle), static_cast<unsigned>(array_size),
        reinterpret_cast<const float4*>(position_in_dev), reinterpret_cast<float4*>(position_out_dev),
        reinterpret_cast<const float4*>(velocity_in_dev), reinterpret_cast<float4*>(velocity_out_dev),
        reinterpret_cast<const float4*>(image_in_dev), reinterpret_cast<float4*>(image_out_dev),
        id_in_dev, id_out_dev, reverse_id_out_dev);
    return check_launch("k_permute");
}

} // extern "C"
