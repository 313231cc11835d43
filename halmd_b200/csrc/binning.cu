// Cell binning: particle → cell index → fixed-size cell placeholders.
//
// Reference: halmd/mdsim/gpu/binning.cpp:166-221 runs compute_cell + a 4-pass radix sort of all
// particles + find_cell_offset + assign_cells (binning_kernel.cu:34-139).  The observable result is
// g_cell[c*cell_size + s] = s-th smallest particle index in cell c, padded with placeholders, and an
// overflow flag when a cell has >= cell_size particles.  Here the same array is produced without a
// global sort: atomic slot claiming per cell (order-free) followed by a warp-local bitonic sort of
// each cell's few entries, which restores the ascending-index order of the stable sort.
#include "common.cuh"
#include <algorithm>

namespace hx {

// cell index of a position, binning_kernel.cu:34-54:  (uint)(r / l + n) % n, x fastest.
// The quotient is the reference build's (-prec-div=false, --fmad=true, -ftz=true: cmake/platform.cmake:75), read off
// the SASS of its compute_cell kernel (oracle/_ref/obj/binning_kernel.o): MUFU.RCP(l), then ONE fused
// FFMA.FTZ rcp * r + float(n) and a truncating conversion — so a particle within an ulp of a cell face lands in the
// cell the reference puts it in.
__device__ __forceinline__ unsigned cell_coord(float r, float len, unsigned n)
{
    return __float2uint_rz(fma_ftz(rcp_approx(len), r, static_cast<float>(n))) % n;
}

// x0, nxl: layers of the grid in x that may hold particles (slab-decomposed ranks; 0, ncell.x otherwise): a particle
// outside them would be missed by the kernels that visit those layers only, and raises *range_error instead
__global__ void k_cell_claim(const float4* __restrict__ pos_hi, unsigned n, vec3 cell_len, uvec3 ncell,
                             unsigned cell_size, unsigned* __restrict__ cell_count,
                             unsigned* __restrict__ g_cell, unsigned* __restrict__ cell_index, unsigned x0, unsigned nxl,
                             int* __restrict__ range_error)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos_hi[i];
    if (__float_as_uint(p.w) == HX_PH) {
        if (cell_index) cell_index[i] = HX_PH;
        return;
    }
    unsigned cx = cell_coord(p.x, cell_len.x, ncell.x);
    unsigned cy = cell_coord(p.y, cell_len.y, ncell.y);
    unsigned cz = cell_coord(p.z, cell_len.z, ncell.z);
    unsigned c = (cz * ncell.y + cy) * ncell.x + cx;
    if (cell_index) cell_index[i] = c;
    if ((cx + ncell.x - x0) % ncell.x >= nxl) atomicOr(range_error, HX_BINNING_OUT_OF_RANGE);
    unsigned s = atomicAdd(&cell_count[c], 1u);
    if (s < cell_size) g_cell[static_cast<size_t>(c) * cell_size + s] = i;
}

// MUFU.RCP of a few host values (see hx_device_rcp)
__global__ void k_device_rcp(const float* __restrict__ x, float* __restrict__ out, unsigned n)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rcp_approx(x[i]);
}

// one warp per cell: sort the claimed entries ascending (placeholders = 0xFFFFFFFF sort last)
__global__ void k_cell_sort(unsigned* __restrict__ g_cell, const unsigned* __restrict__ cell_count,
                            unsigned ncells, unsigned cell_size, unsigned pow2, int* __restrict__ overflow, unsigned nx,
                            unsigned x0, unsigned nxl)
{
    extern __shared__ unsigned s_all[];
    unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    unsigned c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= ncells) return;
    // the c-th cell of the visited layers [x0, x0 + nxl) (periodic) of a grid with nx layers, x fastest
    unsigned x = x0 + c % nxl;
    if (x >= nx) x -= nx;
    c = (c / nxl) * nx + x;
    unsigned cnt = cell_count[c];
    if (cnt >= cell_size) {
        // binning_kernel.cu:124-127 — last placeholder occupied ⇒ overcrowded
        if (lane == 0) atomicOr(overflow, 1);       // (bit 1 may carry HX_BINNING_OUT_OF_RANGE)
        cnt = cell_size;
    }
    if (cnt < 2) return;
    // smallest power of two >= cnt, bounded by pow2 (>= cell_size)
    unsigned m = 2;
    while (m < cnt) m <<= 1;
    unsigned* s = s_all + warp * pow2;
    unsigned* cell = g_cell + static_cast<size_t>(c) * cell_size;
    for (unsigned k = lane; k < m; k += 32) s[k] = (k < cnt) ? cell[k] : HX_PH;
    __syncwarp();
    for (unsigned size = 2; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = lane; t < (m >> 1); t += 32) {
                unsigned lo = 2 * t - (t & (stride - 1));   // index with the `stride` bit clear
                unsigned hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
    for (unsigned k = lane; k < cnt; k += 32) cell[k] = s[k];
}

} // namespace hx

using namespace hx;

static int binning_update_impl(const hx_float4* position_hi_dev, size_t nparticle,
                               const float cell_length[3], const unsigned ncell[3], unsigned cell_size,
                               uint32_t* g_cell_dev, uint32_t* cell_count_dev, uint32_t* cell_index_dev,
                               int* overflow_flag_dev, void* stream, unsigned cell_x0, unsigned cell_nx)
{
    if (!position_hi_dev || !cell_length || !ncell || !g_cell_dev || !cell_count_dev || !overflow_flag_dev
        || cell_size == 0 || cell_size % 32 != 0) {
        set_error("hx_binning_update: invalid argument (cell size must be a multiple of warp size 32)");
        return HX_E_INVALID;
    }
    size_t ncells = static_cast<size_t>(ncell[0]) * ncell[1] * ncell[2];
    if (ncells == 0 || ncells * cell_size >= 0xFFFFFFFFull || nparticle >= 0xFFFFFFFFull) {
        set_error("hx_binning_update: cell grid out of range");
        return HX_E_INVALID;
    }
    if (cell_nx == 0 || cell_nx > ncell[0] || cell_x0 >= ncell[0]) {
        set_error("hx_binning_update_slab: cell layers [%u, +%u) outside the grid", cell_x0, cell_nx);
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err;
    // counters of all cells, placeholders of the visited layers only (two strided blocks when the range wraps)
    err = cudaMemsetAsync(cell_count_dev, 0, ncells * sizeof(uint32_t), st);
    size_t const layer = static_cast<size_t>(cell_size) * sizeof(uint32_t), pitch = ncell[0] * layer;
    size_t const rows = static_cast<size_t>(ncell[1]) * ncell[2];
    unsigned const first = std::min(cell_nx, ncell[0] - cell_x0);
    if (err == cudaSuccess && cell_nx == ncell[0]) err = cudaMemsetAsync(g_cell_dev, 0xFF, ncells * layer, st);
    else if (err == cudaSuccess) {
        err = cudaMemset2DAsync(reinterpret_cast<char*>(g_cell_dev) + cell_x0 * layer, pitch, 0xFF, first * layer, rows, st);
        if (err == cudaSuccess && first < cell_nx) err = cudaMemset2DAsync(g_cell_dev, pitch, 0xFF, (cell_nx - first) * layer, rows, st);
    }
    if (err != cudaSuccess) {
        set_error("hx_binning_update: memset: %s", cudaGetErrorString(err));
        return static_cast<int>(err);
    }
    vec3 cl = make_vec3(cell_length);
    uvec3 nc = {ncell[0], ncell[1], ncell[2]};
    if (nparticle > 0) {
        k_cell_claim<<<div_up(nparticle, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(position_hi_dev), static_cast<unsigned>(nparticle), cl, nc,
            cell_size, cell_count_dev, g_cell_dev, cell_index_dev, cell_x0, cell_nx, overflow_flag_dev);
        int rc = check_launch("k_cell_claim");
        if (rc) return rc;
    }
    unsigned pow2 = 32;
    while (pow2 < cell_size) pow2 <<= 1;
    unsigned warps = 4;
    while (warps > 1 && warps * pow2 * sizeof(unsigned) > 48 * 1024) warps >>= 1;
    if (warps * pow2 * sizeof(unsigned) > 48 * 1024) {
        set_error("hx_binning_update: cell size %u too large", cell_size);
        return HX_E_UNSUPPORTED;
    }
    size_t const visited = static_cast<size_t>(cell_nx) * rows;
    k_cell_sort<<<div_up(visited, warps), warps * 32, warps * pow2 * sizeof(unsigned), st>>>(
        g_cell_dev, cell_count_dev, static_cast<unsigned>(visited), cell_size, pow2, overflow_flag_dev, ncell[0], cell_x0, cell_nx);
    return check_launch("k_cell_sort");
}

extern "C" int hx_binning_update(const hx_float4* position_hi_dev, size_t nparticle,
                                 const float cell_length[3], const unsigned ncell[3], unsigned cell_size,
                                 uint32_t* g_cell_dev, uint32_t* cell_count_dev, uint32_t* cell_index_dev,
                                 int* overflow_flag_dev, void* stream)
{
    return binning_update_impl(position_hi_dev, nparticle, cell_length, ncell, cell_size, g_cell_dev, cell_count_dev,
                               cell_index_dev, overflow_flag_dev, stream, 0, ncell ? ncell[0] : 0);
}

extern "C" int hx_binning_update_slab(const hx_float4* position_hi_dev, size_t nparticle,
                                      const float cell_length[3], const unsigned ncell[3], unsigned cell_size,
                                      unsigned cell_x0, unsigned cell_nx, uint32_t* g_cell_dev, uint32_t* cell_count_dev,
                                      uint32_t* cell_index_dev, int* overflow_flag_dev, void* stream)
{
    return binning_update_impl(position_hi_dev, nparticle, cell_length, ncell, cell_size, g_cell_dev, cell_count_dev,
                               cell_index_dev, overflow_flag_dev, stream, cell_x0, cell_nx);
}

// The divisions of the path by cell and box lengths are r * MUFU.RCP(l) in the reference build (binning_kernel.cu:34-54,
// hilbert_kernel.cu:59); this evaluates that hardware reciprocal for host values so that a host-side restatement of
// the arithmetic (oracle/hx_oracle.c, a host class that wants to predict a cell) can use the device's value.
extern "C" int hx_device_rcp(const float* x_host, float* rcp_host, unsigned n)
{
    if (!x_host || !rcp_host || n == 0 || n > 1024) {
        hx::set_error("hx_device_rcp: invalid argument");
        return HX_E_INVALID;
    }
    float* d = nullptr;
    cudaError_t err = cudaMalloc(&d, 2 * n * sizeof(float));
    if (err == cudaSuccess) err = cudaMemcpy(d, x_host, n * sizeof(float), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        hx::k_device_rcp<<<(n + 127) / 128, 128>>>(d, d + n, n);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpy(rcp_host, d + n, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (err != cudaSuccess) {
        hx::set_error("hx_device_rcp: %s", cudaGetErrorString(err));
        return static_cast<int>(err);
    }
    return HX_OK;
}
