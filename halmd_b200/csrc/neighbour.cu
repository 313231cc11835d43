// Verlet neighbour lists from binned cells.
//
// Reference: halmd/mdsim/gpu/neighbours/from_binning_kernel.cu:99-267 — one block per cell, one
// thread per placeholder; each thread walks all placeholders of the 27 surrounding cells (13 pairs of
// opposite cells, then its own cell) and appends m when rr <= (r_c + skin)^2.
//
// Here: one block per cell stages the *occupied* entries of the 27 cells (same visiting order) in
// shared memory as one packed candidate array; each warp then owns one particle of the home cell at a
// time, tests 32 candidates per step and compacts the accepted ones with ballot/popc, which preserves
// the reference's list order exactly.  Row layout writes are fully coalesced.
#include "common.cuh"

namespace hx {

struct nlist_args {
    const unsigned* g_cell;
    const float4* pos_hi;
    unsigned* g_neighbour;
    unsigned* count_out;
    int* overflow;
    uvec3 ncell;
    unsigned cell_size;
    unsigned size;
    unsigned nbuild;           // lists are built for particles with index < nbuild (all others are partners only)
    size_t stride;
    vec3 L, invL;
    unsigned ntype2;
    int fill_all;
    float rr_cut_skin[HX_MAX_PAIR_TYPES];
};

constexpr unsigned NL_THREADS = 128;
constexpr unsigned NL_WARPS = NL_THREADS / 32;

__device__ __forceinline__ unsigned wrap_cell(int c, int d, int n)
{
    int v = c + d;
    if (v < 0) v += n;
    if (v >= n) v -= n;
    return static_cast<unsigned>(v);
}

// k-th visited cell offset in the reference order: for j0,j1,j2 in {-1,0,1} lexicographic until
// (0,0,0): visit (+j) then (-j); finally the cell itself (k = 26).
__device__ __forceinline__ void visit_offset(unsigned k, int& dx, int& dy, int& dz)
{
    if (k >= 26) { dx = dy = dz = 0; return; }
    unsigned pair = k >> 1;          // 0..12: index into the first 13 of the 27 lexicographic offsets
    int j0 = static_cast<int>(pair / 9) - 1;
    int j1 = static_cast<int>((pair / 3) % 3) - 1;
    int j2 = static_cast<int>(pair % 3) - 1;
    int s = (k & 1u) ? -1 : 1;
    dx = s * j0; dy = s * j1; dz = s * j2;
}

template <bool ROW>
__global__ void __launch_bounds__(NL_THREADS) k_nlist_build(nlist_args a)
{
    extern __shared__ float4 s_dyn[];
    __shared__ unsigned s_cnt[27];
    __shared__ unsigned s_off[28];

    unsigned const cell = blockIdx.x;
    unsigned const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned const cs = a.cell_size;

    // occupied entries sit at the front of each cell (ascending index, placeholders last)
    // empty home cell, or one that holds partners only (their indices follow those of the list owners)
    if (a.g_cell[static_cast<size_t>(cell) * cs] >= a.nbuild) return;

    int const cx = cell % a.ncell.x;
    int const cy = (cell / a.ncell.x) % a.ncell.y;
    int const cz = cell / a.ncell.x / a.ncell.y;

    float4* s_pos = s_dyn;                                             // [27*cs] (x, y, z, species)
    unsigned* s_id = reinterpret_cast<unsigned*>(s_dyn + 27 * cs);     // [27*cs]

    // phase 1: occupancy of the 27 cells
    for (unsigned k = warp; k < 27; k += NL_WARPS) {
        int dx, dy, dz;
        visit_offset(k, dx, dy, dz);
        unsigned c = (wrap_cell(cz, dz, a.ncell.z) * a.ncell.y + wrap_cell(cy, dy, a.ncell.y)) * a.ncell.x
                     + wrap_cell(cx, dx, a.ncell.x);
        const unsigned* src = a.g_cell + static_cast<size_t>(c) * cs;
        unsigned cnt = 0;
        for (unsigned base = 0; base < cs; base += 32) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, src[base + lane] != HX_PH);
            cnt += __popc(m);
            if (m != 0xFFFFFFFFu) break;
        }
        if (lane == 0) s_cnt[k] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned o = 0;
        for (unsigned k = 0; k < 27; ++k) { s_off[k] = o; o += s_cnt[k]; }
        s_off[27] = o;
    }
    __syncthreads();
    // phase 2: stage ids and positions, packed in visiting order
    for (unsigned k = warp; k < 27; k += NL_WARPS) {
        int dx, dy, dz;
        visit_offset(k, dx, dy, dz);
        unsigned c = (wrap_cell(cz, dz, a.ncell.z) * a.ncell.y + wrap_cell(cy, dy, a.ncell.y)) * a.ncell.x
                     + wrap_cell(cx, dx, a.ncell.x);
        const unsigned* src = a.g_cell + static_cast<size_t>(c) * cs;
        unsigned cnt = s_cnt[k], off = s_off[k];
        for (unsigned e = lane; e < cnt; e += 32) {
            unsigned id = src[e];
            s_id[off + e] = id;
            s_pos[off + e] = __ldg(&a.pos_hi[id]);
        }
    }
    __syncthreads();

    unsigned const M = s_off[27];
    unsigned const home_off = s_off[26], home_cnt = s_cnt[26];

    for (unsigned p = warp; p < home_cnt; p += NL_WARPS) {
        unsigned const n = s_id[home_off + p];
        if (n >= a.nbuild) break;                  // ascending indices within a cell: the rest are partners only
        float4 const r1 = s_pos[home_off + p];
        unsigned const type1 = __float_as_uint(r1.w);
        unsigned* row = ROW ? a.g_neighbour + static_cast<size_t>(n) * a.stride : a.g_neighbour + n;
        unsigned count = 0;
        for (unsigned base = 0; base < M; base += 32) {
            unsigned e = base + lane;
            bool ok = false;
            unsigned m = HX_PH;
            if (e < M) {
                m = s_id[e];
                float4 r2 = s_pos[e];
                float dx = minimg(__fsub_rn(r1.x, r2.x), a.L.x, a.invL.x);
                float dy = minimg(__fsub_rn(r1.y, r2.y), a.L.y, a.invL.y);
                float dz = minimg(__fsub_rn(r1.z, r2.z), a.L.z, a.invL.z);
                float rr = norm2(dx, dy, dz);
                unsigned type2 = __float_as_uint(r2.w);
                // note <= : from_binning_kernel.cu:143
                ok = (m != n) && (rr <= a.rr_cut_skin[type1 * a.ntype2 + type2]);
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, ok);
            unsigned slot = count + __popc(mask & ((1u << lane) - 1u));
            if (ok && slot < a.size) {
                if (ROW) row[slot] = m;
                else row[static_cast<size_t>(slot) * a.stride] = m;
            }
            count += __popc(mask);
        }
        if (count >= a.size) {
            // from_binning_kernel.cu:263-266 — list fully occupied ⇒ overcrowded
            count = a.size;
            if (lane == 0) *a.overflow = 1;
        }
        if (ROW) {
            // terminate the row with placeholders up to the next multiple of 32 (or the whole pitch)
            size_t end = a.fill_all ? a.stride : min(static_cast<size_t>((count + 32u) & ~31u), a.stride);
            for (size_t k = count + lane; k < end; k += 32) row[k] = HX_PH;
        }
        if (a.count_out && lane == 0) a.count_out[n] = count;
    }
}

} // namespace hx

using namespace hx;

static int nlist_build_impl(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                            const hx_float4* position_hi_dev, size_t nparticle, size_t nbuild, const float* rr_cut_skin,
                            unsigned ntype1, unsigned ntype2, const float box_length[3],
                            uint32_t* g_neighbour_dev, unsigned size, size_t stride, int layout,
                            int fill_all, uint32_t* count_dev, int* overflow_flag_dev, void* stream)
{
    if (!g_cell_dev || !ncell || !position_hi_dev || !rr_cut_skin || !box_length || !g_neighbour_dev
        || !overflow_flag_dev || size == 0 || cell_size == 0 || cell_size % 32 != 0
        || ntype1 * ntype2 == 0 || ntype1 * ntype2 > HX_MAX_PAIR_TYPES) {
        set_error("hx_nlist_build: invalid argument");
        return HX_E_INVALID;
    }
    if (ncell[0] < 3 || ncell[1] < 3 || ncell[2] < 3) {
        // from_binning.cpp:174-176
        set_error("hx_nlist_build: number of cells per dimension must be at least 3");
        return HX_E_INVALID;
    }
    if (layout == HX_LAYOUT_ROW) {
        if (stride % 32 != 0 || (fill_all ? stride < size : stride < size + 1)) {
            set_error("hx_nlist_build: row pitch must be a multiple of 32 and > size");
            return HX_E_INVALID;
        }
    } else if (layout == HX_LAYOUT_COLUMN) {
        if (stride < (nbuild < nparticle ? nbuild : nparticle)) {
            set_error("hx_nlist_build: column stride must be >= nparticle");
            return HX_E_INVALID;
        }
    } else {
        set_error("hx_nlist_build: unknown layout %d", layout);
        return HX_E_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(stream);
    size_t ncells = static_cast<size_t>(ncell[0]) * ncell[1] * ncell[2];
    size_t smem = static_cast<size_t>(27) * cell_size * (sizeof(float4) + sizeof(unsigned));
    if (smem > 200 * 1024) {
        set_error("hx_nlist_build: cell size %u exceeds shared memory staging capacity", cell_size);
        return HX_E_UNSUPPORTED;
    }
    nlist_args a;
    a.g_cell = g_cell_dev;
    a.pos_hi = reinterpret_cast<const float4*>(position_hi_dev);
    a.g_neighbour = g_neighbour_dev;
    a.count_out = count_dev;
    a.overflow = overflow_flag_dev;
    a.ncell = {ncell[0], ncell[1], ncell[2]};
    a.cell_size = cell_size;
    a.size = size;
    a.nbuild = static_cast<unsigned>(nbuild < nparticle ? nbuild : nparticle);
    a.stride = stride;
    a.L = make_vec3(box_length);
    a.invL = {1.0f / box_length[0], 1.0f / box_length[1], 1.0f / box_length[2]};
    a.ntype2 = ntype2;
    a.fill_all = fill_all;
    for (unsigned i = 0; i < HX_MAX_PAIR_TYPES; ++i) a.rr_cut_skin[i] = (i < ntype1 * ntype2) ? rr_cut_skin[i] : 0.f;

    cudaError_t err;
    static std::atomic<size_t> smem_set[2] = {{0}, {0}};   // opt-in dynamic shared memory, raised on demand
    int const li = (layout == HX_LAYOUT_ROW) ? 1 : 0;
    if (smem > 48 * 1024 && smem_set[li].load() < smem) {
        err = li ? cudaFuncSetAttribute(k_nlist_build<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                 : cudaFuncSetAttribute(k_nlist_build<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) {
            set_error("hx_nlist_build: smem attribute: %s", cudaGetErrorString(err));
            return static_cast<int>(err);
        }
        smem_set[li].store(smem);
    }
    if (layout == HX_LAYOUT_COLUMN) {
        // from_binning.cpp:191 — mark all placeholders as virtual particles
        if ((err = cudaMemsetAsync(g_neighbour_dev, 0xFF, static_cast<size_t>(size) * stride * sizeof(uint32_t), st)) != cudaSuccess) {
            set_error("hx_nlist_build: memset: %s", cudaGetErrorString(err));
            return static_cast<int>(err);
        }
        k_nlist_build<false><<<static_cast<unsigned>(ncells), NL_THREADS, smem, st>>>(a);
    } else {
        k_nlist_build<true><<<static_cast<unsigned>(ncells), NL_THREADS, smem, st>>>(a);
    }
    return check_launch("k_nlist_build");
}

extern "C" int hx_nlist_build(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                              const hx_float4* position_hi_dev, size_t nparticle, const float* rr_cut_skin,
                              unsigned ntype1, unsigned ntype2, const float box_length[3],
                              uint32_t* g_neighbour_dev, unsigned size, size_t stride, int layout,
                              int fill_all, uint32_t* count_dev, int* overflow_flag_dev, void* stream)
{
    return nlist_build_impl(g_cell_dev, ncell, cell_size, position_hi_dev, nparticle, nparticle, rr_cut_skin, ntype1, ntype2,
                            box_length, g_neighbour_dev, size, stride, layout, fill_all, count_dev, overflow_flag_dev, stream);
}

extern "C" int hx_nlist_build2(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                               const hx_float4* position_hi_dev, size_t nparticle, size_t nbuild,
                               const float* rr_cut_skin, unsigned ntype1, unsigned ntype2, const float box_length[3],
                               uint32_t* g_neighbour_dev, unsigned size, size_t stride, int layout, int fill_all,
                               uint32_t* count_dev, int* overflow_flag_dev, void* stream)
{
    return nlist_build_impl(g_cell_dev, ncell, cell_size, position_hi_dev, nparticle, nbuild, rr_cut_skin, ntype1, ntype2,
                            box_length, g_neighbour_dev, size, stride, layout, fill_all, count_dev, overflow_flag_dev, stream);
}
