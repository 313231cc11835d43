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
#include <algorithm>

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
    unsigned ntype1, ntype2;
    int fill_all;
    float rr_cut_skin[HX_MAX_PAIR_TYPES];
    // home cells served by this launch: cell_nx layers of the grid in x starting at layer cell_x0 (periodic); the whole
    // grid for cell_x0 = 0, cell_nx = ncell.x.  A slab-decomposed rank launches blocks for its own layers only.
    unsigned cell_x0, cell_nx;
};

// block -> home cell (x fastest, as the cell index itself)
__device__ __forceinline__ unsigned home_cell(nlist_args const& a, unsigned block, int& cx, int& cy, int& cz)
{
    unsigned const xl = block % a.cell_nx, rest = block / a.cell_nx;
    unsigned x = a.cell_x0 + xl;
    if (x >= a.ncell.x) x -= a.ncell.x;
    cx = static_cast<int>(x);
    cy = static_cast<int>(rest % a.ncell.y);
    cz = static_cast<int>(rest / a.ncell.y);
    return (static_cast<unsigned>(cz) * a.ncell.y + static_cast<unsigned>(cy)) * a.ncell.x + x;
}

constexpr unsigned NL_THREADS = 256;
constexpr unsigned NL_WARPS = NL_THREADS / 32;
constexpr unsigned NL_CELLS_PER_WARP = (27 + NL_WARPS - 1) / NL_WARPS;
constexpr unsigned NL_PAD = 64;                // sentinel slots behind the staged particles
constexpr unsigned NL_MAX_CHUNKS = 4;          // cell_size <= 128 keeps the staging loops fully unrolled

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

// Lists of the home cell's particles: distance tests over the staged candidates, ballot compaction, list stores.
//   IMG   the minimum-image convention is applied.  It is skipped when the staged positions span less than half
//         a box length in every direction: the image number of every pair is then exactly zero and
//         d - 0 * L == d, so not a bit changes
//   TYPED cutoff looked up per species pair
template <bool ROW, bool IMG, bool TYPED>
__device__ __forceinline__ void nlist_lists(nlist_args const& a, const float4* s_pos, const unsigned* s_id, const float* s_rr,
                                            unsigned M, unsigned home_off, unsigned home_cnt)
{
    unsigned const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned const lane_lt = (1u << lane) - 1u;
    float const rr_cut0 = a.rr_cut_skin[0];
    unsigned const Mpad = (M + 63u) & ~63u;          // [M, Mpad) holds NaN sentinels that fail every distance test
    unsigned const size = a.size;
    for (unsigned p = warp; p < home_cnt; p += NL_WARPS) {
        unsigned const n = s_id[home_off + p];
        if (n >= a.nbuild) break;                  // ascending indices within a cell: the rest are partners only
        float4 const r1 = s_pos[home_off + p];
        unsigned* row = ROW ? a.g_neighbour + static_cast<size_t>(n) * a.stride : a.g_neighbour + n;
        if (ROW) asm volatile("" : "+l"(row));     // keep the row base in registers instead of recomputing it per store
        const unsigned* si = s_id + lane;
        const float4* sp = s_pos + lane;
        unsigned count = 0;
        // two chunks per trip: independent shared-memory loads and distance chains in flight
        for (unsigned base = 0; base < Mpad; base += 64) {
            unsigned const m0 = si[base], m1 = si[base + 32];
            float4 const q0 = sp[base], q1 = sp[base + 32];
            float dx0 = __fsub_rn(r1.x, q0.x), dy0 = __fsub_rn(r1.y, q0.y), dz0 = __fsub_rn(r1.z, q0.z);
            float dx1 = __fsub_rn(r1.x, q1.x), dy1 = __fsub_rn(r1.y, q1.y), dz1 = __fsub_rn(r1.z, q1.z);
            if (IMG) {
                dx0 = minimg(dx0, a.L.x, a.invL.x); dy0 = minimg(dy0, a.L.y, a.invL.y); dz0 = minimg(dz0, a.L.z, a.invL.z);
                dx1 = minimg(dx1, a.L.x, a.invL.x); dy1 = minimg(dy1, a.L.y, a.invL.y); dz1 = minimg(dz1, a.L.z, a.invL.z);
            }
            float const cut0 = TYPED ? s_rr[__float_as_uint(r1.w) * a.ntype2 + __float_as_uint(q0.w)] : rr_cut0;
            float const cut1 = TYPED ? s_rr[__float_as_uint(r1.w) * a.ntype2 + __float_as_uint(q1.w)] : rr_cut0;
            bool const ok0 = (m0 != n) && (norm2(dx0, dy0, dz0) <= cut0);          // note <= : from_binning_kernel.cu:143
            bool const ok1 = (m1 != n) && (norm2(dx1, dy1, dz1) <= cut1);
            unsigned const mask0 = __ballot_sync(0xFFFFFFFFu, ok0);
            unsigned const mask1 = __ballot_sync(0xFFFFFFFFu, ok1);
            unsigned const slot0 = count + __popc(mask0 & lane_lt);
            count += __popc(mask0);
            unsigned const slot1 = count + __popc(mask1 & lane_lt);
            count += __popc(mask1);
            if (ROW) {
                if (ok0 && slot0 < size) row[slot0] = m0;
                if (ok1 && slot1 < size) row[slot1] = m1;
            } else {
                if (ok0 && slot0 < size) row[static_cast<size_t>(slot0) * a.stride] = m0;
                if (ok1 && slot1 < size) row[static_cast<size_t>(slot1) * a.stride] = m1;
            }
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

template <bool ROW>
__global__ void __launch_bounds__(NL_THREADS) k_nlist_build(nlist_args a)
{
    extern __shared__ float4 s_dyn[];
    __shared__ unsigned s_cnt[27];
    __shared__ unsigned s_off[28];
    __shared__ unsigned s_cell[27];
    __shared__ float s_rr[HX_MAX_PAIR_TYPES];
    __shared__ float s_ext[NL_WARPS][6];

    int cx, cy, cz;
    unsigned const cell = home_cell(a, blockIdx.x, cx, cy, cz);
    unsigned const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned const cs = a.cell_size;

    // occupied entries sit at the front of each cell (ascending index, placeholders last)
    // empty home cell, or one that holds partners only (their indices follow those of the list owners)
    if (a.g_cell[static_cast<size_t>(cell) * cs] >= a.nbuild) return;

    float4* s_pos = s_dyn;                                             // [27*cs + NL_PAD] (x, y, z, species)
    unsigned* s_id = reinterpret_cast<unsigned*>(s_dyn + 27 * cs + NL_PAD);     // [27*cs + NL_PAD]
    if (threadIdx.x < HX_MAX_PAIR_TYPES) s_rr[threadIdx.x] = a.rr_cut_skin[threadIdx.x];

    // phase 1: occupancy of the 27 cells; the loads of a warp's cells are independent and issued together
    unsigned const nchunk = cs / 32;
#pragma unroll
    for (unsigned kk = 0; kk < NL_CELLS_PER_WARP; ++kk) {
        unsigned const k = warp + kk * NL_WARPS;
        if (k < 27) {
            int dx, dy, dz;
            visit_offset(k, dx, dy, dz);
            unsigned const c = (wrap_cell(cz, dz, a.ncell.z) * a.ncell.y + wrap_cell(cy, dy, a.ncell.y)) * a.ncell.x
                               + wrap_cell(cx, dx, a.ncell.x);
            const unsigned* src = a.g_cell + static_cast<size_t>(c) * cs;
            unsigned cnt = 0;
            if (nchunk <= NL_MAX_CHUNKS) {
                unsigned v[NL_MAX_CHUNKS];
#pragma unroll
                for (unsigned u = 0; u < NL_MAX_CHUNKS; ++u) v[u] = (u < nchunk) ? src[u * 32 + lane] : HX_PH;
#pragma unroll
                for (unsigned u = 0; u < NL_MAX_CHUNKS; ++u) cnt += __popc(__ballot_sync(0xFFFFFFFFu, v[u] != HX_PH));
            } else {
                for (unsigned base = 0; base < cs; base += 32) {
                    unsigned m = __ballot_sync(0xFFFFFFFFu, src[base + lane] != HX_PH);
                    cnt += __popc(m);
                    if (m != 0xFFFFFFFFu) break;
                }
            }
            if (lane == 0) { s_cnt[k] = cnt; s_cell[k] = c; }
        }
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the 27 counts
        unsigned const c = lane < 27 ? s_cnt[lane] : 0u;
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= static_cast<unsigned>(o)) incl += t;
        }
        if (lane < 28) s_off[lane] = incl - c;      // lane 27: total
    }
    __syncthreads();
    // phase 2: stage ids and positions, packed in visiting order; ids (L1/L2 hits now) and the dependent position
    // gathers of all of a warp's cells are in flight together.  Each lane tracks the extent of what it staged.
    if (threadIdx.x < NL_PAD) {
        // sentinels behind the staged entries: the list loop runs over whole pairs of chunks without bounds tests
        unsigned const M = s_off[27];
        float const qnan = __uint_as_float(0x7FC00000u);
        s_id[M + threadIdx.x] = HX_PH;
        s_pos[M + threadIdx.x] = make_float4(qnan, qnan, qnan, 0.f);
    }
    float const big = 3.0e38f;
    float lo_x = big, lo_y = big, lo_z = big, hi_x = -big, hi_y = -big, hi_z = -big;
    auto stage = [&](unsigned slot, unsigned id) {
        float4 const r = __ldg(&a.pos_hi[id]);
        s_id[slot] = id;
        s_pos[slot] = r;
        lo_x = fminf(lo_x, r.x); hi_x = fmaxf(hi_x, r.x);
        lo_y = fminf(lo_y, r.y); hi_y = fmaxf(hi_y, r.y);
        lo_z = fminf(lo_z, r.z); hi_z = fmaxf(hi_z, r.z);
    };
#pragma unroll
    for (unsigned kk = 0; kk < NL_CELLS_PER_WARP; ++kk) {
        unsigned const k = warp + kk * NL_WARPS;
        if (k < 27) {
            const unsigned* src = a.g_cell + static_cast<size_t>(s_cell[k]) * cs;
            unsigned const cnt = s_cnt[k], off = s_off[k];
            if (nchunk <= NL_MAX_CHUNKS) {
#pragma unroll
                for (unsigned u = 0; u < NL_MAX_CHUNKS; ++u) {
                    unsigned const e = u * 32 + lane;
                    if (e < cnt) stage(off + e, src[e]);
                }
            } else {
                for (unsigned e = lane; e < cnt; e += 32) stage(off + e, src[e]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xFFFFFFFFu, lo_x, o)); hi_x = fmaxf(hi_x, __shfl_xor_sync(0xFFFFFFFFu, hi_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xFFFFFFFFu, lo_y, o)); hi_y = fmaxf(hi_y, __shfl_xor_sync(0xFFFFFFFFu, hi_y, o));
        lo_z = fminf(lo_z, __shfl_xor_sync(0xFFFFFFFFu, lo_z, o)); hi_z = fmaxf(hi_z, __shfl_xor_sync(0xFFFFFFFFu, hi_z, o));
    }
    if (lane == 0) {
        s_ext[warp][0] = lo_x; s_ext[warp][1] = lo_y; s_ext[warp][2] = lo_z;
        s_ext[warp][3] = hi_x; s_ext[warp][4] = hi_y; s_ext[warp][5] = hi_z;
    }
    __syncthreads();

    unsigned const M = s_off[27];
    unsigned const home_off = s_off[26], home_cnt = s_cnt[26];
    // extent of the staged positions (the same value in every thread)
    float ext[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        float v = s_ext[0][c];
#pragma unroll
        for (unsigned w = 1; w < NL_WARPS; ++w) v = c < 3 ? fminf(v, s_ext[w][c]) : fmaxf(v, s_ext[w][c]);
        ext[c] = v;
    }
    bool const interior = ext[3] - ext[0] < 0.49f * a.L.x && ext[4] - ext[1] < 0.49f * a.L.y && ext[5] - ext[2] < 0.49f * a.L.z;
    bool const typed = a.ntype2 * a.ntype1 > 1;
    if (typed) {
        if (interior) nlist_lists<ROW, false, true>(a, s_pos, s_id, s_rr, M, home_off, home_cnt);
        else nlist_lists<ROW, true, true>(a, s_pos, s_id, s_rr, M, home_off, home_cnt);
    } else {
        if (interior) nlist_lists<ROW, false, false>(a, s_pos, s_id, s_rr, M, home_off, home_cnt);
        else nlist_lists<ROW, true, false>(a, s_pos, s_id, s_rr, M, home_off, home_cnt);
    }
}

// Cells too large for the shared-memory staging (27 * cell_size entries; strongly inhomogeneous systems, where the
// reference keeps doubling the cell size: binning.cpp:216-219): the same lists straight from global memory.  One block
// per home cell, one warp per home particle at a time, the 27 cells in the reference's visiting order, 32 slots per trip.
template <bool ROW>
__global__ void __launch_bounds__(NL_THREADS) k_nlist_build_direct(nlist_args a)
{
    int cx, cy, cz;
    unsigned const cell = home_cell(a, blockIdx.x, cx, cy, cz);
    unsigned const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned const lane_lt = (1u << lane) - 1u;
    unsigned const cs = a.cell_size;
    const unsigned* const home = a.g_cell + static_cast<size_t>(cell) * cs;
    if (home[0] >= a.nbuild) return;
    bool const typed = a.ntype1 * a.ntype2 > 1;
    for (unsigned p = warp; p < cs; p += NL_WARPS) {
        unsigned const n = home[p];
        if (n >= a.nbuild) break;                  // placeholder, or a partner-only particle (ascending indices)
        float4 const r1 = a.pos_hi[n];
        unsigned* const row = ROW ? a.g_neighbour + static_cast<size_t>(n) * a.stride : a.g_neighbour + n;
        unsigned count = 0;
        for (unsigned k = 0; k < 27; ++k) {
            int dx, dy, dz;
            visit_offset(k, dx, dy, dz);
            unsigned const c = (wrap_cell(cz, dz, a.ncell.z) * a.ncell.y + wrap_cell(cy, dy, a.ncell.y)) * a.ncell.x
                               + wrap_cell(cx, dx, a.ncell.x);
            const unsigned* const src = a.g_cell + static_cast<size_t>(c) * cs;
            for (unsigned base = 0; base < cs; base += 32) {
                unsigned const m = src[base + lane];
                bool const real = m != HX_PH;
                bool ok = false;
                if (real && m != n) {
                    float4 const q = a.pos_hi[m];
                    float const ex = minimg(__fsub_rn(r1.x, q.x), a.L.x, a.invL.x);
                    float const ey = minimg(__fsub_rn(r1.y, q.y), a.L.y, a.invL.y);
                    float const ez = minimg(__fsub_rn(r1.z, q.z), a.L.z, a.invL.z);
                    float const cut = typed ? a.rr_cut_skin[__float_as_uint(r1.w) * a.ntype2 + __float_as_uint(q.w)] : a.rr_cut_skin[0];
                    ok = norm2(ex, ey, ez) <= cut;
                }
                unsigned const mask = __ballot_sync(0xFFFFFFFFu, ok);
                unsigned const slot = count + __popc(mask & lane_lt);
                count += __popc(mask);
                if (ok && slot < a.size) {
                    if (ROW) row[slot] = m;
                    else row[static_cast<size_t>(slot) * a.stride] = m;
                }
                if (!__all_sync(0xFFFFFFFFu, real)) break;       // occupied entries sit at the front of a cell
            }
        }
        if (count >= a.size) {
            count = a.size;
            if (lane == 0) *a.overflow = 1;
        }
        if (ROW) {
            size_t const end = a.fill_all ? a.stride : min(static_cast<size_t>((count + 32u) & ~31u), a.stride);
            for (size_t k = count + lane; k < end; k += 32) row[k] = HX_PH;
        }
        if (a.count_out && lane == 0) a.count_out[n] = count;
    }
}

// ---- private rows ordered by shells of the build-time distance (see force.cu, k_force_shell) -----------------------
struct shell_build_args {
    const unsigned* nl;
    const unsigned* count;
    const float4* pos1;
    const float4* pos2;
    unsigned* rows;
    uint4* ends;
    size_t stride;
    unsigned n, pitch, ntype2;
    vec3 L, invL;
    float t2[HX_MAX_PAIR_TYPES][HX_NSHELL - 1];     // (r_cut_ab + delta[s])^2
};

constexpr unsigned SB_WARPS = 8;
constexpr unsigned SB_MAX = 512;                     // longest list served (entries)
constexpr unsigned SB_SLOTS = SB_MAX / 32;

// One warp per particle.  Pass 1 classifies every entry of the public row (gather, distance as in the list build,
// thresholds) and counts, per slot of 32 entries and shell, with match.any; an exclusive scan over (shell, slot)
// then gives every entry its place: shells in ascending order, entries of a shell in the order of the public row,
// so the result is deterministic.
__global__ void __launch_bounds__(SB_WARPS * 32) k_shell_rows(shell_build_args a)
{
    __shared__ unsigned s_j[SB_WARPS][SB_MAX];
    __shared__ unsigned char s_sr[SB_WARPS][SB_MAX];          // shell | rank within (slot, shell) << 3
    __shared__ unsigned s_cnt[SB_WARPS][HX_NSHELL * SB_SLOTS];   // [shell][slot]
    unsigned const lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned const i = blockIdx.x * SB_WARPS + w;
    if (i >= a.n) return;
    unsigned const lane_lt = (1u << lane) - 1u;
    unsigned const cnt = min(a.count[i], SB_MAX);
    unsigned const nslot = (cnt + 31u) / 32u;
    float4 const r1 = a.pos1[i];
    const float* const t2 = a.t2[__float_as_uint(r1.w) * a.ntype2];
    const unsigned* const src = a.nl + static_cast<size_t>(i) * a.stride;
#pragma unroll
    for (unsigned c = lane; c < HX_NSHELL * SB_SLOTS; c += 32) s_cnt[w][c] = 0;
    __syncwarp();
    // classification of one entry: gather, distance as in the list build, thresholds, rank within (slot, shell)
    auto classify = [&](unsigned slot, unsigned j, float4 const& r2) {
        unsigned const k = slot * 32 + lane;
        bool const valid = k < cnt;
        float const dx = minimg(__fsub_rn(r1.x, r2.x), a.L.x, a.invL.x);
        float const dy = minimg(__fsub_rn(r1.y, r2.y), a.L.y, a.invL.y);
        float const dz = minimg(__fsub_rn(r1.z, r2.z), a.L.z, a.invL.z);
        float const rr = norm2(dx, dy, dz);
        const float* const th = t2 + (HX_NSHELL - 1) * __float_as_uint(r2.w);
        unsigned sh = 0;
#pragma unroll
        for (int q = 0; q < HX_NSHELL - 1; ++q) sh += rr > th[q] ? 1u : 0u;
        if (!valid) sh = HX_NSHELL;                             // (a group of its own, not counted)
        unsigned const m = __match_any_sync(0xFFFFFFFFu, sh);
        unsigned const rank = __popc(m & lane_lt);
        if (valid) {
            if (rank == 0) s_cnt[w][sh * SB_SLOTS + slot] = __popc(m);
            s_j[w][k] = j;
            s_sr[w][k] = static_cast<unsigned char>(sh | (rank << 3));
        }
    };
    // the kernel is latency-bound (index load -> position gather -> arithmetic): the first four slots (128 entries,
    // nearly every list) have their index loads, then their gathers, in flight together
    {
        unsigned jj[4];
        float4 rj[4];
#pragma unroll
        for (unsigned q = 0; q < 4; ++q) jj[q] = (q * 32 + lane < cnt) ? src[q * 32 + lane] : i;
#pragma unroll
        for (unsigned q = 0; q < 4; ++q) rj[q] = a.pos2[jj[q]];
#pragma unroll
        for (unsigned q = 0; q < 4; ++q)
            if (q < nslot) classify(q, jj[q], rj[q]);
    }
    for (unsigned slot = 4; slot < nslot; ++slot) {
        unsigned const k = slot * 32 + lane;
        unsigned const j = k < cnt ? src[k] : i;
        classify(slot, j, a.pos2[j]);
    }
    __syncwarp();
    // exclusive scan over the HX_NSHELL * SB_SLOTS = 128 cells in (shell, slot) order: 4 consecutive cells per lane
    unsigned c[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = s_cnt[w][4 * lane + q]; sum += c[q]; }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned const t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= static_cast<unsigned>(o)) incl += t;
    }
    unsigned base = incl - sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) { s_cnt[w][4 * lane + q] = base; base += c[q]; }
    // end of shell s = inclusive sum through its last cell (lane 4 s + 3 holds cells of shell s only: SB_SLOTS == 16)
    static_assert(SB_SLOTS == 16 && HX_NSHELL == 8, "scan layout");
    unsigned e[HX_NSHELL];
#pragma unroll
    for (int sidx = 0; sidx < HX_NSHELL; ++sidx) e[sidx] = __shfl_sync(0xFFFFFFFFu, incl, 4 * sidx + 3);
    __syncwarp();
    unsigned* const dst = a.rows + shell_row_base(i, a.pitch);
    for (unsigned slot = 0; slot < nslot; ++slot) {
        unsigned const k = slot * 32 + lane;
        if (k < cnt) {
            unsigned const sr = s_sr[w][k];
            unsigned const sh = sr & 7u, rank = sr >> 3;
            dst[shell_row_offset(s_cnt[w][sh * SB_SLOTS + slot] + rank)] = s_j[w][k];
        }
    }
    // behind its entries the row holds the particle's own index — safe to gather, masked by its position — as far as
    // the force kernel walks the rows of this tile of 8 particles: up to the longest of them, in steps of 16 entries
    unsigned const tile = i & ~7u;
    unsigned longest = (lane < 8u && tile + lane < a.n) ? min(a.count[tile + lane], SB_MAX) : 0u;
    longest = __reduce_max_sync(0xFFFFFFFFu, longest);
    unsigned const fill_end = min((longest + 15u) & ~15u, a.pitch);
    for (unsigned k = cnt + lane; k < fill_end; k += 32) dst[shell_row_offset(k)] = i;
    if (lane == 0)
        a.ends[i] = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    // rows of a last, incomplete tile and the slack behind it (read ahead by the force kernel, never used)
    if (i == a.n - 1) {
        size_t const end = static_cast<size_t>((a.n + 7u) / 8u) * 8u * a.pitch;
        for (unsigned g = (a.n & 7u) ? (a.n & 7u) : 8u; g < 8u; ++g)
            for (unsigned k = lane; k < a.pitch; k += 32) a.rows[shell_row_base(a.n - (a.n & 7u) + g, a.pitch) + shell_row_offset(k)] = 0;
        for (unsigned k = lane; k < HX_SHELL_ROWS_SLACK; k += 32) a.rows[end + k] = 0;
    }
}

__global__ void k_count_max(const unsigned* __restrict__ count, unsigned n, unsigned* __restrict__ out)
{
    unsigned m = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, count[i]);
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31u) == 0 && m > 0) atomicMax(out, m);
}

} // namespace hx

using namespace hx;

static int nlist_build_impl(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                            const hx_float4* position_hi_dev, size_t nparticle, size_t nbuild, const float* rr_cut_skin,
                            unsigned ntype1, unsigned ntype2, const float box_length[3],
                            uint32_t* g_neighbour_dev, unsigned size, size_t stride, int layout,
                            int fill_all, uint32_t* count_dev, int* overflow_flag_dev, void* stream,
                            unsigned cell_x0 = 0, unsigned cell_nx = 0)
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
    if (cell_nx == 0) { cell_x0 = 0; cell_nx = ncell[0]; }
    if (cell_x0 >= ncell[0] || cell_nx > ncell[0]) {
        set_error("hx_nlist_build_slab: cell layers [%u, +%u) outside the grid", cell_x0, cell_nx);
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    size_t ncells = static_cast<size_t>(cell_nx) * ncell[1] * ncell[2];        // home cells of this launch
    size_t smem = (static_cast<size_t>(27) * cell_size + NL_PAD) * (sizeof(float4) + sizeof(unsigned));
    bool const direct = smem > 200 * 1024;      // cells too large to stage: lists straight from global memory
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
    a.ntype1 = ntype1;
    a.ntype2 = ntype2;
    a.fill_all = fill_all;
    a.cell_x0 = cell_x0;
    a.cell_nx = cell_nx;
    for (unsigned i = 0; i < HX_MAX_PAIR_TYPES; ++i) a.rr_cut_skin[i] = (i < ntype1 * ntype2) ? rr_cut_skin[i] : 0.f;

    cudaError_t err;
    static std::atomic<size_t> smem_set[2] = {{0}, {0}};   // opt-in dynamic shared memory, raised on demand
    int const li = (layout == HX_LAYOUT_ROW) ? 1 : 0;
    if (!direct && smem > 48 * 1024 && smem_set[li].load() < smem) {
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
        if (direct) k_nlist_build_direct<false><<<static_cast<unsigned>(ncells), NL_THREADS, 0, st>>>(a);
        else k_nlist_build<false><<<static_cast<unsigned>(ncells), NL_THREADS, smem, st>>>(a);
    } else {
        if (direct) k_nlist_build_direct<true><<<static_cast<unsigned>(ncells), NL_THREADS, 0, st>>>(a);
        else k_nlist_build<true><<<static_cast<unsigned>(ncells), NL_THREADS, smem, st>>>(a);
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

extern "C" int hx_nlist_build_slab(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                                   unsigned cell_x0, unsigned cell_nx, const hx_float4* position_hi_dev, size_t nparticle,
                                   size_t nbuild, const float* rr_cut_skin, unsigned ntype1, unsigned ntype2,
                                   const float box_length[3], uint32_t* g_neighbour_dev, unsigned size, size_t stride,
                                   int layout, int fill_all, uint32_t* count_dev, int* overflow_flag_dev, void* stream)
{
    if (cell_nx == 0) { set_error("hx_nlist_build_slab: empty range of cell layers"); return HX_E_INVALID; }
    return nlist_build_impl(g_cell_dev, ncell, cell_size, position_hi_dev, nparticle, nbuild, rr_cut_skin, ntype1, ntype2,
                            box_length, g_neighbour_dev, size, stride, layout, fill_all, count_dev, overflow_flag_dev, stream,
                            cell_x0, cell_nx);
}

extern "C" int hx_shell_rows_plan(float skin, hx_shell_rows* shells)
{
    if (!shells || !(skin > 0)) {
        set_error("hx_shell_rows_plan: invalid argument");
        return HX_E_INVALID;
    }
    for (int s = 0; s < HX_NSHELL; ++s) shells->delta[s] = skin * static_cast<float>(s + 1) / HX_NSHELL;
    return HX_OK;
}

extern "C" size_t hx_shell_rows_words(size_t nparticle, unsigned pitch)
{
    return (nparticle + 7) / 8 * 8 * pitch + HX_SHELL_ROWS_SLACK;
}

extern "C" int hx_count_max(const uint32_t* count_dev, size_t n, uint32_t* max_dev, void* stream)
{
    if (!count_dev || !max_dev || n >= 0xFFFFFFFFull) {
        set_error("hx_count_max: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(max_dev, 0, sizeof(uint32_t), st);
    if (err != cudaSuccess) {
        set_error("hx_count_max: memset: %s", cudaGetErrorString(err));
        return static_cast<int>(err);
    }
    if (n == 0) return HX_OK;
    unsigned const blocks = static_cast<unsigned>(std::min<size_t>((n + 1023) / 1024, 592));
    k_count_max<<<blocks, 256, 0, st>>>(count_dev, static_cast<unsigned>(n), max_dev);
    return check_launch("k_count_max");
}

extern "C" int hx_shell_rows_build(const hx_lj_params* potential, const uint32_t* g_neighbour_dev, size_t stride,
                                   const uint32_t* count_dev, const hx_float4* position1_hi_dev,
                                   const hx_float4* position2_hi_dev, size_t nparticle, const float box_length[3],
                                   const hx_shell_rows* shells, void* stream)
{
    if (!potential || !g_neighbour_dev || !count_dev || !position1_hi_dev || !position2_hi_dev || !box_length || !shells
        || !shells->rows_dev || !shells->ends_dev || shells->pitch < 32 || shells->pitch % 32 != 0 || shells->pitch > SB_MAX
        || nparticle >= 0x7FFFFFFFull || potential->ntype1 * potential->ntype2 == 0
        || potential->ntype1 * potential->ntype2 > HX_MAX_PAIR_TYPES) {
        set_error("hx_shell_rows_build: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    shell_build_args a;
    a.nl = g_neighbour_dev;
    a.count = count_dev;
    a.pos1 = reinterpret_cast<const float4*>(position1_hi_dev);
    a.pos2 = reinterpret_cast<const float4*>(position2_hi_dev);
    a.rows = shells->rows_dev;
    a.ends = reinterpret_cast<uint4*>(shells->ends_dev);
    a.stride = stride;
    a.n = static_cast<unsigned>(nparticle);
    a.pitch = shells->pitch;
    a.ntype2 = potential->ntype2;
    a.L = make_vec3(box_length);
    a.invL = {1.0f / box_length[0], 1.0f / box_length[1], 1.0f / box_length[2]};
    for (unsigned t = 0; t < HX_MAX_PAIR_TYPES; ++t)
        for (int s = 0; s < HX_NSHELL - 1; ++s) {
            float const r = (t < potential->ntype1 * potential->ntype2 ? potential->r_cut[t] : 0.f) + shells->delta[s];
            a.t2[t][s] = r * r;
        }
    k_shell_rows<<<div_up(nparticle, SB_WARPS), SB_WARPS * 32, 0, as_stream(stream)>>>(a);
    return check_launch("k_shell_rows");
}
