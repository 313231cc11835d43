// Slab domain decomposition: device-side helpers for particle migration and ghost (halo) layers.
//
// The reference is single-GPU (doc/usage/multi-gpu.rst.in: independent replicas only); the decomposition is new
// (SURVEY §8e).  Each rank owns the particles whose wrapped x coordinate lies in its slab and keeps ghost copies of
// the neighbouring ranks' particles within r_c + skin of its two faces.  Everything here is deterministic: selections
// are expressed as small integer keys that the caller feeds to the stable radix sort (hx_sort_pairs_stable_u32), so
// that migrants and ghosts keep their relative memory order.  Host orchestration: halmd_b200/decomposition.py.
#include "common.cuh"
#include <cstring>

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
// Owned particles lie in [lo, hi) up to a period of the box: the integrator leaves coordinates in [-L/2, +L/2], and a
// particle at +L/2 belongs to the slab that starts at -L/2 (slab_of folds it).  The distances to the faces are folded
// the same way, or that particle would be sent to the wrong neighbour and missed by the right one.
__global__ void k_dd_ghost_keys(const float4* __restrict__ pos_hi, unsigned n, float lo, float hi, float halo, int side,
                                float box_length_x, unsigned* __restrict__ keys, unsigned* __restrict__ count)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (i < n) {
        float x = pos_hi[i].x;
        float d_lo = __fsub_rn(x, lo), d_hi = __fsub_rn(hi, x);
        // (a period away, not a rounding error of the face coordinates: owned particles may sit an ulp outside [lo, hi))
        float const half = 0.5f * box_length_x;
        if (d_lo > half) { d_lo = __fsub_rn(d_lo, box_length_x); d_hi = __fadd_rn(d_hi, box_length_x); }
        else if (d_hi > half) { d_lo = __fadd_rn(d_lo, box_length_x); d_hi = __fsub_rn(d_hi, box_length_x); }
        bool near_lo = d_lo < halo, near_hi = d_hi < halo;
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


// ---- per-step halo exchange and rebuild vote over peer memory ---------------------------------------------------------
// The ranks of one node map each other's position buffers and inboxes (CUDA IPC).  After the integration kernel of a
// step, k_gather_f4 stores the new positions of this rank's boundary particles straight into the ghost slots of the
// neighbours' position buffers (NVLink stores, no staging, no host), k_dd_signal publishes {step, local maximum squared
// displacement} in every other rank's inbox, and k_dd_wait — the only point where a rank waits for others — spins until
// all inbox entries carry this step, folds the global maximum into the control block (the rebuild vote) and lets the
// force kernel go.  Positions are double-buffered: step k lives in buffer k % 2, so a neighbour that is one kernel ahead
// writes the ghosts of step k + 1 while this rank's force kernel still reads those of step k.
__global__ void k_dd_signal(const unsigned* __restrict__ ctl, unsigned long long* const* __restrict__ peer_inbox, int nranks, int rank,
                            unsigned step)
{
    if (ctl[CTL_HALTED]) return;
    int const r = threadIdx.x;
    if (r >= nranks || r == rank) return;
    // the pushes of this step (earlier kernels of the stream) are performed before the flag becomes visible
    __threadfence_system();
    unsigned long long const word = (static_cast<unsigned long long>(step) << 32) | ctl[CTL_MAX_RR];
    // two inbox rows, used alternately: a peer that is a step ahead writes the other row, and cannot get two steps
    // ahead without this rank's next signal
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(peer_inbox[r] + (step & 1u) * HX_DD_MAX_RANKS + rank), "l"(word) : "memory");
}

__global__ void k_dd_wait(unsigned* __restrict__ ctl, const unsigned long long* __restrict__ inbox, int nranks, int rank, unsigned step,
                          long long timeout_cycles)
{
    if (ctl[CTL_HALTED]) return;
    int const r = threadIdx.x;
    unsigned max_rr = 0;
    bool timed_out = false;
    if (r < nranks && r != rank) {
        long long const t0 = clock64();
        for (;;) {
            unsigned long long word;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(inbox + (step & 1u) * HX_DD_MAX_RANKS + r) : "memory");
            if (static_cast<unsigned>(word >> 32) == step) { max_rr = static_cast<unsigned>(word); break; }
            if (clock64() - t0 > timeout_cycles) { timed_out = true; break; }
            __nanosleep(200);
        }
    }
    if (timed_out) {
        // a peer never arrived: stop the chunk instead of hanging the device; the host reports ctl[3]
        atomicExch(&ctl[CTL_RESERVED], 0xDEADu);
        atomicExch(&ctl[CTL_HALTED], 1u);
        return;
    }
    if (max_rr) atomicMax(&ctl[CTL_MAX_RR], max_rr);
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
                     float box_length_x, uint32_t* keys_dev, uint32_t* count_dev /*1, zeroed here*/, void* stream)
{
    if (!position_hi_dev || !keys_dev || !count_dev || side < 0 || side > 2 || nparticle >= 0x7FFFFFFFull || !(box_length_x > 0)) {
        set_error("hx_dd_ghost_keys: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st);
    if (err != cudaSuccess) { set_error("hx_dd_ghost_keys: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    if (nparticle == 0) return HX_OK;
    k_dd_ghost_keys<<<div_up(nparticle, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(position_hi_dev),
                                                          static_cast<unsigned>(nparticle), lo, hi, halo, side, box_length_x, keys_dev, count_dev);
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


// ---- device memory that other processes of the node can map (cudaMalloc'ed, not sub-allocated) ---------------------
int hx_dd_alloc(size_t bytes, void** out_dev)
{
    if (!out_dev || bytes == 0) { set_error("hx_dd_alloc: invalid argument"); return HX_E_INVALID; }
    cudaError_t err = cudaMalloc(out_dev, bytes);
    if (err == cudaSuccess) err = cudaMemset(*out_dev, 0, bytes);
    if (err != cudaSuccess) { set_error("hx_dd_alloc: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

int hx_dd_free(void* dev)
{
    cudaError_t err = cudaFree(dev);
    if (err != cudaSuccess) { set_error("hx_dd_free: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

int hx_dd_ipc_export(void* dev, unsigned char handle_out[HX_DD_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == HX_DD_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, dev);
    if (err != cudaSuccess) { set_error("hx_dd_ipc_export: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    memcpy(handle_out, &h, sizeof(h));
    return HX_OK;
}

int hx_dd_ipc_open(const unsigned char handle[HX_DD_IPC_HANDLE_BYTES], void** out_dev)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t err = cudaIpcOpenMemHandle(out_dev, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) { set_error("hx_dd_ipc_open: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

int hx_dd_ipc_close(void* dev)
{
    cudaError_t err = cudaIpcCloseMemHandle(dev);
    if (err != cudaSuccess) { set_error("hx_dd_ipc_close: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

// One chunk of speculatively enqueued MD steps of the decomposed loop (see the kernels above); nothing here touches
// the host between steps.  Returns after enqueueing; the caller reads the control block (steps completed, halted).
int hx_dd_enqueue_steps(const hx_dd_steps* a, unsigned chunk, void* stream)
{
    if (!a || !a->position_dev[0] || !a->position_dev[1] || !a->image_dev || !a->velocity_dev || !a->force_dev || !a->r0_dev
        || !a->ctl_dev || !a->potential || a->nranks < 1 || a->nranks > HX_DD_MAX_RANKS || a->rank < 0 || a->rank >= a->nranks
        || (a->current != 0 && a->current != 1) || (!a->shells && !a->g_neighbour_dev)
        || (a->nranks > 1 && (!a->inbox_dev || !a->peer_inbox_dev))) {
        set_error("hx_dd_enqueue_steps: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    int cur = a->current;
    bool const andersen = a->rng != nullptr;
    size_t const ntot = a->n_own + a->n_ghost;
    for (unsigned k = 0; k < chunk; ++k) {
        unsigned const step = static_cast<unsigned>(a->step_base + k + 1);
        hx_float4* const pin = a->position_dev[cur];
        hx_float4* const pout = a->position_dev[cur ^ 1];
        bool const fin = !andersen && (a->pending_finalize || k > 0);
        int rc = verlet_launch(fin, true, pin, a->image_dev, a->velocity_dev, a->force_dev, a->n_own, a->array_size, a->timestep,
                               a->box_length, a->r0_dev, a->ctl_dev, stream, a->n_own, pout);
        if (rc) return rc;
        if (a->nranks > 1) {
            for (int dir = 0; dir < 2; ++dir) {
                if (a->send_count[dir] == 0) continue;
                rc = hx_gather_float4(pout, a->send_index_dev[dir], a->send_count[dir], a->peer_ghost_dev[dir][cur ^ 1], a->ctl_dev, stream);
                if (rc) return rc;
            }
            k_dd_signal<<<1, HX_DD_MAX_RANKS, 0, st>>>(a->ctl_dev, a->peer_inbox_dev, a->nranks, a->rank, step);
            if ((rc = check_launch("k_dd_signal"))) return rc;
            k_dd_wait<<<1, HX_DD_MAX_RANKS, 0, st>>>(a->ctl_dev, a->inbox_dev, a->nranks, a->rank, step, a->timeout_cycles);
            if ((rc = check_launch("k_dd_wait"))) return rc;
        }
        int const flags = HX_FORCE_POSITIONS_REDUCED;
        if (a->shells) {
            rc = hx_pair_force_shells(a->potential, pout, pout, a->shells, a->r0_dev, a->n_own, a->box_length, flags, 1, a->force_dev,
                                      a->ctl_dev, a->skin_half, stream);
        } else {
            rc = hx_pair_force2(a->potential, pout, pout, ntot, a->g_neighbour_dev, a->size, a->stride, HX_LAYOUT_ROW | flags, a->n_own,
                                a->box_length, 1, 0.5f, a->force_dev, nullptr, nullptr, a->array_size, a->ctl_dev, a->skin_half, stream);
        }
        if (rc) return rc;
        if (andersen) {
            rc = hx_andersen_finalize2(a->velocity_dev, a->force_dev, a->n_own, a->array_size, a->timestep, a->sqrt_temperature,
                                       a->coll_prob, a->rng, a->rng_state_dev, a->andersen_scratch_dev, a->andersen_scratch_bytes,
                                       a->ctl_dev, stream);
            if (rc) return rc;
        }
        cur ^= 1;
    }
    return HX_OK;
}

} // extern "C"
