// Host-side module layer: C++ classes mirroring the host classes HALMD binds into Lua, driving the kernel-level
// C ABI (halmd_b200.h).  See include/halmd_b200_modules.h for the module ↔ reference map.
//
// Differences from the reference's host layer that are deliberate (B200-first):
//  * no cudaDeviceSynchronize after every launch; one stream per particle instance;
//  * the neighbour-list rebuild decision is evaluated on the device (control block, guarded kernels) and the
//    host enqueues chunks of steps speculatively (hxm_core_run) instead of reading 64 floats back every step;
//  * finalize(k) and integrate(k+1) are fused with the displacement reduction (bitwise identical results).
#include "common.cuh"
#include "../../include/halmd_b200_modules.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

using namespace hx;

#define HXM_TRY(expr) do { int rc__ = (expr); if (rc__ != 0) return rc__; } while (0)
#define HXM_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    set_error("%s: %s", #expr, cudaGetErrorString(e__)); return static_cast<int>(e__); } } while (0)

namespace {

template <typename T>
int dev_alloc(T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    HXM_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return 0;
}

template <typename T>
void dev_free(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

struct welford { double n = 0, m = 0, v = 0; };

// accumulator<double>::operator()(accumulator const&)  (halmd/numeric/accumulator.hpp:104-118)
void welford_merge(welford& a, welford const& b)
{
    if (b.n > 0) {
        double count = a.n + b.n;
        a.v += b.v;
        double diff = a.m - b.m;
        a.v += diff * diff * a.n * b.n / count;
        a.m = (a.n * a.m + b.n * b.m) / count;
        a.n = count;
    }
}

void welford_add(welford& a, double x)
{
    double t = x - a.m;
    a.n += 1;
    a.m += t / a.n;
    a.v += t * (x - a.m);
}

} // namespace

// ---- mdsim.particle ---------------------------------------------------------------------------------------------
struct hxm_particle {
    size_t n = 0, A = 0;
    unsigned nspecies = 1;
    cudaStream_t st = nullptr;
    float4 *pos = nullptr, *vel = nullptr, *image = nullptr, *force = nullptr;
    float4 *pos_b = nullptr, *vel_b = nullptr, *image_b = nullptr;     // back buffers for rearrange
    float4* pos_c = nullptr;       // second position buffer of the fused force + Verlet chain (hxm_core::run)
    float *en_pot = nullptr, *stress = nullptr;
    uint32_t *id = nullptr, *rid = nullptr, *id_b = nullptr;
    // cache counters (halmd/utility/cache.hpp): bumped on every write access
    uint64_t pos_ver = 1, order_ver = 1;
    uint64_t force_pos_ver = 0;      // position version the force array belongs to
    bool aux_enabled = true;         // particle.cpp:64 — enabled by default to allow sampling of the initial state
    bool aux_valid = false;
    // every position lies in the primary cell [-L/2, L/2]: true once an integration step has reduced them, false
    // after positions were written from outside (HX_FORCE_POSITIONS_REDUCED)
    bool reduced = false;
    double* io_dev = nullptr;        // [3 n] staging of double-precision samples (particle get / set)
    void* scratch = nullptr;         // grow-only device scratch of the observables (wavevectors, density modes)
    size_t scratch_bytes = 0;

    int get_scratch(size_t bytes, void** out)
    {
        if (bytes > scratch_bytes) {
            if (scratch) cudaFree(scratch);
            scratch = nullptr;
            scratch_bytes = 0;
            HXM_CUDA(cudaMalloc(&scratch, bytes));
            scratch_bytes = bytes;
        }
        *out = scratch;
        return 0;
    }

    ~hxm_particle()
    {
        dev_free(pos); dev_free(vel); dev_free(image); dev_free(force); dev_free(pos_b); dev_free(vel_b); dev_free(pos_c);
        dev_free(image_b); dev_free(en_pot); dev_free(stress); dev_free(id); dev_free(rid); dev_free(id_b);
        dev_free(io_dev);
        if (scratch) cudaFree(scratch);
    }
};

struct hxm_potential {
    hx_lj_params p;
};

// ---- mdsim.binning -------------------------------------------------------------------------------------------------
struct hxm_binning {
    hxm_particle* particle = nullptr;
    unsigned ncell[3] = {0, 0, 0};
    float cell_length[3] = {0, 0, 0};
    unsigned cell_size = 0;
    size_t ncells = 0;
    uint32_t *g_cell = nullptr, *cell_count = nullptr;
    int* flag = nullptr;            // device overflow flag
    int* h_flag = nullptr;          // pinned
    uint64_t cache_ver = 0;         // position version of g_cell

    ~hxm_binning()
    {
        dev_free(g_cell); dev_free(cell_count); dev_free(flag);
        if (h_flag) cudaFreeHost(h_flag);
    }

    int set_cell_size(unsigned cs)
    {
        cell_size = cs;
        dev_free(g_cell);
        return dev_alloc(&g_cell, ncells * cell_size);
    }

    // enqueue one binning pass (no overflow handling)
    int enqueue()
    {
        HXM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), particle->st));
        return hx_binning_update(reinterpret_cast<hx_float4*>(particle->pos), particle->n, cell_length, ncell, cell_size,
                                 g_cell, cell_count, nullptr, flag, particle->st);
    }

    // binning::update (binning.cpp:166-221): retry with doubled cell size while overcrowded
    int update()
    {
        for (;;) {
            HXM_TRY(enqueue());
            HXM_CUDA(cudaMemcpyAsync(h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, particle->st));
            HXM_CUDA(cudaStreamSynchronize(particle->st));
            if (*h_flag == 0) break;
            HXM_TRY(set_cell_size(2 * cell_size));
        }
        cache_ver = particle->pos_ver;
        return 0;
    }
};

// ---- mdsim.neighbour = from_binning + max_displacement + sorts.hilbert ---------------------------------------------
struct hxm_neighbour {
    hxm_particle* particle = nullptr;
    hxm_binning* binning = nullptr;
    bool own_binning = false;
    double L[3];
    float Lf[3];
    unsigned ntype1 = 1, ntype2 = 1;
    float rr_cut_skin[HX_MAX_PAIR_TYPES];
    float r_cut_max = 0;
    float r_skin = 0;              // float member as in from_binning.hpp:166
    float nu_cell = 0;             // float member (from_binning.hpp:174)
    int layout = HX_LAYOUT_ROW;
    bool sorting = true;
    unsigned size = 0;
    size_t stride = 0;
    uint32_t *g_neighbour = nullptr, *count = nullptr;
    int *flag = nullptr, *h_flag = nullptr;
    // private shell-ordered copy of the rows for the pair-force kernel (ROW layout only)
    bool use_shells = true;
    hx_shell_rows shells = {};
    size_t shell_words = 0;        // allocated words of shells.rows_dev
    uint32_t* count_max = nullptr; // device word: longest list of the last build
    const hx_lj_params* shell_potential = nullptr;   // cutoffs the shells refer to (set by the force module)
    bool shells_valid = false;
    // max_displacement
    float4* r0 = nullptr;
    uint32_t* ctl = nullptr;       // device control block (HX_CTL_WORDS words), ctl[0] = max rr bits, ctl[4] = second slot
    uint32_t* h_ctl = nullptr;     // pinned mirror
    // sorts.hilbert
    unsigned depth = 0;
    uint32_t *keys = nullptr, *index = nullptr;
    void* sort_scratch = nullptr;
    size_t sort_scratch_bytes = 0;
    // cache observers
    uint64_t order_ver_seen = 0;   // reverse_id cache the lists were built for
    uint64_t disp_pos_ver = 0;     // position version of the last displacement evaluation
    bool disp_exceeded = false;
    size_t updates = 0;

    ~hxm_neighbour()
    {
        if (own_binning) delete binning;
        dev_free(g_neighbour); dev_free(count); dev_free(flag); dev_free(r0); dev_free(ctl); dev_free(keys);
        dev_free(index); dev_free(shells.rows_dev); dev_free(shells.ends_dev); dev_free(count_max);
        if (sort_scratch) cudaFree(sort_scratch);
        if (h_flag) cudaFreeHost(h_flag);
        if (h_ctl) cudaFreeHost(h_ctl);
    }

    // from_binning::set_occupancy (from_binning.cpp:96-119)
    int set_occupancy(double occupancy)
    {
        nu_cell = static_cast<float>(occupancy);
        double volume = L[0] * L[1] * L[2];
        HXM_TRY(hx_nlist_plan(r_cut_max, r_skin, particle->n, volume, occupancy, binning->cell_size, &size));
        stride = (layout == HX_LAYOUT_ROW) ? ((static_cast<size_t>(size) + 32) / 32) * 32 : particle->A;
        dev_free(g_neighbour);
        size_t words = (layout == HX_LAYOUT_ROW) ? particle->n * stride : static_cast<size_t>(size) * stride;
        return dev_alloc(&g_neighbour, words);
    }

    // sorts::hilbert::order (sorts/hilbert.cpp:71-117) + particle::rearrange (particle.cpp:173-218)
    int order()
    {
        hxm_particle* p = particle;
        HXM_TRY(hx_hilbert_keys(reinterpret_cast<hx_float4*>(p->pos), p->n, Lf, depth, keys, p->st));
        HXM_TRY(hx_iota_u32(index, p->n, p->st));
        HXM_TRY(hx_sort_pairs_stable_u32(keys, index, p->n, 3 * depth, sort_scratch, sort_scratch_bytes, p->st));
        HXM_TRY(hx_state_permute(index, p->n, p->A, reinterpret_cast<hx_float4*>(p->pos),
                                 reinterpret_cast<hx_float4*>(p->pos_b), reinterpret_cast<hx_float4*>(p->vel),
                                 reinterpret_cast<hx_float4*>(p->vel_b), reinterpret_cast<hx_float4*>(p->image),
                                 reinterpret_cast<hx_float4*>(p->image_b), p->id, p->id_b, p->rid, p->st));
        std::swap(p->pos, p->pos_b);
        std::swap(p->vel, p->vel_b);
        std::swap(p->image, p->image_b);
        std::swap(p->id, p->id_b);
        p->pos_ver++;
        p->order_ver++;
        return 0;
    }

    // from_binning::update (from_binning.cpp:162-282) preceded by the on_prepend_update slot (sort) and followed by
    // displacement->zero(), as wired in from_binning::g_neighbour (:123-140) and neighbour.lua.in:294
    int update()
    {
        hxm_particle* p = particle;
        if (sorting) HXM_TRY(order());
        for (;;) {
            HXM_TRY(binning->enqueue());
            HXM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), p->st));
            HXM_TRY(hx_nlist_build(binning->g_cell, binning->ncell, binning->cell_size,
                                   reinterpret_cast<hx_float4*>(p->pos), p->n, rr_cut_skin, ntype1, ntype2, Lf,
                                   g_neighbour, size, stride, layout, 0, count, flag, p->st));
            bool const want_shells = use_shells && layout == HX_LAYOUT_ROW && shell_potential;
            if (want_shells) HXM_TRY(hx_count_max(count, p->n, count_max, p->st));
            // one host round-trip per list update for both overflow flags (and the longest list)
            HXM_CUDA(cudaMemcpyAsync(&h_flag[0], binning->flag, sizeof(int), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaMemcpyAsync(&h_flag[1], flag, sizeof(int), cudaMemcpyDeviceToHost, p->st));
            if (want_shells) HXM_CUDA(cudaMemcpyAsync(&h_flag[2], count_max, sizeof(int), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
            if (h_flag[0]) {
                // binning.cpp:216-219 — overcrowded cells: double the cell size
                HXM_TRY(binning->set_cell_size(2 * binning->cell_size));
                // the list size follows the cell size (from_binning::set_occupancy, from_binning.cpp:96-119)
                HXM_TRY(set_occupancy(nu_cell));
                continue;
            }
            if (h_flag[1]) {
                // from_binning.cpp:277-280 — overcrowded lists: halve the occupancy
                HXM_TRY(set_occupancy(nu_cell / 2));
                continue;
            }
            break;
        }
        binning->cache_ver = p->pos_ver;
        HXM_TRY(build_shells());
        // displacement->zero()
        HXM_TRY(hx_displacement_zero(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(r0), p->n, p->st));
        HXM_TRY(hx_ctl_reset(ctl, p->st));
        disp_pos_ver = p->pos_ver;
        disp_exceeded = false;
        order_ver_seen = p->order_ver;
        updates++;
        return 0;
    }

    // private rows for the pair-force kernel: pitch from the longest list of this build (grow-only allocation)
    int build_shells()
    {
        hxm_particle* p = particle;
        shells_valid = false;
        if (!(use_shells && layout == HX_LAYOUT_ROW && shell_potential)) return 0;
        unsigned const longest = static_cast<unsigned>(h_flag[2]);
        unsigned const pitch = ((longest + 32u) / 32u) * 32u;
        if (pitch > 512) return 0;                       // (longer lists are served by the public rows)
        size_t const words = hx_shell_rows_words(p->n, pitch);
        if (words > shell_words) {
            dev_free(shells.rows_dev);
            HXM_TRY(dev_alloc(&shells.rows_dev, words));
            shell_words = words;
        }
        if (!shells.ends_dev) HXM_TRY(dev_alloc(&shells.ends_dev, p->n * HX_NSHELL));
        shells.pitch = pitch;
        HXM_TRY(hx_shell_rows_plan(r_skin, &shells));
        hx_float4* pos = reinterpret_cast<hx_float4*>(p->pos);
        HXM_TRY(hx_shell_rows_build(shell_potential, g_neighbour, stride, count, pos, pos, p->n, Lf, &shells, p->st));
        shells_valid = true;
        return 0;
    }

    // from_binning::g_neighbour (from_binning.cpp:123-140): host-polled variant
    int check(bool* rebuilt)
    {
        hxm_particle* p = particle;
        if (rebuilt) *rebuilt = false;
        bool need = (updates == 0) || (order_ver_seen != p->order_ver);
        if (!need) {
            if (disp_pos_ver != p->pos_ver) {
                HXM_TRY(hx_displacement_max(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(r0), p->n,
                                            Lf, ctl, 1, p->st));
                HXM_CUDA(cudaMemcpyAsync(h_ctl, ctl, sizeof(uint32_t), cudaMemcpyDeviceToHost, p->st));
                HXM_CUDA(cudaStreamSynchronize(p->st));
                disp_exceeded = hx_displacement_exceeds(h_ctl[0], r_skin) != 0;
                disp_pos_ver = p->pos_ver;
            }
            need = disp_exceeded;
        }
        if (need) {
            HXM_TRY(update());
            if (rebuilt) *rebuilt = true;
        }
        return 0;
    }
};

// ---- mdsim.forces.pair_trunc ---------------------------------------------------------------------------------------------
struct hxm_force {
    hxm_potential* potential = nullptr;
    hxm_particle* particle = nullptr;
    hxm_neighbour* neighbour = nullptr;
    float Lf[3];
    float aux_weight = 1;
    // further contributions to the same force array (particle.on_force slots after the first: force_zero = false)
    std::vector<hx_planar_wall> walls;      // mdsim.forces.external with potentials.external.planar_wall

    // weight / 2 for particle1 == particle2 (pair_trunc.hpp:291-294)
    float weight() const { return aux_weight / 2; }

    int compute(bool aux, uint32_t* ctl, float skin_half)
    {
        hxm_particle* p = particle;
        hx_float4* pos = reinterpret_cast<hx_float4*>(p->pos);
        int const layout = neighbour->layout | (p->reduced ? HX_FORCE_POSITIONS_REDUCED : 0);
        if (!aux && neighbour->shells_valid && neighbour->shell_potential == &potential->p) {
            // shell-ordered private rows: walks only the entries that can lie inside the cutoff
            HXM_TRY(hx_pair_force_shells(&potential->p, pos, pos, &neighbour->shells,
                                         reinterpret_cast<hx_float4*>(neighbour->r0), p->n, Lf,
                                         p->reduced ? HX_FORCE_POSITIONS_REDUCED : 0, 1,
                                         reinterpret_cast<hx_float4*>(p->force), ctl, skin_half, p->st));
        } else if (ctl) {
            HXM_TRY(hx_pair_force_guarded(&potential->p, pos, pos, neighbour->g_neighbour, neighbour->size,
                                          neighbour->stride, layout, p->n, Lf, 1, weight(),
                                          reinterpret_cast<hx_float4*>(p->force), aux ? p->en_pot : nullptr,
                                          aux ? p->stress : nullptr, p->A, ctl, skin_half, p->st));
        } else {
            HXM_TRY(hx_pair_force(&potential->p, pos, pos, neighbour->g_neighbour, neighbour->size, neighbour->stride,
                                  layout, p->n, Lf, 1, weight(), reinterpret_cast<hx_float4*>(p->force),
                                  aux ? p->en_pot : nullptr, aux ? p->stress : nullptr, p->A, p->st));
        }
        // (a halted chunk: the guarded pair kernel raised the flag before this kernel reads it, in stream order)
        for (hx_planar_wall const& w : walls) {
            HXM_TRY(hx_external_planar_wall(&w, pos, p->n, Lf, 0, reinterpret_cast<hx_float4*>(p->force),
                                            aux ? p->en_pot : nullptr, aux ? p->stress : nullptr, p->A, ctl, p->st));
        }
        return 0;
    }

    // particle::update_force_ (particle.cpp:221-253) + pair_trunc::check_cache/apply (pair_trunc.hpp:175-215)
    int ensure(bool with_aux)
    {
        hxm_particle* p = particle;
        bool force_dirty = p->force_pos_ver != p->pos_ver;
        bool aux_dirty = !p->aux_valid || force_dirty;
        if (force_dirty || (with_aux && aux_dirty)) {
            if (with_aux && aux_dirty) p->aux_enabled = true;
            HXM_TRY(neighbour->check(nullptr));
            bool aux = p->aux_enabled;
            HXM_TRY(compute(aux, nullptr, 0.f));
            p->force_pos_ver = p->pos_ver;
            p->aux_valid = aux;
            p->aux_enabled = false;
        }
        return 0;
    }
};

// ---- integrators -------------------------------------------------------------------------------------------------------------
struct hxm_integrator {
    enum kind_t { VERLET, ANDERSEN, HOOVER } kind = VERLET;
    hxm_particle* particle = nullptr;
    hxm_force* force = nullptr;
    float Lf[3];
    double timestep = 0.001;
    // Andersen
    double temperature = 1, rate = 0;
    float sqrt_temperature = 1, coll_prob = 0;
    hx_rand48 rng;
    uint16_t* rng_state = nullptr;
    unsigned char* andersen_scratch = nullptr;
    // Nose-Hoover chain (verlet_nvt_hoover.hpp:100-160): heat-bath variables on the host
    hx_hoover_chain chain = {};
    double resonance_frequency = 5, en_nhc = 0;
    double* d_hoover = nullptr;      // device / pinned scratch of the kinetic-energy reduction
    double* h_hoover = nullptr;

    ~hxm_integrator()
    {
        dev_free(rng_state); dev_free(andersen_scratch); dev_free(d_hoover);
        if (h_hoover) cudaFreeHost(h_hoover);
    }

    // verlet_nvt_hoover::propagate_chain (verlet_nvt_hoover.cpp:185-222): returns the velocity scale factor
    int propagate_chain(double* scale)
    {
        hxm_particle* p = particle;
        HXM_TRY(hx_thermodynamics(reinterpret_cast<hx_float4*>(p->vel), nullptr, nullptr, p->A, p->n, d_hoover, p->st));
        HXM_CUDA(cudaMemcpyAsync(h_hoover, d_hoover, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->st));
        HXM_CUDA(cudaStreamSynchronize(p->st));
        // sum m v^2 = 2 E_kin of all particles drives the chain
        *scale = hx_hoover_chain_propagate(&chain, h_hoover[0], timestep);
        return 0;
    }

    void update_andersen()
    {
        // verlet_nvt_andersen.cpp:52-76: sqrt_temperature_ = sqrt(T), coll_prob_ = coll_rate * timestep (floats)
        sqrt_temperature = static_cast<float>(std::sqrt(temperature));
        coll_prob = static_cast<float>(rate * timestep);
    }

    int integrate()
    {
        hxm_particle* p = particle;
        HXM_TRY(force->ensure(false));      // read_cache(particle->force()) before invalidating the positions
        if (kind == HOOVER) {
            // verlet_nvt_hoover::integrate (verlet_nvt_hoover.cpp:103-135)
            double scale = 1;
            HXM_TRY(propagate_chain(&scale));
            HXM_TRY(hx_hoover_integrate(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->image),
                                        reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->A,
                                        static_cast<float>(timestep), scale, Lf, p->st));
            p->pos_ver++;
            p->reduced = true;
            return 0;
        }
        HXM_TRY(hx_verlet_integrate(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->image),
                                    reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->A,
                                    static_cast<float>(timestep), Lf, p->st));
        p->pos_ver++;
        p->reduced = true;
        return 0;
    }

    int finalize_kernel(uint32_t* ctl)
    {
        hxm_particle* p = particle;
        if (kind == HOOVER) {
            // verlet_nvt_hoover::finalize (verlet_nvt_hoover.cpp:140-183): half-kick, chain, rescale, chain energy
            HXM_TRY(hx_verlet_finalize(reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->A,
                                       static_cast<float>(timestep), p->st));
            double scale = 1;
            HXM_TRY(propagate_chain(&scale));
            HXM_TRY(hx_hoover_rescale(reinterpret_cast<hx_float4*>(p->vel), p->A, scale, p->st));
            en_nhc = hx_hoover_chain_energy(&chain, p->n);
            return 0;
        }
        if (kind == VERLET) {
            if (ctl) return hx_verlet_finalize_guarded(reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force),
                                                       p->A, static_cast<float>(timestep), ctl, p->st);
            return hx_verlet_finalize(reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->A,
                                      static_cast<float>(timestep), p->st);
        }
        return hx_andersen_finalize2(reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->n, p->A,
                                     static_cast<float>(timestep), sqrt_temperature, coll_prob, &rng, rng_state,
                                     andersen_scratch, hx_andersen_scratch_bytes(p->n), ctl, p->st);
    }

    int finalize()
    {
        HXM_TRY(force->ensure(false));      // triggers the (lazy) force computation for the new positions
        return finalize_kernel(nullptr);
    }
};

// ---- core / sampler -------------------------------------------------------------------------------------------------------------
struct hxm_core {
    hxm_integrator* integrator = nullptr;
    uint64_t step = 0;
    unsigned min_chunk = 2, max_chunk = 64;
    // NVE: velocity-Verlet update in the tail of the shell force kernel.  Off by default: measured on B200 at config 2 the
    // chain is 3.6 % slower than separate launches (75.8 vs 73.0 us per step) — the force kernel is issue-bound, so the
    // integrator's instructions (one coordinate per lane, 8 particles per warp) add to its critical path, while the
    // stand-alone k_verlet is latency-bound and overlaps the tail of the force kernel under programmatic dependent launch
    bool fused_verlet = false;
    unsigned last_interval = 16;      // steps between the last two list updates (chunk length predictor)
    uint64_t steps_since_update = 0;
    uint64_t stat_launches = 0, stat_chunks = 0, stat_syncs = 0;
    double* d_thermo = nullptr;
    double* h_thermo = nullptr;

    ~hxm_core()
    {
        dev_free(d_thermo);
        if (h_thermo) cudaFreeHost(h_thermo);
    }

    // core::mdstep (core.cpp:30-63)
    int mdstep()
    {
        HXM_TRY(integrator->integrate());
        HXM_TRY(integrator->finalize());
        step++;
        return 0;
    }

    // sampler::run (sampler.cpp:49-83) without observables: `steps` MD steps.
    int run(uint64_t steps)
    {
        if (steps == 0) return 0;
        hxm_integrator* in = integrator;
        hxm_force* f = in->force;
        hxm_neighbour* nb = f->neighbour;
        hxm_particle* p = in->particle;
        uint64_t const launches0 = hx_launch_count();
        float const dt = static_cast<float>(in->timestep);
        // float(r_skin / 2) as compared in from_binning.cpp:130-131
        float const skin_half = static_cast<float>(nb->r_skin / 2);
        bool const fuse = (in->kind == hxm_integrator::VERLET);
        if (in->kind == hxm_integrator::HOOVER) {
            // the chain is propagated on the host from the kinetic energy before and after every kick: no speculation
            for (uint64_t k = 0; k < steps; ++k) HXM_TRY(mdstep());
            stat_launches += hx_launch_count() - launches0;
            return 0;
        }

        HXM_TRY(f->ensure(false));                 // forces for the current positions (first step only)
        bool pending_finalize = false;             // velocities still lack the second half-kick of the last step
        uint64_t remaining = steps;
        while (remaining > 0) {
            unsigned predicted = last_interval > steps_since_update ? static_cast<unsigned>(last_interval - steps_since_update) : 1;
            unsigned chunk = std::min<uint64_t>(std::max(min_chunk, std::min(max_chunk, predicted + 2)), remaining);
            // reset halted flag, step counter (and the second displacement slot), keep the running maximum displacement
            HXM_CUDA(cudaMemsetAsync(nb->ctl + CTL_HALTED, 0, (CTL_MAX_RR_B - CTL_HALTED + 1) * sizeof(uint32_t), p->st));
            p->reduced = true;      // the integration kernel reduces every position before the force kernel reads it
            // NVE with private shell rows and no further force contributions: all steps of the chunk but the first
            // integrate in the tail of the force kernel of the step before (one launch per step)
            bool const chain = fuse && fused_verlet && p->pos_c && chunk > 1 && f->walls.empty() && nb->shells_valid
                               && nb->shell_potential == &f->potential->p;
            if (chain) {
                HXM_TRY(hx_verlet_finalize_integrate(
                    reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->image),
                    reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->n, p->A, dt, in->Lf,
                    pending_finalize ? 1 : 0, reinterpret_cast<hx_float4*>(nb->r0), nb->ctl, p->st));
                float4* const buf[2] = {p->pos, p->pos_c};
                unsigned const word[2] = {CTL_MAX_RR, CTL_MAX_RR_B};
                hx_verlet_tail t;
                t.velocity_dev = reinterpret_cast<hx_float4*>(p->vel);
                t.image_dev = reinterpret_cast<hx_float4*>(p->image);
                t.array_size = p->A;
                t.timestep = dt;
                int cur = 0;
                for (unsigned k = 0; k < chunk; ++k) {
                    bool const last = k + 1 == chunk;
                    t.position_out_dev = last ? nullptr : reinterpret_cast<hx_float4*>(buf[cur ^ 1]);
                    t.max_rr_word_in = word[cur];
                    HXM_TRY(hx_pair_force_shells_verlet(&f->potential->p, reinterpret_cast<hx_float4*>(buf[cur]), &nb->shells,
                                                        reinterpret_cast<hx_float4*>(nb->r0), p->n, f->Lf, HX_FORCE_POSITIONS_REDUCED,
                                                        last ? reinterpret_cast<hx_float4*>(p->force) : nullptr, nb->ctl, skin_half,
                                                        &t, p->st));
                    if (!last) cur ^= 1;
                }
            } else {
                for (unsigned k = 0; k < chunk; ++k) {
                    HXM_TRY(hx_verlet_finalize_integrate(
                        reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->image),
                        reinterpret_cast<hx_float4*>(p->vel), reinterpret_cast<hx_float4*>(p->force), p->n, p->A, dt, in->Lf,
                        (fuse && (pending_finalize || k > 0)) ? 1 : 0, reinterpret_cast<hx_float4*>(nb->r0), nb->ctl, p->st));
                    HXM_TRY(f->compute(false, nb->ctl, skin_half));
                    if (!fuse) HXM_TRY(in->finalize_kernel(nb->ctl));
                }
            }
            HXM_CUDA(cudaMemcpyAsync(nb->h_ctl, nb->ctl, HX_CTL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
            stat_chunks++;
            stat_syncs++;
            if (chain) {
                // fused launches completed before the chunk ended or halted: an odd number leaves the current positions in
                // the second buffer and the running maximum displacement in the second slot
                uint64_t const fused_done = std::min<uint64_t>(nb->h_ctl[CTL_STEPS], chunk - 1);
                if (fused_done & 1u) {
                    std::swap(p->pos, p->pos_c);
                    HXM_CUDA(cudaMemcpyAsync(nb->ctl + CTL_MAX_RR, nb->ctl + CTL_MAX_RR_B, sizeof(uint32_t), cudaMemcpyDeviceToDevice, p->st));
                }
            }
            uint64_t done = nb->h_ctl[CTL_STEPS];
            bool halted = nb->h_ctl[CTL_HALTED] != 0;
            p->pos_ver += done + (halted ? 1 : 0);
            steps_since_update += done;
            if (halted) {
                // positions of step done+1 are in place; lists are stale: update them and complete the step
                last_interval = static_cast<unsigned>(std::max<uint64_t>(steps_since_update + 1, 1));
                HXM_TRY(nb->update());
                stat_syncs++;
                HXM_TRY(f->compute(false, nullptr, 0.f));
                if (!fuse) HXM_TRY(in->finalize_kernel(nullptr));
                done += 1;
                steps_since_update = 0;
            } else if (steps_since_update > last_interval) {
                last_interval = static_cast<unsigned>(steps_since_update);
            }
            p->force_pos_ver = p->pos_ver;
            p->aux_valid = false;
            nb->disp_pos_ver = p->pos_ver;      // the device-side test covered these positions
            nb->disp_exceeded = false;
            if (fuse) pending_finalize = true;
            remaining -= done;
            step += done;
        }
        if (pending_finalize) HXM_TRY(in->finalize_kernel(nullptr));
        stat_launches += hx_launch_count() - launches0;
        return 0;
    }

    int thermodynamics(double out[10])
    {
        hxm_integrator* in = integrator;
        hxm_force* f = in->force;
        hxm_particle* p = in->particle;
        HXM_TRY(f->ensure(true));
        HXM_TRY(hx_thermodynamics(reinterpret_cast<hx_float4*>(p->vel), p->en_pot, p->stress, p->A, p->n, d_thermo, p->st));
        HXM_CUDA(cudaMemcpyAsync(h_thermo, d_thermo, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->st));
        HXM_CUDA(cudaStreamSynchronize(p->st));
        double const N = static_cast<double>(p->n);
        double const volume = f->neighbour->L[0] * f->neighbour->L[1] * f->neighbour->L[2];
        double en_kin = 0.5 * h_thermo[0] / N;
        double en_pot = h_thermo[5] / N;
        double virial = h_thermo[6] / N;
        double density = N / volume;
        double temp = 2 * en_kin / 3;                  // thermodynamics.hpp:73-76
        out[0] = en_kin;
        out[1] = en_pot;
        out[2] = virial;
        out[3] = temp;
        out[4] = density * (temp + virial / 3);        // thermodynamics.hpp:78-81
        out[5] = en_kin + en_pot;
        for (int d = 0; d < 3; ++d) out[6 + d] = h_thermo[1 + d] / h_thermo[4];
        out[9] = density;
        return 0;
    }
};

// ---- blocking scheme ------------------------------------------------------------------------------------------------------------
struct hxm_blocking {
    // device buffers of retired samples are reused: cudaMalloc / cudaFree of a 16 MB buffer costs tens of milliseconds,
    // far more than the sampling kernels themselves
    struct buffer_pool {
        std::vector<float4*> free_list;
        ~buffer_pool() { for (float4* p : free_list) cudaFree(p); }
        int get(float4** out, size_t n)
        {
            if (!free_list.empty()) { *out = free_list.back(); free_list.pop_back(); return 0; }
            return dev_alloc(out, n);
        }
    };
    struct sample {
        std::shared_ptr<buffer_pool> pool;
        float4* positions = nullptr;         // unfolded positions in id order (device)
        float4* velocities = nullptr;        // velocities in id order (device)
        std::vector<double> rho;             // density modes (re, im) per wavevector
        uint64_t step = 0;
        ~sample()
        {
            if (positions) pool->free_list.push_back(positions);
            if (velocities) pool->free_list.push_back(velocities);
        }
    };
    std::shared_ptr<buffer_pool> pool = std::make_shared<buffer_pool>();
    hxm_core* core = nullptr;
    double timestep = 0;                     // time step the sampling intervals were computed for
    unsigned block_size = 10, separation = 10;
    std::vector<uint64_t> interval, origin;
    std::vector<std::deque<std::shared_ptr<sample>>> blocks;
    std::shared_ptr<sample> current;         // sample acquired at the current step (shared between levels)
    // correlations
    std::vector<welford> msd;                // [level][lag]
    std::vector<welford> mqd, vacf;          // [level][lag]: mean-quartic displacement, velocity autocorrelation
    std::vector<welford> isf;                // [level][lag][shell]
    unsigned nq = 0, nshell = 0;
    std::vector<unsigned> shell_offsets;
    float* d_q = nullptr;
    double *d_rho = nullptr, *d_msd = nullptr;
    void* d_msd_scratch = nullptr;
    size_t msd_scratch_bytes = 0;
    const float4** d_ptrs = nullptr;
    double norm = 1;
    bool trace_on = false;                   // record (level, first step, second step) of every correlated sample pair
    std::vector<uint64_t> trace;

    ~hxm_blocking()
    {
        dev_free(d_q); dev_free(d_rho); dev_free(d_msd); dev_free(d_ptrs);
        if (d_msd_scratch) cudaFree(d_msd_scratch);
    }

    // phase_space position sampler + density_mode::acquire, cached per step
    int acquire(std::shared_ptr<sample>& out)
    {
        hxm_particle* p = core->integrator->particle;
        if (current && current->step == core->step) { out = current; return 0; }
        auto s = std::make_shared<sample>();
        s->pool = pool;
        s->step = core->step;
        HXM_TRY(pool->get(&s->positions, p->n));
        float* Lf = core->integrator->Lf;
        HXM_TRY(hx_sample_positions(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->image), p->rid, p->n,
                                    Lf, reinterpret_cast<hx_float4*>(s->positions), p->st));
        HXM_TRY(pool->get(&s->velocities, p->n));
        HXM_TRY(hx_sample_velocities(reinterpret_cast<hx_float4*>(p->vel), p->rid, p->n,
                                     reinterpret_cast<hx_float4*>(s->velocities), p->st));
        if (nq > 0) {
            HXM_TRY(hx_density_modes(reinterpret_cast<hx_float4*>(p->pos), p->n, d_q, nq, d_rho, p->st));
            s->rho.resize(2 * nq);
            HXM_CUDA(cudaMemcpyAsync(s->rho.data(), d_rho, 2 * nq * sizeof(double), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
        }
        current = s;
        out = s;
        return 0;
    }

    // correlation<tcf>::compute (correlation.hpp:208-224) for both correlation functions, then
    // blocking_scheme::process (blocking_scheme.cpp:176-200)
    int process(unsigned level)
    {
        hxm_particle* p = core->integrator->particle;
        auto& block = blocks[level];
        unsigned nl = static_cast<unsigned>(block.size());
        if (nl > 0) {
            std::vector<const float4*> ptrs(nl);
            for (unsigned k = 0; k < nl; ++k) ptrs[k] = block[k]->positions;
            HXM_CUDA(cudaMemcpyAsync(d_ptrs, ptrs.data(), nl * sizeof(float4*), cudaMemcpyHostToDevice, p->st));
            HXM_TRY(hx_msd_block(reinterpret_cast<const hx_float4*>(block[0]->positions),
                                 reinterpret_cast<const hx_float4* const*>(d_ptrs), nl, p->n, d_msd, d_msd_scratch,
                                 msd_scratch_bytes, p->st));
            std::vector<double> h(3 * nl);
            HXM_CUDA(cudaMemcpyAsync(h.data(), d_msd, 3 * nl * sizeof(double), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
            for (unsigned k = 0; k < nl; ++k) {
                welford w;
                w.n = h[3 * k]; w.m = h[3 * k + 1]; w.v = h[3 * k + 2];
                welford_merge(msd[level * block_size + k], w);
            }
            // mean-quartic displacement on the same samples, velocity autocorrelation on the velocity samples
            for (int kind = HX_CORR_MQD; kind <= HX_CORR_VACF; ++kind) {
                if (kind == HX_CORR_VACF) {
                    for (unsigned k = 0; k < nl; ++k) ptrs[k] = block[k]->velocities;
                    HXM_CUDA(cudaMemcpyAsync(d_ptrs, ptrs.data(), nl * sizeof(float4*), cudaMemcpyHostToDevice, p->st));
                }
                HXM_TRY(hx_correlate_block(kind, reinterpret_cast<const hx_float4*>(ptrs[0]),
                                           reinterpret_cast<const hx_float4* const*>(d_ptrs), nl, p->n, d_msd, d_msd_scratch,
                                           msd_scratch_bytes, p->st));
                HXM_CUDA(cudaMemcpyAsync(h.data(), d_msd, 3 * nl * sizeof(double), cudaMemcpyDeviceToHost, p->st));
                HXM_CUDA(cudaStreamSynchronize(p->st));
                auto& acc = kind == HX_CORR_MQD ? mqd : vacf;
                for (unsigned k = 0; k < nl; ++k) {
                    welford w;
                    w.n = h[3 * k]; w.m = h[3 * k + 1]; w.v = h[3 * k + 2];
                    welford_merge(acc[level * block_size + k], w);
                }
            }
            if (nq > 0) {
                // intermediate_scattering_function.hpp:90-113
                static_assert(sizeof(welford) == 3 * sizeof(double), "welford is (count, mean, M2)");
                for (unsigned k = 0; k < nl; ++k)
                    HXM_TRY(hx_isf_accumulate(block[0]->rho.data(), block[k]->rho.data(), shell_offsets.data(), nshell, norm,
                                              &isf[(level * block_size + k) * nshell].n));
            }
            if (trace_on)
                for (unsigned k = 0; k < nl; ++k) {
                    trace.push_back(level); trace.push_back(block[0]->step); trace.push_back(block[k]->step);
                }
        }
        uint64_t skip = std::max<uint64_t>((separation + interval[level] - 1) / interval[level], 1);
        origin[level] += skip * interval[level];
        for (uint64_t k = 0; k < skip && !block.empty(); ++k) block.pop_front();
        return 0;
    }

    // blocking_scheme::sample (blocking_scheme.cpp:129-161)
    int do_sample()
    {
        // blocking_scheme.cpp:136-139
        if (core->integrator->timestep != timestep) {
            set_error("blocking scheme does not allow variable time step");
            return HX_E_INVALID;
        }
        uint64_t const step = core->step;
        for (unsigned i = 0; i < interval.size(); ++i) {
            if (step >= origin[i] && (step - origin[i]) % interval[i] == 0) {
                std::shared_ptr<sample> s;
                HXM_TRY(acquire(s));
                blocks[i].push_back(s);
                if (blocks[i].size() == block_size) HXM_TRY(process(i));
            }
        }
        return 0;
    }

    bool due(uint64_t step) const
    {
        for (unsigned i = 0; i < interval.size(); ++i)
            if (step >= origin[i] && (step - origin[i]) % interval[i] == 0) return true;
        return false;
    }
};

// =====================================================================================================================
extern "C" {

int hxm_particle_create(size_t nparticle, unsigned nspecies, void* stream, hxm_particle** out)
{
    if (!out || nparticle == 0 || nparticle >= 0x7FFFFFFFull) { set_error("hxm_particle_create: invalid argument"); return HX_E_INVALID; }
    if (hx_device_count() < 1) return HX_E_NOGPU;
    auto p = std::unique_ptr<hxm_particle>(new hxm_particle);
    p->n = nparticle;
    p->A = hx_array_size(nparticle);
    p->nspecies = std::max(nspecies, 1u);
    p->st = as_stream(stream);
    size_t A = p->A;
    HXM_TRY(dev_alloc(&p->pos, 2 * A)); HXM_TRY(dev_alloc(&p->vel, 2 * A)); HXM_TRY(dev_alloc(&p->image, A));
    HXM_TRY(dev_alloc(&p->force, A)); HXM_TRY(dev_alloc(&p->pos_b, 2 * A)); HXM_TRY(dev_alloc(&p->vel_b, 2 * A));
    HXM_TRY(dev_alloc(&p->image_b, A)); HXM_TRY(dev_alloc(&p->en_pot, A)); HXM_TRY(dev_alloc(&p->stress, 6 * A));
    HXM_TRY(dev_alloc(&p->id, A)); HXM_TRY(dev_alloc(&p->rid, A)); HXM_TRY(dev_alloc(&p->id_b, A));
    HXM_TRY(hx_state_init(reinterpret_cast<hx_float4*>(p->pos), reinterpret_cast<hx_float4*>(p->vel),
                          reinterpret_cast<hx_float4*>(p->image), reinterpret_cast<hx_float4*>(p->force), p->id, p->rid,
                          p->n, A, p->st));
    HXM_CUDA(cudaMemsetAsync(p->en_pot, 0, A * sizeof(float), p->st));
    HXM_CUDA(cudaMemsetAsync(p->stress, 0, 6 * A * sizeof(float), p->st));
    HXM_CUDA(cudaStreamSynchronize(p->st));
    *out = p.release();
    return 0;
}

void hxm_particle_destroy(hxm_particle* p) { delete p; }
size_t hxm_particle_nparticle(const hxm_particle* p) { return p->n; }
size_t hxm_particle_array_size(const hxm_particle* p) { return p->A; }
unsigned hxm_particle_nspecies(const hxm_particle* p) { return p->nspecies; }

static int particle_float4_io(hxm_particle* p, float4* dev, bool lo_plane, bool set, bool tag, void* host, int ncomp)
{
    // stage through a host float4 buffer
    size_t n = p->n, A = p->A;
    std::vector<float4> buf(n);
    HXM_CUDA(cudaMemcpyAsync(buf.data(), dev, n * sizeof(float4), cudaMemcpyDeviceToHost, p->st));
    HXM_CUDA(cudaStreamSynchronize(p->st));
    if (!set) {
        if (tag) {
            uint32_t* h = static_cast<uint32_t*>(host);
            for (size_t i = 0; i < n; ++i) std::memcpy(&h[i], &buf[i].w, 4);
        } else {
            float* h = static_cast<float*>(host);
            for (size_t i = 0; i < n; ++i) {
                h[ncomp * i] = buf[i].x;
                h[ncomp * i + 1] = buf[i].y;
                h[ncomp * i + 2] = buf[i].z;
            }
        }
        return 0;
    }
    if (tag) {
        const uint32_t* h = static_cast<const uint32_t*>(host);
        for (size_t i = 0; i < n; ++i) std::memcpy(&buf[i].w, &h[i], 4);
    } else {
        const float* h = static_cast<const float*>(host);
        for (size_t i = 0; i < n; ++i) {
            buf[i].x = h[ncomp * i];
            buf[i].y = h[ncomp * i + 1];
            buf[i].z = h[ncomp * i + 2];
        }
    }
    HXM_CUDA(cudaMemcpyAsync(dev, buf.data(), n * sizeof(float4), cudaMemcpyHostToDevice, p->st));
    if (lo_plane && !tag) HXM_CUDA(cudaMemsetAsync(dev + A, 0, n * sizeof(float4), p->st));   // low words are lost on set
    HXM_CUDA(cudaStreamSynchronize(p->st));
    return 0;
}

// double <-> double-single on the device: one 24-byte-per-particle transfer each way instead of staging both float4
// planes through the host; the split is dsdeq (dsfun.hpp:122-126), the fourth components (species / mass) stay
__global__ void k_ds_split(float4* __restrict__ dev, unsigned A, unsigned n, const double* __restrict__ src)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double const x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    float const hx_ = static_cast<float>(x), hy = static_cast<float>(y), hz = static_cast<float>(z);
    float4 h = dev[i];
    h.x = hx_; h.y = hy; h.z = hz;
    dev[i] = h;
    dev[i + A] = make_float4(static_cast<float>(x - hx_), static_cast<float>(y - hy), static_cast<float>(z - hz), 0.f);
}

__global__ void k_ds_join(const float4* __restrict__ dev, unsigned A, unsigned n, double* __restrict__ dst)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 const h = dev[i], l = dev[i + A];
    dst[3 * i] = static_cast<double>(h.x) + l.x;
    dst[3 * i + 1] = static_cast<double>(h.y) + l.y;
    dst[3 * i + 2] = static_cast<double>(h.z) + l.z;
}

static int particle_ds_io(hxm_particle* p, float4* dev, bool set, double* host)
{
    size_t const n = p->n;
    if (n == 0) return 0;
    if (!p->io_dev) HXM_TRY(dev_alloc(&p->io_dev, 3 * n));
    unsigned const blocks = static_cast<unsigned>((n + 255) / 256);
    if (set) {
        HXM_CUDA(cudaMemcpyAsync(p->io_dev, host, 3 * n * sizeof(double), cudaMemcpyHostToDevice, p->st));
        k_ds_split<<<blocks, 256, 0, p->st>>>(dev, static_cast<unsigned>(p->A), static_cast<unsigned>(n), p->io_dev);
        HXM_CUDA(cudaGetLastError());
    } else {
        k_ds_join<<<blocks, 256, 0, p->st>>>(dev, static_cast<unsigned>(p->A), static_cast<unsigned>(n), p->io_dev);
        HXM_CUDA(cudaGetLastError());
        HXM_CUDA(cudaMemcpyAsync(host, p->io_dev, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, p->st));
    }
    HXM_CUDA(cudaStreamSynchronize(p->st));
    return 0;
}

static int particle_io(hxm_particle* p, const char* name_c, void* host, bool set)
{
    if (!p || !name_c || !host) { set_error("hxm_particle_get/set: null argument"); return HX_E_INVALID; }
    std::string name(name_c);
    size_t n = p->n, A = p->A;
    if (name == "position" || name == "species") {
        HXM_TRY(particle_float4_io(p, p->pos, true, set, name == "species", host, 3));
        if (set) { p->pos_ver++; p->reduced = false; }
        return 0;
    }
    if (name == "velocity" || name == "mass") {
        if (name == "mass") {
            std::vector<float4> buf(n);
            HXM_CUDA(cudaMemcpyAsync(buf.data(), p->vel, n * sizeof(float4), cudaMemcpyDeviceToHost, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
            float* h = static_cast<float*>(host);
            if (!set) { for (size_t i = 0; i < n; ++i) h[i] = buf[i].w; return 0; }
            for (size_t i = 0; i < n; ++i) buf[i].w = h[i];
            HXM_CUDA(cudaMemcpyAsync(p->vel, buf.data(), n * sizeof(float4), cudaMemcpyHostToDevice, p->st));
            HXM_CUDA(cudaStreamSynchronize(p->st));
            return 0;
        }
        return particle_float4_io(p, p->vel, true, set, false, host, 3);
    }
    if (name == "position_ds") {
        HXM_TRY(particle_ds_io(p, p->pos, set, static_cast<double*>(host)));
        if (set) { p->pos_ver++; p->reduced = false; }
        return 0;
    }
    if (name == "velocity_ds") return particle_ds_io(p, p->vel, set, static_cast<double*>(host));
    if (name == "image") {
        HXM_TRY(particle_float4_io(p, p->image, false, set, false, host, 3));
        return 0;
    }
    if (name == "force") {
        if (set) { set_error("particle: 'force' is read-only"); return HX_E_INVALID; }
        return particle_float4_io(p, p->force, false, false, false, host, 3);
    }
    if (name == "id" || name == "reverse_id") {
        uint32_t* dev = name == "id" ? p->id : p->rid;
        if (set) { set_error("particle: '%s' is read-only", name_c); return HX_E_INVALID; }
        HXM_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->st));
        HXM_CUDA(cudaStreamSynchronize(p->st));
        return 0;
    }
    if (name == "potential_energy") {
        if (set) { set_error("particle: 'potential_energy' is read-only"); return HX_E_INVALID; }
        HXM_CUDA(cudaMemcpyAsync(host, p->en_pot, n * sizeof(float), cudaMemcpyDeviceToHost, p->st));
        HXM_CUDA(cudaStreamSynchronize(p->st));
        return 0;
    }
    if (name == "potential_stress_tensor") {
        if (set) { set_error("particle: 'potential_stress_tensor' is read-only"); return HX_E_INVALID; }
        std::vector<float> buf(6 * A);
        HXM_CUDA(cudaMemcpyAsync(buf.data(), p->stress, 6 * A * sizeof(float), cudaMemcpyDeviceToHost, p->st));
        HXM_CUDA(cudaStreamSynchronize(p->st));
        float* h = static_cast<float*>(host);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 6; ++k) h[6 * i + k] = buf[k * A + i];     // SoA → row-major (force_kernel.hpp:60-95)
        return 0;
    }
    set_error("particle: unknown array '%s'", name_c);
    return HX_E_INVALID;
}

int hxm_particle_set(hxm_particle* p, const char* name, const void* host_data) { return particle_io(p, name, const_cast<void*>(host_data), true); }
int hxm_particle_get(hxm_particle* p, const char* name, void* host_data) { return particle_io(p, name, host_data, false); }

void* hxm_particle_device_ptr(hxm_particle* p, const char* name_c)
{
    std::string name(name_c ? name_c : "");
    if (name == "position") { p->reduced = false; return p->pos; }      // (the caller may write through the pointer)
    if (name == "velocity") return p->vel;
    if (name == "image") return p->image;
    if (name == "force") return p->force;
    if (name == "id") return p->id;
    if (name == "reverse_id") return p->rid;
    if (name == "potential_energy") return p->en_pot;
    if (name == "potential_stress_tensor") return p->stress;
    set_error("particle: unknown array '%s'", name_c ? name_c : "(null)");
    return nullptr;
}

int hxm_particle_aux_enable(hxm_particle* p)
{
    if (!p) return HX_E_INVALID;
    p->aux_enabled = true;
    return 0;
}

int hxm_potential_lj_create(unsigned ntype1, unsigned ntype2, const float* epsilon, const float* sigma, int truncation,
                            const float* cutoff_sigma, float h, hxm_potential** out)
{
    if (!out) return HX_E_INVALID;
    auto p = std::unique_ptr<hxm_potential>(new hxm_potential);
    HXM_TRY(hx_potential_lj_init(&p->p, ntype1, ntype2, epsilon, sigma, truncation, cutoff_sigma, h));
    *out = p.release();
    return 0;
}

int hxm_potential_lj_hard_core_create(unsigned ntype1, unsigned ntype2, const float* epsilon, const float* sigma,
                                      const float* core_sigma, int truncation, const float* cutoff_sigma, float h,
                                      hxm_potential** out)
{
    if (!out || !core_sigma) return HX_E_INVALID;
    auto p = std::unique_ptr<hxm_potential>(new hxm_potential);
    HXM_TRY(hx_potential_lj_init_hard_core(&p->p, ntype1, ntype2, epsilon, sigma, core_sigma, truncation, cutoff_sigma, h));
    *out = p.release();
    return 0;
}

void hxm_potential_destroy(hxm_potential* p) { delete p; }

int hxm_potential_r_cut(const hxm_potential* p, float* r_cut_out)
{
    if (!p || !r_cut_out) return HX_E_INVALID;
    for (unsigned i = 0; i < p->p.ntype1 * p->p.ntype2; ++i) r_cut_out[i] = p->p.r_cut[i];
    return 0;
}

const hx_lj_params* hxm_potential_params(const hxm_potential* p) { return p ? &p->p : nullptr; }

int hxm_binning_create(hxm_particle* particle, const double box_length[3], const float* r_cut, unsigned ntype1,
                       unsigned ntype2, double skin, double occupancy, hxm_binning** out)
{
    if (!particle || !box_length || !r_cut || !out || ntype1 * ntype2 == 0) { set_error("hxm_binning_create: invalid argument"); return HX_E_INVALID; }
    auto b = std::unique_ptr<hxm_binning>(new hxm_binning);
    b->particle = particle;
    float r_cut_max = *std::max_element(r_cut, r_cut + ntype1 * ntype2);
    if (occupancy <= 0) occupancy = 0.5;          // binning.cpp defaults::occupancy
    unsigned cs = 0;
    HXM_TRY(hx_binning_plan(box_length, particle->n, r_cut_max, static_cast<float>(skin), occupancy, b->ncell, b->cell_length, &cs));
    b->ncells = static_cast<size_t>(b->ncell[0]) * b->ncell[1] * b->ncell[2];
    HXM_TRY(dev_alloc(&b->cell_count, b->ncells));
    HXM_TRY(dev_alloc(&b->flag, 1));
    HXM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_flag), sizeof(int)));
    HXM_TRY(b->set_cell_size(cs));
    *out = b.release();
    return 0;
}

void hxm_binning_destroy(hxm_binning* b) { delete b; }

int hxm_binning_ncell(const hxm_binning* b, unsigned ncell[3])
{
    if (!b || !ncell) return HX_E_INVALID;
    for (int d = 0; d < 3; ++d) ncell[d] = b->ncell[d];
    return 0;
}

unsigned hxm_binning_cell_size(const hxm_binning* b) { return b ? b->cell_size : 0; }

int hxm_binning_update(hxm_binning* b, uint32_t* g_cell_host_out)
{
    if (!b) return HX_E_INVALID;
    if (b->cache_ver != b->particle->pos_ver) HXM_TRY(b->update());      // binning::g_cell (binning.cpp:152-160)
    if (g_cell_host_out) {
        HXM_CUDA(cudaMemcpyAsync(g_cell_host_out, b->g_cell, b->ncells * b->cell_size * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, b->particle->st));
        HXM_CUDA(cudaStreamSynchronize(b->particle->st));
    }
    return 0;
}

int hxm_neighbour_create(hxm_particle* particle, const double box_length[3], const float* r_cut, unsigned ntype1,
                         unsigned ntype2, double skin, double occupancy, int layout, int disable_sorting,
                         hxm_binning* binning, hxm_neighbour** out)
{
    if (!particle || !box_length || !r_cut || !out || ntype1 * ntype2 == 0 || ntype1 * ntype2 > HX_MAX_PAIR_TYPES
        || (layout != HX_LAYOUT_COLUMN && layout != HX_LAYOUT_ROW)) {
        set_error("hxm_neighbour_create: invalid argument");
        return HX_E_INVALID;
    }
    auto nb = std::unique_ptr<hxm_neighbour>(new hxm_neighbour);
    nb->particle = particle;
    nb->ntype1 = ntype1;
    nb->ntype2 = ntype2;
    nb->layout = layout;
    nb->sorting = !disable_sorting;
    nb->r_skin = static_cast<float>(skin);
    for (int d = 0; d < 3; ++d) { nb->L[d] = box_length[d]; nb->Lf[d] = static_cast<float>(box_length[d]); }
    nb->r_cut_max = *std::max_element(r_cut, r_cut + ntype1 * ntype2);
    for (unsigned i = 0; i < ntype1 * ntype2; ++i) {
        // rr_cut_skin_(i, j) = std::pow(r_cut(i, j) + r_skin_, 2): float sum, double pow, float store (from_binning.cpp:78-82)
        nb->rr_cut_skin[i] = static_cast<float>(std::pow(r_cut[i] + nb->r_skin, 2));
    }
    if (binning) {
        nb->binning = binning;
    } else {
        // neighbour.lua.in:201-206 passes the (possibly nil) occupancy argument on to mdsim.binning
        hxm_binning* b = nullptr;
        HXM_TRY(hxm_binning_create(particle, box_length, r_cut, ntype1, ntype2, skin, occupancy, &b));
        nb->binning = b;
        nb->own_binning = true;
    }
    if (nb->binning->ncell[0] < 3 || nb->binning->ncell[1] < 3 || nb->binning->ncell[2] < 3) {
        set_error("neighbour: number of cells per dimension must be at least 3 (from_particle fallback is out of scope)");
        return HX_E_UNSUPPORTED;
    }
    size_t n = particle->n;
    HXM_TRY(dev_alloc(&nb->count, n));
    HXM_TRY(dev_alloc(&nb->flag, 1));
    HXM_TRY(dev_alloc(&nb->r0, n));
    HXM_TRY(dev_alloc(&nb->ctl, HX_CTL_WORDS));
    HXM_TRY(dev_alloc(&nb->keys, particle->A));
    HXM_TRY(dev_alloc(&nb->index, particle->A));
    nb->sort_scratch_bytes = hx_sort_scratch_bytes(n);
    HXM_CUDA(cudaMalloc(&nb->sort_scratch, nb->sort_scratch_bytes));
    HXM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&nb->h_flag), 4 * sizeof(int)));
    HXM_TRY(dev_alloc(&nb->count_max, 1));
    HXM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&nb->h_ctl), HX_CTL_WORDS * sizeof(uint32_t)));
    HXM_CUDA(cudaMemsetAsync(nb->ctl, 0, HX_CTL_WORDS * sizeof(uint32_t), particle->st));
    nb->depth = hx_hilbert_depth(box_length);
    if (occupancy <= 0) occupancy = 0.4;          // from_binning.cpp:284-287
    HXM_TRY(nb->set_occupancy(occupancy));
    *out = nb.release();
    return 0;
}

void hxm_neighbour_destroy(hxm_neighbour* nb) { delete nb; }
unsigned hxm_neighbour_size(const hxm_neighbour* nb) { return nb->size; }
size_t hxm_neighbour_stride(const hxm_neighbour* nb) { return nb->stride; }
int hxm_neighbour_layout(const hxm_neighbour* nb) { return nb->layout; }
size_t hxm_neighbour_updates(const hxm_neighbour* nb) { return nb->updates; }
const void* hxm_neighbour_device_ptr(const hxm_neighbour* nb) { return nb->g_neighbour; }

const void* hxm_neighbour_named_ptr(const hxm_neighbour* nb, const char* name_c)
{
    std::string name(name_c ? name_c : "");
    if (!nb) return nullptr;
    if (name == "g_neighbour") return nb->g_neighbour;
    if (name == "count") return nb->count;
    if (name == "r0") return nb->r0;
    if (name == "ctl") return nb->ctl;
    if (name == "shell_rows") return nb->shells_valid ? &nb->shells : nullptr;      // host pointer to the descriptor
    set_error("neighbour: unknown array '%s'", name_c ? name_c : "(null)");
    return nullptr;
}

int hxm_neighbour_ctl_read(hxm_neighbour* nb, uint32_t ctl_out[4])
{
    if (!nb || !ctl_out) return HX_E_INVALID;
    HXM_CUDA(cudaMemcpyAsync(nb->h_ctl, nb->ctl, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, nb->particle->st));
    HXM_CUDA(cudaStreamSynchronize(nb->particle->st));
    for (int k = 0; k < 4; ++k) ctl_out[k] = nb->h_ctl[k];
    return 0;
}

int hxm_neighbour_set_shell_rows(hxm_neighbour* nb, int enable)
{
    if (!nb) return HX_E_INVALID;
    nb->use_shells = enable != 0;
    if (!nb->use_shells) nb->shells_valid = false;
    else if (nb->updates > 0) nb->order_ver_seen = 0;      // rebuilt (with the private rows) at the next check
    return 0;
}

int hxm_neighbour_check(hxm_neighbour* nb, int* rebuilt)
{
    if (!nb) return HX_E_INVALID;
    bool r = false;
    HXM_TRY(nb->check(&r));
    if (rebuilt) *rebuilt = r ? 1 : 0;
    return 0;
}

double hxm_neighbour_mean_length(hxm_neighbour* nb)
{
    size_t n = nb->particle->n;
    std::vector<uint32_t> c(n);
    if (cudaMemcpyAsync(c.data(), nb->count, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, nb->particle->st) != cudaSuccess) return -1;
    cudaStreamSynchronize(nb->particle->st);
    double s = 0;
    for (uint32_t v : c) s += v;
    return s / static_cast<double>(n);
}

int hxm_neighbour_get(hxm_neighbour* nb, uint32_t* list_host, uint32_t* count_host)
{
    if (!nb) return HX_E_INVALID;
    hxm_particle* p = nb->particle;
    if (list_host) {
        size_t words = (nb->layout == HX_LAYOUT_ROW) ? p->n * nb->stride : static_cast<size_t>(nb->size) * nb->stride;
        HXM_CUDA(cudaMemcpyAsync(list_host, nb->g_neighbour, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->st));
    }
    if (count_host) HXM_CUDA(cudaMemcpyAsync(count_host, nb->count, p->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->st));
    HXM_CUDA(cudaStreamSynchronize(p->st));
    return 0;
}

int hxm_force_pair_trunc_create(hxm_potential* potential, hxm_particle* particle, const double box_length[3],
                                hxm_neighbour* neighbour, float aux_weight, hxm_force** out)
{
    if (!potential || !particle || !box_length || !neighbour || !out || neighbour->particle != particle) {
        set_error("hxm_force_pair_trunc_create: invalid argument");
        return HX_E_INVALID;
    }
    if (potential->p.ntype1 < particle->nspecies || potential->p.ntype2 < particle->nspecies) {
        // pair_trunc.hpp:157-163
        set_error("size of potential coefficients less than number of particle species");
        return HX_E_INVALID;
    }
    auto f = std::unique_ptr<hxm_force>(new hxm_force);
    f->potential = potential;
    f->particle = particle;
    f->neighbour = neighbour;
    f->aux_weight = aux_weight;
    for (int d = 0; d < 3; ++d) f->Lf[d] = static_cast<float>(box_length[d]);
    // the first pair force bound to the lists decides the cutoffs of their private shell rows
    if (!neighbour->shell_potential && !potential->p.hard_core) neighbour->shell_potential = &potential->p;
    *out = f.release();
    return 0;
}

void hxm_force_destroy(hxm_force* f) { delete f; }

int hxm_force_add_planar_wall(hxm_force* f, const hx_planar_wall* wall)
{
    if (!f || !wall || wall->nwall == 0 || wall->nwall > HX_MAX_WALLS || wall->nspecies < f->particle->nspecies
        || wall->nspecies > HX_MAX_WALL_SPECIES) {
        set_error("forces.external: invalid planar_wall potential (nspecies %u, particle has %u)", wall ? wall->nspecies : 0,
                  f ? f->particle->nspecies : 0);
        return HX_E_INVALID;
    }
    f->walls.push_back(*wall);
    f->particle->force_pos_ver = 0;      // force cache is stale
    f->particle->aux_valid = false;
    return 0;
}

int hxm_force_apply(hxm_force* f)
{
    if (!f) return HX_E_INVALID;
    return f->ensure(f->particle->aux_enabled);
}

static int integrator_create(hxm_integrator::kind_t kind, hxm_particle* particle, hxm_force* force,
                             const double box_length[3], double timestep, hxm_integrator** out)
{
    if (!particle || !force || !box_length || !out || force->particle != particle || !(timestep > 0)) {
        set_error("hxm_integrator_create: invalid argument");
        return HX_E_INVALID;
    }
    auto i = std::unique_ptr<hxm_integrator>(new hxm_integrator);
    i->kind = kind;
    i->particle = particle;
    i->force = force;
    i->timestep = timestep;
    for (int d = 0; d < 3; ++d) i->Lf[d] = static_cast<float>(box_length[d]);
    *out = i.release();
    return 0;
}

int hxm_integrator_verlet_create(hxm_particle* particle, hxm_force* force, const double box_length[3], double timestep,
                                 hxm_integrator** out)
{
    return integrator_create(hxm_integrator::VERLET, particle, force, box_length, timestep, out);
}

int hxm_integrator_andersen_create(hxm_particle* particle, hxm_force* force, const double box_length[3], double timestep,
                                   double temperature, double rate, unsigned seed, hxm_integrator** out)
{
    hxm_integrator* i = nullptr;
    HXM_TRY(integrator_create(hxm_integrator::ANDERSEN, particle, force, box_length, timestep, &i));
    std::unique_ptr<hxm_integrator> guard(i);
    i->temperature = temperature;
    i->rate = rate;
    i->update_andersen();
    // random/gpu/random.cpp:94-101 — 32 blocks x 256 threads
    HXM_TRY(dev_alloc(&i->rng_state, 3 * 32 * 256));
    HXM_TRY(dev_alloc(&i->andersen_scratch, hx_andersen_scratch_bytes(particle->n)));
    HXM_TRY(hx_rand48_seed(&i->rng, 32, 256, seed, i->rng_state, particle->st));
    *out = guard.release();
    return 0;
}

void hxm_integrator_destroy(hxm_integrator* i) { delete i; }
int hxm_integrator_integrate(hxm_integrator* i) { return i ? i->integrate() : HX_E_INVALID; }
int hxm_integrator_finalize(hxm_integrator* i) { return i ? i->finalize() : HX_E_INVALID; }

int hxm_integrator_set_timestep(hxm_integrator* i, double timestep)
{
    if (!i || !(timestep > 0)) return HX_E_INVALID;
    i->timestep = timestep;
    i->update_andersen();
    return 0;
}

int hxm_integrator_set_temperature(hxm_integrator* i, double temperature)
{
    if (!i || !(temperature >= 0)) return HX_E_INVALID;
    i->temperature = temperature;
    if (i->kind == hxm_integrator::HOOVER) return hx_hoover_chain_set_temperature(&i->chain, temperature, i->resonance_frequency);
    i->update_andersen();
    return 0;
}

int hxm_integrator_hoover_create(hxm_particle* particle, hxm_force* force, const double box_length[3], double timestep,
                                 double temperature, double resonance_frequency, hxm_integrator** out)
{
    if (!(temperature > 0) || !(resonance_frequency > 0)) { set_error("integrators.verlet_nvt_hoover: invalid argument"); return HX_E_INVALID; }
    hxm_integrator* i = nullptr;
    HXM_TRY(integrator_create(hxm_integrator::HOOVER, particle, force, box_length, timestep, &i));
    std::unique_ptr<hxm_integrator> guard(i);
    i->temperature = temperature;
    i->resonance_frequency = resonance_frequency;
    HXM_TRY(hx_hoover_chain_init(&i->chain, particle->n, 3, temperature, resonance_frequency));
    HXM_TRY(dev_alloc(&i->d_hoover, HX_THERMO_DOUBLES));
    HXM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&i->h_hoover), 8 * sizeof(double)));
    *out = guard.release();
    return 0;
}

int hxm_integrator_hoover_set_mass(hxm_integrator* i, const double mass[2])
{
    if (!i || !mass || i->kind != hxm_integrator::HOOVER || !(mass[0] > 0) || !(mass[1] > 0)) return HX_E_INVALID;
    i->chain.mass_xi[0] = mass[0];
    i->chain.mass_xi[1] = mass[1];
    return 0;
}

/* out = (xi[0], xi[1], v_xi[0], v_xi[1], mass_xi[0], mass_xi[1], en_nhc per particle) */
int hxm_integrator_hoover_state(const hxm_integrator* i, double out[7])
{
    if (!i || !out || i->kind != hxm_integrator::HOOVER) return HX_E_INVALID;
    out[0] = i->chain.xi[0]; out[1] = i->chain.xi[1]; out[2] = i->chain.v_xi[0]; out[3] = i->chain.v_xi[1];
    out[4] = i->chain.mass_xi[0]; out[5] = i->chain.mass_xi[1]; out[6] = i->en_nhc;
    return 0;
}

int hxm_core_create(hxm_integrator* integrator, hxm_core** out)
{
    if (!integrator || !out) return HX_E_INVALID;
    auto c = std::unique_ptr<hxm_core>(new hxm_core);
    c->integrator = integrator;
    HXM_TRY(dev_alloc(&c->d_thermo, HX_THERMO_DOUBLES));
    HXM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_thermo), 8 * sizeof(double)));
    *out = c.release();
    return 0;
}

void hxm_core_destroy(hxm_core* c) { delete c; }
int hxm_core_mdstep(hxm_core* c) { return c ? c->mdstep() : HX_E_INVALID; }
int hxm_core_run(hxm_core* c, uint64_t steps) { return c ? c->run(steps) : HX_E_INVALID; }
uint64_t hxm_core_step(const hxm_core* c) { return c ? c->step : 0; }

int hxm_core_set_chunk(hxm_core* c, unsigned min_chunk, unsigned max_chunk)
{
    if (!c || min_chunk == 0 || max_chunk < min_chunk) return HX_E_INVALID;
    c->min_chunk = min_chunk;
    c->max_chunk = max_chunk;
    return 0;
}

int hxm_core_set_fused_verlet(hxm_core* c, int enable)
{
    if (!c) return HX_E_INVALID;
    hxm_particle* p = c->integrator->particle;
    if (enable && !p->pos_c) {
        // second position buffer of the chain, allocated on first use; the padding entries behind nparticle are never
        // written by the kernels, so both buffers hold the same values there
        HXM_TRY(dev_alloc(&p->pos_c, 2 * p->A));
        HXM_CUDA(cudaMemcpyAsync(p->pos_c, p->pos, 2 * p->A * sizeof(float4), cudaMemcpyDeviceToDevice, p->st));
    }
    c->fused_verlet = enable != 0;
    return 0;
}

int hxm_core_stats(const hxm_core* c, uint64_t* launches, uint64_t* chunks, uint64_t* syncs)
{
    if (!c) return HX_E_INVALID;
    if (launches) *launches = c->stat_launches;
    if (chunks) *chunks = c->stat_chunks;
    if (syncs) *syncs = c->stat_syncs;
    return 0;
}

int hxm_core_thermodynamics(hxm_core* c, double out[10]) { return (c && out) ? c->thermodynamics(out) : HX_E_INVALID; }

// observables.density_mode:acquire (halmd/observables/gpu/density_mode.cpp:70-131) for the group of all particles
int hxm_density_modes(hxm_particle* p, const float* wavevector_host, unsigned nq, double* rho_host)
{
    if (!p || !wavevector_host || !rho_host || nq == 0) { set_error("density_mode: invalid argument"); return HX_E_INVALID; }
    void* scratch = nullptr;
    size_t const rho_bytes = sizeof(double) * 2 * nq;
    HXM_TRY(p->get_scratch(rho_bytes + sizeof(float) * 3 * nq, &scratch));
    double* rho = static_cast<double*>(scratch);
    float* q = reinterpret_cast<float*>(static_cast<char*>(scratch) + rho_bytes);
    HXM_CUDA(cudaMemcpyAsync(q, wavevector_host, sizeof(float) * 3 * nq, cudaMemcpyHostToDevice, p->st));
    HXM_TRY(hx_density_modes(reinterpret_cast<const hx_float4*>(p->pos), p->n, q, nq, rho, p->st));
    HXM_CUDA(cudaMemcpyAsync(rho_host, rho, rho_bytes, cudaMemcpyDeviceToHost, p->st));
    HXM_CUDA(cudaStreamSynchronize(p->st));
    return 0;
}

int hxm_blocking_create(hxm_core* core, double max_lag_time, double resolution, unsigned block_size, unsigned shift,
                        unsigned separation, const float* wavevectors, unsigned nq, const unsigned* shell_offsets,
                        unsigned nshell, hxm_blocking** out)
{
    // blocking_scheme.cpp:40-125
    if (!core || !out) return HX_E_INVALID;
    if (block_size < 2) { set_error("Minimum size of coarse-graining blocks is 2."); return HX_E_INVALID; }
    if (shift == 0) { set_error("Coarse-graining shift must be non-zero."); return HX_E_INVALID; }
    if (shift >= block_size) { set_error("Coarse-graining shift must be less than block size."); return HX_E_INVALID; }
    if (!(resolution > 0)) { set_error("Resolution must be a positive value."); return HX_E_INVALID; }
    if (nq > 0 && (!wavevectors || !shell_offsets || nshell == 0)) { set_error("hxm_blocking_create: invalid wavevector shells"); return HX_E_INVALID; }
    auto b = std::unique_ptr<hxm_blocking>(new hxm_blocking);
    b->core = core;
    b->block_size = block_size;
    b->separation = separation;
    double const timestep = core->integrator->timestep;
    b->timestep = timestep;
    uint64_t max_interval = static_cast<uint64_t>(std::round(max_lag_time / timestep));
    uint64_t s = static_cast<uint64_t>(std::round(resolution / timestep));
    if (s == 0) { set_error("Resolution must be at least one time step."); return HX_E_INVALID; }
    while (s <= max_interval) {
        b->interval.push_back(s);
        if (shift > 1) {
            uint64_t s_odd = s * shift;
            if (s_odd <= max_interval) b->interval.push_back(s_odd);
        }
        s *= block_size;
    }
    if (b->interval.empty()) { set_error("Sampling interval of dynamic correlations is too large."); return HX_E_INVALID; }
    size_t levels = b->interval.size();
    b->origin.assign(levels, core->step);
    b->blocks.resize(levels);
    b->msd.assign(levels * block_size, welford());
    b->mqd.assign(levels * block_size, welford());
    b->vacf.assign(levels * block_size, welford());
    hxm_particle* p = core->integrator->particle;
    b->norm = static_cast<double>(p->n);
    b->nq = nq;
    b->nshell = nq > 0 ? nshell : 0;
    if (nq > 0) {
        b->shell_offsets.assign(shell_offsets, shell_offsets + nshell + 1);
        if (b->shell_offsets.back() != nq) { set_error("hxm_blocking_create: shell offsets do not cover the wavevectors"); return HX_E_INVALID; }
        b->isf.assign(levels * block_size * nshell, welford());
        HXM_TRY(dev_alloc(&b->d_q, 3 * nq));
        HXM_TRY(dev_alloc(&b->d_rho, 2 * nq));
        HXM_CUDA(cudaMemcpy(b->d_q, wavevectors, 3 * nq * sizeof(float), cudaMemcpyHostToDevice));
    }
    HXM_TRY(dev_alloc(&b->d_msd, 3 * block_size));
    HXM_TRY(dev_alloc(&b->d_ptrs, block_size));
    b->msd_scratch_bytes = hx_msd_scratch_bytes(block_size);
    HXM_CUDA(cudaMalloc(&b->d_msd_scratch, b->msd_scratch_bytes));
    *out = b.release();
    return 0;
}

void hxm_blocking_destroy(hxm_blocking* b) { delete b; }
unsigned hxm_blocking_levels(const hxm_blocking* b) { return b ? static_cast<unsigned>(b->interval.size()) : 0; }
unsigned hxm_blocking_block_size(const hxm_blocking* b) { return b ? b->block_size : 0; }

int hxm_blocking_intervals(const hxm_blocking* b, uint64_t* intervals_out)
{
    if (!b || !intervals_out) return HX_E_INVALID;
    std::copy(b->interval.begin(), b->interval.end(), intervals_out);
    return 0;
}

int hxm_blocking_sample(hxm_blocking* b) { return b ? b->do_sample() : HX_E_INVALID; }

int hxm_blocking_finalise(hxm_blocking* b)
{
    // blocking_scheme::finalise (blocking_scheme.cpp:163-174)
    if (!b) return HX_E_INVALID;
    for (unsigned i = 0; i < b->interval.size(); ++i)
        while (!b->blocks[i].empty()) HXM_TRY(b->process(i));
    return 0;
}

int hxm_blocking_msd(hxm_blocking* b, double* out)
{
    if (!b || !out) return HX_E_INVALID;
    for (size_t i = 0; i < b->msd.size(); ++i) { out[3 * i] = b->msd[i].n; out[3 * i + 1] = b->msd[i].m; out[3 * i + 2] = b->msd[i].v; }
    return 0;
}

int hxm_blocking_mqd(hxm_blocking* b, double* out)
{
    if (!b || !out) return HX_E_INVALID;
    for (size_t i = 0; i < b->mqd.size(); ++i) { out[3 * i] = b->mqd[i].n; out[3 * i + 1] = b->mqd[i].m; out[3 * i + 2] = b->mqd[i].v; }
    return 0;
}

int hxm_blocking_vacf(hxm_blocking* b, double* out)
{
    if (!b || !out) return HX_E_INVALID;
    for (size_t i = 0; i < b->vacf.size(); ++i) { out[3 * i] = b->vacf[i].n; out[3 * i + 1] = b->vacf[i].m; out[3 * i + 2] = b->vacf[i].v; }
    return 0;
}

int hxm_blocking_isf(hxm_blocking* b, double* out)
{
    if (!b || !out) return HX_E_INVALID;
    for (size_t i = 0; i < b->isf.size(); ++i) { out[3 * i] = b->isf[i].n; out[3 * i + 1] = b->isf[i].m; out[3 * i + 2] = b->isf[i].v; }
    return 0;
}

int hxm_blocking_trace(hxm_blocking* b, int enable, uint64_t* events_out, size_t capacity, size_t* nevents_out)
{
    if (!b) return HX_E_INVALID;
    b->trace_on = enable != 0;
    size_t const n = b->trace.size() / 3;
    if (nevents_out) *nevents_out = n;
    if (events_out) {
        if (capacity < n) { set_error("hxm_blocking_trace: capacity %zu < %zu events", capacity, n); return HX_E_INVALID; }
        std::copy(b->trace.begin(), b->trace.end(), events_out);
    }
    if (!enable) b->trace.clear();
    return 0;
}

int hxm_core_run_sampling(hxm_core* c, uint64_t steps, hxm_blocking* b)
{
    if (!c || !b || b->core != c) return HX_E_INVALID;
    uint64_t const end = c->step + steps;
    // sampler::run: sample the initial state, then after every step that is due (sampler.cpp:63-80)
    if (b->due(c->step)) HXM_TRY(b->do_sample());
    while (c->step < end) {
        uint64_t next = end;
        for (unsigned i = 0; i < b->interval.size(); ++i) {
            uint64_t o = b->origin[i], iv = b->interval[i];
            uint64_t due = (c->step < o) ? o : o + ((c->step - o) / iv + 1) * iv;
            next = std::min(next, due);
        }
        HXM_TRY(c->run(next - c->step));
        if (b->due(c->step)) HXM_TRY(b->do_sample());
    }
    return 0;
}

} // extern "C"
