// refsim.cpp — the reference's OWN GPU step loop, replayed on device-resident arrays (oracle/_ref/libhalmd_ref_gpu.so).
// TEST INFRASTRUCTURE ONLY: used by tests/ (energy drift of the reference's kernels) and by bench.py's `gpu_reference`
// leg (the recompiled-CUDA baseline on the same B200); never linked into or called by the product.
//
// One MD step is what core::mdstep (halmd/mdsim/core.cpp:30-63) triggers through the host classes, with the
// reference's launch configurations and its cuda::thread::synchronize() after every launch:
//   verlet::integrate                      gpu/integrators/verlet.cpp:64-90
//   from_binning::g_neighbour (cache test) gpu/neighbours/from_binning.cpp:123-140
//     max_displacement::compute            gpu/max_displacement.cpp:77-106 (kernel + blocking copy of 64 floats)
//     on rebuild: sorts::hilbert::order    gpu/sorts/hilbert.cpp:71-117  (map, gen_index, radix_sort, particle::rearrange
//                                          gpu/particle.cpp:173-218 incl. the reverse-id sort)
//                 binning::update          gpu/binning.cpp:166-221
//                 from_binning::update     gpu/neighbours/from_binning.cpp:162-282
//                 max_displacement::zero   gpu/max_displacement.cpp:60-72
//   pair_trunc::apply -> compute_          gpu/forces/pair_trunc.hpp:218-272 (default column lists)
//   verlet::finalize                       gpu/integrators/verlet.cpp:95-121
// The sorts are the reference's own radix sort, host orchestration included (halmd/algorithm/gpu/radix_sort.hpp,
// scan.hpp), compiled in place; ref_compat/halmd/utility/gpu/device.hpp stands in for the Lua-bound device module.
#include <halmd/config.hpp>
#include <halmd/algorithm/gpu/iota.hpp>
#include <halmd/algorithm/gpu/radix_sort.hpp>
#include <halmd/mdsim/gpu/binning_kernel.hpp>
#include <halmd/mdsim/gpu/forces/pair_trunc_kernel.hpp>
#include <halmd/mdsim/gpu/integrators/verlet_kernel.hpp>
#include <halmd/mdsim/gpu/max_displacement_kernel.hpp>
#include <halmd/mdsim/gpu/neighbours/from_binning_kernel.hpp>
#include <halmd/mdsim/gpu/particle_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/lennard_jones_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/truncations/smooth_r4_kernel.hpp>
#include <halmd/mdsim/gpu/sorts/hilbert_kernel.hpp>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#include "refgpu_util.hpp"

namespace {

typedef cuda::memory::device::vector<unsigned int> uvector;
typedef cuda::memory::device::vector<float4> f4vector;
namespace lj = potentials::pair::lennard_jones_kernel;
namespace tr = potentials::pair::truncations;
typedef tr::smooth_r4_kernel::smooth_r4<lj::lennard_jones> potential_type;

struct ref_sim {
    size_t n, A;
    float L[3], cell_length[3], dt, skin, rr_cut_skin, rri_smooth;
    unsigned ncell[3], cell_size, size, depth;
    size_t ncells;
    // particle arrays (double-single: high words [0, A), low words [A, 2A)), swapped with the back buffers by rearrange
    std::unique_ptr<f4vector> pos, vel, img, pos_b, vel_b, img_b, force, r0;
    std::unique_ptr<uvector> id, id_b, rid, g_cell, cell_index, cell_perm, cell_offset, nl;
    cuda::memory::device::vector<float> rr_block, d_rrcs;
    cuda::memory::device::vector<int> d_ret;
    cuda::memory::device::vector<float2> d_lj;
    cuda::memory::device::vector<float4> d_sm;
    ref_sim() : rr_block(64), d_rrcs(1), d_ret(1), d_lj(1), d_sm(1) {}
    std::unique_ptr<lintex<float>> t_rrcs;
    std::unique_ptr<lintex<float2>> t_lj;
    std::unique_ptr<lintex<float4>> t_sm;
    size_t steps = 0, rebuilds = 0;
    bool lists_valid = false;
};

template <typename T>
T* raw(cuda::memory::device::vector<T>& v) { return &*v.begin(); }

fixed_vector<float, 3> box(ref_sim const& s) { return vec3<fixed_vector<float, 3>>(s.L); }

// sorts::hilbert::order + particle::rearrange
int order(ref_sim& s)
{
    typedef particle_wrapper<3, dsfloat> pwrapper;
    typedef pwrapper::ptr_type ptr_type;
    dim3 const grid(static_cast<unsigned>(s.A / 128)), block(128);
    uvector g_index(s.n), g_map(s.n);
    g_index.reserve(s.A);
    g_map.reserve(s.A);
    auto& hk = sorts::hilbert_wrapper<3>::kernel;
    RG_TRY(
        hk.depth.set(s.depth);
        hk.map.configure(grid, block);
        hk.map(raw(*s.pos), raw(g_map), box(s));
        cuda::thread::synchronize();
        hk.gen_index.configure(grid, block);
        hk.gen_index(raw(g_index));
        cuda::thread::synchronize();
        halmd::radix_sort(g_map.begin(), g_map.end(), g_index.begin());
    )
    {
        lintex<float4> t_r(raw(*s.pos), 2 * s.A), t_v(raw(*s.vel), 2 * s.A), t_img(raw(*s.img), s.A);
        lintex<unsigned int> t_id(raw(*s.id), s.A);
        ptr_type pos = {raw(*s.pos_b), raw(*s.pos_b) + s.A}, vel = {raw(*s.vel_b), raw(*s.vel_b) + s.A};
        RG_TRY(
            pwrapper::kernel.rearrange.configure(grid, block);
            pwrapper::kernel.rearrange(t_r.t, t_img.t, t_v.t, t_id.t, raw(g_index), pos, raw(*s.img_b), vel, raw(*s.id_b),
                                       static_cast<unsigned>(s.n));
            cuda::thread::synchronize();
        )
    }
    std::swap(s.pos, s.pos_b);
    std::swap(s.vel, s.vel_b);
    std::swap(s.img, s.img_b);
    RG_CHECK(cudaMemcpy(raw(*s.id), raw(*s.id_b), s.A * sizeof(unsigned int), cudaMemcpyDeviceToDevice));
    RG_TRY(
        halmd::iota(s.rid->begin(), s.rid->begin() + s.A, 0);
        halmd::radix_sort(s.id_b->begin(), s.id_b->begin() + s.n, s.rid->begin());
    )
    return 0;
}

// binning::update + from_binning::update + max_displacement::zero
int rebuild(ref_sim& s)
{
    int rc;
    if ((rc = order(s))) return rc;
    dim3 const grid(static_cast<unsigned>(s.A / 128)), block(128);
    unsigned const nbox = static_cast<unsigned>(s.n);
    fixed_vector<unsigned int, 3> nc;
    for (int d = 0; d < 3; ++d) nc[d] = s.ncell[d];
    auto& bk = binning_wrapper<3>::kernel;
    int h_ret = 0;
    RG_TRY(
        bk.compute_cell.configure(grid, block);
        bk.compute_cell(raw(*s.pos), raw(*s.cell_index), vec3<fixed_vector<float, 3>>(s.cell_length), nc);
        bk.gen_index.configure(grid, block);
        bk.gen_index(raw(*s.cell_perm), nbox);
        halmd::radix_sort(s.cell_index->begin(), s.cell_index->end(), s.cell_perm->begin());
        cuda::memset(s.cell_offset->begin(), s.cell_offset->end(), 0xFF);
        bk.find_cell_offset.configure(grid, block);
        bk.find_cell_offset(raw(*s.cell_index), raw(*s.cell_offset), nbox);
        cuda::memset(s.d_ret.begin(), s.d_ret.end(), 0);
        bk.assign_cells.configure(dim3(static_cast<unsigned>(s.ncells)), dim3(s.cell_size));
        bk.assign_cells(raw(s.d_ret), raw(*s.cell_index), raw(*s.cell_offset), raw(*s.cell_perm), raw(*s.g_cell), nbox, s.cell_size);
        cuda::thread::synchronize();
    )
    RG_CHECK(cudaMemcpy(&h_ret, raw(s.d_ret), sizeof(int), cudaMemcpyDeviceToHost));
    if (h_ret) { fprintf(stderr, "refsim: overcrowded cells\n"); return -2; }
    {
        lintex<float4> t_r(raw(*s.pos), s.A);
        auto& nk = neighbours::from_binning_wrapper<3>::kernel.normal;
        RG_TRY(
            cuda::memset(s.nl->begin(), s.nl->end(), 0xFF);
            cuda::memset(s.d_ret.begin(), s.d_ret.end(), 0);
            nk.update_neighbours.configure(dim3(static_cast<unsigned>(s.ncells)), dim3(s.cell_size), s.cell_size * (2 + 3) * sizeof(int));
            nk.update_neighbours(s.t_rrcs->t, t_r.t, t_r.t, raw(s.d_ret), raw(*s.nl), s.size, static_cast<unsigned>(s.A), raw(*s.g_cell),
                                 raw(*s.g_cell), 1u, 1u, static_cast<unsigned>(s.ncells * s.cell_size), nc, box(s));
            cuda::thread::synchronize();
        )
    }
    RG_CHECK(cudaMemcpy(&h_ret, raw(s.d_ret), sizeof(int), cudaMemcpyDeviceToHost));
    if (h_ret) { fprintf(stderr, "refsim: overcrowded neighbour lists\n"); return -3; }
    RG_CHECK(cudaMemcpy(raw(*s.r0), raw(*s.pos), s.n * sizeof(float4), cudaMemcpyDeviceToDevice));
    s.lists_valid = true;
    s.rebuilds++;
    return 0;
}

int check_lists(ref_sim& s)
{
    if (s.lists_valid) {
        auto& dk = max_displacement_wrapper<3>::kernel;
        std::vector<float> h(64);
        RG_TRY(
            dk.displacement.configure(dim3(64), dim3(128));
            dk.displacement(raw(*s.pos), raw(*s.r0), raw(s.rr_block), static_cast<unsigned>(s.n), box(s));
            cuda::thread::synchronize();
        )
        RG_CHECK(cudaMemcpy(h.data(), raw(s.rr_block), 64 * sizeof(float), cudaMemcpyDeviceToHost));
        float const d = std::sqrt(*std::max_element(h.begin(), h.end()));
        if (!(d > s.skin / 2)) return 0;
    }
    return rebuild(s);
}

int compute_force(ref_sim& s)
{
    typedef forces::pair_trunc_wrapper<3, potential_type> wrapper_type;
    lintex<float4> t_r2(raw(*s.pos), s.A);
    lj::lennard_jones const plain(s.t_lj->t);
    RG_TRY(
        tr::smooth_r4_wrapper<lj::lennard_jones>::rri_smooth.set(s.rri_smooth);
        potential_type pot(plain, s.t_sm->t);
        auto& f = wrapper_type::kernel.compute;
        f.configure(dim3(static_cast<unsigned>(s.A / 128)), dim3(128));
        f(pot, raw(*s.pos), t_r2.t, raw(*s.force), raw(*s.nl), s.size, static_cast<unsigned>(s.A), nullptr, nullptr, 1u, 1u, box(s),
          true, 0.5f);
        cuda::thread::synchronize();
    )
    return 0;
}

int mdstep(ref_sim& s)
{
    typedef integrators::verlet_wrapper<3, dsfloat> vwrapper;
    typedef vwrapper::ptr_type ptr_type;
    dim3 const grid(static_cast<unsigned>(s.A / 128)), block(128);
    int rc;
    {
        ptr_type pos = {raw(*s.pos), raw(*s.pos) + s.A}, vel = {raw(*s.vel), raw(*s.vel) + s.A};
        RG_TRY(
            vwrapper::wrapper.integrate.configure(grid, block);
            vwrapper::wrapper.integrate(pos, raw(*s.img), vel, raw(*s.force), s.dt, box(s));
            cuda::thread::synchronize();
        )
    }
    if ((rc = check_lists(s))) return rc;
    if ((rc = compute_force(s))) return rc;
    {
        ptr_type vel = {raw(*s.vel), raw(*s.vel) + s.A};
        RG_TRY(
            vwrapper::wrapper.finalize.configure(grid, block);
            vwrapper::wrapper.finalize(vel, raw(*s.force), s.dt);
            cuda::thread::synchronize();
        )
    }
    s.steps++;
    return 0;
}

} // namespace

extern "C" {

// position8 / velocity8: float4[2 A] double-single arrays in the particle packing (species tag / mass in w); one
// species, Lennard-Jones (epsilon, sigma2) with smooth_r4 (r_cut, rr_cut, en_cut, h^-2); cell grid, cell size and
// list size as planned by the host classes (binning.cpp:81-101, from_binning.cpp:96-119) — passed in by the caller
int refsim_create(size_t nparticle, size_t array_size, const float* position8, const float* velocity8, const float box_length[3],
                  float epsilon, float sigma2, float r_cut, float rr_cut, float en_cut, float rri_smooth, float skin, float timestep,
                  const unsigned ncell[3], const float cell_length[3], unsigned cell_size, unsigned size, unsigned depth, void** out)
{
    if (!out || array_size % 128 != 0 || nparticle > array_size) return static_cast<int>(cudaErrorInvalidValue);
    try {
        std::unique_ptr<ref_sim> s(new ref_sim);
        s->n = nparticle;
        s->A = array_size;
        for (int d = 0; d < 3; ++d) { s->L[d] = box_length[d]; s->ncell[d] = ncell[d]; s->cell_length[d] = cell_length[d]; }
        s->dt = timestep;
        s->skin = skin;
        s->rri_smooth = rri_smooth;
        s->cell_size = cell_size;
        s->size = size;
        s->depth = depth;
        s->ncells = static_cast<size_t>(ncell[0]) * ncell[1] * ncell[2];
        size_t const A = array_size;
        s->pos.reset(new f4vector(2 * A)); s->vel.reset(new f4vector(2 * A)); s->img.reset(new f4vector(A));
        s->pos_b.reset(new f4vector(2 * A)); s->vel_b.reset(new f4vector(2 * A)); s->img_b.reset(new f4vector(A));
        s->force.reset(new f4vector(A)); s->r0.reset(new f4vector(A));
        s->id.reset(new uvector(A)); s->id_b.reset(new uvector(A)); s->rid.reset(new uvector(A));
        s->g_cell.reset(new uvector(s->ncells * cell_size));
        s->cell_index.reset(new uvector(nparticle)); s->cell_index->reserve(A);
        s->cell_perm.reset(new uvector(nparticle)); s->cell_perm->reserve(A);
        s->cell_offset.reset(new uvector(s->ncells));
        s->nl.reset(new uvector(static_cast<size_t>(size) * A));
        RG_CHECK(cudaMemcpy(raw(*s->pos), position8, 2 * A * sizeof(float4), cudaMemcpyHostToDevice));
        RG_CHECK(cudaMemcpy(raw(*s->vel), velocity8, 2 * A * sizeof(float4), cudaMemcpyHostToDevice));
        RG_CHECK(cudaMemset(raw(*s->img), 0, A * sizeof(float4)));
        RG_CHECK(cudaMemset(raw(*s->force), 0, A * sizeof(float4)));
        halmd::iota(s->id->begin(), s->id->begin() + A, 0);
        halmd::iota(s->rid->begin(), s->rid->begin() + A, 0);
        float const rrcs = static_cast<float>(std::pow(r_cut + skin, 2));      // from_binning.cpp:78-82
        float2 const h_lj = make_float2(epsilon, sigma2);
        float4 const h_sm = make_float4(r_cut, rr_cut, en_cut, 0.f);
        RG_CHECK(cudaMemcpy(raw(s->d_rrcs), &rrcs, sizeof(float), cudaMemcpyHostToDevice));
        RG_CHECK(cudaMemcpy(raw(s->d_lj), &h_lj, sizeof(float2), cudaMemcpyHostToDevice));
        RG_CHECK(cudaMemcpy(raw(s->d_sm), &h_sm, sizeof(float4), cudaMemcpyHostToDevice));
        s->t_rrcs.reset(new lintex<float>(raw(s->d_rrcs), 1));
        s->t_lj.reset(new lintex<float2>(raw(s->d_lj), 1));
        s->t_sm.reset(new lintex<float4>(raw(s->d_sm), 1));
        int rc;
        // the first force evaluation builds the lists (particle::update_force_ -> neighbour cache miss)
        if ((rc = check_lists(*s))) return rc;
        if ((rc = compute_force(*s))) return rc;
        *out = s.release();
    } catch (std::exception const& e) {
        fprintf(stderr, "refsim_create: %s\n", e.what());
        return -1;
    }
    return 0;
}

void refsim_destroy(void* h) { delete static_cast<ref_sim*>(h); }

// runs `steps` MD steps; device time (CUDA events around the whole loop, host round trips included) in *ms
int refsim_run(void* h, size_t steps, float* ms, size_t* rebuilds)
{
    ref_sim& s = *static_cast<ref_sim*>(h);
    cudaEvent_t e0, e1;
    RG_CHECK(cudaEventCreate(&e0));
    RG_CHECK(cudaEventCreate(&e1));
    size_t const r0 = s.rebuilds;
    RG_CHECK(cudaEventRecord(e0));
    for (size_t k = 0; k < steps; ++k) {
        int rc = mdstep(s);
        if (rc) return rc;
    }
    RG_CHECK(cudaEventRecord(e1));
    RG_CHECK(cudaEventSynchronize(e1));
    float t = 0;
    RG_CHECK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = t;
    if (rebuilds) *rebuilds = s.rebuilds - r0;
    return 0;
}

// state in the current (Hilbert) order plus the particle ids
int refsim_get(void* h, float* position8, float* velocity8, float* image4, float* force4, uint32_t* id)
{
    ref_sim& s = *static_cast<ref_sim*>(h);
    if (position8) RG_CHECK(cudaMemcpy(position8, raw(*s.pos), 2 * s.A * sizeof(float4), cudaMemcpyDeviceToHost));
    if (velocity8) RG_CHECK(cudaMemcpy(velocity8, raw(*s.vel), 2 * s.A * sizeof(float4), cudaMemcpyDeviceToHost));
    if (image4) RG_CHECK(cudaMemcpy(image4, raw(*s.img), s.A * sizeof(float4), cudaMemcpyDeviceToHost));
    if (force4) RG_CHECK(cudaMemcpy(force4, raw(*s.force), s.A * sizeof(float4), cudaMemcpyDeviceToHost));
    if (id) RG_CHECK(cudaMemcpy(id, raw(*s.id), s.A * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
