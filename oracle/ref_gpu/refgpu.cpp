// refgpu.cpp — host-array launchers for the REFERENCE's own CUDA kernels (oracle/_ref/libhalmd_ref_gpu.so).
// TEST INFRASTRUCTURE ONLY: loaded by tests/ (and nothing else) to compare the product's kernels with the reference's.
//
// The reference's kernel translation units (halmd/mdsim/gpu/**/*_kernel.cu) are compiled by nvcc where they lie under
// /root/reference — unmodified, nothing copied — for sm_100a with the reference's own flags
// (-ftz=true -prec-div=false -prec-sqrt=false --fmad=true, cmake/platform.cmake:75); see oracle/Makefile.  This file is
// the host side, compiled by g++ like the reference's host classes: it fills device arrays and textures the way those
// classes do and launches through the same `*_wrapper::kernel` cuda::function objects (file:line cited per function).
// Every entry point takes HOST arrays and is synchronous.  Return value: 0 or a cudaError_t.
#include <halmd/config.hpp>
#include <halmd/mdsim/gpu/binning_kernel.hpp>
#include <halmd/mdsim/gpu/forces/external_kernel.hpp>
#include <halmd/mdsim/gpu/forces/pair_trunc_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/external/planar_wall_kernel.hpp>
#include <halmd/mdsim/gpu/integrators/verlet_kernel.hpp>
#include <halmd/mdsim/gpu/integrators/verlet_nvt_andersen_kernel.hpp>
#include <halmd/mdsim/gpu/integrators/verlet_nvt_hoover_kernel.hpp>
#include <halmd/random/gpu/rand48_kernel.cuh>
#include <halmd/mdsim/gpu/max_displacement_kernel.hpp>
#include <halmd/mdsim/gpu/neighbours/from_binning_kernel.hpp>
#include <halmd/mdsim/gpu/particle_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/adapters/hard_core_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/lennard_jones_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/truncations/force_shifted_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/truncations/sharp_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/truncations/shifted_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/truncations/smooth_r4_kernel.hpp>
#include <halmd/mdsim/gpu/sorts/hilbert_kernel.hpp>
#include <halmd/algorithm/gpu/reduce_kernel.hpp>
#include <halmd/numeric/accumulator.hpp>
#include <halmd/observables/dynamics/mean_quartic_displacement.hpp>
#include <halmd/observables/dynamics/mean_square_displacement.hpp>
#include <halmd/observables/dynamics/velocity_autocorrelation.hpp>
#include <halmd/observables/gpu/density_mode_kernel.hpp>
#include <halmd/observables/gpu/dynamics/tagged_particle.hpp>
#include <halmd/observables/gpu/thermodynamics_kernel.hpp>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "refgpu_util.hpp"

extern "C" float refgpu_last_kernel_ms() { return g_last_ms; }

extern "C" {

// from_binning::update (from_binning.cpp:186-240): lists pre-set to 0xFF, one block per cell, cell_size threads,
// cell_size * (2 + dimension) * sizeof(int) bytes of shared memory.
// layout 0: g_neighbour[k * stride + i] (default); layout 1: g_neighbour[i * size + k] (unroll_force_loop)
int refgpu_nlist_build(const uint32_t* g_cell, const unsigned ncell[3], unsigned cell_size, const float* position4,
                       size_t nposition, const float* rr_cut_skin, unsigned ntype1, unsigned ntype2, const float box_length[3],
                       uint32_t* g_neighbour, unsigned size, size_t stride, int layout, int* overcrowded)
{
    size_t const ncells = static_cast<size_t>(ncell[0]) * ncell[1] * ncell[2];
    size_t const nl_len = static_cast<size_t>(size) * stride;
    dbuf<unsigned int> d_cell(g_cell, ncells * cell_size);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), nposition);
    dbuf<float> d_rr(rr_cut_skin, ntype1 * ntype2);
    dbuf<unsigned int> d_nl(nl_len);
    dbuf<int> d_ret(1);
    if (!d_cell.p || !d_pos.p || !d_rr.p || !d_nl.p || !d_ret.p) return static_cast<int>(cudaErrorMemoryAllocation);
    RG_CHECK(cudaMemset(d_nl.p, 0xFF, nl_len * sizeof(unsigned int)));
    RG_CHECK(cudaMemset(d_ret.p, 0, sizeof(int)));
    lintex<float> t_rr(d_rr.p, ntype1 * ntype2);
    lintex<float4> t_r(d_pos.p, nposition);
    fixed_vector<unsigned int, 3> nc;
    for (int d = 0; d < 3; ++d) nc[d] = ncell[d];
    auto* kernel = &neighbours::from_binning_wrapper<3>::kernel.normal;
    if (layout) kernel = &neighbours::from_binning_wrapper<3>::kernel.unroll_force_loop;
    RG_TRY(
        kernel->update_neighbours.configure(dim3(static_cast<unsigned>(ncells)), dim3(cell_size), cell_size * (2 + 3) * sizeof(int));
        kernel->update_neighbours(t_rr.t, t_r.t, t_r.t, d_ret.p, d_nl.p, size, static_cast<unsigned>(stride), d_cell.p, d_cell.p,
                                  ntype1, ntype2, static_cast<unsigned>(ncells * cell_size), nc,
                                  vec3<fixed_vector<float, 3>>(box_length));
    )
    RG_CHECK(cudaDeviceSynchronize());
    int ret = 0;
    RG_CHECK(static_cast<cudaError_t>(d_ret.down(&ret)));
    if (overcrowded) *overcrowded = ret;
    return d_nl.down(g_neighbour);
}

// sorts::hilbert (hilbert.cpp:59, 88-101): depth symbol, then `map` over the whole particle array
int refgpu_hilbert_map(const float* position4, size_t array_size, const float box_length[3], unsigned depth, uint32_t* keys)
{
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), array_size);
    dbuf<unsigned int> d_key(array_size);
    if (!d_pos.p || !d_key.p) return static_cast<int>(cudaErrorMemoryAllocation);
    auto& kernel = sorts::hilbert_wrapper<3>::kernel;
    RG_TRY(
        kernel.depth.set(depth);
        kernel.map.configure(dim3(static_cast<unsigned>(array_size / 128)), dim3(128));
        kernel.map(d_pos.p, d_key.p, vec3<fixed_vector<float, 3>>(box_length));
    )
    RG_CHECK(cudaDeviceSynchronize());
    return d_key.down(keys);
}

// max_displacement::compute (max_displacement.cpp:77-106): block maxima on the device, maximum over blocks on the host
int refgpu_displacement_max(const float* position4, const float* r0_4, size_t nparticle, const float box_length[3], float* max_rr)
{
    unsigned const blocks = 64, threads = 128;
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), nparticle);
    dbuf<float4> d_r0(reinterpret_cast<const float4*>(r0_4), nparticle);
    dbuf<float> d_rr(blocks);
    if (!d_pos.p || !d_r0.p || !d_rr.p) return static_cast<int>(cudaErrorMemoryAllocation);
    auto& kernel = max_displacement_wrapper<3>::kernel;
    RG_TRY(
        kernel.displacement.configure(dim3(blocks), dim3(threads));
        kernel.displacement(d_pos.p, d_r0.p, d_rr.p, static_cast<unsigned>(nparticle), vec3<fixed_vector<float, 3>>(box_length));
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<float> h(blocks);
    RG_CHECK(static_cast<cudaError_t>(d_rr.down(h.data())));
    *max_rr = *std::max_element(h.begin(), h.end());
    return 0;
}

// binning::update (binning.cpp:176-183): cell index of every slot of the particle array
int refgpu_compute_cell(const float* position4, size_t array_size, const float cell_length[3], const unsigned ncell[3],
                        uint32_t* cell)
{
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), array_size);
    dbuf<unsigned int> d_cell(array_size);
    if (!d_pos.p || !d_cell.p) return static_cast<int>(cudaErrorMemoryAllocation);
    fixed_vector<unsigned int, 3> nc;
    for (int d = 0; d < 3; ++d) nc[d] = ncell[d];
    auto& kernel = binning_wrapper<3>::kernel;
    RG_TRY(
        kernel.compute_cell.configure(dim3(static_cast<unsigned>(array_size / 128)), dim3(128));
        kernel.compute_cell(d_pos.p, d_cell.p, vec3<fixed_vector<float, 3>>(cell_length), nc);
    )
    RG_CHECK(cudaDeviceSynchronize());
    return d_cell.down(cell);
}

// binning::update (binning.cpp:166-221): compute_cell, gen_index, stable sort of (cell, index) — the reference's
// radix_sort is a stable LSD sort (test/unit/algorithm/radix_sort.cpp compares it with std::stable_sort); it is
// replaced by std::stable_sort on the host here —, find_cell_offset, assign_cells.  g_cell[ncells * cell_size].
int refgpu_binning_update(const float* position4, size_t nparticle, size_t array_size, const float cell_length[3],
                          const unsigned ncell[3], unsigned cell_size, uint32_t* g_cell, int* overcrowded)
{
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    size_t const ncells = static_cast<size_t>(ncell[0]) * ncell[1] * ncell[2];
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), array_size);
    dbuf<unsigned int> d_index(array_size), d_perm(array_size), d_offset(ncells), d_cell(ncells * cell_size);
    dbuf<int> d_ret(1);
    if (!d_pos.p || !d_index.p || !d_perm.p || !d_offset.p || !d_cell.p || !d_ret.p) return static_cast<int>(cudaErrorMemoryAllocation);
    fixed_vector<unsigned int, 3> nc;
    for (int d = 0; d < 3; ++d) nc[d] = ncell[d];
    auto& kernel = binning_wrapper<3>::kernel;
    dim3 const grid(static_cast<unsigned>(array_size / 128)), block(128);
    unsigned const nbox = static_cast<unsigned>(nparticle);
    RG_TRY(
        kernel.compute_cell.configure(grid, block);
        kernel.compute_cell(d_pos.p, d_index.p, vec3<fixed_vector<float, 3>>(cell_length), nc);
        kernel.gen_index.configure(grid, block);
        kernel.gen_index(d_perm.p, nbox);
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<unsigned int> h_index(array_size), h_perm(array_size), order(array_size);
    RG_CHECK(static_cast<cudaError_t>(d_index.down(h_index.data())));
    RG_CHECK(static_cast<cudaError_t>(d_perm.down(h_perm.data())));
    for (size_t i = 0; i < array_size; ++i) order[i] = static_cast<unsigned int>(i);
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return h_index[a] < h_index[b]; });
    std::vector<unsigned int> s_index(array_size), s_perm(array_size);
    for (size_t i = 0; i < array_size; ++i) { s_index[i] = h_index[order[i]]; s_perm[i] = h_perm[order[i]]; }
    RG_CHECK(cudaMemcpy(d_index.p, s_index.data(), array_size * sizeof(unsigned int), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemcpy(d_perm.p, s_perm.data(), array_size * sizeof(unsigned int), cudaMemcpyHostToDevice));
    RG_CHECK(cudaMemset(d_offset.p, 0xFF, ncells * sizeof(unsigned int)));
    RG_CHECK(cudaMemset(d_ret.p, 0, sizeof(int)));
    RG_TRY(
        kernel.find_cell_offset.configure(grid, block);
        kernel.find_cell_offset(d_index.p, d_offset.p, nbox);
        kernel.assign_cells.configure(dim3(static_cast<unsigned>(ncells)), dim3(cell_size));
        kernel.assign_cells(d_ret.p, d_index.p, d_offset.p, d_perm.p, d_cell.p, nbox, cell_size);
    )
    RG_CHECK(cudaDeviceSynchronize());
    int ret = 0;
    RG_CHECK(static_cast<cudaError_t>(d_ret.down(&ret)));
    if (overcrowded) *overcrowded = ret;
    return d_cell.down(g_cell);
}

// integrators::verlet_nvt_andersen::finalize (verlet_nvt_andersen.cpp:109-138): launched with the execution dimensions
// of the random number generator; rand48 state (ushort3 per RNG thread) and leapfrog constants a, c (16-bit words, low
// word first) are passed in and the advanced state is returned.
int refgpu_andersen_finalize(float* velocity8, const float* force4, size_t nparticle, size_t array_size, float timestep,
                             float sqrt_temperature, float coll_prob, unsigned blocks, unsigned threads,
                             const uint16_t a[3], const uint16_t c[3], uint16_t* state)
{
    typedef integrators::verlet_nvt_andersen_wrapper<3, dsfloat, random::gpu::rand48_rng> wrapper_type;
    typedef wrapper_type::ptr_type ptr_type;
    size_t const nthread = static_cast<size_t>(blocks) * threads;
    dbuf<float4> d_vel(reinterpret_cast<const float4*>(velocity8), 2 * array_size);
    dbuf<float4> d_f(reinterpret_cast<const float4*>(force4), array_size);
    dbuf<ushort3> d_state(reinterpret_cast<const ushort3*>(state), nthread);
    if (!d_vel.p || !d_f.p || !d_state.p) return static_cast<int>(cudaErrorMemoryAllocation);
    random::gpu::rand48_rng rng;
    rng.a = make_uint48(a[0], a[1], a[2]);
    rng.c = make_uint48(c[0], c[1], c[2]);
    rng.g_state = d_state.p;
    ptr_type vel = {d_vel.p, d_vel.p + array_size};
    RG_TRY(
        wrapper_type::kernel.finalize.configure(dim3(blocks), dim3(threads));
        wrapper_type::kernel.finalize(vel, d_f.p, timestep, sqrt_temperature, coll_prob, static_cast<unsigned>(nparticle),
                                      static_cast<unsigned>(array_size), rng);
    )
    RG_CHECK(cudaDeviceSynchronize());
    int rc;
    if ((rc = d_vel.down(reinterpret_cast<float4*>(velocity8)))) return rc;
    return d_state.down(reinterpret_cast<ushort3*>(state));
}

// integrators::verlet (verlet.cpp:64-121), double-single arrays: high words [0, A), low words [A, 2A)
int refgpu_verlet(int do_finalize, int do_integrate, float* position8, float* image4, float* velocity8, const float* force4,
                  size_t array_size, float timestep, const float box_length[3])
{
    typedef integrators::verlet_wrapper<3, dsfloat> wrapper_type;
    typedef wrapper_type::ptr_type ptr_type;
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position8), 2 * array_size);
    dbuf<float4> d_vel(reinterpret_cast<const float4*>(velocity8), 2 * array_size);
    dbuf<float4> d_img(reinterpret_cast<const float4*>(image4), array_size);
    dbuf<float4> d_f(reinterpret_cast<const float4*>(force4), array_size);
    if (!d_pos.p || !d_vel.p || !d_img.p || !d_f.p) return static_cast<int>(cudaErrorMemoryAllocation);
    ptr_type pos = {d_pos.p, d_pos.p + array_size}, vel = {d_vel.p, d_vel.p + array_size};
    dim3 const grid(static_cast<unsigned>(array_size / 128)), block(128);
    RG_TRY(
        if (do_finalize) {
            wrapper_type::wrapper.finalize.configure(grid, block);
            wrapper_type::wrapper.finalize(vel, d_f.p, timestep);
        }
        if (do_integrate) {
            wrapper_type::wrapper.integrate.configure(grid, block);
            wrapper_type::wrapper.integrate(pos, d_img.p, vel, d_f.p, timestep, vec3<fixed_vector<float, 3>>(box_length));
        }
    )
    RG_CHECK(cudaDeviceSynchronize());
    int rc;
    if ((rc = d_pos.down(reinterpret_cast<float4*>(position8)))) return rc;
    if ((rc = d_vel.down(reinterpret_cast<float4*>(velocity8)))) return rc;
    return d_img.down(reinterpret_cast<float4*>(image4));
}

// integrators::verlet_nvt_hoover (verlet_nvt_hoover.cpp:103-183): which = 0 integrate(scale), 1 finalize, 2 rescale(scale);
// the host passes its double-precision scale factor through dsfloat's converting constructor
int refgpu_hoover(int which, float* position8, float* image4, float* velocity8, const float* force4, size_t array_size,
                  float timestep, double scale, const float box_length[3])
{
    typedef integrators::verlet_nvt_hoover_wrapper<3, dsfloat> wrapper_type;
    typedef wrapper_type::ptr_type ptr_type;
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position8), 2 * array_size);
    dbuf<float4> d_vel(reinterpret_cast<const float4*>(velocity8), 2 * array_size);
    dbuf<float4> d_img(reinterpret_cast<const float4*>(image4), array_size);
    dbuf<float4> d_f(reinterpret_cast<const float4*>(force4), array_size);
    if (!d_pos.p || !d_vel.p || !d_img.p || !d_f.p) return static_cast<int>(cudaErrorMemoryAllocation);
    ptr_type pos = {d_pos.p, d_pos.p + array_size}, vel = {d_vel.p, d_vel.p + array_size};
    dim3 const grid(static_cast<unsigned>(array_size / 128)), block(128);
    RG_TRY(
        if (which == 0) {
            wrapper_type::kernel.integrate.configure(grid, block);
            wrapper_type::kernel.integrate(pos, d_img.p, vel, d_f.p, timestep, dsfloat(scale), vec3<fixed_vector<float, 3>>(box_length));
        }
        else if (which == 1) {
            wrapper_type::kernel.finalize.configure(grid, block);
            wrapper_type::kernel.finalize(vel, d_f.p, timestep);
        }
        else {
            wrapper_type::kernel.rescale.configure(grid, block);
            wrapper_type::kernel.rescale(vel, dsfloat(scale));
        }
    )
    RG_CHECK(cudaDeviceSynchronize());
    int rc;
    if ((rc = d_pos.down(reinterpret_cast<float4*>(position8)))) return rc;
    if ((rc = d_vel.down(reinterpret_cast<float4*>(velocity8)))) return rc;
    return d_img.down(reinterpret_cast<float4*>(image4));
}

} // extern "C"

namespace {

namespace lj = potentials::pair::lennard_jones_kernel;
namespace tr = potentials::pair::truncations;

// pair_trunc::compute_ / compute_aux_ (pair_trunc.hpp:217-337): one thread per slot of the particle array (default lists)
// or nparallel_particles threads per slot (unroll_force_loop)
template <typename potential_type>
int launch_force(potential_type const& pot, const float4* pos, cudaTextureObject_t t_r2, float4* force, const unsigned int* nl,
                 unsigned size, unsigned stride, int layout, float* en_pot, float* stress, unsigned ntype, const float* L,
                 bool force_zero, float aux_weight, size_t array_size)
{
    typedef forces::pair_trunc_wrapper<3, potential_type> wrapper_type;
    auto& k = wrapper_type::kernel;
    fixed_vector<float, 3> const box = vec3<fixed_vector<float, 3>>(L);
    bool const aux = en_pot != nullptr;
    RG_TRY(
        if (layout) {
            size_t const threads = array_size * k.nparallel_particles;
            auto& f = aux ? k.compute_aux_unroll_force_loop : k.compute_unroll_force_loop;
            f.configure(dim3(static_cast<unsigned>(threads / 256)), dim3(256));
            f(pot, pos, t_r2, force, nl, size, en_pot, stress, ntype, ntype, box, force_zero, aux_weight);
        } else {
            auto& f = aux ? k.compute_aux : k.compute;
            f.configure(dim3(static_cast<unsigned>(array_size / 128)), dim3(128));
            f(pot, pos, t_r2, force, nl, size, stride, en_pot, stress, ntype, ntype, box, force_zero, aux_weight);
        }
    )
    RG_CHECK(cudaDeviceSynchronize());
    return 0;
}

} // namespace

// truncation: 0 sharp, 1 shifted, 2 smooth_r4, 3 force_shifted (force_cut table passed through `en_pot`-independent
// argument force_cut, may be NULL otherwise).  Parameter tables [ntype * ntype] as lennard_jones.cpp:60-105 and
// truncations/{sharp,shifted,smooth_r4}.hpp upload them: (epsilon, sigma2); rr_cut; (rr_cut, en_cut);
// (r_cut, rr_cut, en_cut); rri_smooth = h^-2 (smooth_r4.hpp:97).
// force (in/out, array_size float4); en_pot (array_size) and stress (6 * array_size) may both be NULL (no aux).
static int pair_force_impl(int truncation, unsigned ntype, const float* epsilon, const float* sigma2, const float* r_cut,
                           const float* rr_cut, const float* en_cut, const float* force_cut, const float* r_core,
                           float rri_smooth, const float* position4, size_t array_size, const uint32_t* g_neighbour,
                           unsigned size, size_t stride, int layout, const float box_length[3], int force_zero,
                           float aux_weight, float* force4, float* en_pot, float* stress)
{
    if (array_size % 256 != 0) return static_cast<int>(cudaErrorInvalidValue);
    unsigned const np = ntype * ntype;
    std::vector<float2> h_lj(np), h_sh(np);
    std::vector<float4> h_sm(np), h_fs(np);
    std::vector<float> h_sp(np), h_core(np);
    for (unsigned t = 0; t < np; ++t) {
        h_lj[t] = make_float2(epsilon[t], sigma2[t]);                 // lennard_jones_kernel::EPSILON, SIGMA2
        h_sm[t] = make_float4(r_cut[t], rr_cut[t], en_cut[t], 0.f);   // smooth_r4_kernel::R_CUT, RR_CUT, EN_CUT
        h_sh[t] = make_float2(rr_cut[t], en_cut[t]);                  // shifted_kernel::RR_CUT, EN_CUT
        h_sp[t] = rr_cut[t];
        // force_shifted_kernel::R_CUT, RR_CUT, EN_CUT, FORCE_CUT
        h_fs[t] = make_float4(r_cut[t], rr_cut[t], en_cut[t], force_cut ? force_cut[t] : 0.f);
        h_core[t] = r_core ? r_core[t] : 0.f;                         // adapters/hard_core.hpp:64-72
    }
    dbuf<float2> d_lj(h_lj.data(), np), d_sh(h_sh.data(), np);
    dbuf<float4> d_sm(h_sm.data(), np), d_fs(h_fs.data(), np);
    dbuf<float> d_sp(h_sp.data(), np), d_core(h_core.data(), np);
    size_t const nl_len = static_cast<size_t>(size) * stride;
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), array_size);
    dbuf<unsigned int> d_nl(g_neighbour, nl_len);
    dbuf<float4> d_f(reinterpret_cast<const float4*>(force4), array_size);
    bool const aux = en_pot != nullptr;
    dbuf<float> d_en(en_pot, aux ? array_size : 0), d_st(stress, aux ? 6 * array_size : 0);
    if (!d_pos.p || !d_nl.p || !d_f.p || !d_en.p || !d_st.p) return static_cast<int>(cudaErrorMemoryAllocation);
    lintex<float2> t_lj(d_lj.p, np), t_sh(d_sh.p, np);
    lintex<float4> t_sm(d_sm.p, np), t_fs(d_fs.p, np), t_r2(d_pos.p, array_size);
    lintex<float> t_sp(d_sp.p, np), t_core(d_core.p, np);
    float* const en = aux ? d_en.p : nullptr;
    float* const st = aux ? d_st.p : nullptr;
    unsigned const pitch = static_cast<unsigned>(stride);
    // truncation<parent>: the same four wrappers over the plain potential or its hard-core adapter
    auto run = [&](auto const& parent) -> int {
        typedef typename std::decay<decltype(parent)>::type parent_type;
        if (truncation == 2) {
            RG_TRY(tr::smooth_r4_wrapper<parent_type>::rri_smooth.set(rri_smooth))
            tr::smooth_r4_kernel::smooth_r4<parent_type> pot(parent, t_sm.t);
            return launch_force(pot, d_pos.p, t_r2.t, d_f.p, d_nl.p, size, pitch, layout, en, st, ntype, box_length, force_zero != 0, aux_weight, array_size);
        } else if (truncation == 3) {
            tr::force_shifted_kernel::force_shifted<parent_type> pot(parent, t_fs.t);
            return launch_force(pot, d_pos.p, t_r2.t, d_f.p, d_nl.p, size, pitch, layout, en, st, ntype, box_length, force_zero != 0, aux_weight, array_size);
        } else if (truncation == 1) {
            tr::shifted_kernel::shifted<parent_type> pot(parent, t_sh.t);
            return launch_force(pot, d_pos.p, t_r2.t, d_f.p, d_nl.p, size, pitch, layout, en, st, ntype, box_length, force_zero != 0, aux_weight, array_size);
        }
        tr::sharp_kernel::sharp<parent_type> pot(parent, t_sp.t);
        return launch_force(pot, d_pos.p, t_r2.t, d_f.p, d_nl.p, size, pitch, layout, en, st, ntype, box_length, force_zero != 0, aux_weight, array_size);
    };
    lj::lennard_jones const plain(t_lj.t);
    int rc;
    if (r_core) {
        // In this version of the reference the force kernels call potential.fetch_param() (pair_trunc_kernel.cuh:118),
        // which hard_core does not override: its fetch() (hard_core_kernel.cuh:37-44, the only place that reads the
        // core-radius texture) is never called and r_core_ keeps whatever the host-side functor object contained when
        // it was passed as a kernel argument.  The harness therefore stores the radius the host class uploads
        // (adapters/hard_core.hpp:60-72) into that member before the launch — one value for all species pairs, since
        // nothing in the kernel could select another — through a layout-identical mirror of the private members.
        typedef potentials::pair::adapters::hard_core_kernel::hard_core<lj::lennard_jones> adapted_type;
        struct mirror : lj::lennard_jones {
            mirror(lj::lennard_jones const& p) : lj::lennard_jones(p) {}
            float r_core;
            cudaTextureObject_t t_param;
        };
        static_assert(sizeof(mirror) == sizeof(adapted_type), "hard_core layout");
        for (unsigned t = 1; t < np; ++t) if (h_core[t] != h_core[0]) return static_cast<int>(cudaErrorInvalidValue);
        adapted_type adapted(plain, t_core.t);
        reinterpret_cast<mirror&>(adapted).r_core = h_core[0];
        rc = run(static_cast<adapted_type const&>(adapted));
    } else {
        rc = run(plain);
    }
    if (rc) return rc;
    if ((rc = d_f.down(reinterpret_cast<float4*>(force4)))) return rc;
    if (aux) {
        if ((rc = d_en.down(en_pot))) return rc;
        if ((rc = d_st.down(stress))) return rc;
    }
    return 0;
}

extern "C" int refgpu_pair_force(int truncation, unsigned ntype, const float* epsilon, const float* sigma2, const float* r_cut,
                                 const float* rr_cut, const float* en_cut, const float* force_cut, float rri_smooth,
                                 const float* position4,
                                 size_t array_size, const uint32_t* g_neighbour, unsigned size, size_t stride, int layout,
                                 const float box_length[3], int force_zero, float aux_weight, float* force4, float* en_pot,
                                 float* stress)
{
    return pair_force_impl(truncation, ntype, epsilon, sigma2, r_cut, rr_cut, en_cut, force_cut, nullptr, rri_smooth, position4,
                           array_size, g_neighbour, size, stride, layout, box_length, force_zero, aux_weight, force4, en_pot,
                           stress);
}

// the same through adapters::hard_core<lennard_jones> (adapters/hard_core_kernel.hpp:37-84); r_core = core * sigma
extern "C" int refgpu_pair_force_hard_core(int truncation, unsigned ntype, const float* epsilon, const float* sigma2,
                                           const float* r_cut, const float* rr_cut, const float* en_cut,
                                           const float* force_cut, const float* r_core, float rri_smooth,
                                           const float* position4, size_t array_size, const uint32_t* g_neighbour,
                                           unsigned size, size_t stride, int layout, const float box_length[3],
                                           int force_zero, float aux_weight, float* force4, float* en_pot, float* stress)
{
    if (!r_core) return static_cast<int>(cudaErrorInvalidValue);
    return pair_force_impl(truncation, ntype, epsilon, sigma2, r_cut, rr_cut, en_cut, force_cut, r_core, rri_smooth, position4,
                           array_size, g_neighbour, size, stride, layout, box_length, force_zero, aux_weight, force4, en_pot,
                           stress);
}

// forces::external<3, float, planar_wall> (external.hpp:120-190) with potentials::external::planar_wall
// (planar_wall.cpp:40-112): geometry texture (normal, offset), potential texture [species * nwall + wall] =
// (epsilon, sigma, wetting, cutoff), symbols nwall and smoothing; one thread per slot of the particle array.
// Normals must be normalised by the caller (the host class does it before the upload).
extern "C" int refgpu_planar_wall(unsigned nwall, unsigned nspecies, const float* normal_offset4 /* [nwall] float4 */,
                                  const float* potential4 /* [nspecies * nwall] float4 */, float smoothing,
                                  const float* position4, size_t array_size, const float box_length[3], int force_zero,
                                  float* force4, float* en_pot, float* stress)
{
    using namespace potentials::external;
    typedef planar_wall_kernel::planar_wall<3> potential_type;
    typedef forces::external_wrapper<3, potential_type> wrapper_type;
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<float4> d_geo(reinterpret_cast<const float4*>(normal_offset4), nwall);
    dbuf<float4> d_pot(reinterpret_cast<const float4*>(potential4), nspecies * nwall);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), array_size);
    dbuf<float4> d_f(reinterpret_cast<const float4*>(force4), array_size);
    bool const aux = en_pot != nullptr;
    dbuf<float> d_en(en_pot, aux ? array_size : 0), d_st(stress, aux ? 6 * array_size : 0);
    if (!d_geo.p || !d_pot.p || !d_pos.p || !d_f.p || !d_en.p || !d_st.p) return static_cast<int>(cudaErrorMemoryAllocation);
    lintex<float4> t_geo(d_geo.p, nwall), t_pot(d_pot.p, nspecies * nwall);
    RG_TRY(
        planar_wall_wrapper::nwall.set(static_cast<int>(nwall));
        planar_wall_wrapper::smoothing.set(smoothing);
        potential_type pot(t_geo.t, t_pot.t);
        auto& f = aux ? wrapper_type::kernel.compute_aux : wrapper_type::kernel.compute;
        f.configure(dim3(static_cast<unsigned>(array_size / 128)), dim3(128));
        f(pot, d_pos.p, d_f.p, aux ? d_en.p : nullptr, aux ? d_st.p : nullptr, vec3<fixed_vector<float, 3>>(box_length),
          force_zero != 0);
    )
    RG_CHECK(cudaDeviceSynchronize());
    int rc;
    if ((rc = d_f.down(reinterpret_cast<float4*>(force4)))) return rc;
    if (aux) {
        if ((rc = d_en.down(en_pot))) return rc;
        if ((rc = d_st.down(stress))) return rc;
    }
    return 0;
}

// observables::gpu::density_mode::acquire (density_mode.cpp:85-131): block sums of exp(i q.r) per wavevector, then one
// block per wavevector finalises; the host stores (rho.x, -rho.y).  The group is all particles (identity index).
extern "C" int refgpu_density_modes(const float* position4, size_t nparticle, const float* q3, unsigned nq, double* rho2)
{
    unsigned const blocks = 32, threads = 256;
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position4), nparticle);
    std::vector<float4> h_q(nq);
    for (unsigned k = 0; k < nq; ++k) h_q[k] = make_float4(q3[3 * k], q3[3 * k + 1], q3[3 * k + 2], 0.f);
    dbuf<float4> d_q(h_q.data(), nq);
    std::vector<unsigned int> h_idx(nparticle);
    for (size_t i = 0; i < nparticle; ++i) h_idx[i] = static_cast<unsigned int>(i);
    dbuf<unsigned int> d_idx(h_idx.data(), nparticle);
    dbuf<float2> d_block(static_cast<size_t>(nq) * blocks), d_rho(nq);
    if (!d_pos.p || !d_q.p || !d_idx.p || !d_block.p || !d_rho.p) return static_cast<int>(cudaErrorMemoryAllocation);
    lintex<float4> t_q(d_q.p, nq);
    auto& kernel = observables::gpu::density_mode_wrapper<3>::kernel;
    RG_TRY(
        kernel.compute.configure(dim3(blocks), dim3(threads));
        kernel.compute(t_q.t, d_pos.p, d_idx.p, static_cast<int>(nparticle), d_block.p, static_cast<int>(nq));
        kernel.finalise.configure(dim3(nq), dim3(threads));
        kernel.finalise(d_block.p, d_rho.p, static_cast<int>(nq), static_cast<int>(blocks));
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<float2> h(nq);
    RG_CHECK(static_cast<cudaError_t>(d_rho.down(h.data())));
    for (unsigned k = 0; k < nq; ++k) { rho2[2 * k] = h[k].x; rho2[2 * k + 1] = -h[k].y; }
    return 0;
}

// particle::rearrange (particle.cpp:173-218): gather positions / velocities (double-single: hi and lo planes), images
// and ids by a permutation through textures; reverse ids are rebuilt by the caller in the reference (particle.cpp:213)
extern "C" int refgpu_rearrange(const uint32_t* index, size_t nparticle, size_t array_size, const float* position8,
                                const float* velocity8, const float* image4, const uint32_t* id, float* position8_out,
                                float* velocity8_out, float* image4_out, uint32_t* id_out)
{
    typedef particle_wrapper<3, dsfloat> wrapper_type;
    typedef wrapper_type::ptr_type ptr_type;
    if (array_size % 128 != 0) return static_cast<int>(cudaErrorInvalidValue);
    dbuf<unsigned int> d_index(index, nparticle);
    dbuf<float4> d_pos(reinterpret_cast<const float4*>(position8), 2 * array_size);
    dbuf<float4> d_vel(reinterpret_cast<const float4*>(velocity8), 2 * array_size);
    dbuf<float4> d_img(reinterpret_cast<const float4*>(image4), array_size);
    dbuf<unsigned int> d_id(id, array_size);
    dbuf<float4> o_pos(2 * array_size), o_vel(2 * array_size), o_img(array_size);
    dbuf<unsigned int> o_id(array_size);
    if (!d_index.p || !d_pos.p || !d_vel.p || !d_img.p || !d_id.p || !o_pos.p || !o_vel.p || !o_img.p || !o_id.p)
        return static_cast<int>(cudaErrorMemoryAllocation);
    // texFetch<float4, dsfloat> reads the low word at index + texture size / 2 (dsfloat_cuda_vector layout)
    lintex<float4> t_r(d_pos.p, 2 * array_size), t_v(d_vel.p, 2 * array_size), t_img(d_img.p, array_size);
    lintex<unsigned int> t_id(d_id.p, array_size);
    ptr_type pos = {o_pos.p, o_pos.p + array_size}, vel = {o_vel.p, o_vel.p + array_size};
    RG_TRY(
        wrapper_type::kernel.rearrange.configure(dim3(static_cast<unsigned>(array_size / 128)), dim3(128));
        wrapper_type::kernel.rearrange(t_r.t, t_img.t, t_v.t, t_id.t, d_index.p, pos, o_img.p, vel, o_id.p,
                                       static_cast<unsigned>(nparticle));
    )
    RG_CHECK(cudaDeviceSynchronize());
    int rc;
    if ((rc = o_pos.down(reinterpret_cast<float4*>(position8_out)))) return rc;
    if ((rc = o_vel.down(reinterpret_cast<float4*>(velocity8_out)))) return rc;
    if ((rc = o_img.down(reinterpret_cast<float4*>(image4_out)))) return rc;
    return o_id.down(id_out);
}

// observables::gpu::dynamics::mean_square_displacement::operator() (mean_square_displacement.cpp:55-68): the
// reduction kernel instantiated in mean_square_displacement_kernel.cu over zip(first, second) with the
// tagged_particle<mean_square_displacement<3, float>, dsfloat> accumulator, 32 blocks x 256 threads
// (defaults, :45-52), block results merged on the host like reduction::operator() (algorithm/gpu/reduce.hpp:101-112).
// out = (count, mean, variance) of |r2 - r1|^2 over the particles.
extern "C" int refgpu_msd(const float* first4, const float* second4, size_t nparticle, double out[3])
{
    typedef observables::gpu::dynamics::tagged_particle<observables::dynamics::mean_square_displacement<3, float>, dsfloat> acc_type;
    unsigned const blocks = 32, threads = 256;
    dbuf<float4> d_a(reinterpret_cast<const float4*>(first4), nparticle), d_b(reinterpret_cast<const float4*>(second4), nparticle);
    dbuf<acc_type> d_block(blocks);
    if (!d_a.p || !d_b.p || !d_block.p) return static_cast<int>(cudaErrorMemoryAllocation);
    auto& kernel = reduction_kernel<acc_type>::kernel;
    RG_TRY(
        kernel.reduce.configure(dim3(blocks), dim3(threads));
        acc_type::iterator first(std::make_tuple(static_cast<float4 const*>(d_a.p), static_cast<float4 const*>(d_b.p)));
        kernel.reduce(first, static_cast<std::ptrdiff_t>(nparticle), d_block.p, acc_type());
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<acc_type> h(blocks);
    RG_CHECK(static_cast<cudaError_t>(d_block.down(h.data())));
    acc_type total;
    for (unsigned i = 0; i < blocks; ++i) total(h[i]);
    out[0] = static_cast<double>(count(total()));
    out[1] = static_cast<double>(mean(total()));
    out[2] = static_cast<double>(variance(total()));
    return 0;
}

namespace {

template <typename correlation_function>
int reduce_correlation(const float* first4, const float* second4, size_t nparticle, double out[3])
{
    typedef observables::gpu::dynamics::tagged_particle<correlation_function, dsfloat> acc_type;
    unsigned const blocks = 32, threads = 256;
    dbuf<float4> d_a(reinterpret_cast<const float4*>(first4), nparticle), d_b(reinterpret_cast<const float4*>(second4), nparticle);
    dbuf<acc_type> d_block(blocks);
    if (!d_a.p || !d_b.p || !d_block.p) return static_cast<int>(cudaErrorMemoryAllocation);
    auto& kernel = reduction_kernel<acc_type>::kernel;
    RG_TRY(
        kernel.reduce.configure(dim3(blocks), dim3(threads));
        typename acc_type::iterator first(std::make_tuple(static_cast<float4 const*>(d_a.p), static_cast<float4 const*>(d_b.p)));
        kernel.reduce(first, static_cast<std::ptrdiff_t>(nparticle), d_block.p, acc_type());
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<acc_type> h(blocks);
    RG_CHECK(static_cast<cudaError_t>(d_block.down(h.data())));
    acc_type total;
    for (unsigned i = 0; i < blocks; ++i) total(h[i]);
    out[0] = static_cast<double>(count(total()));
    out[1] = static_cast<double>(mean(total()));
    out[2] = static_cast<double>(variance(total()));
    return 0;
}

} // namespace

// mean_quartic_displacement_kernel.cu / velocity_autocorrelation_kernel.cu: the same reduction as refgpu_msd with the
// other two tagged-particle correlation functions (kind 1 / 2)
extern "C" int refgpu_correlate(int kind, const float* first4, const float* second4, size_t nparticle, double out[3])
{
    if (kind == 1) return reduce_correlation<observables::dynamics::mean_quartic_displacement<3, float>>(first4, second4, nparticle, out);
    if (kind == 2) return reduce_correlation<observables::dynamics::velocity_autocorrelation<3, float>>(first4, second4, nparticle, out);
    return reduce_correlation<observables::dynamics::mean_square_displacement<3, float>>(first4, second4, nparticle, out);
}

namespace {

// reduce(&*unordered.begin(), &*unordered.end(), acc)  (mdsim/gpu/particle_group.hpp:133-247 through
// algorithm/gpu/reduce.hpp:60-104): the reduction kernel over the particle indices of the group with the default
// 16 x 1024 execution configuration, block results merged on the host in block order
template <typename acc_type>
int reduce_group(const unsigned int* d_group, size_t count, acc_type const& acc, acc_type& total)
{
    unsigned const blocks = 16, threads = 1024;
    dbuf<acc_type> d_block(blocks);
    if (!d_block.p) return static_cast<int>(cudaErrorMemoryAllocation);
    auto& kernel = reduction_kernel<acc_type>::kernel;
    RG_TRY(
        kernel.reduce.configure(dim3(blocks), dim3(threads));
        kernel.reduce(d_group, static_cast<std::ptrdiff_t>(count), d_block.p, acc);
    )
    RG_CHECK(cudaDeviceSynchronize());
    std::vector<char> raw_bytes(blocks * sizeof(acc_type));
    RG_CHECK(cudaMemcpy(raw_bytes.data(), d_block.p, raw_bytes.size(), cudaMemcpyDeviceToHost));
    acc_type const* h = reinterpret_cast<acc_type const*>(raw_bytes.data());
    total = acc;
    for (unsigned i = 0; i < blocks; ++i) total(h[i]);
    return 0;
}

} // namespace

// observables::gpu::thermodynamics (thermodynamics.cpp:64-200) for the group of all particles: the accumulators of
// observables/gpu/thermodynamics_kernel.{hpp,cu} — kinetic_energy, velocity_of_centre_of_mass, potential_energy, virial —
// over velocity (float4[A], mass in w), potential energy (float[A]) and the stress tensor (float[6 A], stride A).
// out = (sum m v^2, sum m v_x, sum m v_y, sum m v_z, sum m, sum en_pot, sum of the virial trace)
extern "C" int refgpu_thermodynamics(const float* velocity4, const float* en_pot, const float* stress, size_t array_size,
                                     size_t nparticle, double out[7])
{
    using namespace observables::gpu;
    dbuf<float4> d_vel(reinterpret_cast<const float4*>(velocity4), array_size);
    dbuf<float> d_en(en_pot, array_size), d_st(stress, 6 * array_size);
    std::vector<unsigned int> h_group(nparticle);
    for (size_t i = 0; i < nparticle; ++i) h_group[i] = static_cast<unsigned int>(i);
    dbuf<unsigned int> d_group(h_group.data(), nparticle);
    if (!d_vel.p || !d_en.p || !d_st.p || !d_group.p) return static_cast<int>(cudaErrorMemoryAllocation);
    lintex<float4> t_v(d_vel.p, array_size);
    lintex<float> t_en(d_en.p, array_size), t_st(d_st.p, 6 * array_size);
    int rc;
    kinetic_energy<3, dsfloat> ekin(t_v.t);
    if ((rc = reduce_group(d_group.p, nparticle, kinetic_energy<3, dsfloat>(t_v.t), ekin))) return rc;
    out[0] = 2 * ekin();
    velocity_of_centre_of_mass<3, dsfloat> vcm(t_v.t);
    if ((rc = reduce_group(d_group.p, nparticle, velocity_of_centre_of_mass<3, dsfloat>(t_v.t), vcm))) return rc;
    fixed_vector<double, 3> mv;
    double m;
    std::tie(mv, m) = vcm();
    out[1] = mv[0]; out[2] = mv[1]; out[3] = mv[2]; out[4] = m;
    potential_energy<dsfloat> epot(t_en.t);
    if ((rc = reduce_group(d_group.p, nparticle, potential_energy<dsfloat>(t_en.t), epot))) return rc;
    out[5] = epot();
    virial<3, dsfloat> vir(t_st.t, static_cast<unsigned int>(array_size));
    if ((rc = reduce_group(d_group.p, nparticle, virial<3, dsfloat>(t_st.t, static_cast<unsigned int>(array_size)), vir))) return rc;
    out[6] = vir();
    return 0;
}
