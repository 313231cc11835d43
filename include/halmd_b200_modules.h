/*
 * halmd_b200_modules.h — C ABI of the host-side module layer.
 *
 * The C++ classes behind these handles (halmd_b200/csrc/modules.cu) mirror, one to one, the host classes that
 * HALMD binds into Lua; each drives the kernel-level entry points of halmd_b200.h exactly where the reference's
 * class drives its `*_wrapper::kernel`.  Keyword arguments of the Lua constructors map to the arguments here.
 *
 *   Lua module (lua/halmd/...)                        reference class (file:line)                       handle
 *   mdsim.box{length}                                 mdsim/box.cpp:37-56                               (3 doubles)
 *   mdsim.particle{dimension,particles,species}       mdsim/gpu/particle.cpp:56-160                     hxm_particle
 *   mdsim.potentials.pair.lennard_jones{...}:truncate potentials/pair/lennard_jones.cpp:44-70,
 *                                                     truncations/smooth_r4.hpp:55-99                   hxm_potential
 *   mdsim.binning{box,particle,r_cut,skin,occupancy}  mdsim/gpu/binning.cpp:47-221                      hxm_binning
 *   mdsim.max_displacement{box,particle}              mdsim/gpu/max_displacement.cpp:40-106             (inside hxm_neighbour)
 *   mdsim.sorts.hilbert{box,particle}                 mdsim/gpu/sorts/hilbert.cpp:40-117                (inside hxm_neighbour)
 *   mdsim.neighbour{particle,box,r_cut,skin,...}      mdsim/gpu/neighbours/from_binning.cpp:47-282      hxm_neighbour
 *   mdsim.forces.pair{box,particle,potential,neighbour} mdsim/gpu/forces/pair_trunc.hpp:175-370         hxm_force
 *   mdsim.integrators.verlet{box,particle,timestep}   mdsim/gpu/integrators/verlet.cpp:40-151           hxm_integrator
 *   mdsim.integrators.verlet_nvt_andersen{...}        mdsim/gpu/integrators/verlet_nvt_andersen.cpp:36-143  hxm_integrator
 *   mdsim.core / observables.sampler:run(steps)       mdsim/core.cpp:30-63, observables/sampler.cpp:49-83   hxm_core
 *   observables.thermodynamics{box,group,force}       observables/gpu/thermodynamics.cpp                hxm_core_thermodynamics
 *   observables.dynamics.blocking_scheme + correlation observables/dynamics/blocking_scheme.cpp:40-200,
 *                                                     observables/dynamics/correlation.hpp:208-224      hxm_blocking
 *
 * All functions return 0 on success, a cudaError_t (>0) or an HX_E_* code (<0); hx_last_error_string() has the
 * message.  Handles are opaque; every module works on the CUDA stream of its particle instance.
 */
#ifndef HALMD_B200_MODULES_H
#define HALMD_B200_MODULES_H

#include "halmd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hxm_particle hxm_particle;
typedef struct hxm_potential hxm_potential;
typedef struct hxm_binning hxm_binning;
typedef struct hxm_neighbour hxm_neighbour;
typedef struct hxm_force hxm_force;
typedef struct hxm_integrator hxm_integrator;
typedef struct hxm_core hxm_core;
typedef struct hxm_blocking hxm_blocking;

/* ---- mdsim.particle ---------------------------------------------------------------------------------------- */
int hxm_particle_create(size_t nparticle, unsigned nspecies, void* stream, hxm_particle** out);
void hxm_particle_destroy(hxm_particle* p);
size_t hxm_particle_nparticle(const hxm_particle* p);
size_t hxm_particle_array_size(const hxm_particle* p);
unsigned hxm_particle_nspecies(const hxm_particle* p);
/* particle.data[name] (particle.cpp:127-140): host arrays in the reference's host formats.
 * names: "position" (n x 3 float; set clears the low words), "position_ds" (n x 3 double, hi+lo),
 * "velocity", "velocity_ds", "species" (n uint32), "mass" (n float), "image" (n x 3 float), "force" (n x 3 float),
 * "id", "reverse_id" (n uint32), "potential_energy" (n float), "potential_stress_tensor" (n x 6 float).
 * Arrays are indexed by particle slot (memory order), like particle:get_* in the reference. */
int hxm_particle_set(hxm_particle* p, const char* name, const void* host_data);
int hxm_particle_get(hxm_particle* p, const char* name, void* host_data);
/* raw device pointer of an array ("position", "velocity", "image", "force", "id", "reverse_id",
 * "potential_energy", "potential_stress_tensor") for zero-copy interop */
void* hxm_particle_device_ptr(hxm_particle* p, const char* name);
/* particle:aux_enable() (particle.cpp:162-167) */
int hxm_particle_aux_enable(hxm_particle* p);

/* ---- potential ------------------------------------------------------------------------------------------------ */
int hxm_potential_lj_create(unsigned ntype1, unsigned ntype2, const float* epsilon, const float* sigma,
                            int truncation, const float* cutoff_sigma, float h, hxm_potential** out);
/* potential:modify({"hard_core", radius = core_sigma}):truncate{...} (adapters/modify.lua.in, adapters/hard_core.hpp) */
int hxm_potential_lj_hard_core_create(unsigned ntype1, unsigned ntype2, const float* epsilon, const float* sigma,
                                      const float* core_sigma, int truncation, const float* cutoff_sigma, float h,
                                      hxm_potential** out);
void hxm_potential_destroy(hxm_potential* p);
/* r_cut matrix (cutoff * sigma), row-major ntype1 x ntype2 */
int hxm_potential_r_cut(const hxm_potential* p, float* r_cut_out);
const hx_lj_params* hxm_potential_params(const hxm_potential* p);

/* ---- mdsim.binning ---------------------------------------------------------------------------------------------- */
int hxm_binning_create(hxm_particle* particle, const double box_length[3], const float* r_cut, unsigned ntype1,
                       unsigned ntype2, double skin, double occupancy, hxm_binning** out);
void hxm_binning_destroy(hxm_binning* b);
int hxm_binning_ncell(const hxm_binning* b, unsigned ncell[3]);
unsigned hxm_binning_cell_size(const hxm_binning* b);
/* binning:g_cell() — updates if the positions changed; copies ncells*cell_size uint32 to host when out != NULL */
int hxm_binning_update(hxm_binning* b, uint32_t* g_cell_host_out);

/* ---- mdsim.neighbour (from_binning + max_displacement + sorts.hilbert) ------------------------------------------- */
/* algorithm/layout: HX_LAYOUT_COLUMN (reference default) or HX_LAYOUT_ROW (unroll_force_loop layout);
 * occupancy <= 0 selects the reference default 0.4 (from_binning.cpp:284-287); disable_sorting as in neighbour.lua.in:162 */
int hxm_neighbour_create(hxm_particle* particle, const double box_length[3], const float* r_cut, unsigned ntype1,
                         unsigned ntype2, double skin, double occupancy, int layout, int disable_sorting,
                         hxm_binning* binning /* may be NULL: constructed with the same occupancy rule as Lua */,
                         hxm_neighbour** out);
void hxm_neighbour_destroy(hxm_neighbour* nb);
unsigned hxm_neighbour_size(const hxm_neighbour* nb);
size_t hxm_neighbour_stride(const hxm_neighbour* nb);
int hxm_neighbour_layout(const hxm_neighbour* nb);
/* neighbour:g_neighbour() semantics: rebuild iff positions moved > skin/2 or particles were reordered; returns 1 in
 * *rebuilt when an update ran */
int hxm_neighbour_check(hxm_neighbour* nb, int* rebuilt);
/* private shell-ordered rows for the pair-force kernel (halmd_b200.h, hx_shell_rows_build; default on for the ROW
 * layout): enable = 0 makes the force module walk the public lists; a change takes effect with the next list update */
int hxm_neighbour_set_shell_rows(hxm_neighbour* nb, int enable);
/* statistics: number of list updates so far; mean list length of the last update */
size_t hxm_neighbour_updates(const hxm_neighbour* nb);
double hxm_neighbour_mean_length(hxm_neighbour* nb);
/* copy lists to host (size*stride or n*stride uint32) and per-particle counts (n uint32); either may be NULL */
int hxm_neighbour_get(hxm_neighbour* nb, uint32_t* list_host, uint32_t* count_host);
const void* hxm_neighbour_device_ptr(const hxm_neighbour* nb);
/* further arrays of the module: "g_neighbour", "count" (uint32[n]), "r0" (float4[n], positions at the last update),
 * "ctl" (the 4-word control block of halmd_b200.h) as device pointers; "shell_rows" as a HOST pointer to the
 * hx_shell_rows descriptor of the private rows (NULL while they are not in use) */
const void* hxm_neighbour_named_ptr(const hxm_neighbour* nb, const char* name);
/* synchronising read of the control block (ctl[0] = float bits of the largest squared displacement since the last update) */
int hxm_neighbour_ctl_read(hxm_neighbour* nb, uint32_t ctl_out[4]);

/* ---- mdsim.forces.pair (pair_trunc) ------------------------------------------------------------------------------- */
int hxm_force_pair_trunc_create(hxm_potential* potential, hxm_particle* particle, const double box_length[3],
                                hxm_neighbour* neighbour, float aux_weight, hxm_force** out);
void hxm_force_destroy(hxm_force* f);
/* force:apply() through particle->force() laziness: recomputes only if positions changed (pair_trunc.hpp:175-215) */
int hxm_force_apply(hxm_force* f);
/* mdsim.forces.external{particle, box, potential = potentials.external.planar_wall{...}} (external.lua.in,
 * halmd/mdsim/gpu/forces/external.hpp:120-190): a further contribution to the particle's force array, evaluated
 * after the pair force of `f` with force_zero = false (particle.cpp:221-253). */
int hxm_force_add_planar_wall(hxm_force* f, const hx_planar_wall* wall);

/* ---- integrators ---------------------------------------------------------------------------------------------------- */
int hxm_integrator_verlet_create(hxm_particle* particle, hxm_force* force, const double box_length[3], double timestep,
                                 hxm_integrator** out);
int hxm_integrator_andersen_create(hxm_particle* particle, hxm_force* force, const double box_length[3],
                                   double timestep, double temperature, double rate, unsigned seed,
                                   hxm_integrator** out);
/* mdsim.integrators.verlet_nvt_hoover{box, particle, timestep, temperature, resonance_frequency}
 * (halmd/mdsim/gpu/integrators/verlet_nvt_hoover.cpp:40-222): Nose-Hoover chain of length 2; the chain is propagated on
 * the host twice per step.  hoover_state: (xi[0], xi[1], v_xi[0], v_xi[1], mass_xi[0], mass_xi[1], en_nhc per particle). */
int hxm_integrator_hoover_create(hxm_particle* particle, hxm_force* force, const double box_length[3], double timestep,
                                 double temperature, double resonance_frequency, hxm_integrator** out);
int hxm_integrator_hoover_state(const hxm_integrator* i, double out[7]);
int hxm_integrator_hoover_set_mass(hxm_integrator* i, const double mass[2]);
void hxm_integrator_destroy(hxm_integrator* i);
int hxm_integrator_integrate(hxm_integrator* i);
int hxm_integrator_finalize(hxm_integrator* i);
int hxm_integrator_set_timestep(hxm_integrator* i, double timestep);
int hxm_integrator_set_temperature(hxm_integrator* i, double temperature);

/* ---- core / sampler ----------------------------------------------------------------------------------------------------- */
int hxm_core_create(hxm_integrator* integrator, hxm_core** out);
void hxm_core_destroy(hxm_core* c);
/* one step exactly as core::mdstep (core.cpp:30-63): integrate, (lazy force), finalize; host-polled rebuild decision */
int hxm_core_mdstep(hxm_core* c);
/* sampler:run(steps): device-resident rebuild decision, chunked speculative enqueue (see halmd_b200.h) */
int hxm_core_run(hxm_core* c, uint64_t steps);
uint64_t hxm_core_step(const hxm_core* c);
/* chunk length bounds for hxm_core_run (default 2..64, adaptive) */
int hxm_core_set_chunk(hxm_core* c, unsigned min_chunk, unsigned max_chunk);
/* NVE with private shell rows: integrate in the tail of the pair-force kernel (hx_pair_force_shells_verlet): one launch
 * per step instead of two.  Results are bit-identical either way; default off (measured 3.6 % slower at config 2 on B200:
 * the pair-force kernel is issue-bound and the stand-alone integrator overlaps its tail). */
int hxm_core_set_fused_verlet(hxm_core* c, int enable);
/* statistics of the last run: kernels launched, chunks, host synchronisations */
int hxm_core_stats(const hxm_core* c, uint64_t* launches, uint64_t* chunks, uint64_t* syncs);
/* observables.thermodynamics: out[0]=E_kin/N, [1]=E_pot/N, [2]=virial/N, [3]=temperature, [4]=pressure,
 * [5]=E_tot/N, [6..8]=v_cm, [9]=density  (observables/thermodynamics.hpp:73-86); forces are re-evaluated with
 * auxiliary variables if needed (particle.cpp:221-253) */
int hxm_core_thermodynamics(hxm_core* c, double out[10]);
/* observables.density_mode:acquire for all particles (halmd/observables/gpu/density_mode.cpp:70-131):
 * rho_host[2*k + {0,1}] = (sum cos q_k.r, -sum sin q_k.r); wavevector_host: float[3 * nq] */
int hxm_density_modes(hxm_particle* p, const float* wavevector_host, unsigned nq, double* rho_host);

/* ---- observables.dynamics.blocking_scheme + correlations ------------------------------------------------------------------ */
/* blocking_scheme(clock, max_lag, resolution, block_size, shift, separation) with timestep from the integrator;
 * correlations: mean_square_displacement (always) and intermediate_scattering_function for nq wavevectors grouped in
 * nshell shells (shell_offsets[nshell+1], norm = number of particles) when nq > 0. */
int hxm_blocking_create(hxm_core* core, double max_lag_time, double resolution, unsigned block_size, unsigned shift,
                        unsigned separation, const float* wavevectors, unsigned nq, const unsigned* shell_offsets,
                        unsigned nshell, hxm_blocking** out);
void hxm_blocking_destroy(hxm_blocking* b);
unsigned hxm_blocking_levels(const hxm_blocking* b);
unsigned hxm_blocking_block_size(const hxm_blocking* b);
int hxm_blocking_intervals(const hxm_blocking* b, uint64_t* intervals_out);
/* blocking_scheme:sample() at the current step, :finalise() at the end */
int hxm_blocking_sample(hxm_blocking* b);
int hxm_blocking_finalise(hxm_blocking* b);
/* results: (count, mean, M2) per [level][lag] for MSD; per [level][lag][shell] for ISF */
int hxm_blocking_msd(hxm_blocking* b, double* out);
int hxm_blocking_isf(hxm_blocking* b, double* out);
/* observables.dynamics.mean_quartic_displacement / velocity_autocorrelation on the same blocking scheme:
 * out[(level * block_size + lag) * 3 + {0,1,2}] = (count, mean, M2) like hxm_blocking_msd */
int hxm_blocking_mqd(hxm_blocking* b, double* out);
int hxm_blocking_vacf(hxm_blocking* b, double* out);
/* diagnostics for parity tests: with enable != 0 every correlated sample pair is recorded as (level, step of the first
 * sample of the block, step of the later sample) in the order correlation<>::compute visits them
 * (correlation.hpp:208-224); events_out (3 * capacity words) may be NULL to query *nevents_out only.  enable == 0 stops and clears. */
int hxm_blocking_trace(hxm_blocking* b, int enable, uint64_t* events_out, size_t capacity, size_t* nevents_out);
/* run `steps` steps sampling the blocking scheme at every step that a level is due (sampler on_sample slot) */
int hxm_core_run_sampling(hxm_core* c, uint64_t steps, hxm_blocking* b);

#ifdef __cplusplus
}
#endif
#endif /* HALMD_B200_MODULES_H */
