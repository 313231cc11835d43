/*
 * halmd_b200.h — C ABI of the Blackwell-native MD step loop (kernel level).
 *
 * Every entry point replaces one call site where a HALMD host class invokes its
 * `*_wrapper::kernel.<fn>` cuda::function (layer L3→L2 of the reference).  The
 * reference interface that each function stands in for is cited as
 * `file:line` relative to the halmd-org/halmd source tree.
 *
 * Conventions
 *  - plain C types only; all pointers marked `_dev` are device pointers owned by
 *    the caller; `stream` is a `cudaStream_t` passed as `void*` (NULL = default);
 *  - every call is asynchronous on `stream`; there are no hidden device syncs;
 *  - return value: 0 = ok, >0 = cudaError_t of the failed launch/runtime call,
 *    <0 = contract violation (HX_E_*); `hx_last_error_string()` gives the text;
 *  - overflow of cell / neighbour placeholders is reported through device-side
 *    int flags (0 = ok, 1 = overcrowded), never handled behind the caller's back
 *    (retry contract of halmd/mdsim/gpu/binning.cpp:173-220 and
 *    halmd/mdsim/gpu/neighbours/from_binning.cpp:185-281);
 *  - particle state uses HALMD's packing (halmd/mdsim/gpu/particle.cpp:56-160,
 *    halmd/utility/gpu/dsfloat_cuda_vector.hpp:30-100): `position` and
 *    `velocity` are float4[2*A] — high words (x,y,z,tag) in [0,A), low words
 *    (x,y,z,0) in [A,2A); tag = species bits / mass; slots >= n are ghosts with
 *    species 0xFFFFFFFF.
 */
#ifndef HALMD_B200_H
#define HALMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HX_PLACEHOLDER 0xFFFFFFFFu

/* error codes (<0: contract violations) */
#define HX_OK 0
#define HX_E_INVALID (-1)   /* invalid argument */
#define HX_E_UNSUPPORTED (-2)
#define HX_E_NOGPU (-3)     /* no CUDA device / extension unusable: never falls back to CPU */

/* neighbour-list memory layouts (halmd/mdsim/gpu/neighbours/from_binning_kernel.cu:145-149) */
#define HX_LAYOUT_COLUMN 0  /* g_neighbour[k * stride + i]  (reference default)          */
#define HX_LAYOUT_ROW    1  /* g_neighbour[i * stride + k]  (reference "unroll_force_loop",
                               stride = row pitch >= size)                               */
/* OR-ed into the `layout` argument of the pair-force entry points: the caller vouches that every position passed
 * (position1 and position2) lies in the primary cell [-L/2, L/2], as left by the integrator's periodic reduction
 * (verlet_kernel.cu:60-62, box_kernel.cuh:37-45).  Particles farther than r_cut from all faces then skip the
 * minimum-image arithmetic (same results: their pairs within the cutoff have image 0). */
#define HX_FORCE_POSITIONS_REDUCED 0x100
/* further per-call bits of `layout` for the ROW layout (tuning; defaults = packed f32x2 arithmetic, 4 lanes):
 * HX_FORCE_SCALAR selects the scalar instruction stream, HX_FORCE_LANES(l) l = 2, 4 or 8 lanes cooperating on one
 * particle (the reference's unroll_force_loop variant uses a whole warp, pair_trunc_kernel.cuh:54-154) */
#define HX_FORCE_SCALAR 0x200
#define HX_FORCE_LANES(l) ((l) << 12)

/* potential truncations (halmd/mdsim/gpu/potentials/pair/truncations/) */
#define HX_TRUNC_SHARP     0
#define HX_TRUNC_SHIFTED   1
#define HX_TRUNC_SMOOTH_R4 2
#define HX_TRUNC_FORCE_SHIFTED 3   /* truncations/force_shifted_kernel.cuh:52-63 */

#define HX_MAX_PAIR_TYPES 16 /* ntype1*ntype2 <= 16 */

typedef struct hx_float4 { float x, y, z, w; } hx_float4;

/* ---- library -------------------------------------------------------------------------- */
int hx_version(void);
const char* hx_last_error_string(void);
/* number of CUDA devices visible, or HX_E_NOGPU */
int hx_device_count(void);
/* kernels launched by this library since load (all streams); used for `gpu_launches` */
uint64_t hx_launch_count(void);

/* ---- a1: particle container ------------------------------------------------------------
 * halmd/mdsim/gpu/particle.cpp:69-113 — array_size = n rounded up to 1024; initial values. */
size_t hx_array_size(size_t nparticle);
int hx_state_init(hx_float4* position_dev /*2A*/, hx_float4* velocity_dev /*2A*/,
                  hx_float4* image_dev /*A*/, hx_float4* force_dev /*A*/,
                  uint32_t* id_dev /*A*/, uint32_t* reverse_id_dev /*A*/,
                  size_t nparticle, size_t array_size, void* stream);

/* ---- a2/a3: cell binning ---------------------------------------------------------------
 * halmd/mdsim/gpu/binning.cpp:47-116 (sizing), :166-221 (update),
 * halmd/mdsim/gpu/binning_kernel.cu:34-139. */
int hx_binning_plan(const double box_length[3], size_t nparticle, float r_cut_max, float skin,
                    double occupancy, unsigned ncell[3], float cell_length[3],
                    unsigned* cell_size);
/* scratch: cell_count_dev[ncells] (overwritten). cell_index_dev (n entries) may be NULL.
 * g_cell_dev[ncells*cell_size]: ascending particle index per cell, padded with placeholders. */
int hx_binning_update(const hx_float4* position_hi_dev, size_t nparticle,
                      const float cell_length[3], const unsigned ncell[3], unsigned cell_size,
                      uint32_t* g_cell_dev, uint32_t* cell_count_dev, uint32_t* cell_index_dev,
                      int* overflow_flag_dev, void* stream);

/* MUFU.RCP (rcp.approx.ftz.f32) of n <= 1024 host values, synchronously: the reference build (-prec-div=false) divides
 * by cell and box lengths as r * MUFU.RCP(l) (binning_kernel.cu:34-54, hilbert_kernel.cu:59), and so do the kernels
 * here; a host-side restatement of that arithmetic needs the device's reciprocal of those few constants. */
int hx_device_rcp(const float* x_host, float* rcp_host, unsigned n);

/* ---- a4: Hilbert space-filling-curve order ---------------------------------------------
 * halmd/mdsim/gpu/sorts/hilbert.cpp:51-54 (depth), :71-117; hilbert_kernel.cu:38-69;
 * halmd/mdsim/sorts/hilbert_kernel.hpp:70-142,205-230. */
unsigned hx_hilbert_depth(const double box_length[3]);
int hx_hilbert_keys(const hx_float4* position_hi_dev, size_t nslots, const float box_length[3],
                    unsigned depth, uint32_t* keys_dev, void* stream);
int hx_iota_u32(uint32_t* out_dev, size_t n, void* stream);
/* stable LSD radix sort of (key,value) pairs, in place (halmd/algorithm/gpu/radix_sort.hpp:94-184).
 * key_bits: number of significant key bits (rounded up to a multiple of 8, <= 32). */
size_t hx_sort_scratch_bytes(size_t n);
int hx_sort_pairs_stable_u32(uint32_t* keys_dev, uint32_t* values_dev, size_t n,
                             unsigned key_bits, void* scratch_dev, size_t scratch_bytes,
                             void* stream);

/* ---- a5: particle::rearrange -----------------------------------------------------------
 * halmd/mdsim/gpu/particle.cpp:173-218, particle_kernel.cu:45-68.
 * out[i] = in[index[i]] for i < n (identity for n <= i < A) for position (hi,lo),
 * velocity (hi,lo), image, id; reverse_id_out[id_out[i]] = i.  Force is NOT permuted. */
int hx_state_permute(const uint32_t* index_dev, size_t nparticle, size_t array_size,
                     const hx_float4* position_in_dev, hx_float4* position_out_dev,
                     const hx_float4* velocity_in_dev, hx_float4* velocity_out_dev,
                     const hx_float4* image_in_dev, hx_float4* image_out_dev,
                     const uint32_t* id_in_dev, uint32_t* id_out_dev,
                     uint32_t* reverse_id_out_dev, void* stream);

/* ---- a6: max_displacement --------------------------------------------------------------
 * halmd/mdsim/gpu/max_displacement.cpp:60-106, max_displacement_kernel.cu:36-68. */
int hx_displacement_zero(const hx_float4* position_hi_dev, hx_float4* r0_dev, size_t nparticle,
                         void* stream);
/* writes max_i |minimg(r_i - r0_i)|^2 as float bits into *max_rr_bits_dev (atomicMax on the
 * IEEE bit pattern; caller zeroes it, or pass reset=1). sqrt and the comparison with skin/2
 * are hx_displacement_exceeds(). */
int hx_displacement_max(const hx_float4* position_hi_dev, const hx_float4* r0_dev,
                        size_t nparticle, const float box_length[3], uint32_t* max_rr_bits_dev,
                        int reset, void* stream);
/* host helper: float(sqrt(rr)) > float(skin/2)  (from_binning.cpp:130-131) */
int hx_displacement_exceeds(uint32_t max_rr_bits, double skin);

/* ---- a7: Verlet neighbour lists from cells ---------------------------------------------
 * halmd/mdsim/gpu/neighbours/from_binning.cpp:96-119 (sizing), :162-282 (update),
 * from_binning_kernel.cu:99-267. */
int hx_nlist_plan(float r_cut_max, float skin, size_t nparticle2, double box_volume,
                  double occupancy, unsigned cell_size, unsigned* size);
/* rr_cut_skin[ntype1*ntype2] = (r_cut + skin)^2 as float.  layout COLUMN: list is memset to
 * placeholders first, stride = array_size.  layout ROW: stride = pitch (multiple of 32 and
 * >= size+1 unless fill_all); each row is terminated by placeholders up to the next multiple
 * of 32 (fill_all=1: up to `stride`).  count_dev (n entries, may be NULL) receives list lengths.
 * overflow flag is raised when a list reaches `size` entries (from_binning_kernel.cu:263-266). */
int hx_nlist_build(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                   const hx_float4* position_hi_dev, size_t nparticle,
                   const float* rr_cut_skin, unsigned ntype1, unsigned ntype2,
                   const float box_length[3], uint32_t* g_neighbour_dev, unsigned size,
                   size_t stride, int layout, int fill_all, uint32_t* count_dev,
                   int* overflow_flag_dev, void* stream);

/* ---- a9-a11: Lennard-Jones potential with truncation ------------------------------------
 * halmd/mdsim/gpu/potentials/pair/lennard_jones.cpp:44-70, lennard_jones_kernel.hpp:43-56,
 * truncations/smooth_r4.hpp:55-99, smooth_r4_kernel.cuh:53-76, shifted_kernel.cuh:49-57.
 * Derived tables are computed in float with the reference's formulas. */
typedef struct hx_lj_params {
    unsigned ntype1, ntype2;
    int truncation;            /* HX_TRUNC_* */
    float rri_smooth;          /* h^-2 (smooth_r4) */
    float epsilon[HX_MAX_PAIR_TYPES];
    float sigma2[HX_MAX_PAIR_TYPES];
    float r_cut[HX_MAX_PAIR_TYPES];   /* cutoff * sigma */
    float rr_cut[HX_MAX_PAIR_TYPES];
    float en_cut[HX_MAX_PAIR_TYPES];  /* U(r_cut), 0 for sharp */
    float force_cut[HX_MAX_PAIR_TYPES]; /* |F(r_cut)| = fval(r_cut) * r_cut (force_shifted.hpp:69-70), else 0 */
    int hard_core;             /* adapters::hard_core<lennard_jones> (adapters/hard_core.hpp:46-100) */
    float r_core[HX_MAX_PAIR_TYPES];  /* core * sigma; U(r) = U_LJ(r - r_core) */
} hx_lj_params;

int hx_potential_lj_init(hx_lj_params* p, unsigned ntype1, unsigned ntype2, const float* epsilon,
                         const float* sigma, int truncation, const float* cutoff_sigma, float h);
/* potential:modify({"hard_core", radius = core_sigma}) before truncate (lennard_jones.lua.in:120-130):
 * core_sigma[ntype1 * ntype2] in units of sigma (NULL: plain potential); en_cut / force_cut are those of the
 * adapted potential at r_cut (shifted.hpp / force_shifted.hpp evaluate potential_type::operator()). */
int hx_potential_lj_init_hard_core(hx_lj_params* p, unsigned ntype1, unsigned ntype2, const float* epsilon,
                                   const float* sigma, const float* core_sigma, int truncation,
                                   const float* cutoff_sigma, float h);

/* ---- a8: truncated pair force ------------------------------------------------------------
 * halmd/mdsim/gpu/forces/pair_trunc.hpp:193-337, pair_trunc_kernel.cuh:54-255.
 * en_pot_dev / stress_dev NULL <=> compute_ (no aux); stress is SoA [6][stress_stride]
 * (xx,yy,zz,xy,xz,yz; halmd/mdsim/force_kernel.hpp:36-95).  aux_weight is the effective weight
 * (already halved for particle1 == particle2, pair_trunc.hpp:291-294). */
int hx_pair_force(const hx_lj_params* potential, const hx_float4* position1_hi_dev,
                  const hx_float4* position2_hi_dev, const uint32_t* g_neighbour_dev,
                  unsigned size, size_t stride, int layout, size_t nparticle,
                  const float box_length[3], int force_zero, float aux_weight,
                  hx_float4* force_dev, float* en_pot_dev, float* stress_dev,
                  size_t stress_stride, void* stream);

/* ---- a8, private list format: rows ordered by shells of the pair distance at build time ---------------------------
 * The reference walks every stored neighbour and discards those outside r_cut (pair_trunc_kernel.cuh:205-224: 94.5
 * stored, 55.3 inside at config 2).  hx_shell_rows_build copies a ROW-layout list (which stays untouched and
 * bit-identical to the reference's) into private rows in which the entries of a particle are grouped by
 * r_build - r_cut_ab in (delta[s-1], delta[s]] (stable within a shell), records the end of every shell per particle,
 * and fills the rest of each row with the particle's own index.  hx_pair_force_shells walks, for particle i, only the
 * shells with delta[s-1] < |r_i - r0_i| + sqrt(ctl[0]) (+ rounding margin): by the triangle inequality no other entry can
 * lie inside the cutoff, so forces equal those of hx_pair_force up to the order of the compensated summation.
 * r0_dev (positions at the list build) and ctl_dev (ctl[0] = running maximum squared displacement of ALL particles
 * whose indices appear in the rows, ghosts included) may be NULL: the whole row is walked then.  ctl_dev also guards
 * the launch as in hx_pair_force_guarded.  No aux variables (use hx_pair_force for en_pot / stress) and no hard_core.
 * Layout of rows_dev: tiles of 8 consecutive particles (the particles of one warp, 4 lanes each), rows interleaved in
 * granules of 8 entries so that a warp reads 256 contiguous bytes per half-step: entry k of particle i is word
 *   (i / 8) * 8 * pitch + (k / 8) * 64 + (i % 8) * 8 + slot(k % 8),   slot(e) = e < 4 ? 2 e : 2 (e - 4) + 1
 * (the 64-bit word of a lane holds entries e and e + 4: one gather request covers 4 consecutive entries per particle).
 * rows_dev: hx_shell_rows_words(nparticle, pitch) words (tiles + slack read ahead by the kernel; build initialises all);
 * ends_dev: nparticle * HX_NSHELL uint16, 16-byte aligned; pitch: multiple of 32, > longest list, <= 512. */
#define HX_NSHELL 8
typedef struct hx_shell_rows {
    uint32_t* rows_dev;
    uint16_t* ends_dev;
    unsigned pitch;
    float delta[HX_NSHELL];   /* ascending, delta[HX_NSHELL-1] >= skin */
} hx_shell_rows;
/* delta[s] = skin * (s + 1) / HX_NSHELL */
int hx_shell_rows_plan(float skin, hx_shell_rows* shells);
size_t hx_shell_rows_words(size_t nparticle, unsigned pitch);
/* largest entry of count_dev[0..n) into *max_dev (the caller picks the pitch from it) */
int hx_count_max(const uint32_t* count_dev, size_t n, uint32_t* max_dev, void* stream);
int hx_shell_rows_build(const hx_lj_params* potential, const uint32_t* g_neighbour_dev, size_t stride,
                        const uint32_t* count_dev, const hx_float4* position1_hi_dev,
                        const hx_float4* position2_hi_dev, size_t nparticle, const float box_length[3],
                        const hx_shell_rows* shells, void* stream);
int hx_pair_force_shells(const hx_lj_params* potential, const hx_float4* position1_hi_dev,
                         const hx_float4* position2_hi_dev, const hx_shell_rows* shells, const hx_float4* r0_dev,
                         size_t nparticle, const float box_length[3], int flags /* HX_FORCE_POSITIONS_REDUCED */, int force_zero,
                         hx_float4* force_dev, uint32_t* ctl_dev, float skin_half, void* stream);

/* The same kernel with the velocity-Verlet update of the particle fused into its tail, for the device-resident step
 * loop: after the force on particle i is summed, the second half-kick of this step (verlet_kernel.cu:73-98) and the first
 * half-step of the next one (verlet_kernel.cu:37-68: half-kick, drift, periodic wrap + image) are applied to i with the
 * arithmetic of hx_verlet_finalize_integrate, and its squared displacement from r0 (max_displacement_kernel.cu:36-68)
 * is accumulated.  The new positions go to position_out_dev (hi plane, lo plane at + array_size), which must not be the
 * buffer the kernel gathers from; velocities and images are updated in place; the force array is not written.
 * The control block (HX_CTL_WORDS words) holds two slots of the running maximum squared displacement, words 0 and 4: the
 * launch tests slot max_rr_word_in like hx_pair_force_guarded and accumulates into the other slot (carrying the running
 * maximum over), so that blocks which start late never see displacements of this launch.  position_out_dev == NULL: plain
 * guarded force kernel (forces to force_dev) that tests slot max_rr_word_in — the last launch of a chunk. */
#define HX_CTL_WORDS 8
typedef struct hx_verlet_tail {
    hx_float4* position_out_dev;
    hx_float4* velocity_dev;
    hx_float4* image_dev;
    size_t array_size;
    float timestep;
    unsigned max_rr_word_in;    /* 0 or 4 */
} hx_verlet_tail;
int hx_pair_force_shells_verlet(const hx_lj_params* potential, const hx_float4* position_dev, const hx_shell_rows* shells,
                                const hx_float4* r0_dev, size_t nparticle, const float box_length[3], int flags,
                                hx_float4* force_dev, uint32_t* ctl_dev, float skin_half, const hx_verlet_tail* tail,
                                void* stream);

/* ---- f1: external one-body force, planar Lennard-Jones 9-3 walls -----------------------------------------------
 * halmd/mdsim/gpu/forces/external.hpp:120-190, external_kernel.cuh:49-103,
 * halmd/mdsim/gpu/potentials/external/planar_wall.cpp:40-112, planar_wall_kernel.cu:43-105.
 * Parameter tables are indexed [species * nwall + wall] like the reference's parameter texture; surface normals are
 * normalised by hx_planar_wall_init (planar_wall.cpp:83-86).  The kernel adds to force / en_pot when force_zero == 0
 * (external_kernel.cuh:78-99: the stress contribution of the wall is not implemented in the reference — zero).
 * ctl_dev (may be NULL): returns at once when the control block is halted (chunked stepping). */
#define HX_MAX_WALLS 8
#define HX_MAX_WALL_SPECIES 4
typedef struct hx_planar_wall {
    unsigned nwall, nspecies;
    float smoothing;
    float surface_normal[HX_MAX_WALLS][3];
    float offset[HX_MAX_WALLS];
    float epsilon[HX_MAX_WALL_SPECIES * HX_MAX_WALLS];
    float sigma[HX_MAX_WALL_SPECIES * HX_MAX_WALLS];
    float wetting[HX_MAX_WALL_SPECIES * HX_MAX_WALLS];
    float cutoff[HX_MAX_WALL_SPECIES * HX_MAX_WALLS];
} hx_planar_wall;
/* epsilon, sigma, wetting, cutoff: [nwall][nspecies] row-major as in the Lua module (planar_wall.lua.in) */
int hx_planar_wall_init(hx_planar_wall* w, unsigned nwall, unsigned nspecies, const float* offset,
                        const float* surface_normal /* [nwall][3] */, const float* epsilon, const float* sigma,
                        const float* wetting, const float* cutoff, float smoothing);
int hx_external_planar_wall(const hx_planar_wall* wall, const hx_float4* position_hi_dev, size_t nparticle,
                            const float box_length[3], int force_zero, hx_float4* force_dev, float* en_pot_dev,
                            float* stress_dev, size_t stress_stride, const uint32_t* ctl_dev, void* stream);

/* ---- a12: velocity-Verlet in double-single arithmetic -------------------------------------
 * halmd/mdsim/gpu/integrators/verlet.cpp:64-121, verlet_kernel.cu:37-98,
 * halmd/numeric/mp/dsfun.hpp:142-207. */
int hx_verlet_integrate(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                        const hx_float4* force_dev, size_t array_size, float timestep,
                        const float box_length[3], void* stream);
int hx_verlet_finalize(hx_float4* velocity_dev, const hx_float4* force_dev, size_t array_size,
                       float timestep, void* stream);
/* fused: finalize(step k) [if do_finalize] + integrate(step k+1) + max-displacement reduction
 * against r0 (bitwise identical to the separate calls).  ctl_dev is a 4-word control block (see
 * "step-loop support" below): the kernel exits at once if ctl[1] (halted) is set and folds the
 * maximum squared displacement into ctl[0].  r0_dev and ctl_dev may be NULL (r0 needs ctl). */
int hx_verlet_finalize_integrate(hx_float4* position_dev, hx_float4* image_dev,
                                 hx_float4* velocity_dev, const hx_float4* force_dev,
                                 size_t nparticle, size_t array_size, float timestep,
                                 const float box_length[3], int do_finalize,
                                 const hx_float4* r0_dev, uint32_t* ctl_dev, void* stream);
int hx_verlet_finalize_guarded(hx_float4* velocity_dev, const hx_float4* force_dev,
                               size_t array_size, float timestep, uint32_t* ctl_dev, void* stream);

/* ---- f3: Nose-Hoover chain thermostat (halmd/mdsim/gpu/integrators/verlet_nvt_hoover_kernel.cu:37-125) ------------
 * integrate: v *= scale (double-single), then the first velocity-Verlet half-step; rescale: v *= scale.  The second
 * half-step is hx_verlet_finalize.  `scale` is the factor returned by the host-side chain propagation
 * (verlet_nvt_hoover.cpp:185-222), converted to double-single like halmd::dsfloat(double). */
int hx_hoover_integrate(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                        const hx_float4* force_dev, size_t array_size, float timestep, double scale,
                        const float box_length[3], void* stream);
int hx_hoover_rescale(hx_float4* velocity_dev, size_t array_size, double scale, void* stream);
/* Host side of the thermostat (no device work): the chain of two heat-bath variables of
 * verlet_nvt_hoover.hpp:100-160.  init: set_temperature (verlet_nvt_hoover.cpp:79-95), masses
 * (dof T / omega^2, T / omega^2) with omega = 2 pi resonance_frequency, dof = dimension * nparticle, chain at rest.
 * propagate: propagate_chain (verlet_nvt_hoover.cpp:185-222) for en_kin_2 = sum m v^2 over all particles; returns the
 * velocity scale factor.  energy: en_nhc per particle (verlet_nvt_hoover.cpp:175-182). */
typedef struct hx_hoover_chain {
    double xi[2], v_xi[2], mass_xi[2];
    double temperature, en_kin_target_2, dof;
} hx_hoover_chain;
int hx_hoover_chain_init(hx_hoover_chain* chain, size_t nparticle, unsigned dimension, double temperature,
                         double resonance_frequency);
int hx_hoover_chain_set_temperature(hx_hoover_chain* chain, double temperature, double resonance_frequency);
double hx_hoover_chain_propagate(hx_hoover_chain* chain, double en_kin_2, double timestep);
double hx_hoover_chain_energy(const hx_hoover_chain* chain, size_t nparticle);

/* ---- a13: Andersen thermostat with rand48 ---------------------------------------------------
 * halmd/random/gpu/rand48.hpp:57-81, rand48_kernel.cu:42-100, rand48_kernel.cuh:40-79,
 * halmd/mdsim/gpu/integrators/verlet_nvt_andersen_kernel.cu:96-171. */
typedef struct hx_rand48 {
    unsigned blocks, threads;      /* 32 x 256 by default (random/gpu/random.cpp:94-101) */
    uint16_t a[3], c[3];           /* leapfrog multiplier / addend (x = low word) */
} hx_rand48;
/* state_dev: uint16_t[3 * blocks*threads] (ushort3 per RNG thread) */
int hx_rand48_seed(hx_rand48* rng, unsigned blocks, unsigned threads, unsigned seed,
                   uint16_t* state_dev, void* stream);
int hx_rand48_uniform(const hx_rand48* rng, uint16_t* state_dev, float* out_dev, size_t n,
                      void* stream);
int hx_andersen_finalize(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle,
                         size_t array_size, float timestep, float sqrt_temperature,
                         float coll_prob, const hx_rand48* rng, uint16_t* state_dev,
                         void* stream);

/* The same thermostat step in two passes (generator bookkeeping on the RNG grid, then one thread per particle): the
 * same generator stream and bit-identical velocities at a fraction of the time for large systems.  scratch_dev:
 * hx_andersen_scratch_bytes(nparticle) bytes; ctl_dev may be NULL (see hx_andersen_finalize_guarded). */
size_t hx_andersen_scratch_bytes(size_t nparticle);
int hx_andersen_finalize2(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle,
                          size_t array_size, float timestep, float sqrt_temperature, float coll_prob,
                          const hx_rand48* rng, uint16_t* state_dev, void* scratch_dev, size_t scratch_bytes,
                          const uint32_t* ctl_dev, void* stream);

/* ---- a14-a16: observables -------------------------------------------------------------------
 * halmd/observables/gpu/phase_space_kernel.cu:39-66 (unfolded positions in id order),
 * halmd/observables/gpu/dynamics/mean_square_displacement.cpp:55-68 + tagged_particle.hpp,
 * halmd/observables/gpu/density_mode_kernel.cu:46-108. */
int hx_sample_positions(const hx_float4* position_hi_dev, const hx_float4* image_dev,
                        const uint32_t* reverse_id_dev, size_t nparticle,
                        const float box_length[3], hx_float4* sample_dev, void* stream);
int hx_sample_velocities(const hx_float4* velocity_hi_dev, const uint32_t* reverse_id_dev,
                         size_t nparticle, hx_float4* sample_dev, void* stream);
/* correlate `first` with nlags samples: out_dev[3*l + {0,1,2}] = (count, mean, M2) of
 * |R_l - R_first|^2 over particles, as doubles (mergeable into accumulator<double>,
 * halmd/numeric/accumulator.hpp:104-118).  samples_dev: device array of nlags device pointers. */
size_t hx_msd_scratch_bytes(unsigned nlags);
int hx_msd_block(const hx_float4* first_dev, const hx_float4* const* samples_dev, unsigned nlags,
                 size_t nparticle, double* out_dev, void* scratch_dev, size_t scratch_bytes,
                 void* stream);
/* the same accumulation for the other tagged-particle correlation functions of halmd/observables/dynamics:
 * HX_CORR_MSD |dr|^2, HX_CORR_MQD |dr|^4 (mean_quartic_displacement.hpp:36-44), HX_CORR_VACF v1.v2 on velocity
 * samples (velocity_autocorrelation.hpp:36-39); scratch as for hx_msd_block */
#define HX_CORR_MSD 0
#define HX_CORR_MQD 1
#define HX_CORR_VACF 2
int hx_correlate_block(int kind, const hx_float4* first_dev, const hx_float4* const* samples_dev, unsigned nlags,
                       size_t nparticle, double* out_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
/* rho_dev[2*q + {0,1}] = (sum cos q.r, -sum sin q.r) as doubles; q_dev: float[3*nq] */
int hx_density_modes(const hx_float4* position_hi_dev, size_t nparticle, const float* q_dev,
                     unsigned nq, double* rho_dev, void* stream);
/* observables.utility.wavevector{box, wavenumber, tolerance, max_count, filter} (halmd/observables/utility/wavevector.cpp:40-123,
 * halmd/algorithm/host/pick_lattice_points.hpp:62-160), host side, no device work: wavevectors n (h, k, l) 2 pi / L with
 * coprime (h, k, l) whose magnitude matches wavenumber[i] within the relative tolerance, at most max_count per wavenumber,
 * preferring small h + k + l.  vectors_out: 3 * capacity doubles, grouped by wavenumber; shell_offsets_out: nwavenumber + 1
 * offsets (an empty shell marks a wavenumber the caller must drop).  filter (3 entries 0 / 1) may be NULL. */
int hx_wavevector_pick(const double* wavenumber, unsigned nwavenumber, const double box_length[3], double tolerance,
                       unsigned max_count, const unsigned* filter, double* vectors_out, unsigned capacity,
                       unsigned* shell_offsets_out);

/* intermediate_scattering_function::operator() (halmd/observables/dynamics/intermediate_scattering_function.hpp:90-113), host
 * side, no device work: for every wavevector shell s (wavevectors shell_offsets[s] .. shell_offsets[s + 1]) the values
 * Re[rho1 conj(rho2)] / norm are accumulated (numeric/accumulator.hpp:56-132) and the shell's accumulator is merged into
 * acc[3 * s + {0,1,2}] = (count, mean, M2), in place.  rho: (re, im) per wavevector. */
int hx_isf_accumulate(const double* rho1, const double* rho2, const unsigned* shell_offsets, unsigned nshell, double norm,
                      double* acc);

/* thermodynamic sums over particles i < n, as doubles:
 * out[0]=sum m v^2, out[1..3]=sum m v, out[4]=sum m, out[5]=sum en_pot, out[6]=sum virial trace
 * (halmd/observables/gpu/thermodynamics_kernel.cu:35-97).  en_pot/stress may be NULL.
 * out_dev holds HX_THERMO_DOUBLES doubles: the 8 result words, then scratch for the block partials, which are
 * combined in block order (results do not depend on the scheduling of blocks). */
#define HX_THERMO_BLOCKS 296
#define HX_THERMO_DOUBLES (8 + 8 * HX_THERMO_BLOCKS)
int hx_thermodynamics(const hx_float4* velocity_hi_dev, const float* en_pot_dev,
                      const float* stress_dev, size_t stress_stride, size_t nparticle,
                      double* out_dev /* HX_THERMO_DOUBLES */, void* stream);

/* ---- step-loop support (device-resident rebuild decision) ------------------------------------
 * The reference copies 64 floats to the host every step to decide whether the neighbour lists
 * must be rebuilt (max_displacement.cpp:77-106).  Here the decision stays on the device: a
 * 4-word control block
 *   ctl[0] max squared displacement since the last rebuild (float bits, atomicMax),
 *   ctl[1] halted flag, ctl[2] steps completed in this chunk, ctl[3] reserved
 * lets the host enqueue a chunk of steps speculatively.  The guarded force kernel evaluates
 * float(sqrt(ctl[0])) > skin/2 (from_binning.cpp:130-131); when it trips it sets ctl[1] and all
 * remaining guarded kernels of the chunk return immediately.  The host then rebuilds the lists,
 * completes the interrupted step and enqueues the next chunk ("chunked stepping"). */
int hx_ctl_reset(uint32_t* ctl_dev, void* stream);
int hx_pair_force_guarded(const hx_lj_params* potential, const hx_float4* position1_hi_dev,
                          const hx_float4* position2_hi_dev, const uint32_t* g_neighbour_dev,
                          unsigned size, size_t stride, int layout, size_t nparticle,
                          const float box_length[3], int force_zero, float aux_weight,
                          hx_float4* force_dev, float* en_pot_dev, float* stress_dev,
                          size_t stress_stride, uint32_t* ctl_dev, float skin_half, void* stream);
int hx_andersen_finalize_guarded(hx_float4* velocity_dev, const hx_float4* force_dev,
                                 size_t nparticle, size_t array_size, float timestep,
                                 float sqrt_temperature, float coll_prob, const hx_rand48* rng,
                                 uint16_t* state_dev, const uint32_t* ctl_dev, void* stream);

/* ---- range-limited variants (arrays that hold ghost particles behind the owned ones) ------------------------------
 * Same kernels as hx_nlist_build / hx_pair_force_guarded / hx_verlet_finalize_integrate with explicit counts:
 * lists are built for particles < nbuild only (the others are partners), partner indices are valid below nparticle2,
 * and the integrator touches slots < nprocess only (array_size stays the offset of the low-word planes).
 * ctl_dev may be NULL in hx_pair_force2 (unguarded). */
int hx_nlist_build2(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size,
                    const hx_float4* position_hi_dev, size_t nparticle, size_t nbuild, const float* rr_cut_skin,
                    unsigned ntype1, unsigned ntype2, const float box_length[3], uint32_t* g_neighbour_dev,
                    unsigned size, size_t stride, int layout, int fill_all, uint32_t* count_dev,
                    int* overflow_flag_dev, void* stream);
/* Slab-decomposed ranks hold particles in a few layers of the (global) cell grid only: cell_nx layers in x starting
 * at layer cell_x0 (periodic).  These variants place placeholders, sort cells and launch list blocks for those layers
 * only — work per rank proportional to its own particles, not to the global grid.  A particle outside the layers sets
 * *overflow_flag_dev to HX_BINNING_OUT_OF_RANGE in hx_binning_update_slab (the caller's layer arithmetic is wrong). */
#define HX_BINNING_OUT_OF_RANGE 2
int hx_binning_update_slab(const hx_float4* position_hi_dev, size_t nparticle, const float cell_length[3],
                           const unsigned ncell[3], unsigned cell_size, unsigned cell_x0, unsigned cell_nx,
                           uint32_t* g_cell_dev, uint32_t* cell_count_dev, uint32_t* cell_index_dev,
                           int* overflow_flag_dev, void* stream);
int hx_nlist_build_slab(const uint32_t* g_cell_dev, const unsigned ncell[3], unsigned cell_size, unsigned cell_x0,
                        unsigned cell_nx, const hx_float4* position_hi_dev, size_t nparticle, size_t nbuild,
                        const float* rr_cut_skin, unsigned ntype1, unsigned ntype2, const float box_length[3],
                        uint32_t* g_neighbour_dev, unsigned size, size_t stride, int layout, int fill_all,
                        uint32_t* count_dev, int* overflow_flag_dev, void* stream);
int hx_pair_force2(const hx_lj_params* potential, const hx_float4* position1_hi_dev,
                   const hx_float4* position2_hi_dev, size_t nparticle2, const uint32_t* g_neighbour_dev,
                   unsigned size, size_t stride, int layout, size_t nparticle, const float box_length[3],
                   int force_zero, float aux_weight, hx_float4* force_dev, float* en_pot_dev, float* stress_dev,
                   size_t stress_stride, uint32_t* ctl_dev, float skin_half, void* stream);
int hx_verlet_finalize_integrate_range(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                                       const hx_float4* force_dev, size_t nparticle, size_t nprocess,
                                       size_t array_size, float timestep, const float box_length[3],
                                       int do_finalize, int do_integrate, const hx_float4* r0_dev,
                                       uint32_t* ctl_dev, void* stream);

/* ---- slab domain decomposition (new; the reference is single-GPU, doc/usage/multi-gpu.rst.in) ---------------------
 * Rank r of nranks owns wrapped x in [x0 + r*width, x0 + (r+1)*width).  Selections are produced as small integer
 * keys for hx_sort_pairs_stable_u32 so that migrants and ghosts keep their memory order (deterministic).
 * hx_dd_owner_keys: 0 stays, 1 to rank-1, 2 to rank+1 (1 for both when nranks == 2), 3 lost; counts_dev[4] histogram.
 * hx_dd_ghost_keys: 0 if within `halo` of the low (side 0), high (side 1) or either (side 2) face, else 1; the distances to the
 * faces are taken modulo box_length_x (a coordinate of +L/2 belongs to the slab that starts at -L/2).
 * hx_dd_permute: out[i] = in[index[i]], i < nparticle, for position/velocity (hi and lo planes), image and id. */
int hx_dd_owner_keys(const hx_float4* position_hi_dev, size_t nparticle, float x0, float width, int nranks, int rank,
                     uint32_t* keys_dev, uint32_t* counts_dev, void* stream);
int hx_dd_ghost_keys(const hx_float4* position_hi_dev, size_t nparticle, float lo, float hi, float halo, int side,
                     float box_length_x, uint32_t* keys_dev, uint32_t* count_dev, void* stream);
int hx_dd_permute(const uint32_t* index_dev, size_t nparticle, size_t array_size, const hx_float4* position_in_dev,
                  hx_float4* position_out_dev, const hx_float4* velocity_in_dev, hx_float4* velocity_out_dev,
                  const hx_float4* image_in_dev, hx_float4* image_out_dev, const uint32_t* id_in_dev,
                  uint32_t* id_out_dev, void* stream);
/* dst[i] = src[index[i]]; the float4 variant returns at once when ctl_dev (may be NULL) is halted */
int hx_gather_float4(const hx_float4* src_dev, const uint32_t* index_dev, size_t n, hx_float4* dst_dev,
                     const uint32_t* ctl_dev, void* stream);
int hx_gather_u32(const uint32_t* src_dev, const uint32_t* index_dev, size_t n, uint32_t* dst_dev, void* stream);


/* ---- decomposed step loop: halo exchange and rebuild vote over peer memory (NVLink), no host and no library collective
 * between the kernels of a step.  Position buffers and inboxes are allocated with hx_dd_alloc (plain cudaMalloc, so that
 * they can be exported), exported / mapped with the ipc functions (cudaIpc*), one process per GPU.
 * hx_dd_enqueue_steps enqueues `chunk` MD steps, each:
 *   fused Verlet kernel: position_dev[current] -> position_dev[current ^ 1] (owned particles, displacement maximum),
 *   hx_gather_float4 of the send lists into peer_ghost_dev[dir][current ^ 1] (the neighbours' ghost slots, peer memory),
 *   signal: {step, local maximum} into inbox slot [rank] of every other rank; wait: until all inbox slots of this
 *   rank carry the step, global maximum folded into ctl[0] (spins at most timeout_cycles, then halts with ctl[3] =
 *   0xDEAD), guarded pair force (private shell rows if `shells`, else the public rows) on position_dev[current ^ 1],
 *   [Andersen half-kick if rng != NULL]; current ^= 1.
 * A halted control block turns every later kernel of the chunk into a no-op on every rank at the same step (all
 * ranks see the same global maximum).  step numbers are global: step_base = number of the last completed step. */
#define HX_DD_MAX_RANKS 32
#define HX_DD_IPC_HANDLE_BYTES 64
typedef struct hx_dd_steps {
    hx_float4* position_dev[2];
    int current;
    hx_float4 *image_dev, *velocity_dev, *force_dev;
    const hx_float4* r0_dev;
    size_t n_own, n_ghost, array_size;
    float timestep;
    float box_length[3];
    int nranks, rank;
    const uint32_t* send_index_dev[2];          /* owned particles pushed to the left / right neighbour */
    size_t send_count[2];
    hx_float4* peer_ghost_dev[2][2];            /* [direction][buffer]: this rank's block in the neighbour's buffers */
    unsigned long long* const* peer_inbox_dev;  /* device array [nranks] of mapped inbox pointers */
    const unsigned long long* inbox_dev;        /* own inbox: 2 rows (step parity) of HX_DD_MAX_RANKS words, zeroed at allocation */
    long long timeout_cycles;
    uint32_t* ctl_dev;
    float skin_half;
    uint64_t step_base;
    int pending_finalize;                       /* velocities still lack the second half-kick of the last step */
    const hx_lj_params* potential;
    const hx_shell_rows* shells;                /* NULL: walk g_neighbour_dev */
    const uint32_t* g_neighbour_dev;
    unsigned size;
    size_t stride;
    const hx_rand48* rng;                       /* NULL: NVE */
    uint16_t* rng_state_dev;
    void* andersen_scratch_dev;
    size_t andersen_scratch_bytes;
    float sqrt_temperature, coll_prob;
} hx_dd_steps;
int hx_dd_alloc(size_t bytes, void** out_dev);
int hx_dd_free(void* dev);
int hx_dd_ipc_export(void* dev, unsigned char handle_out[HX_DD_IPC_HANDLE_BYTES]);
int hx_dd_ipc_open(const unsigned char handle[HX_DD_IPC_HANDLE_BYTES], void** out_dev);
int hx_dd_ipc_close(void* dev);
int hx_dd_enqueue_steps(const hx_dd_steps* steps, unsigned chunk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HALMD_B200_H */
