// Library plumbing + host-side sizing formulas of the reference modules.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <iterator>
#include <map>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <algorithm>

namespace hx {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return static_cast<int>(err);
    }
    return HX_OK;
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_version(void) { return 100; }

const char* hx_last_error_string(void) { return g_error; }

int hx_device_count(void)
{
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        set_error("no CUDA device: %s", cudaGetErrorString(err));
        return HX_E_NOGPU;
    }
    return n;
}

uint64_t hx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// halmd/mdsim/gpu/particle.cpp:69-88 — round up to a multiple of max threads per block (1024)
size_t hx_array_size(size_t nparticle)
{
    size_t const block = 1024;
    size_t a = nparticle > 0 ? (nparticle + block - 1) / block : 1;
    return a * block;
}

// halmd/mdsim/gpu/binning.cpp:81-116 — cell grid and placeholders per cell
int hx_binning_plan(const double box_length[3], size_t nparticle, float r_cut_max, float skin,
                    double occupancy, unsigned ncell[3], float cell_length[3], unsigned* cell_size)
{
    if (!box_length || !ncell || !cell_length || !cell_size || occupancy <= 0 || nparticle == 0) {
        set_error("hx_binning_plan: invalid argument");
        return HX_E_INVALID;
    }
    double const warp_size = 32;
    double nwarps = nparticle / (occupancy * warp_size);
    double volume = 1.;
    for (int d = 0; d < 3; ++d) volume *= box_length[d];
    size_t ncells = 1;
    for (int d = 0; d < 3; ++d) {
        // upper bound: cell_length > r_c + r_skin (float arithmetic as in the reference members)
        unsigned ncell_max = static_cast<unsigned>(box_length[d] / static_cast<double>(r_cut_max + skin));
        // fixed_vector<double> = float length * pow(...)
        unsigned want = static_cast<unsigned>(std::ceil(box_length[d] * std::pow(nwarps / volume, 1. / 3)));
        ncell[d] = std::min(want, ncell_max);
        if (ncell[d] == 0) {
            set_error("hx_binning_plan: box too small for cutoff");
            return HX_E_INVALID;
        }
        cell_length[d] = static_cast<float>(box_length[d]) / static_cast<float>(ncell[d]);
        ncells *= ncell[d];
    }
    *cell_size = static_cast<unsigned>(warp_size * static_cast<size_t>(std::ceil(nwarps / ncells)));
    return HX_OK;
}

// halmd/mdsim/gpu/neighbours/from_binning.cpp:96-119 — placeholders per neighbour list
int hx_nlist_plan(float r_cut_max, float skin, size_t nparticle2, double box_volume,
                  double occupancy, unsigned cell_size, unsigned* size)
{
    if (!size || occupancy <= 0 || box_volume <= 0) {
        set_error("hx_nlist_plan: invalid argument");
        return HX_E_INVALID;
    }
    float const unit_sphere3 = static_cast<float>(4 * M_PI / 3);
    float neighbour_sphere = static_cast<float>(unit_sphere3 * std::pow(r_cut_max + skin, 3));
    float density = static_cast<float>(nparticle2 / box_volume);
    size_t s = static_cast<size_t>(std::ceil(neighbour_sphere * (density / occupancy)));
    *size = static_cast<unsigned>(std::max<size_t>(s, cell_size));
    return HX_OK;
}

// halmd/mdsim/gpu/sorts/hilbert.cpp:51-54
unsigned hx_hilbert_depth(const double box_length[3])
{
    float max_length = static_cast<float>(std::max(box_length[0], std::max(box_length[1], box_length[2])));
    unsigned depth = static_cast<unsigned>(std::ceil(std::log(max_length) / M_LN2));
    return std::min(10u, depth);
}

// from_binning.cpp:130-131 — float(displacement) > float(skin / 2), displacement = sqrt(float)
int hx_displacement_exceeds(uint32_t max_rr_bits, double skin)
{
    float rr;
    std::memcpy(&rr, &max_rr_bits, 4);
    float d = std::sqrt(rr);
    return d > static_cast<float>(static_cast<float>(skin) / 2) ? 1 : 0;
}

// lennard_jones.cpp:44-70, smooth_r4.hpp:55-91 (float matrices, float formulas)
int hx_potential_lj_init(hx_lj_params* p, unsigned ntype1, unsigned ntype2, const float* epsilon,
                         const float* sigma, int truncation, const float* cutoff_sigma, float h)
{
    return hx_potential_lj_init_hard_core(p, ntype1, ntype2, epsilon, sigma, nullptr, truncation, cutoff_sigma, h);
}

int hx_potential_lj_init_hard_core(hx_lj_params* p, unsigned ntype1, unsigned ntype2, const float* epsilon,
                                   const float* sigma, const float* core_sigma, int truncation,
                                   const float* cutoff_sigma, float h)
{
    if (!p || !epsilon || !sigma || !cutoff_sigma || ntype1 == 0 || ntype2 == 0
        || ntype1 * ntype2 > HX_MAX_PAIR_TYPES) {
        set_error("hx_potential_lj_init: invalid argument (ntype1*ntype2 must be in 1..%d)", HX_MAX_PAIR_TYPES);
        return HX_E_INVALID;
    }
    if (truncation < HX_TRUNC_SHARP || truncation > HX_TRUNC_FORCE_SHIFTED) {
        set_error("hx_potential_lj_init: unknown truncation %d", truncation);
        return HX_E_UNSUPPORTED;
    }
    std::memset(p, 0, sizeof(*p));
    p->ntype1 = ntype1;
    p->ntype2 = ntype2;
    p->truncation = truncation;
    p->hard_core = core_sigma != nullptr;
    // rri_smooth_(std::pow(h, -2)): pow(float, int) is evaluated in double, stored as float
    p->rri_smooth = (truncation == HX_TRUNC_SMOOTH_R4)
        ? static_cast<float>(std::pow(static_cast<double>(h), -2)) : 0.f;
    for (unsigned i = 0; i < ntype1 * ntype2; ++i) {
        float eps = epsilon[i];
        float sigma2 = sigma[i] * sigma[i];
        float r_cut = sigma[i] * cutoff_sigma[i];
        float rr_cut = r_cut * r_cut;
        // host-side operator() of the float potential, evaluated at rr_cut; through the hard-core adapter
        // (adapters/hard_core.hpp:86-94) at (sqrt(rr_cut) - r_core)^2, the force scaled by r_s / r
        float r_core = core_sigma ? core_sigma[i] * sigma[i] : 0.f;
        if (core_sigma && !(r_core >= 0.f && r_core < r_cut)) {
            set_error("hx_potential_lj_init_hard_core: core radius must lie in [0, r_cut)");
            return HX_E_INVALID;
        }
        float r = std::sqrt(rr_cut);
        float r_s = r - r_core;
        float rri = sigma2 / (core_sigma ? r_s * r_s : rr_cut);
        float r6i = rri * rri * rri;
        float eps_r6i = eps * r6i;
        float en_cut = 4 * eps_r6i * (r6i - 1);
        float fval_cut = 48 * rri * eps_r6i * (r6i - 0.5f) / sigma2;
        if (core_sigma) fval_cut *= r_s / r;
        p->r_core[i] = r_core;
        p->epsilon[i] = eps;
        p->sigma2[i] = sigma2;
        p->r_cut[i] = r_cut;
        p->rr_cut[i] = rr_cut;
        p->en_cut[i] = (truncation == HX_TRUNC_SHARP) ? 0.f : en_cut;
        // force_shifted.hpp:69-70 — force_cut = fval(rr_cut) * r_cut
        p->force_cut[i] = (truncation == HX_TRUNC_FORCE_SHIFTED) ? fval_cut * r_cut : 0.f;
    }
    return HX_OK;
}

/* ---- Nose-Hoover chain, host side (verlet_nvt_hoover.cpp:79-95, :175-222) ---------------------------------------- */
int hx_hoover_chain_set_temperature(hx_hoover_chain* c, double temperature, double resonance_frequency)
{
    if (!c || !(temperature > 0) || !(resonance_frequency > 0)) {
        set_error("hx_hoover_chain_set_temperature: invalid argument");
        return HX_E_INVALID;
    }
    c->temperature = temperature;
    c->en_kin_target_2 = c->dof * temperature;
    // follow Martyna et al. [J. Chem. Phys. 97, 2635 (1992)] for the masses
    double const omega_sq = std::pow(2 * M_PI * resonance_frequency, 2);
    c->mass_xi[0] = c->dof * temperature / omega_sq;
    c->mass_xi[1] = temperature / omega_sq;
    return HX_OK;
}

int hx_hoover_chain_init(hx_hoover_chain* c, size_t nparticle, unsigned dimension, double temperature,
                         double resonance_frequency)
{
    if (!c || nparticle == 0 || dimension == 0) {
        set_error("hx_hoover_chain_init: invalid argument");
        return HX_E_INVALID;
    }
    std::memset(c, 0, sizeof(*c));
    c->dof = static_cast<double>(dimension) * static_cast<double>(nparticle);
    return hx_hoover_chain_set_temperature(c, temperature, resonance_frequency);
}

double hx_hoover_chain_propagate(hx_hoover_chain* c, double en_kin_2, double timestep)
{
    double const dt2 = timestep / 2, dt4 = timestep / 4, dt8 = timestep / 8;
    double* const xi = c->xi;
    double* const v = c->v_xi;
    double const* const m = c->mass_xi;
    double const T = c->temperature;
    // head of the chain
    v[1] += (m[0] * v[0] * v[0] - T) / m[1] * dt4;
    double const t = std::exp(-v[1] * dt8);
    v[0] *= t;
    v[0] += (en_kin_2 - c->en_kin_target_2) / m[0] * dt4;
    v[0] *= t;
    // heat-bath positions
    xi[0] += v[0] * dt2;
    xi[1] += v[1] * dt2;
    // the factor the kernels apply to the velocities; the kinetic energy follows with its square
    double const s = std::exp(-v[0] * dt2);
    en_kin_2 *= s * s;
    // tail of the chain, mirror image of the head
    v[0] *= t;
    v[0] += (en_kin_2 - c->en_kin_target_2) / m[0] * dt4;
    v[0] *= t;
    v[1] += (m[0] * v[0] * v[0] - T) / m[1] * dt4;
    return s;
}

double hx_hoover_chain_energy(const hx_hoover_chain* c, size_t nparticle)
{
    double en = c->temperature * (c->dof * c->xi[0] + c->xi[1]);
    for (int i = 0; i < 2; ++i) en += c->mass_xi[i] * c->v_xi[i] * c->v_xi[i] / 2;
    return en / static_cast<double>(nparticle);
}


// observables.utility.wavevector (halmd/observables/utility/wavevector.cpp:40-123) on top of
// algorithm::host::pick_lattice_points_from_shell (halmd/algorithm/host/pick_lattice_points.hpp:62-160): reciprocal
// lattice vectors n (h, k, l) * 2 pi / L with coprime Miller indices whose magnitude matches one of the wavenumbers
// within the relative tolerance, at most max_count per wavenumber, generated in the order of increasing h + k + l
// (index tuple (h, h + k, h + k + l) counted like an odometer) — that order decides which points a full shell keeps.
// Output: vectors grouped by wavenumber in the order of `wavenumber` (wavenumbers without a vector get an empty shell:
// shell_offsets[i] == shell_offsets[i + 1]; the caller drops them as wavevector.cpp:103-118 does).
int hx_wavevector_pick(const double* wavenumber, unsigned nwavenumber, const double box_length[3], double tolerance,
                       unsigned max_count, const unsigned* filter, double* vectors_out, unsigned capacity,
                       unsigned* shell_offsets_out)
{
    if (!wavenumber || !box_length || !vectors_out || !shell_offsets_out || nwavenumber == 0) {
        set_error("hx_wavevector_pick: invalid argument");
        return HX_E_INVALID;
    }
    unsigned const one[3] = {1, 1, 1};
    if (!filter) filter = one;
    for (int d = 0; d < 3; ++d)
        if (filter[d] > 1) { set_error("filter values must be 0 or 1"); return HX_E_INVALID; }
    double unit[3];
    for (int d = 0; d < 3; ++d) unit[d] = 2 * M_PI / box_length[d];
    double r_max = 0;
    for (unsigned i = 0; i < nwavenumber; ++i) r_max = std::max(r_max, wavenumber[i]);
    r_max *= 1 + tolerance;
    unsigned long long sum_max[3];
    for (int d = 0; d < 3; ++d) sum_max[d] = (unit[d] == 0) ? 0 : static_cast<unsigned int>(r_max / unit[d]);
    sum_max[1] += sum_max[0];
    sum_max[2] += sum_max[1];
    // wavevector.cpp:78-87 collects the points in a multimap keyed by the wavenumber through std::inserter(map, end):
    // the hint of every insertion is the position behind the previous one, and an element goes "as close as possible
    // to the position just prior to the hint" — which fixes the order of the vectors inside a shell (the first
    // wavenumber keeps the generation order, later ones mostly do not).  The same container and inserter are used here.
    struct vec3d { double v[3]; };
    std::multimap<unsigned, vec3d> picked;
    auto out = std::inserter(picked, picked.end());
    std::vector<unsigned> count(nwavenumber, 0);
    auto gcd = [](unsigned long long a, unsigned long long b) { while (b) { unsigned long long t = b; b = a % b; a = t; } return a; };
    unsigned long long idx[3] = {0, 0, 0};
    while (idx[2] <= sum_max[2]) {
        // (the reference forms h, k, l in unsigned arithmetic; tuples with idx[j] < idx[j-1] wrap to huge indices whose
        // direction never matches a wavenumber: they are skipped here)
        if (idx[1] >= idx[0] && idx[2] >= idx[1]) {
            unsigned long long const hkl[3] = {idx[0], idx[1] - idx[0], idx[2] - idx[1]};
            bool const pass = hkl[0] * filter[0] == hkl[0] && hkl[1] * filter[1] == hkl[1] && hkl[2] * filter[2] == hkl[2];
            if (pass && gcd(hkl[0], gcd(hkl[1], hkl[2])) == 1) {
                double const r0[3] = {unit[0] * hkl[0], unit[1] * hkl[1], unit[2] * hkl[2]};
                double const r0_norm = std::sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
                // (directions longer than the largest radius match nothing: n * r0_norm > r * (1 + tolerance) for n >= 1)
                for (unsigned i = 0; r0_norm <= r_max && i < nwavenumber; ++i) {
                    if (count[i] < max_count) {
                        unsigned int const n = static_cast<unsigned int>(std::floor(wavenumber[i] / r0_norm + 0.5));
                        if (n > 0 && std::abs(n * r0_norm - wavenumber[i]) <= wavenumber[i] * tolerance) {
                            *out++ = std::make_pair(i, vec3d{{n * r0[0], n * r0[1], n * r0[2]}});
                            ++count[i];
                        }
                    }
                }
            }
        }
        ++idx[0];
        for (int j = 0; j < 2; ++j) {
            if (idx[j] <= sum_max[j]) break;
            idx[j] = 0;
            ++idx[j + 1];
        }
    }
    unsigned total = 0;
    for (unsigned i = 0; i < nwavenumber; ++i) {
        shell_offsets_out[i] = total;
        if (total + count[i] > capacity) { set_error("hx_wavevector_pick: output capacity %u too small", capacity); return HX_E_INVALID; }
        auto const range = picked.equal_range(i);
        for (auto it = range.first; it != range.second; ++it, ++total)
            for (int d = 0; d < 3; ++d) vectors_out[3 * static_cast<size_t>(total) + d] = it->second.v[d];
    }
    shell_offsets_out[nwavenumber] = total;
    if (total == 0) {
        set_error("Constraints on wavevectors are incompatible with geometry of simulation box.");
        return HX_E_INVALID;
    }
    return HX_OK;
}

// intermediate_scattering_function.hpp:90-113 with the accumulator of numeric/accumulator.hpp:56-132 (value update
// :64-70, merge :104-118) in the reference's operation order
int hx_isf_accumulate(const double* rho1, const double* rho2, const unsigned* shell_offsets, unsigned nshell, double norm,
                      double* acc)
{
    if (!rho1 || !rho2 || !shell_offsets || !acc) { set_error("hx_isf_accumulate: invalid argument"); return HX_E_INVALID; }
    for (unsigned s = 0; s < nshell; ++s) {
        double n = 0, m = 0, v = 0;
        for (unsigned i = shell_offsets[s]; i < shell_offsets[s + 1]; ++i) {
            double const re = rho1[2 * i] * rho2[2 * i] + rho1[2 * i + 1] * rho2[2 * i + 1];
            double const x = re / norm;
            double const t = x - m;
            n += 1;
            m += t / n;
            v += t * (x - m);
        }
        if (n > 0) {
            double* const a = acc + 3 * static_cast<size_t>(s);
            double const count = a[0] + n;
            a[2] += v;
            double const diff = a[1] - m;
            a[2] += diff * diff * a[0] * n / count;
            a[1] = (a[0] * a[1] + n * m) / count;
            a[0] = count;
        }
    }
    return HX_OK;
}

} // extern "C"
