// ref_shim.cpp — thin extern "C" window onto the REFERENCE's own header-only numerics, compiled from the
// sources where they lie under /root/reference (never copied).  Output: oracle/_ref/libhalmd_ref.so.
// Used only by tests to validate oracle/hx_oracle.c (dsfun, Hilbert map, LJ functor, lattice).
#include <cuda_runtime.h>   // cudaTextureObject_t used by the kernel headers
#include <halmd/config.hpp>
#include <halmd/numeric/blas/fixed_vector.hpp>
#include <halmd/numeric/mp/dsfun.hpp>
#include <halmd/mdsim/sorts/hilbert_kernel.hpp>
#include <halmd/mdsim/gpu/potentials/pair/lennard_jones_kernel.hpp>
#include <halmd/mdsim/positions/lattice_primitive.hpp>
#include <halmd/numeric/accumulator.hpp>
#include <halmd/algorithm/host/pick_lattice_points.hpp>
#include <iterator>
#include <map>
#include <tuple>
#include <vector>

using namespace halmd;

extern "C" {

void ref_dsadd(const float* a, const float* b, float* c) { detail::numeric::mp::dsadd(c[0], c[1], a[0], a[1], b[0], b[1]); }
void ref_dssub(const float* a, const float* b, float* c) { detail::numeric::mp::dssub(c[0], c[1], a[0], a[1], b[0], b[1]); }
void ref_dsmul(const float* a, const float* b, float* c) { detail::numeric::mp::dsmul(c[0], c[1], a[0], a[1], b[0], b[1]); }

unsigned ref_hilbert_map(float x, float y, float z, unsigned depth)
{
    fixed_vector<float, 3> r;
    r[0] = x; r[1] = y; r[2] = z;
    return mdsim::sorts::hilbert_kernel::map(r, depth);
}

void ref_lj_compute_f(float rr, float sigma2, float epsilon, float* fval, float* en)
{
    float f, e;
    tie(f, e) = mdsim::gpu::potentials::pair::lennard_jones_kernel::compute(rr, sigma2, epsilon);
    *fval = f; *en = e;
}

void ref_lj_compute_d(double rr, double sigma2, double epsilon, double* fval, double* en)
{
    double f, e;
    tie(f, e) = mdsim::gpu::potentials::pair::lennard_jones_kernel::compute(rr, sigma2, epsilon);
    *fval = f; *en = e;
}

// fcc lattice site of index n for a lattice of shape (nx, ny, nz) unit cells (lattice_primitive.hpp:121-129)
void ref_fcc_site(unsigned n, unsigned nx, unsigned ny, unsigned nz, float* out)
{
    typedef fixed_vector<float, 3> vec;
    typedef fixed_vector<unsigned int, 3> shape;
    shape s; s[0] = nx; s[1] = ny; s[2] = nz;
    close_packed_lattice<vec, shape> lat(s);
    vec r = lat(n);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

// Welford accumulator (numeric/accumulator.hpp:56-132): feed values, then merge a second batch
void ref_accumulate(const double* x, unsigned n1, unsigned n2, double* out)
{
    accumulator<double> a, b;
    for (unsigned i = 0; i < n1; ++i) a(x[i]);
    for (unsigned i = n1; i < n1 + n2; ++i) b(x[i]);
    a(b);
    out[0] = static_cast<double>(count(a)); out[1] = mean(a); out[2] = variance(a);
}


// observables::utility::wavevector<3> (observables/utility/wavevector.cpp:40-123): the reference's
// pick_lattice_points_from_shell (algorithm/host/pick_lattice_points.hpp:62-160) called with the arguments and the
// multimap the constructor uses; vectors grouped by wavenumber in equal_range order, empty shells kept as empty ranges.
int ref_wavevector_pick(const double* wavenumber, unsigned nwavenumber, const double box_length[3], double tolerance,
                        unsigned max_count, const unsigned* filter, double* vectors_out, unsigned capacity,
                        unsigned* shell_offsets_out)
{
    typedef fixed_vector<double, 3> vector_type;
    typedef fixed_vector<unsigned int, 3> filter_type;
    std::vector<double> q(wavenumber, wavenumber + nwavenumber);
    vector_type L;
    filter_type f(1);
    for (int d = 0; d < 3; ++d) { L[d] = box_length[d]; if (filter) f[d] = filter[d]; }
    auto first = q.begin();
    auto last = q.end();
    std::multimap<decltype(first), vector_type> wavevector_map;
    algorithm::host::pick_lattice_points_from_shell(
        first, last, std::inserter(wavevector_map, wavevector_map.end()), element_div(vector_type(2 * M_PI), L), tolerance,
        max_count, f);
    unsigned total = 0;
    for (auto q_it = q.begin(); q_it != q.end(); ++q_it) {
        shell_offsets_out[q_it - q.begin()] = total;
        auto range = wavevector_map.equal_range(q_it);
        for (auto it = range.first; it != range.second; ++it) {
            if (total >= capacity) return -1;
            for (int d = 0; d < 3; ++d) vectors_out[3 * total + d] = it->second[d];
            ++total;
        }
    }
    shell_offsets_out[nwavenumber] = total;
    return 0;
}

}
