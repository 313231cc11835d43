// ref_dynamics.cpp — TEST INFRASTRUCTURE ONLY.  The reference's own host-side dynamics classes compiled in place (their .cpp
// files are #included from /root/reference, nothing is copied): mdsim::clock, observables::dynamics::blocking_scheme
// (blocking_scheme.cpp:34-200: sampling intervals, sample / process / finalise, the origin + drop rule),
// observables::samples::blocking_scheme<T> (the per-level sample buffers), observables::dynamics::correlation<tcf>
// (correlation.hpp:208-224: first entry of a block against all entries), observables::utility::wavevector<3>
// (wavevector.cpp:40-123) and observables::dynamics::intermediate_scattering_function<3>
// (intermediate_scattering_function.hpp:90-113).  Boost, Lua and luaponte are absent from this image; the stand-ins under
// oracle/ref_compat/ (multi_array, circular_buffer, optional, array, the logger and a do-nothing binding DSL) supply
// what these files name.  Part of oracle/_ref/libhalmd_ref.so; used only by tests/.
#include <halmd/config.hpp>
#include <halmd/mdsim/clock.cpp>
#include <halmd/observables/dynamics/blocking_scheme.cpp>
#include <halmd/observables/utility/wavevector.cpp>
#include <halmd/observables/dynamics/correlation.hpp>
#include <halmd/observables/dynamics/intermediate_scattering_function.hpp>
#include <halmd/utility/raw_array.hpp>

#include <cstdint>
#include <memory>
#include <vector>

using namespace halmd;

namespace {

// a sample is the step at which it was taken; the correlation functor records which pairs of samples it is handed
struct step_tcf {
    typedef uint64_t sample_type;
    typedef double result_type;
    std::vector<uint64_t>* log;
    template <typename Accumulator>
    void operator()(sample_type const& first, sample_type const& second, Accumulator& result) const
    {
        log->push_back(first);
        log->push_back(second);
        result(static_cast<double>(second - first));
    }
};

} // namespace

extern "C" {

// Drive the reference's blocking scheme like observables::sampler does (sampler.cpp:63-80: sample the initial state,
// then advance the clock and sample after every step), with one sample buffer and one correlation function connected.
//   intervals_out[levels], events_out: (level, first step, second step) per correlated pair in call order,
//   counts_out[levels * block_size]: accumulator count per (level, lag); mean_out: mean lag in steps per (level, lag)
// Returns the number of levels, or -1 - (error) on an exception / insufficient capacity.
int ref_blocking_schedule(double timestep, double max_lag_time, double resolution, unsigned block_size, unsigned shift,
                          unsigned separation, uint64_t start_step, uint64_t nsteps, int finalise, uint64_t* intervals_out,
                          unsigned max_levels, uint64_t* events_out, uint64_t max_events, uint64_t* nevents_out, double* counts_out,
                          double* mean_out, double* time_out)
{
    try {
        auto clock = std::make_shared<mdsim::clock>();
        clock->set_timestep(timestep);
        for (uint64_t i = 0; i < start_step; ++i) clock->advance();
        auto bs = std::make_shared<observables::dynamics::blocking_scheme>(
            clock, max_lag_time, resolution, block_size, shift, separation, std::make_shared<logger>());
        unsigned const levels = bs->count();
        if (levels > max_levels) return -2;
        typedef observables::samples::blocking_scheme<uint64_t> samples_type;
        auto samples = std::make_shared<samples_type>(
            [clock]() { return std::make_shared<uint64_t const>(clock->step()); }, levels, block_size);
        std::vector<uint64_t> log;
        auto tcf = std::make_shared<step_tcf>();
        tcf->log = &log;
        typedef observables::dynamics::correlation<step_tcf> correlation_type;
        auto corr = std::make_shared<correlation_type>(tcf, samples, samples, std::make_shared<logger>());
        bs->on_sample(samples);
        // record the level of every compute() call through a second, level-logging correlation
        struct level_logger : observables::dynamics::correlation_base {
            std::vector<unsigned> levels;
            std::vector<size_t> log_size;
            std::vector<uint64_t>* log;
            void compute(unsigned int level) { levels.push_back(level); log_size.push_back(log->size()); }
        };
        auto ll = std::make_shared<level_logger>();
        ll->log = &log;
        bs->on_correlate(ll);           // (slots run in connection order: the level is logged before the pairs)
        bs->on_correlate(corr);
        bs->sample();
        for (uint64_t i = 0; i < nsteps; ++i) {
            clock->advance();
            bs->sample();
        }
        if (finalise) bs->finalise();
        // events: (level, first, second)
        uint64_t nev = 0;
        for (size_t c = 0; c < ll->levels.size(); ++c) {
            size_t const b = ll->log_size[c], e = c + 1 < ll->levels.size() ? ll->log_size[c + 1] : log.size();
            for (size_t k = b; k < e; k += 2) {
                if (nev >= max_events) return -3;
                events_out[3 * nev] = ll->levels[c];
                events_out[3 * nev + 1] = log[k];
                events_out[3 * nev + 2] = log[k + 1];
                ++nev;
            }
        }
        *nevents_out = nev;
        auto const& count = correlation_type::get_count(corr)();
        auto const& mean = correlation_type::get_mean(corr)();
        for (unsigned i = 0; i < levels; ++i) {
            for (unsigned j = 0; j < block_size; ++j) {
                counts_out[i * block_size + j] = static_cast<double>(count[i][j]);
                mean_out[i * block_size + j] = mean[i][j];
                time_out[i * block_size + j] = bs->time()[i][j];
            }
        }
        // sampling intervals = time grid / time step at lag 1 (the member itself is private)
        for (unsigned i = 0; i < levels; ++i) intervals_out[i] = static_cast<uint64_t>(std::llround(bs->time()[i][1] / timestep));
        return static_cast<int>(levels);
    }
    catch (std::exception const&) {
        return -1;
    }
}

// The reference's ISF functor on two density-mode samples (re, im per wavevector) for the wavevector shells the
// reference's wavevector class picks; `nsamples` consecutive sample pairs are accumulated into the same output
// accumulators as correlation<>::compute does over time.  out[3 * nshell]: count, mean, variance per shell.
int ref_isf_accumulate(const double* wavenumber, unsigned nwavenumber, const double box_length[3], double tolerance,
                       unsigned max_count, const double* rho1, const double* rho2, unsigned nsamples, double norm,
                       unsigned* nq_out, unsigned* nshell_out, double* out)
{
    try {
        typedef observables::utility::wavevector<3> wavevector_type;
        typedef observables::dynamics::intermediate_scattering_function<3> isf_type;
        wavevector_type::vector_type L;
        for (int d = 0; d < 3; ++d) L[d] = box_length[d];
        auto wv = std::make_shared<wavevector_type>(std::vector<double>(wavenumber, wavenumber + nwavenumber), L, tolerance,
                                                    max_count, wavevector_type::filter_type(1));
        unsigned const nq = wv->value().size(), nshell = wv->shell().size();
        *nq_out = nq;
        *nshell_out = nshell;
        if (!rho1) return 0;
        isf_type isf(wv, norm);
        boost::multi_array<accumulator<double>, 1> result(boost::extents[nshell]);
        for (unsigned s = 0; s < nsamples; ++s) {
            isf_type::sample_type a(nq), b(nq);
            for (unsigned i = 0; i < nq; ++i) {
                a[i][0] = rho1[2 * (s * nq + i)]; a[i][1] = rho1[2 * (s * nq + i) + 1];
                b[i][0] = rho2[2 * (s * nq + i)]; b[i][1] = rho2[2 * (s * nq + i) + 1];
            }
            isf(a, b, result.whole());
        }
        for (unsigned k = 0; k < nshell; ++k) {
            out[3 * k] = static_cast<double>(count(result[k]));
            out[3 * k + 1] = mean(result[k]);
            out[3 * k + 2] = count(result[k]) > 1 ? variance(result[k]) : 0;
        }
        return 0;
    }
    catch (std::exception const&) {
        return -1;
    }
}

// wavevectors and shells of the reference's wavevector class (empty shells discarded, wavevector.cpp:88-120)
int ref_wavevector_class(const double* wavenumber, unsigned nwavenumber, const double box_length[3], double tolerance,
                         unsigned max_count, double* vectors_out, unsigned capacity, unsigned* shells_out, double* kept_wavenumbers_out)
{
    try {
        typedef observables::utility::wavevector<3> wavevector_type;
        wavevector_type::vector_type L;
        for (int d = 0; d < 3; ++d) L[d] = box_length[d];
        wavevector_type wv(std::vector<double>(wavenumber, wavenumber + nwavenumber), L, tolerance, max_count,
                           wavevector_type::filter_type(1));
        if (wv.value().size() > capacity) return -2;
        for (size_t i = 0; i < wv.value().size(); ++i)
            for (int d = 0; d < 3; ++d) vectors_out[3 * i + d] = wv.value()[i][d];
        for (size_t s = 0; s < wv.shell().size(); ++s) {
            shells_out[2 * s] = wv.shell()[s].first;
            shells_out[2 * s + 1] = wv.shell()[s].second;
            kept_wavenumbers_out[s] = wv.wavenumber()[s];
        }
        return static_cast<int>(wv.shell().size());
    }
    catch (std::exception const&) {
        return -1;
    }
}

}
