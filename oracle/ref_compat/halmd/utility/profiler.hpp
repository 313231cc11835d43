// TEST INFRASTRUCTURE ONLY: stand-in for halmd/utility/profiler.hpp (its timers need Boost.System): runtime accumulators
// exist, scoped timers do nothing.
#ifndef HALMD_UTILITY_PROFILER_HPP
#define HALMD_UTILITY_PROFILER_HPP
#include <lua.hpp>
#include <halmd/numeric/accumulator.hpp>
namespace halmd {
namespace utility {
class profiler {
public:
    typedef accumulator<double> accumulator_type;
    struct scoped_timer_type { explicit scoped_timer_type(accumulator_type&) {} };
};
} // namespace utility
} // namespace halmd
#endif
