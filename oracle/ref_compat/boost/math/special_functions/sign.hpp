// Minimal stand-in for <boost/math/special_functions/sign.hpp> (Boost is not installed here): the reference's
// halmd/mdsim/sorts/hilbert_kernel.hpp:26-29 only needs boost::math::signbit on the host.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <typename T> inline int signbit(T x) { return std::signbit(x) ? 1 : 0; }
}}
