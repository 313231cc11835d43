// TEST INFRASTRUCTURE ONLY: boost::array as std::array (Boost is absent)
#ifndef HALMD_REF_COMPAT_BOOST_ARRAY_HPP
#define HALMD_REF_COMPAT_BOOST_ARRAY_HPP
#include <array>
#include <cstddef>
namespace boost { template <typename T, std::size_t N> using array = std::array<T, N>; }
#endif
