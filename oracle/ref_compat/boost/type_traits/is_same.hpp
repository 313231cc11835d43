// TEST INFRASTRUCTURE ONLY: boost::is_same on top of the standard trait.
#pragma once
#include <type_traits>
namespace boost { template <typename A, typename B> struct is_same : std::is_same<A, B> {}; }
