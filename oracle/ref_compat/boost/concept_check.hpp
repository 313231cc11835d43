// TEST INFRASTRUCTURE ONLY (oracle/Makefile): Boost is absent; the reference's host algorithm compiled in place
// (halmd/algorithm/host/pick_lattice_points.hpp) only names these concepts inside BOOST_CONCEPT_REQUIRES.
#pragma once
namespace boost {
template <typename T> struct InputIterator {};
template <typename T, typename V> struct OutputIterator {};
} // namespace boost
