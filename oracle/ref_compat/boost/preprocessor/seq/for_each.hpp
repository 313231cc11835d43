// Minimal stand-in for <boost/preprocessor/seq/for_each.hpp> (Boost is not installed in this image): enough of
// BOOST_PP_SEQ_FOR_EACH(macro, data, seq) for the sequences of up to four elements that
// halmd/mdsim/gpu/potentials/pair/truncations/truncations.cuh iterates over.  TEST INFRASTRUCTURE ONLY.
#pragma once
#define HXPP_CAT(a, b) HXPP_CAT_I(a, b)
#define HXPP_CAT_I(a, b) a##b
// length of (a)(b)...: the counters chain HXPP_N0 -> HXPP_N1 -> ... once per element
#define HXPP_SIZE(seq) HXPP_CAT(HXPP_SIZE_, HXPP_N0 seq)
#define HXPP_N0(x) HXPP_N1
#define HXPP_N1(x) HXPP_N2
#define HXPP_N2(x) HXPP_N3
#define HXPP_N3(x) HXPP_N4
#define HXPP_SIZE_HXPP_N0 0
#define HXPP_SIZE_HXPP_N1 1
#define HXPP_SIZE_HXPP_N2 2
#define HXPP_SIZE_HXPP_N3 3
#define HXPP_SIZE_HXPP_N4 4
// first element and rest of a sequence
#define HXPP_HEAD(seq) HXPP_HEAD_I(HXPP_HEAD_II seq)
#define HXPP_HEAD_II(x) x, HXPP_NIL
#define HXPP_NIL(x)
#define HXPP_HEAD_I(...) HXPP_HEAD_III(__VA_ARGS__)
#define HXPP_HEAD_III(x, ...) x
#define HXPP_TAIL(seq) HXPP_EAT seq
#define HXPP_EAT(x)
#define HXPP_FE_0(m, d, seq)
#define HXPP_FE_1(m, d, seq) m(1, d, HXPP_HEAD(seq))
#define HXPP_FE_2(m, d, seq) m(1, d, HXPP_HEAD(seq)) HXPP_FE_1(m, d, HXPP_TAIL(seq))
#define HXPP_FE_3(m, d, seq) m(1, d, HXPP_HEAD(seq)) HXPP_FE_2(m, d, HXPP_TAIL(seq))
#define HXPP_FE_4(m, d, seq) m(1, d, HXPP_HEAD(seq)) HXPP_FE_3(m, d, HXPP_TAIL(seq))
#define BOOST_PP_SEQ_FOR_EACH(m, d, seq) HXPP_CAT(HXPP_FE_, HXPP_SIZE(seq))(m, d, seq)
