// TEST INFRASTRUCTURE ONLY: BOOST_CONCEPT_REQUIRES(concepts, (return_type)) reduced to the return type.
#pragma once
namespace ref_compat {
template <typename F> struct unparenthesise;
template <typename R> struct unparenthesise<void(R)> { typedef R type; };
template <> struct unparenthesise<void()> { typedef void type; };
} // namespace ref_compat
#define BOOST_CONCEPT_REQUIRES(models, result) typename ::ref_compat::unparenthesise<void result>::type
