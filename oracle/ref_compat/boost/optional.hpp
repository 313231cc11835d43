// TEST INFRASTRUCTURE ONLY: the part of boost::optional mdsim/clock.hpp uses (Boost is absent)
#ifndef HALMD_REF_COMPAT_BOOST_OPTIONAL_HPP
#define HALMD_REF_COMPAT_BOOST_OPTIONAL_HPP
namespace boost {
template <typename T>
class optional {
public:
    optional() : set_(false), value_() {}
    optional(T const& v) : set_(true), value_(v) {}
    optional& operator=(T const& v) { set_ = true; value_ = v; return *this; }
    explicit operator bool() const { return set_; }
    bool operator!() const { return !set_; }
    T const& operator*() const { return value_; }
    T& operator*() { return value_; }
private:
    bool set_;
    T value_;
};
} // namespace boost
#endif
