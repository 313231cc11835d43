// TEST INFRASTRUCTURE ONLY: the part of boost::circular_buffer observables/samples/blocking_scheme.hpp uses
// (fixed capacity, push_back, pop_front, full / empty / size, iteration) on a std::deque (Boost is absent)
#ifndef HALMD_REF_COMPAT_BOOST_CIRCULAR_BUFFER_HPP
#define HALMD_REF_COMPAT_BOOST_CIRCULAR_BUFFER_HPP
#include <cstddef>
#include <deque>
namespace boost {
template <typename T>
class circular_buffer {
public:
    typedef typename std::deque<T>::iterator iterator;
    typedef typename std::deque<T>::const_iterator const_iterator;
    explicit circular_buffer(std::size_t capacity = 0) : capacity_(capacity) {}
    void push_back(T const& v) { if (capacity_ == 0) return; if (d_.size() == capacity_) d_.pop_front(); d_.push_back(v); }
    void pop_front() { d_.pop_front(); }
    void clear() { d_.clear(); }
    bool full() const { return d_.size() == capacity_; }
    bool empty() const { return d_.empty(); }
    std::size_t size() const { return d_.size(); }
    std::size_t capacity() const { return capacity_; }
    iterator begin() { return d_.begin(); }
    iterator end() { return d_.end(); }
    const_iterator begin() const { return d_.begin(); }
    const_iterator end() const { return d_.end(); }
    T const& operator[](std::size_t i) const { return d_[i]; }
private:
    std::size_t capacity_;
    std::deque<T> d_;
};
} // namespace boost
#endif
