// TEST INFRASTRUCTURE ONLY: the part of boost::multi_array the reference's blocking scheme and correlation classes use
// (row-major storage, resize from boost::extents[..] or an extent array, nested operator[] views with begin() / end(),
// origin(), num_elements(), shape()) — Boost is absent.
#ifndef HALMD_REF_COMPAT_BOOST_MULTI_ARRAY_HPP
#define HALMD_REF_COMPAT_BOOST_MULTI_ARRAY_HPP
#include <array>
#include <cstddef>
#include <iterator>
#include <vector>
namespace boost {
namespace detail_ma {
template <std::size_t N> struct extent_gen {
    std::array<std::size_t, N> e;
    extent_gen<N + 1> operator[](std::size_t n) const
    {
        extent_gen<N + 1> r;
        for (std::size_t i = 0; i < N; ++i) r.e[i] = e[i];
        r.e[N] = n;
        return r;
    }
};
template <> struct extent_gen<0> {
    extent_gen<1> operator[](std::size_t n) const { extent_gen<1> r; r.e[0] = n; return r; }
};
// view of rank R into row-major storage; rank 0 is the element itself
template <typename T, std::size_t R> class view;
template <typename T, std::size_t R> struct view_ref { typedef view<T, R> type; static type make(T* p, std::size_t const* s) { return type(p, s); } };
template <typename T> struct view_ref<T, 0> { typedef T& type; static type make(T* p, std::size_t const*) { return *p; } };
template <typename T, std::size_t R>
class view_iterator {
public:
    typedef std::forward_iterator_tag iterator_category;
    typedef typename view_ref<T, R - 1>::type reference;
    typedef std::ptrdiff_t difference_type;
    typedef void value_type;
    typedef void pointer;
    view_iterator(T* p, std::size_t const* shape, std::size_t stride) : p_(p), shape_(shape), stride_(stride) {}
    reference operator*() const { return view_ref<T, R - 1>::make(p_, shape_); }
    view_iterator& operator++() { p_ += stride_; return *this; }
    view_iterator operator++(int) { view_iterator t(*this); p_ += stride_; return t; }
    bool operator==(view_iterator const& o) const { return p_ == o.p_; }
    bool operator!=(view_iterator const& o) const { return p_ != o.p_; }
private:
    T* p_;
    std::size_t const* shape_;      // shape of the sub-views this iterator yields
    std::size_t stride_;
};
template <typename T, std::size_t R>
class view {
public:
    typedef view_iterator<T, R> iterator;
    typedef std::size_t size_type;
    view(T* p, std::size_t const* shape) : p_(p), shape_(shape) {}
    std::size_t stride() const { std::size_t s = 1; for (std::size_t i = 1; i < R; ++i) s *= shape_[i]; return s; }
    typename view_ref<T, R - 1>::type operator[](std::size_t i) const { return view_ref<T, R - 1>::make(p_ + i * stride(), shape_ + 1); }
    iterator begin() const { return iterator(p_, shape_ + 1, stride()); }
    iterator end() const { return iterator(p_ + shape_[0] * stride(), shape_ + 1, stride()); }
    std::size_t size() const { return shape_[0]; }
private:
    T* p_;
    std::size_t const* shape_;
};
} // namespace detail_ma
static detail_ma::extent_gen<0> const extents = detail_ma::extent_gen<0>();

template <typename T, std::size_t N>
class multi_array {
public:
    typedef std::size_t size_type;
    typedef T element;
    typedef detail_ma::view_iterator<T, N> iterator;
    multi_array() { shape_.fill(0); }
    explicit multi_array(detail_ma::extent_gen<N> const& e) { resize(e); }
    void resize(detail_ma::extent_gen<N> const& e) { resize(e.e); }
    template <typename Extents> void resize(Extents const& e)
    {
        std::size_t n = 1;
        for (std::size_t i = 0; i < N; ++i) { shape_[i] = e[i]; n *= shape_[i]; }
        data_.assign(n, T());
    }
    std::size_t const* shape() const { return shape_.data(); }
    std::size_t num_elements() const { return data_.size(); }
    std::size_t size() const { return shape_[0]; }
    T* origin() { return data_.data(); }
    T const* origin() const { return data_.data(); }
    T* data() { return data_.data(); }
    T const* data() const { return data_.data(); }
    detail_ma::view<T, N> whole() { return detail_ma::view<T, N>(data_.data(), shape_.data()); }
    typename detail_ma::view_ref<T, N - 1>::type operator[](std::size_t i) { return whole()[i]; }
    typename detail_ma::view_ref<T const, N - 1>::type operator[](std::size_t i) const
    {
        return detail_ma::view<T const, N>(data_.data(), shape_.data())[i];
    }
    iterator begin() { return whole().begin(); }
    iterator end() { return whole().end(); }
private:
    std::array<std::size_t, N> shape_;
    std::vector<T> data_;
};
} // namespace boost
#endif
