// Shim for the tiny part of boost::multi_array the reference touches outside its Python bridge:
// `cmat` (types.hpp:53) constructed from boost::extents[a][b] in matrix.cpp:44-48.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
namespace detail_shim {
struct extent2 { std::size_t a, b; };
struct extent1 { std::size_t a; extent2 operator[](std::size_t b) const { return extent2{a, b}; } };
struct extent0 { extent1 operator[](std::size_t a) const { return extent1{a}; } };
}
static const detail_shim::extent0 extents = detail_shim::extent0();
template <typename T, std::size_t N>
class multi_array {
    std::size_t r_, c_;
    std::vector<T> d_;
  public:
    multi_array() : r_(0), c_(0) {}
    multi_array(const detail_shim::extent2& e) : r_(e.a), c_(e.b), d_(e.a * e.b, T()) {}
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    const std::size_t* shape() const { return &r_; }
};
}
