// Functional stand-in for boost::numeric::ublas::matrix<double> (dense, row-major) covering exactly
// the members the reference's varf.cpp / matrix.cpp / tensorize.cpp / project.cpp use:
// (r,c[,init]) construction, (i,j) access, size1/size2, A + B, A * scalar, and the
// begin1()/end1() -> begin()/end() -> index1()/index2()/operator* traversal (varf.cpp:342-347).
// Test infrastructure only (oracle build); never linked into the product library.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <iostream>
#include <memory>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
struct row_major {};

template <typename T>
class matrix {
    std::size_t r_, c_;
    std::vector<T> d_;
  public:
    matrix() : r_(0), c_(0) {}
    matrix(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c, T()) {}
    matrix(std::size_t r, std::size_t c, T init) : r_(r), c_(c), d_(r * c, init) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    T& operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    const T& operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
    std::vector<T>& data() { return d_; }
    const std::vector<T>& data() const { return d_; }

    class const_iterator2 {
        const matrix* m_; std::size_t i_, j_;
      public:
        const_iterator2(const matrix* m, std::size_t i, std::size_t j) : m_(m), i_(i), j_(j) {}
        std::size_t index1() const { return i_; }
        std::size_t index2() const { return j_; }
        const T& operator*() const { return (*m_)(i_, j_); }
        const_iterator2& operator++() { ++j_; return *this; }
        bool operator!=(const const_iterator2& o) const { return j_ != o.j_ || i_ != o.i_; }
    };
    class const_iterator1 {
        const matrix* m_; std::size_t i_;
      public:
        const_iterator1(const matrix* m, std::size_t i) : m_(m), i_(i) {}
        const_iterator2 begin() const { return const_iterator2(m_, i_, 0); }
        const_iterator2 end() const { return const_iterator2(m_, i_, m_->size2()); }
        const_iterator1& operator++() { ++i_; return *this; }
        bool operator!=(const const_iterator1& o) const { return i_ != o.i_; }
    };
    typedef const_iterator1 iterator1;
    typedef const_iterator2 iterator2;
    const_iterator1 begin1() const { return const_iterator1(this, 0); }
    const_iterator1 end1() const { return const_iterator1(this, r_); }
};

template <typename T>
matrix<T> operator+(const matrix<T>& a, const matrix<T>& b) {
    matrix<T> r(a.size1(), a.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t j = 0; j < a.size2(); j++) r(i, j) = a(i, j) + b(i, j);
    return r;
}
template <typename T>
matrix<T> operator*(const matrix<T>& a, T s) {
    matrix<T> r(a.size1(), a.size2());
    for (std::size_t i = 0; i < a.size1(); i++)
        for (std::size_t j = 0; j < a.size2(); j++) r(i, j) = a(i, j) * s;
    return r;
}
}}}
