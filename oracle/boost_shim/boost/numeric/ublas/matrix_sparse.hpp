// Functional stand-in for boost::numeric::ublas::compressed_matrix<double, row_major>, covering the
// members the reference uses (tensorize.cpp:197-283, varf.cpp:235-261,290-376, sparse.cpp:28-100,
// project.cpp:78-109): (s1,s2[,nnz]) construction, m(i,j) as l-value (inserts) and r-value (0 when
// absent), size1/size2/nnz, row-wise traversal in increasing column order, A + B (structural union),
// A * scalar.  Storage: one ordered map per row.  Test infrastructure only (oracle build).
#pragma once
#include <cstddef>
#include <map>
#include <vector>
#include "matrix.hpp"
namespace boost { namespace numeric { namespace ublas {

template <typename T, typename L = row_major>
class compressed_matrix {
    typedef std::map<std::size_t, T> row_t;
    std::size_t r_, c_;
    std::vector<row_t> rows_;
  public:
    compressed_matrix() : r_(0), c_(0) {}
    compressed_matrix(std::size_t r, std::size_t c, std::size_t /*nnz*/ = 0) : r_(r), c_(c), rows_(r) {}
    std::size_t size1() const { return r_; }
    std::size_t size2() const { return c_; }
    std::size_t nnz() const {
        std::size_t n = 0;
        for (const row_t& r : rows_) n += r.size();
        return n;
    }

    class reference {
        compressed_matrix* m_; std::size_t i_, j_;
      public:
        reference(compressed_matrix* m, std::size_t i, std::size_t j) : m_(m), i_(i), j_(j) {}
        reference& operator=(T v) { m_->rows_[i_][j_] = v; return *this; }
        operator T() const {
            typename row_t::const_iterator it = m_->rows_[i_].find(j_);
            return it == m_->rows_[i_].end() ? T() : it->second;
        }
    };
    reference operator()(std::size_t i, std::size_t j) { return reference(this, i, j); }
    T operator()(std::size_t i, std::size_t j) const {
        typename row_t::const_iterator it = rows_[i].find(j);
        return it == rows_[i].end() ? T() : it->second;
    }

    class const_iterator2 {
        std::size_t i_; typename row_t::const_iterator it_;
      public:
        const_iterator2(std::size_t i, typename row_t::const_iterator it) : i_(i), it_(it) {}
        std::size_t index1() const { return i_; }
        std::size_t index2() const { return it_->first; }
        const T& operator*() const { return it_->second; }
        const_iterator2& operator++() { ++it_; return *this; }
        bool operator!=(const const_iterator2& o) const { return it_ != o.it_; }
    };
    class const_iterator1 {
        const compressed_matrix* m_; std::size_t i_;
      public:
        const_iterator1(const compressed_matrix* m, std::size_t i) : m_(m), i_(i) {}
        const_iterator2 begin() const { return const_iterator2(i_, m_->rows_[i_].begin()); }
        const_iterator2 end() const { return const_iterator2(i_, m_->rows_[i_].end()); }
        const_iterator1& operator++() { ++i_; return *this; }
        bool operator!=(const const_iterator1& o) const { return i_ != o.i_; }
    };
    typedef const_iterator1 iterator1;
    typedef const_iterator2 iterator2;
    const_iterator1 begin1() const { return const_iterator1(this, 0); }
    const_iterator1 end1() const { return const_iterator1(this, r_); }

    template <typename U, typename M> friend compressed_matrix<U, M> operator+(const compressed_matrix<U, M>&, const compressed_matrix<U, M>&);
    template <typename U, typename M> friend compressed_matrix<U, M> operator*(const compressed_matrix<U, M>&, U);
};

template <typename T, typename L>
compressed_matrix<T, L> operator+(const compressed_matrix<T, L>& a, const compressed_matrix<T, L>& b) {
    compressed_matrix<T, L> r(a);
    for (std::size_t i = 0; i < b.rows_.size(); i++)
        for (typename std::map<std::size_t, T>::const_iterator it = b.rows_[i].begin(); it != b.rows_[i].end(); ++it)
            r.rows_[i][it->first] += it->second;
    return r;
}
template <typename T, typename L>
compressed_matrix<T, L> operator*(const compressed_matrix<T, L>& a, T s) {
    compressed_matrix<T, L> r(a);
    for (std::size_t i = 0; i < r.rows_.size(); i++)
        for (typename std::map<std::size_t, T>::iterator it = r.rows_[i].begin(); it != r.rows_[i].end(); ++it)
            it->second *= s;
    return r;
}
}}}
