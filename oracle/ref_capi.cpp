// TEST INFRASTRUCTURE ONLY -- C-ABI wrapper around the reference's OWN C++ sources.
//
// This file contains no algorithm: every function forwards to hermite::* as compiled from
// /root/reference/cpp/src/hermite/*.cpp (unmodified, where they lie) against the header shim in
// oracle/boost_shim/.  The resulting oracle/_ref/libhermite_ref.so is the strongest oracle available:
// the reference's own object code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  See oracle/Makefile.
#include <cstring>
#include <string>
#include <vector>

#include "hermite/discretize.hpp"
#include "hermite/hermite.hpp"
#include "hermite/inner.hpp"
#include "hermite/iterators.hpp"
#include "hermite/matrix.hpp"
#include "hermite/project.hpp"
#include "hermite/sparse.hpp"
#include "hermite/tensorize.hpp"
#include "hermite/transform.hpp"
#include "hermite/types.hpp"
#include "hermite/varf.hpp"

using namespace hermite;

namespace {
mat unflatten(const double* flat, const int* lens, int dim) {
    mat out(dim);
    size_t off = 0;
    for (int d = 0; d < dim; d++) {
        out[d].assign(flat + off, flat + off + lens[d]);
        off += lens[d];
    }
    return out;
}
mat to_mat(const double* a, int rows, int cols) {
    mat m(rows, vec(cols));
    for (int i = 0; i < rows; i++) m[i].assign(a + (size_t)i * cols, a + (size_t)(i + 1) * cols);
    return m;
}
void from_mat(const mat& m, double* out) {
    size_t k = 0;
    for (const vec& r : m)
        for (double v : r) out[k++] = v;
}
ivec to_ivec(const int* a, int n) { return ivec(a, a + n); }
imat to_imat(const int* flat, const int* lens, int n) {
    imat out(n);
    size_t off = 0;
    for (int i = 0; i < n; i++) {
        out[i].assign(flat + off, flat + off + lens[i]);
        off += lens[i];
    }
    return out;
}
}  // namespace

extern "C" {

// hermite.cpp:29-41
void ref_hermite_eval(double x, unsigned degree, double* out) {
    vec v(degree + 2, 0.);  // +1 slot: the reference writes values[1] even for degree 0
    hermite_eval(x, degree, v);
    memcpy(out, v.data(), (degree + 1) * sizeof(double));
}

// discretize.cpp:112-126 (writes /tmp/<hash>.cpp, compiles it with the host c++ and dlopen()s the result)
int ref_discretize(const char* body, int dim, const int* n_pts, const double* nodes, const double* translation,
                   const double* dilation, double* out, int out_len) {
    vec r = discretize_from_string(body, unflatten(nodes, n_pts, dim), vec(translation, translation + dim),
                                   to_mat(dilation, dim, dim));
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}

// transform.cpp:151-168
int ref_transform(unsigned degree, const double* f, int n_f, int dim, const int* n_pts, const double* nodes,
                  const double* weights, const int* do_fourier, int forward, const char* index_set,
                  double* out, int out_len) {
    vec r = transform(degree, vec(f, f + n_f), unflatten(nodes, n_pts, dim), unflatten(weights, n_pts, dim),
                      to_ivec(do_fourier, dim), forward != 0, index_set);
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}

// transform.cpp:170-181
double ref_integrate(const double* f, int n_f, int dim, const int* n_pts, const double* nodes,
                     const double* weights, const int* do_fourier) {
    return integrate(vec(f, f + n_f), unflatten(nodes, n_pts, dim), unflatten(weights, n_pts, dim),
                     to_ivec(do_fourier, dim));
}

// iterators.cpp: s_size / s_get_degree / s_list / s_index
static int set_id(const char* s) {
    std::string n(s);
    if (n == "triangle") return 0;
    if (n == "cross") return 1;
    if (n == "cross_nc") return 2;
    if (n == "cube") return 3;
    if (n == "rectangle") return 4;
    return -1;
}
unsigned ref_index_size(const char* set, unsigned dim, unsigned degree) {
    switch (set_id(set)) {
        case 0: return Triangle_iterator::s_size(dim, degree);
        case 1: return Cross_iterator::s_size(dim, degree);
        case 2: return Cross_iterator_nc::s_size(dim, degree);
        case 3: return Cube_iterator::s_size(dim, degree);
        case 4: return Rectangle_iterator::s_size(dim, degree);
    }
    return 0;
}
unsigned ref_index_get_degree(const char* set, unsigned dim, unsigned n_polys) {
    switch (set_id(set)) {
        case 0: return Triangle_iterator::s_get_degree(dim, n_polys);
        case 1: return Cross_iterator::s_get_degree(dim, n_polys);
        case 2: return Cross_iterator_nc::s_get_degree(dim, n_polys);
        case 3: return Cube_iterator::s_get_degree(dim, n_polys);
        case 4: return Rectangle_iterator::s_get_degree(dim, n_polys);
    }
    return 0;
}
// writes size*dim unsigned ints (row-major list); returns the number of multi-indices
unsigned ref_index_list(const char* set, unsigned dim, unsigned degree, unsigned* out, unsigned cap) {
    imat l;
    switch (set_id(set)) {
        case 0: l = Triangle_iterator::s_list(dim, degree); break;
        case 1: l = Cross_iterator::s_list(dim, degree); break;
        case 2: l = Cross_iterator_nc::s_list(dim, degree); break;
        case 3: l = Cube_iterator::s_list(dim, degree); break;
        case 4: l = Rectangle_iterator::s_list(dim, degree); break;
        default: return 0;
    }
    if (l.size() <= cap)
        for (size_t i = 0; i < l.size(); i++)
            for (unsigned d = 0; d < dim; d++) out[i * dim + d] = l[i][d];
    return (unsigned)l.size();
}
unsigned ref_triangle_index(const unsigned* m, unsigned dim) { return Triangle_iterator::s_index(ivec(m, m + dim)); }
unsigned ref_cube_index(const unsigned* m, unsigned dim) { return Cube_iterator::s_index(ivec(m, m + dim)); }

// varf.cpp:49-94 ; out is (2N+1)*(N+1)*(N+1)
void ref_triple_products(int degree, double* out) {
    cube t = triple_products_1d(degree);
    size_t k = 0;
    for (const mat& m : t)
        for (const vec& r : m)
            for (double v : r) out[k++] = v;
}
void ref_triple_products_fourier(int degree, double* out) {
    cube t = triple_products_fourier(degree);
    size_t k = 0;
    for (const mat& m : t)
        for (const vec& r : m)
            for (double v : r) out[k++] = v;
}

// varf.cpp:271-288, dense result (the module exposes varf<boost_mat>, module.cpp:101)
int ref_varf(unsigned degree, const double* f, int n_f, int dim, const int* n_pts, const double* nodes,
             const double* weights, const int* do_fourier, const char* index_set, double* out, int n_polys) {
    boost_mat r = varf<boost_mat>(degree, vec(f, f + n_f), unflatten(nodes, n_pts, dim),
                                  unflatten(weights, n_pts, dim), to_ivec(do_fourier, dim), index_set);
    if ((int)r.size1() != n_polys) return (int)r.size1();
    for (int i = 0; i < n_polys; i++)
        for (int j = 0; j < n_polys; j++) out[(size_t)i * n_polys + j] = r(i, j);
    return 0;
}
// sparse result (varf_sp, module.cpp:102): returns nnz; writes (row, col, val) triplets up to cap
int ref_varf_sp(unsigned degree, const double* f, int n_f, int dim, const int* n_pts, const double* nodes,
                const double* weights, const int* do_fourier, const char* index_set, int* rows, int* cols,
                double* vals, int cap) {
    spmat r = varf<spmat>(degree, vec(f, f + n_f), unflatten(nodes, n_pts, dim), unflatten(weights, n_pts, dim),
                          to_ivec(do_fourier, dim), index_set);
    mat rcv = row_col_val(r);
    int nnz = (int)rcv[0].size();
    for (int i = 0; i < nnz && i < cap; i++) {
        rows[i] = (int)rcv[0][i];
        cols[i] = (int)rcv[1][i];
        vals[i] = rcv[2][i];
    }
    return nnz;
}

// varf.cpp:290-376 (boost_mat instantiation, module.cpp:103)
void ref_varfd(unsigned dim, unsigned degree, unsigned direction, const double* var, int n, unsigned do_fourier,
               const char* index_set, double* out) {
    boost_mat v(n, n, 0.);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) v(i, j) = var[(size_t)i * n + j];
    boost_mat r = varfd<boost_mat>(dim, degree, direction, v, do_fourier, index_set);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[(size_t)i * n + j] = r(i, j);
}

// tensorize.cpp:124-152 ; inputs flattened with lens[]; dirs flattened with dir_lens[]
int ref_tensorize_vecs_dirs(int k, const double* inputs, const int* lens, const int* dirs, const int* dir_lens,
                            const char* index_set, double* out, int out_len) {
    vec r = tensorize_vecs_dirs(unflatten(inputs, lens, k), to_imat(dirs, dir_lens, k), index_set);
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}
int ref_tensorize_vec_id(const double* input, int len, unsigned dim, unsigned dir, const char* index_set,
                         double* out, int out_len) {
    vec r = tensorize_vec_id(vec(input, input + len), dim, dir, index_set);
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}
// tensorize.cpp:285-357 ; k square matrices of sizes[i]
int ref_tensorize_mats_dirs(int k, const double* inputs, const int* sizes, const int* dirs, const int* dir_lens,
                            const char* index_set, double* out, int n_out) {
    std::vector<mat> in(k);
    size_t off = 0;
    for (int i = 0; i < k; i++) {
        in[i] = to_mat(inputs + off, sizes[i], sizes[i]);
        off += (size_t)sizes[i] * sizes[i];
    }
    mat r = tensorize_mats_dirs<mat, mat>(in, to_imat(dirs, dir_lens, k), index_set);
    if ((int)r.size() != n_out) return (int)r.size();
    from_mat(r, out);
    return 0;
}
int ref_tensorize_mat_id(const double* input, int n, unsigned dim, unsigned dir, const char* index_set, double* out,
                         int n_out) {
    mat r = tensorize_mat_id<mat, mat>(to_mat(input, n, n), dim, dir, index_set);
    if ((int)r.size() != n_out) return (int)r.size();
    from_mat(r, out);
    return 0;
}

// project.cpp:60-75,111-129
int ref_project_vec(const double* input, int len, unsigned dim, const int* dirs, int n_dirs, const char* index_set,
                    double* out, int out_len) {
    vec r = project_vec_nd(vec(input, input + len), dim, to_ivec(dirs, n_dirs), index_set);
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}
int ref_project_mat(const double* input, int n, unsigned dim, const int* dirs, int n_dirs, const char* index_set,
                    double* out, int n_out) {
    mat r = project_mat_nd<mat>(to_mat(input, n, n), dim, to_ivec(dirs, n_dirs), index_set);
    if ((int)r.size() != n_out) return (int)r.size();
    from_mat(r, out);
    return 0;
}

// inner.cpp:219-233
int ref_inner(const double* s1, int n1, const double* s2, int n2, const int* d1, int nd1, const int* d2, int nd2,
              const char* index_set, double* out, int out_len) {
    vec r = inner(vec(s1, s1 + n1), vec(s2, s2 + n2), to_ivec(d1, nd1), to_ivec(d2, nd2), index_set);
    if ((int)r.size() != out_len) return (int)r.size();
    memcpy(out, r.data(), r.size() * sizeof(double));
    return 0;
}

}  // extern "C"
