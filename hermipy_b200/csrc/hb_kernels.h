// Launchers of the helper kernels in hb_kernels.cu (internal, not part of the C ABI).
#pragma once
#include "hb_internal.h"

#define HB_MAX_DIM 8
#define HB_MAX_FACTORS 8

namespace hb {

int launch_basis_tables(const double* x, const double* w, int n, int P, int fourier, double* plain, double* wgt,
                        long long ld_ik, double* plainT, double* wgtT, long long ld_ki, cudaStream_t s);
int launch_psi_table(const double* x, const double* w, const double* sqrt_w, int n, int P, int fourier, double* psi,
                     long long ld, cudaStream_t s);
int launch_pair_table(const double* plain, long long ld_ik, const double* w, int n, int P, int Pp, double* out,
                      cudaStream_t s);
int launch_triple_products(int N, int Pp, double* T, cudaStream_t s);
int launch_blocked_pair_table(int mode, const double* src, long long ld, long long plane, const double* w, int rows,
                              const int* slots, int nb, double* out, cudaStream_t s);
int launch_axis_parity_split(const double* in, long long in_stride, int n, int nh, long long cols, int batch,
                             double* be, double* bo, cudaStream_t s);
int launch_axis_parity_merge(const double* ge, const double* go, int n, int nh, long long cols, int batch, double* out,
                             long long out_stride, cudaStream_t s);
int launch_rows_parity_split(const double* f, long long f_stride, int n, int nh, double* fe, double* fo,
                             long long h_stride, int batch, cudaStream_t s);
int launch_rows_parity_merge(const double* ge, const double* go, long long h_stride, int n, int nh, int n_pad,
                             double* f, long long f_stride, int batch, cudaStream_t s);
int launch_strided_copy(const double* src, long long src_ld, int r0, int rs, int c0, int cs, double* dst,
                        long long dst_ld, long long rows, int cols, int batch, long long src_bstride,
                        long long dst_bstride, cudaStream_t s);
int launch_parity_split(const double* f, int n, int nh, int seg, long long pitch, double* out, cudaStream_t s);

// Parity split / merge of a d-dimensional grid function along up to three axes AT ONCE (one pass over the data).
// "Natural" array: axis a has n[a] entries nat_stride[a] apart.  "Sorted" array: a split axis is indexed by
// (parity p, h) with h < nh[a] = ceil(n[a]/2) counting the non-negative nodes x = n - nh + h; the entry sits at
// h * srt_stride[a] + p * srt_par[a].  Other axes keep their index (srt_stride[a] apart).
//   split:  sorted[p..][h..] = sum over the sign combinations of (+-1)^(p.s) natural[x(+-)..]   (even part: f(x)+f(-x))
//   merge:  natural[x(+-)..] = sum over parities of (+-1)^(p.s) sorted[p..][h..]               (f(+-x) = g_e +- g_o)
// The middle node of an odd rule is taken once: its odd part is 0 * f (NaN / inf stay visible, as in 0 * f * w * phi).
struct ParityNDArgs {
    int d = 0;
    int n[HB_MAX_DIM] = {0}, nh[HB_MAX_DIM] = {0}, split[HB_MAX_DIM] = {0};
    long long nat_stride[HB_MAX_DIM] = {0}, srt_stride[HB_MAX_DIM] = {0}, srt_par[HB_MAX_DIM] = {0};
    long long batch = 1, nat_bstride = 0, srt_bstride = 0;
    int nat_pad_last = 0;   // merge: columns n[d-1] .. nat_pad_last-1 of every natural row are zero-filled
};
int launch_nd_parity_split(const ParityNDArgs& a, const double* natural, double* sorted, cudaStream_t s);
int launch_nd_parity_merge(const ParityNDArgs& a, const double* sorted, double* natural, cudaStream_t s);
int launch_deinterleave_rows(const double* src, long long src_ld, int Pe, int Po, double* even, long long lde, double* odd,
                             long long ldo, long long rows, cudaStream_t s);
int launch_repack_rows(const double* src, long long src_ld, double* dst, long long dst_ld, long long rows, int cols,
                       cudaStream_t s);
int launch_gather_lut(const double* src, long long src_stride, const int* lut, double* dst, long long dst_stride,
                      long long n_dst, int batch, cudaStream_t s);

struct TensorizeArgs {
    int k;
    const double* in[HB_MAX_FACTORS];
    const int* map[HB_MAX_FACTORS];
    int size[HB_MAX_FACTORS];
};
int launch_tensorize(const TensorizeArgs& a, bool matrices, double* out, int n_out, cudaStream_t s);
int launch_project(const double* in, int n_in, const int* sel, bool matrix, double* out, int n_out, cudaStream_t s);
int launch_varfd(const double* var, long long ldv, const int* src, const double* fac, double* out, long long ldo,
                 int rows, int n, cudaStream_t s);
// A[r][c] = full[sum_a (m_a(r) * Pp_a + n_a(c)) * stride_a]  (d-dimensional dense varf, last step)
int launch_varf_gather_nd(int dim, int n_polys, int row_begin, int row_end, const int* Pp, const long long* stride,
                          const int* midx, const double* full, double* out, long long ldo, cudaStream_t s);
// out[b][i] = a[b][i] / (f[i] == 0 ? 1e-300 : f[i]) (op 0) or a[b][i] * f[i] (op 1): weighting steps of the object API
int launch_pointwise(int op, const double* a, long long a_stride, const double* b, double* out, long long o_stride,
                     long long n, int batch, cudaStream_t s);
// y = A x for a row-major M x N matrix resident in HBM (hb_matvec.cu)
int launch_matvec(int M, int N, const double* A, long long lda, const double* x, double* y, cudaStream_t s);

struct VarfSparseArgs {
    int dim, n_polys, n_kept;
    int row_begin, row_end;
    long long ldo;
    const int* midx;        // n_polys x dim multi-indices (device)
    const int* kept_alpha;  // n_kept x dim
    const double* kept_val; // n_kept
    const double* T[HB_MAX_DIM];  // per-axis triple-product table [k][i][Pp]
    long long plane[HB_MAX_DIM];  // (N+1)*Pp
    int Pp[HB_MAX_DIM];
};
int launch_varf_sparse(const VarfSparseArgs& a, double* out, cudaStream_t s);
struct VarfBandArgs {
    int amax[HB_MAX_DIM];     // largest retained alpha per axis = half band width
    int extent[HB_MAX_DIM];   // bounding box of the index set
    int pitch_last;           // padded extent of the last axis in lut_box
    long long width_total;    // prod (2*amax+1)
    const int* lut_box;       // padded lexicographic box -> position (-1 absent)
};
struct CsrResult;
// the band of launch_varf_band as a CSR result (rows relative to a.row_begin), no dense matrix (hb_sparse.cu)
int varf_band_csr(const VarfSparseArgs& a, const VarfBandArgs& b, CsrResult& out, cudaStream_t s);
int launch_varf_band(const VarfSparseArgs& a, const VarfBandArgs& b, double* out, cudaStream_t s,
                     bool zero_filled = false);   // zero_filled: the caller has already cleared the rows

int launch_inner(const double* s1, const double* s2, const int* ind1, const int* ind2, int n_res, int n_inner,
                 double* out, cudaStream_t s);
int launch_count_nnz(const double* A, long long ld, int n_rows, int n_cols, long long* row_nnz, cudaStream_t s);
int launch_fill_csr(const double* A, long long ld, int n_rows, int n_cols, const long long* indptr, int* indices,
                    double* data, cudaStream_t s);

}  // namespace hb
