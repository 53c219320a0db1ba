/* hermite_b200.h -- C ABI of libhermite_b200.so
 *
 * B200-native (sm_100a CUDA) replacement for the numerical core of urbainvaes/hermipy.  Every entry
 * point below replaces one function that the reference registers in its Boost.Python module
 * `hermite_cpp` (reference: cpp/src/hermite/python/module.cpp:45-168) and that the reference's
 * hermipy/core.py:125-348 calls; the citation next to each declaration names that function.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point data is IEEE double, all index data int32/uint32;
 *   - `index_set` is one of "triangle", "cross", "cross_nc", "cube", "rectangle" (transform.cpp:160-166);
 *   - grids are row-major with the LAST axis fastest (Grid_iterator, iterators.cpp:36-66);
 *   - `nodes` / `weights` are the per-axis arrays concatenated (axis 0 first), lengths in n_pts[dim];
 *   - every function returns an int status (HB_OK or an HB_ERR_* code) and never calls exit();
 *     the conditions on which the reference prints a message and exit()s map to error codes;
 *   - functions without the _dev suffix take HOST pointers and include the host<->device copies;
 *     hb_plan_* functions take DEVICE pointers (resident data) and a CUDA stream.
 *   - there is no CPU fallback: with no usable CUDA device every compute entry point returns HB_ERR_CUDA.
 */
#ifndef HERMITE_B200_H
#define HERMITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------------ */
#define HB_OK 0
#define HB_ERR_INVALID 1        /* bad argument (reference: message + exit(1), e.g. tensorize.cpp:50,64,79) */
#define HB_ERR_INDEX_SET 2      /* unknown index set ("Invalid index set!" + exit(1), transform.cpp:166) */
#define HB_ERR_FOURIER_DEGREE 3 /* odd degree on a Fourier axis (transform.cpp:92-96, varf.cpp:103-107) */
#define HB_ERR_DEGREE 4         /* n_polys matches no degree ("Can't find x" + exit(0), lib.cpp:93-99) */
#define HB_ERR_CUDA 5           /* CUDA runtime/driver failure or no device; see hb_last_error() */
#define HB_ERR_ALIGN 6          /* device operand violates the 16-byte pitch rule of the TMA path */
#define HB_ERR_NOMEM 7          /* device or caller buffer too small */

const char* hb_last_error(void);  /* thread-local text of the last failure ("" if none) */
const char* hb_version(void);
int hb_device_count(int* count);  /* HB_ERR_CUDA if no driver */

/* ---- multi-index sets (host only, usable without a GPU) ------------------------------------------
 * reference: {triangle,cross,cross_nc,cube,rectangle}_{size,get_degree,list_indices}, triangle_index,
 * cube_index  (module.cpp:143-166; iterators.cpp:106-596). */
int hb_index_size(const char* index_set, int dim, int degree, long long* size);
int hb_index_get_degree(const char* index_set, int dim, long long n_polys, int* degree);
int hb_index_list(const char* index_set, int dim, int degree, uint32_t* out, size_t out_capacity /* entries */);
int hb_triangle_index(const uint32_t* m, int dim, uint32_t* index);
int hb_cube_index(const uint32_t* m, int dim, uint32_t* index);
/* positions of `count` multi-indices (count x dim, row-major) in the enumeration of the set; -1 where the
 * multi-index is not in the set (the reference looks one up at a time in its hash map, lib.cpp:76-84) */
int hb_index_positions(const char* index_set, int dim, int degree, const uint32_t* m, size_t count, int32_t* pos);

/* ---- Hermite functions by three-term recurrence ---------------------------------------------------
 * reference: hermite_eval (hermite.cpp:29-41).  out[i*(degree+1) + k] = He_k(x_i)/sqrt(k!). */
int hb_hermite_eval(const double* x, int n, int degree, double* out);

/* ---- discretize --------------------------------------------------------------------------------------
 * reference: discretize(function_body, nodes, translation, dilation) (module.cpp:88,
 * discretize.cpp:112-126).  `body` is a C expression in v[0..dim-1] (what sympy.ccode produces after
 * hermipy/function.py:118-123); it is evaluated at translation + dilation * node for every node of the
 * tensor grid (dilation: dim x dim row-major; NULL = identity, translation NULL = 0).  The expression
 * is compiled at run time with NVRTC for sm_100a (the reference shells out to the host C++ compiler). */
int hb_discretize(const char* body, int dim, const int* n_pts, const double* nodes, const double* translation,
                  const double* dilation, double* out, size_t out_len);

/* ---- transform ---------------------------------------------------------------------------------------
 * reference: transform(degree, f_grid, nodes, weights, do_fourier, forward, index_set)
 * (module.cpp:92, transform.cpp:151-168).  forward: f has prod(n_pts) entries, out has |I| entries;
 * backward: f has |I| entries, out has prod(n_pts).  hb_transform_batch does `batch` independent
 * functions in one call (inputs/outputs stored contiguously one after another); the reference has no
 * batched entry point (callers loop in Python). */
int hb_transform(int dim, const int* n_pts, int degree, const double* f, const double* nodes,
                 const double* weights, const int* do_fourier, int forward, const char* index_set, double* out,
                 size_t out_len);
int hb_transform_batch(int dim, const int* n_pts, int degree, int batch, const double* f, const double* nodes,
                       const double* weights, const int* do_fourier, int forward, const char* index_set,
                       double* out, size_t out_len_each);
/* reference: integrate(f_grid, nodes, weights, do_fourier) (module.cpp:91, transform.cpp:170-181) */
int hb_integrate(int dim, const int* n_pts, const double* f, const double* nodes, const double* weights,
                 const int* do_fourier, double* result);

/* ---- triple products and varf ----------------------------------------------------------------------
 * reference: triple_products(degree) / triple_products_fourier(degree) (module.cpp:98-99,
 * varf.cpp:49-162): out is (2*degree+1) x (degree+1) x (degree+1), row-major. */
int hb_triple_products(int degree, double* out);
int hb_triple_products_fourier(int degree, double* out);

#define HB_VARF_REFERENCE 0 /* the reference's algebra: thresholded sum over alpha of Hf[alpha] * (x) T[alpha_d] */
#define HB_VARF_GRAM 1      /* quadrature Gram form sum_x w f phi_m phi_n (dense FP64 tensor-core contraction) */
/* reference: varf(degree, f_grid, nodes, weights, do_fourier, index_set) -> boost_mat (module.cpp:101,
 * varf.cpp:164-288).  out is n_polys x n_polys row-major. */
int hb_varf(int dim, const int* n_pts, int degree, const double* f, const double* nodes, const double* weights,
            const int* do_fourier, const char* index_set, int mode, double* out, size_t n_polys);
/* reference: varf_sp(...) -> sparse_matrix + row_col_val (module.cpp:102,133; sparse.cpp:28-43).
 * Two-step: hb_varf_sp computes and returns nnz; hb_sparse_fetch copies the CSR arrays of the last
 * sparse result of this process (indptr has n_rows+1 entries).  For polynomial-like f (few retained alpha: a
 * narrow band of triple products per row) the band is emitted straight as CSR -- no n_polys^2 buffer, no
 * zero-fill; otherwise the dense matrix is assembled on the device and compacted. */
int hb_varf_sp(int dim, const int* n_pts, int degree, const double* f, const double* nodes, const double* weights,
               const int* do_fourier, const char* index_set, int mode, long long* nnz);
int hb_sparse_fetch(long long* indptr, int* indices, double* data);
/* Sparse in, sparse out: the spmat instantiations of the reference (tensorize.cpp:197-283 folded over the factors
 * :286-332; project.cpp:78-109; varfd varf.cpp:290-376).  Inputs are CSR triplets (indptr int64 with n+1 entries,
 * indices int32, data); the result is CSR with sorted columns and exact zeros dropped (sparse.cpp:52), left on the
 * device like the result of hb_varf_sp: *nnz (and *n_out, its number of rows) are returned, hb_sparse_fetch copies
 * it out.  Work and memory are proportional to the stored entries, never to n_polys^2.
 * The sparse result is ONE per process ("last result"): pair each call with its hb_sparse_fetch under a lock when
 * several host threads use the library. */
int hb_tensorize_mats_sp(int k, const long long* const* indptr, const int* const* indices, const double* const* data,
                         const int* sizes, const int* dirs_flat, const int* dir_lens, const char* index_set,
                         long long* n_out, long long* nnz);
int hb_project_mat_sp(const long long* indptr, const int* indices, const double* data, size_t n_polys, int dim,
                      const int* dirs, int n_dirs, const char* index_set, long long* n_out, long long* nnz);
int hb_varfd_sp(int dim, int degree, int direction, const long long* indptr, const int* indices, const double* data,
                size_t n_polys, int do_fourier, const char* index_set, long long* nnz);
/* The object API's transform / eval / varf in ONE call each, so that the weighting steps run as device kernels and
 * nothing but the inputs and the result crosses PCIe (reference: hermipy/quad.py:287-342, where the function and the
 * factor of the quadrature are discretised, divided / multiplied with NumPy and only then handed to hermite_cpp).
 * Expressions are C in v[i] (Function.ccode()), evaluated at translation + dilation @ node.
 *   hb_transform_weighted: coefficients of  f / factor  (f: batch grid functions on the host, or ONE expression;
 *                          factor_body may be NULL = 1; the division is f / (factor == 0 ? 1e-300 : factor), quad.py:299)
 *   hb_eval_weighted:      values of the series on eval_nodes (the quadrature's nodes mapped into the frame of the
 *                          series) times the factor evaluated on the quadrature's own nodes (quad.py:316-328)
 *   hb_varf_function[_sp]: hb_varf[_sp] of an expression */
int hb_transform_weighted(int dim, const int* n_pts, int degree, int batch, const double* f, const char* f_body,
                          const char* factor_body, const double* nodes, const double* weights, const double* translation,
                          const double* dilation, const int* do_fourier, const char* index_set, double* out,
                          size_t out_len_each);
int hb_eval_weighted(int dim, const int* n_pts, int degree, const double* coeffs, size_t n_polys, const double* eval_nodes,
                     const double* nodes, const double* weights, const char* factor_body, const double* translation,
                     const double* dilation, const int* do_fourier, const char* index_set, double* out, size_t out_len);
int hb_varf_function(int dim, const int* n_pts, int degree, const char* f_body, const double* translation,
                     const double* dilation, const double* nodes, const double* weights, const int* do_fourier,
                     const char* index_set, int mode, double* out, size_t n_polys);
int hb_varf_function_sp(int dim, const int* n_pts, int degree, const char* f_body, const double* translation,
                        const double* dilation, const double* nodes, const double* weights, const int* do_fourier,
                        const char* index_set, int mode, long long* nnz);
/* reference: varfd(dim, degree, direction, var, do_fourier, index_set) (module.cpp:103-104, varf.cpp:290-456) */
int hb_varfd(int dim, int degree, int direction, const double* var, size_t n_polys, int do_fourier,
             const char* index_set, double* out);

/* ---- tensorize / project / inner ---------------------------------------------------------------------
 * reference: tensorize_vecs_dirs / tensorize_mats_dirs (module.cpp:116-131, tensorize.cpp:86-381).
 * k factors; factor j lives on directions dirs[j] (flattened into dirs_flat with lengths dir_lens[k]). */
int hb_tensorize_vecs(int k, const double* const* inputs, const int* lens, const int* dirs_flat,
                      const int* dir_lens, const char* index_set, double* out, size_t out_len);
int hb_tensorize_mats(int k, const double* const* inputs, const int* sizes, const int* dirs_flat,
                      const int* dir_lens, const char* index_set, double* out, size_t n_out);
/* reference: project_vec_nd / project_mat_nd (module.cpp:108-113, project.cpp:35-129) */
int hb_project_vec(const double* input, size_t len, int dim, const int* dirs, int n_dirs, const char* index_set,
                   double* out, size_t out_len);
int hb_project_mat(const double* input, size_t n, int dim, const int* dirs, int n_dirs, const char* index_set,
                   double* out, size_t n_out);
/* reference: inner(s1, s2, dirs1, dirs2, index_set) (module.cpp:95, inner.cpp:36-233) */
int hb_inner(const double* s1, size_t n1, const double* s2, size_t n2, const int* dirs1, int nd1, const int* dirs2,
             int nd2, const char* index_set, double* out, size_t out_len);

/* ---- resident-data API (device pointers) -------------------------------------------------------------
 * A plan owns the per-axis basis tables, index look-up tables and workspace for one
 * (nodes, weights, degree, index_set) combination, like an FFT plan.  Device layout of a grid
 * function: [n_0]...[n_{d-2}][pitch] doubles with pitch = hb_plan_grid_pitch() (n_{d-1} rounded up to
 * an even count: 16-byte rows for TMA); batch entries are `stride` doubles apart (even).  Coefficient
 * vectors are dense in index-set order, batch entries `stride` apart (even when dim == 1). */
typedef struct hb_plan hb_plan;
int hb_plan_create(hb_plan** plan, int dim, const int* n_pts, int degree, const double* nodes,
                   const double* weights, const int* do_fourier, const char* index_set);
void hb_plan_destroy(hb_plan* plan);
long long hb_plan_n_polys(const hb_plan* plan);
long long hb_plan_grid_pitch(const hb_plan* plan);   /* padded length of the last grid axis */
long long hb_plan_grid_size(const hb_plan* plan);    /* doubles per grid function incl. padding */
int hb_plan_forward(hb_plan* plan, const double* d_f, long long f_stride, int batch, double* d_coef,
                    long long coef_stride, void* cuda_stream);
int hb_plan_backward(hb_plan* plan, const double* d_coef, long long coef_stride, int batch, double* d_f,
                     long long f_stride, void* cuda_stream);
/* evaluate an expression on the plan's grid straight into the padded device layout (no host round trip) */
int hb_plan_discretize(hb_plan* plan, const char* body, const double* translation, const double* dilation,
                       double* d_f, void* cuda_stream);
/* varf of one grid function into the owned rows [row_begin,row_end) of the n_polys x n_polys matrix
 * (row-major, pitch ldo >= n_polys); row sharding across GPUs = one call per rank with its own range. */
int hb_plan_varf(hb_plan* plan, const double* d_f, int mode, double* d_out, long long ldo, long long row_begin,
                 long long row_end, void* cuda_stream);
/* Gram-form varf at high degree: sqrt(w_i) of one axis in extended range (host array of n entries).  A weight below
 * 1e-308 is 0 in double precision, but the Gram tables only need psi_k = sqrt(w) phi_k, and sqrt(w) is representable
 * while w > 1e-616: at degree 400 the basis functions reach |x| ~ 40, where w ~ 1e-348 -- with plain double weights the
 * quadrature silently loses the rows above degree ~290.  Drops the tables built from the weights so far. */
int hb_plan_set_sqrt_weights(hb_plan* plan, int axis, const double* sqrt_w);
/* the owned rows [row_begin,row_end) of the reference-algebra varf as CSR (rows relative to row_begin), for
 * polynomial-like f only (HB_ERR_INVALID otherwise); result fetched with hb_sparse_fetch.  Row sharding of a
 * sparse matrix across GPUs = one call per rank with its own range. */
int hb_plan_varf_sp(hb_plan* plan, const double* d_f, int mode, long long row_begin, long long row_end,
                    long long* nnz, void* cuda_stream);
/* device-side timing (CUDA events on the launch stream) of every GEMM launch of the last hb_plan_*
 * call, for roofline reporting.  tag: 1+axis forward stage, 11+axis backward stage, 21/22 varf stage
 * 1/2; padded_flops = 2*M*N*K*batch the launch was asked for (table padding included). */
int hb_plan_set_timing(hb_plan* plan, int on);
int hb_plan_timing_count(const hb_plan* plan);
int hb_plan_timing_get(hb_plan* plan, int i, int* tag, double* ms, double* padded_flops);
/* building blocks for multi-GPU composition: the FP64 tensor-core GEMM on device operands
 * (row-major; A is [M][K], or [K][M] if a_mmajor; B is [K][N]; a batch stride of 0 shares the operand;
 * leading dimensions even, bases 16-byte aligned) and a plan's per-axis basis tables
 * (kind 0: phi [node][k], 1: w*phi [node][k], 2: phi [k][node], 3: w*phi [k][node]). */
int hb_dev_dgemm(int M, int N, int K, int batch, const double* A, long long lda, long long strideA, int a_mmajor,
                 const double* B, long long ldb, long long strideB, double* C, long long ldc, long long strideC,
                 void* cuda_stream);
/* same GEMM, but output rows [i*block_rows, (i+1)*block_rows) are stored at blocks[i] + (row % block_rows)*ldc
 * (+ b*strideC + col).  The block pointers may be peer-GPU addresses mapped over NVLink, which turns the
 * epilogue into the all-to-all of the sharded transform (hermipy_b200/dist.py). */
int hb_dev_dgemm_scatter(int M, int N, int K, int batch, const double* A, long long lda, long long strideA,
                         int a_mmajor, const double* B, long long ldb, long long strideB, double* const* blocks,
                         int n_blocks, int block_rows, long long ldc, long long strideC, void* cuda_stream);
/* The general form of the two calls above.  Two-level batch: batch * batch_lo GEMMs, entry (b_hi, b_lo) uses
 * A + b_hi*strideA + b_lo*strideA_lo (likewise B, C); a stride of 0 shares the operand at that level.  The low
 * level is how a parity-split transform stage runs its even and odd halves in one launch.  With n_blocks > 0 the
 * output row r of entry (b_hi, b_lo) goes to destination row  d = row_offset + b_lo*row_offset_lo + r,  stored at
 * blocks[d / block_rows] + (d % block_rows)*ldc + b_hi*strideC + b_lo*strideC_lo  (C is ignored). */
typedef struct hb_gemm_desc {
    int M, N, K, batch, batch_lo, a_mmajor;
    const double* A; long long lda, strideA, strideA_lo;
    const double* B; long long ldb, strideB, strideB_lo;
    double* C;       long long ldc, strideC, strideC_lo;
    int n_blocks, block_rows, row_offset, row_offset_lo;
    double* const* blocks;
    /* Exchange flags (optional; all zero = off): device-side completion signalling for a GEMM whose epilogue stores
     * into peer GPUs, so that no separate barrier kernel sits between it and its consumer.
     * Producer side: when the whole grid has finished -- its stores performed and fenced system-wide -- the last CTA
     * increments the 64-bit epoch counter signal_state[0] (local memory; signal_state[1] is scratch, both start at
     * zero) and writes the new epoch to every signal_flags[i], i < n_signal (peer addresses: this rank's slot in each
     * peer's flag array).  Consumer side: before its first operand load the kernel waits until every
     * wait_flags[i], i < n_wait (local memory, written by the peers) has reached the epoch stored at wait_epoch
     * (local memory: the signal_state[0] of this rank's own producer launch of the same exchange, which precedes the
     * consumer in stream order).  The epochs live in device memory, so a captured CUDA graph replays correctly. */
    unsigned long long* const* signal_flags; int n_signal; unsigned long long* signal_state;
    const unsigned long long* wait_flags;    int n_wait;   const unsigned long long* wait_epoch;
} hb_gemm_desc;
int hb_dev_dgemm_ex(const hb_gemm_desc* desc, void* cuda_stream);
/* Parity split (merge = 0) / merge (merge = 1) of a d-dimensional grid function along up to three axes in one pass
 * (reference: none -- the sum-factorised transform on a symmetric node set contracts even and odd parts
 * separately).  Natural array: axis a has n[a] entries nat_stride[a] apart.  Sorted array: a split
 * axis is indexed by (parity p, h), h < ceil(n[a]/2) counting the nodes x = n[a] - ceil(n[a]/2) + h, stored at
 * h*srt_stride[a] + p*srt_par[a]; other axes keep their index.  nat_pad_last: the merge zero-fills the natural
 * rows from n[d-1] up to this pitch. */
int hb_dev_parity_pass(int merge, int d, const int* n, const int* split, const long long* nat_stride,
                       const long long* srt_stride, const long long* srt_par, long long batch, long long nat_bstride,
                       long long srt_bstride, int nat_pad_last, const double* in, double* out, void* cuda_stream);
/* per-parity basis tables of a (symmetric, non-Fourier) axis, both parities in one buffer `stride` doubles apart,
 * the odd entry zero-padded to the shape of the even one; nh = ceil(n/2) non-negative nodes, Pe = ceil(P/2):
 * kind 0: w*phi [h][k'] (nh x Pe), 1: w*phi [k'][h] (Pe x nh), 2: phi [h][k'], 3: phi [k'][h].
 * Returns HB_ERR_INVALID for an axis whose nodes / weights are not bit-symmetric. */
int hb_plan_parity_table(hb_plan* plan, int axis, int kind, const double** ptr, long long* rows, long long* cols,
                         long long* ld, long long* stride, void* cuda_stream);
/* consumers of an assembled matrix that stays in HBM (reference: hermipy/varf.py:127-134 `Varf.__call__`,
 * the only contact the Krylov solvers of varf.py:170-237 have with the matrix): y = A x, A row-major M x N
 * with leading dimension lda (any parity), x and y device vectors.  HBM-bound, deterministic summation. */
int hb_dev_matvec(int M, int N, const double* A, long long lda, const double* x, double* y, void* cuda_stream);
/* hb_varfd on a resident row block: var / out are n_rows x n_polys device matrices (distinct buffers) */
int hb_dev_varfd(int dim, int degree, int direction, const double* var, long long ldv, long long n_rows,
                 size_t n_polys, int do_fourier, const char* index_set, double* out, long long ldo,
                 void* cuda_stream);
int hb_plan_table(const hb_plan* plan, int axis, int kind, const double** ptr, long long* rows, long long* cols,
                  long long* ld);
/* after hb_plan_varf(HB_VARF_REFERENCE): retained alpha count and path (0 = few-alpha assembly, 1 = dense GEMM) */
int hb_plan_varf_info(const hb_plan* plan, int* n_kept, int* path);
/* number of kernels launched by this library on the calling thread since the last reset */
long long hb_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
