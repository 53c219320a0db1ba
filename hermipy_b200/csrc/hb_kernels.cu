// HBM-bound helper kernels of the hot path: basis tables by three-term recurrence, exact triple
// products, pitch repacking, index-map gathers (tensorize / project / varfd), and the "few retained
// alpha" varf assembly.  All FP64; products and sums that decide reference parity are written with
// explicit __dmul_rn/__dadd_rn so that nvcc does not contract them into FMAs the reference (compiled
// for baseline x86-64, no FMA) does not have.
#include "hb_kernels.h"
#include "hb_plan.h"

#include <algorithm>
#include <cmath>

namespace hb {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// gridDim.y (batch entry / matrix row of several kernels here) is limited to 65535: larger counts -- a 3-D transform
// of many functions, any 4-D grid, matrices of more than 65535 rows -- are launched in slices
static const int kMaxGridY = 65535;

// ------------------------------------------------------------------------------------------------
// Basis tables.  Reference: hermite_eval (cpp/src/hermite/hermite.cpp:29-41, REC_A/REC_B
// hermite.hpp:27-28) called per node from _transform (transform.cpp:77-111), Fourier branch :90-109.
//   plain[i][k] = phi_k(x_i)              (inverse transform / eval)
//   wgt  [i][k] = w_i * phi_k(x_i)        (forward transform: the reference forms f*w*phi, :121-139)
// and their transposes [k][i].  One thread per node; the k-recurrence runs in registers in exactly the
// reference's operation order ((A*x)*v_k - B*v_{k-1}); rows are written coalesced for the [k][i]
// copies.  w_i == 0 gives wgt = 0 even where phi overflowed (the reference would produce 0*inf = NaN).
// Fourier axes (no recurrence): one thread per node.
__global__ void fourier_tables_kernel(const double* __restrict__ x, const double* __restrict__ w, int n, int P,
                                      double* plain, double* wgt, long long ld_ik, double* plainT, double* wgtT,
                                      long long ld_ki, double seed_scale_unused) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = x[i];
    const double wi = w ? w[i] : 1.0;
    auto put = [&](int k, double v) {
        const double vw = (wi == 0.0) ? 0.0 : __dmul_rn(wi, v);
        if (plain) plain[(long long)i * ld_ik + k] = v;
        if (wgt) wgt[(long long)i * ld_ik + k] = vw;
        if (plainT) plainT[(long long)k * ld_ki + i] = v;
        if (wgtT) wgtT[(long long)k * ld_ki + i] = vw;
    };
    // transform.cpp:98-109 ; COS(k)=2k, SIN(k)=2k-1 (transform.hpp:55-56)
    put(0, 1.0 / sqrt(2 * M_PI));
    for (int k = 1; 2 * k < P; k++) {
        put(2 * k, cos(k * xi) / sqrt(M_PI));
        put(2 * k - 1, sin(k * xi) / sqrt(M_PI));
    }
}

// Hermite axes: one CTA of 4 warps per 32 nodes.  The recurrence coefficients REC_A(k) = 1/sqrt(k+1),
// REC_B(k) = sqrt(k)/sqrt(k+1) are formed once per CTA in shared memory (not a square root and two divisions per step
// per node).  Warp 0 runs the recurrence, lane = node, in registers -- in exactly the reference's operation order,
// ((A*x)*v_k) - (B*v_{k-1}), no FMA contraction, so the tables are bit-identical to hermite_eval -- 32 degrees at a
// time into one of two 32 x 33 shared-memory tiles; warps 1-3 write the previous tile out while warp 0 is already on
// the next 32 degrees.  BOTH layouts leave coalesced: the [k][i] tables row by row (32 consecutive nodes), the [i][k]
// tables node by node (32 consecutive degrees, read transposed from the tile).
// seed_mode 1: the recurrence starts from sqrt(w_i) (or the caller's extended-range sw[i]) instead of 1:
// psi_k = sqrt(w) phi_k of the Gram-form varf, bounded where phi alone overflows.
// HBM-write-bound in principle (8 n P bytes per table, once per plan); in practice the serial chain of P steps per
// node sets the time (see profiles/r02_tables_full_summary.txt).
__global__ void __launch_bounds__(128)
    hermite_tables_kernel(const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ sw,
                          int seed_mode, int n, int P, double* plain, double* wgt, long long ld_ik, double* plainT,
                          double* wgtT, long long ld_ki) {
    extern __shared__ double sh[];
    double* ra = sh;                 // [P]
    double* rb = sh + P;             // [P]
    double* tiles = sh + 2 * P;      // [2][32][33]
    double* wsh = tiles + 2 * 32 * 33;   // [32] weights of this CTA's nodes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, i0 = blockIdx.x * 32;
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const double sk1 = sqrt((double)(k + 1));
        ra[k] = __ddiv_rn(1.0, sk1);
        rb[k] = __ddiv_rn(sqrt((double)k), sk1);
    }
    if (threadIdx.x < 32) wsh[lane] = (i0 + lane < n && w && !seed_mode) ? w[i0 + lane] : 1.0;
    __syncthreads();
    const int i = i0 + lane;
    const bool live = i < n;
    const double xi = live ? x[i] : 0.0;
    const double s0 = !seed_mode ? 1.0 : (!live ? 0.0 : (sw ? sw[i] : sqrt(w[i])));
    double vm = 0.0, v = 0.0;       // phi_{k-1}, phi_k (warp 0)
    const int n_chunks = (P + 31) / 32;
    for (int c = 0; c <= n_chunks; c++) {
        if (warp == 0 && c < n_chunks) {
            double* tile = tiles + (c & 1) * 32 * 33;
            const int k0 = c * 32, kn = min(32, P - k0);
            for (int kk = 0; kk < kn; kk++) {
                const int k = k0 + kk;
                double nv;
                if (k == 0) nv = s0;
                else if (k == 1) nv = seed_mode ? __dmul_rn(xi, v) : xi;
                else nv = __dsub_rn(__dmul_rn(__dmul_rn(ra[k - 1], xi), v), __dmul_rn(rb[k - 1], vm));
                vm = v;
                v = nv;
                tile[kk * 33 + lane] = nv;
            }
        } else if (warp > 0 && c > 0) {
            // write out chunk c - 1: rows kk = warp - 1, warp + 2, ... of the tile for the [k][i] layouts,
            // nodes ii = warp - 1, warp + 2, ... for the [i][k] layouts
            const double* tile = tiles + ((c - 1) & 1) * 32 * 33;
            const int k0 = (c - 1) * 32, kn = min(32, P - k0);
            if (plainT || wgtT) {
                for (int kk = warp - 1; kk < kn; kk += 3) {
                    if (!live) break;
                    const double val = tile[kk * 33 + lane];
                    const double wi = wsh[lane];
                    if (plainT) plainT[(long long)(k0 + kk) * ld_ki + i] = val;
                    if (wgtT) wgtT[(long long)(k0 + kk) * ld_ki + i] = (wi == 0.0) ? 0.0 : __dmul_rn(wi, val);
                }
            }
            if ((plain || wgt) && lane < kn) {
                for (int ii = warp - 1; ii < 32 && i0 + ii < n; ii += 3) {
                    const double val = tile[lane * 33 + ii];
                    const double wn = wsh[ii];
                    if (plain) plain[(long long)(i0 + ii) * ld_ik + k0 + lane] = val;
                    if (wgt) wgt[(long long)(i0 + ii) * ld_ik + k0 + lane] = (wn == 0.0) ? 0.0 : __dmul_rn(wn, val);
                }
            }
        }
        __syncthreads();
    }
}

static int launch_hermite_tables(const double* x, const double* w, const double* sw, int seed_mode, int n, int P,
                                 double* plain, double* wgt, long long ld_ik, double* plainT, double* wgtT,
                                 long long ld_ki, cudaStream_t s) {
    const size_t smem = ((size_t)2 * P + 2 * 32 * 33 + 32) * 8;
    if (smem > 48 * 1024) {   // degree > ~2 500: opt in to more dynamic shared memory (cheap, per device: set every time)
        if (smem > 200 * 1024) { set_error("basis tables: degree %d is beyond the shared-memory coefficient table", P - 1); return HB_ERR_INVALID; }
        if (cudaFuncSetAttribute(hermite_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return HB_ERR_CUDA;
    }
    hermite_tables_kernel<<<cdiv(n, 32), 128, smem, s>>>(x, w, sw, seed_mode, n, P, plain, wgt, ld_ik, plainT, wgtT, ld_ki);
    return launch_status();
}

int launch_basis_tables(const double* x, const double* w, int n, int P, int fourier, double* plain, double* wgt,
                        long long ld_ik, double* plainT, double* wgtT, long long ld_ki, cudaStream_t s) {
    if (fourier) {
        fourier_tables_kernel<<<cdiv(n, 64), 64, 0, s>>>(x, w, n, P, plain, wgt, ld_ik, plainT, wgtT, ld_ki, 0.0);
        return launch_status();
    }
    return launch_hermite_tables(x, w, nullptr, 0, n, P, plain, wgt, ld_ik, plainT, wgtT, ld_ki, s);
}

// Weighted basis psi_k(x_i) = sqrt(w_i) * phi_k(x_i), the operand of the Gram-form varf.  The recurrence is
// linear, so it is simply seeded with sqrt(w_i) instead of 1: psi is bounded (|psi| < 1 for a Gauss-Hermite
// rule) where phi alone reaches 1e159 at degree 400 and overflows beyond degree ~780, and the pair products
// psi_m psi_n = w phi_m phi_n never leave the double range (SURVEY.md findings 4 and 7).  Nodes whose weight
// underflowed to 0 contribute 0, as in the weighted table of the transform.
// `sw` != nullptr: sqrt(w_i) supplied by the caller in extended range (a weight below 1e-308 underflows to 0 in double
// precision although its square root -- all the table needs -- is representable while w > 1e-616 and psi_k is O(1) there:
// at degree 400 the functions live out to |x| ~ 40, where w ~ 1e-348).
__global__ void psi_table_kernel(const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ sw,
                                 int n, int P, int fourier, double* __restrict__ psi, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = x[i], s = sw ? sw[i] : sqrt(w[i]);
    double* row = psi + (long long)i * ld;
    if (fourier) {
        row[0] = s / sqrt(2 * M_PI);
        for (int k = 1; 2 * k < P; k++) {
            row[2 * k] = s * (cos(k * xi) / sqrt(M_PI));
            row[2 * k - 1] = s * (sin(k * xi) / sqrt(M_PI));
        }
        return;
    }
    double vm = s, v = xi * s;
    row[0] = vm;
    if (P > 1) row[1] = v;
    for (int k = 1; k + 1 < P; k++) {
        const double sk1 = sqrt((double)(k + 1));
        const double a = __ddiv_rn(1.0, sk1), b = __ddiv_rn(sqrt((double)k), sk1);
        const double nv = __dsub_rn(__dmul_rn(__dmul_rn(a, xi), v), __dmul_rn(b, vm));
        vm = v;
        v = nv;
        row[k + 1] = v;
    }
}

int launch_psi_table(const double* x, const double* w, const double* sw, int n, int P, int fourier, double* psi,
                     long long ld, cudaStream_t s) {
    if (fourier) {
        psi_table_kernel<<<cdiv(n, 64), 64, 0, s>>>(x, w, sw, n, P, fourier, psi, ld);
        return launch_status();
    }
    return launch_hermite_tables(x, w, sw, 1, n, P, psi, nullptr, ld, nullptr, nullptr, 0, s);
}

// Products of pairs of basis functions on the nodes, the operand of the Gram-form varf:
//   psi2[x][m*Pp + n] = scale_x * phi_m(x) * phi_n(x),  scale_x = w_x (left/outer table) or 1.
__global__ void pair_table_kernel(const double* __restrict__ plain, long long ld_ik, const double* __restrict__ w,
                                  int n, int P, int Pp, double* out) {
    const int x = blockIdx.y;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)P * Pp) return;
    const int m = (int)(idx / Pp), nn = (int)(idx - (long long)m * Pp);
    double v = 0.0;
    if (nn < P) {
        v = plain[(long long)x * ld_ik + m] * plain[(long long)x * ld_ik + nn];
        if (w) v = (w[x] == 0.0) ? 0.0 : v * w[x];
    }
    out[(long long)x * P * Pp + idx] = v;
}

int launch_pair_table(const double* plain, long long ld_ik, const double* w, int n, int P, int Pp, double* out,
                      cudaStream_t s) {
    dim3 grid(cdiv((long long)P * Pp, 256), n);
    pair_table_kernel<<<grid, 256, 0, s>>>(plain, ld_ik, w, n, P, Pp, out);
    return launch_status();
}

// Blocked pair tables for the two-stage dense varf (dim 2).  Pairs (i, j) of basis indices are laid out
// in 8x8 blocks, p(i,j) = ((i>>3)*nb + (j>>3))*64 + (i&7)*8 + (j&7), nb = ceil(P/8), so that a 64-row GEMM
// tile is exactly one block: skipping by symmetry (i <= j blocks only) and by total degree then works
// at a granularity of 8 instead of the tile width.  Entries of padding slots are zero.
//   mode 0: out[x][p] = w_x * phi_i(x) * phi_j(x)      (Gram form; src = phi[x][k] with pitch ld, w may be null)
//   mode 1: out[k][p] = T[k][i][j]                     (reference algebra; src = T with plane/pitch)
// `slots` maps the 8*nb slots of an axis to basis indices (-1 = padding).  For a Hermite axis the even indices
// come first and the odd ones start on a block boundary, so every 8x8 block has one parity of i + j.
__global__ void blocked_pair_table_kernel(int mode, const double* __restrict__ src, long long ld, long long plane,
                                          const double* __restrict__ w, const int* __restrict__ slots, int nb,
                                          double* __restrict__ out) {
    const int x = blockIdx.y;
    const long long np = (long long)nb * nb * 64;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int blk = (int)(p >> 6), bi = blk / nb, bj = blk - bi * nb;
    const int i = slots[8 * bi + (int)((p >> 3) & 7)], j = slots[8 * bj + (int)(p & 7)];   // -1: padding slot
    double v = 0.0;
    if (i >= 0 && j >= 0) {
        if (mode == 0) {
            v = src[(long long)x * ld + i] * src[(long long)x * ld + j];
            if (w) v = (w[x] == 0.0) ? 0.0 : v * w[x];
        } else {
            v = src[(long long)x * plane + (long long)i * ld + j];
        }
    }
    out[(long long)x * np + p] = v;
}

int launch_blocked_pair_table(int mode, const double* src, long long ld, long long plane, const double* w, int rows,
                              const int* slots, int nb, double* out, cudaStream_t s) {
    dim3 grid(cdiv((long long)nb * nb * 64, 256), rows);
    blocked_pair_table_kernel<<<grid, 256, 0, s>>>(mode, src, ld, plane, w, slots, nb, out);
    return launch_status();
}

// Even / odd parts of a grid function along its first axis, for a symmetric node set (x_i = -x_{n-1-i}):
//   out[h][c]        = f[lo+h][c] + f[n-1-lo-h][c]      (h < nh; the middle node of an odd rule is taken once)
//   out[seg + h][c]  = f[lo+h][c] - f[n-1-lo-h][c]      (0 for the middle node)
// with lo = n - nh the first non-negative node.  Rows nh..seg-1 of the even segment are left untouched (zero).
__global__ void parity_split_kernel(const double* __restrict__ f, int n, int nh, int seg, long long pitch,
                                    double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nh * pitch) return;
    const int h = (int)(idx / pitch);
    const long long c = idx - (long long)h * pitch;
    const int xp = n - nh + h, xm = n - 1 - xp;
    const double a = f[(long long)xp * pitch + c];
    const double b = (xm == xp) ? 0.0 : f[(long long)xm * pitch + c];
    out[(long long)h * pitch + c] = a + b;
    out[(long long)(seg + h) * pitch + c] = (xm == xp) ? a * 0.0 : a - b;   // keeps a NaN / inf of f visible, as 0 * f
}

int launch_parity_split(const double* f, int n, int nh, int seg, long long pitch, double* out, cudaStream_t s) {
    parity_split_kernel<<<cdiv((long long)nh * pitch, 256), 256, 0, s>>>(f, n, nh, seg, pitch, out);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Exact Hermite triple products T[k][i][j] = <phi_i phi_j, phi_k>, k <= 2N, i,j <= N.
// Reference: triple_products_1d, cpp/src/hermite/varf.cpp:49-94.  The recursion raises i at fixed j
// (three k-ranges :77-84, then symmetrisation :88-91), so columns j are independent: one CTA per j,
// threads over k, levels i in sequence with the two previous levels kept in shared memory.  Each
// entry is accumulated in the reference's order (term1, then term2, then term3) without FMA
// contraction.  Output layout [k][i][Pp] with Pp >= P even (pad columns are zero).
__global__ void triple_products_kernel(int N, int Pp, double* __restrict__ T) {
    extern __shared__ double sh[];
    const int K = 2 * N + 1;
    double* lev0 = sh;            // level i-2
    double* lev1 = sh + (K + 2);  // level i-1   (one guard cell on each side)
    double* lev2 = sh + 2 * (K + 2);
    const int j = blockIdx.x;
    const long long plane = (long long)(N + 1) * Pp;
    for (int k = threadIdx.x; k < K + 2; k += blockDim.x) lev0[k] = lev1[k] = lev2[k] = 0.0;
    __syncthreads();
    // level 0: T[j][0][j] = 1
    if (threadIdx.x == 0) {
        lev1[1 + j] = 1.0;
        T[(long long)j * plane + 0 * Pp + j] = 1.0;
        T[(long long)j * plane + (long long)j * Pp + 0] = 1.0;
    }
    __syncthreads();
    for (int i = 1; i <= j; i++) {
        const double ai = __ddiv_rn(1.0, sqrt((double)i));                    // a[i-1]
        const double bi = __ddiv_rn(sqrt((double)(i - 1)), sqrt((double)i));  // b[i-1]
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            double v = 0.0;
            if (k >= j - i && k <= j + i) {
                const double sk1 = sqrt((double)(k + 1));
                const double ak = __ddiv_rn(1.0, sk1);
                const double r = __ddiv_rn(ai, ak);
                if (k <= j + i - 2) v = __dadd_rn(v, __dmul_rn(r, lev1[1 + k + 1]));
                if (k >= j - i + 2 && k <= j + i - 2) v = __dadd_rn(v, __dmul_rn(-bi, lev0[1 + k]));
                if (k >= j - i + 2 && k >= 1) {
                    const double bk = __ddiv_rn(sqrt((double)k), sk1);
                    v = __dadd_rn(v, __dmul_rn(__dmul_rn(r, bk), lev1[1 + k - 1]));
                }
                T[(long long)k * plane + (long long)i * Pp + j] = v;
                T[(long long)k * plane + (long long)j * Pp + i] = v;
            }
            lev2[1 + k] = v;
        }
        __syncthreads();
        double* t = lev0; lev0 = lev1; lev1 = lev2; lev2 = t;
    }
}

int launch_triple_products(int N, int Pp, double* T, cudaStream_t s) {
    const size_t bytes = (size_t)(2 * N + 1) * (N + 1) * Pp * 8;
    if (cudaMemsetAsync(T, 0, bytes, s) != cudaSuccess) return HB_ERR_CUDA;
    const int smem = 3 * (2 * N + 3) * 8;
    triple_products_kernel<<<N + 1, 256, smem, s>>>(N, Pp, T);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Parity helpers of the 1-D transform on a symmetric node set (x_i = -x_{n-1-i}, w_i = w_{n-1-i}): phi_k(-x) =
// (-1)^k phi_k(x), so even coefficients only see the even part of f and odd ones the odd part, both contracted
// over the nh = ceil(n/2) non-negative nodes lo .. n-1 (lo = n - nh; the node x = 0 of an odd rule is taken once).
//   split : fe[b][h] = f[b][lo+h] + f[b][n-1-lo-h],  fo[b][h] = f[b][lo+h] - f[b][n-1-lo-h]
//   merge : f[b][lo+h] = ge[b][h] + go[b][h],        f[b][n-1-lo-h] = ge[b][h] - go[b][h]
__global__ void rows_parity_split_kernel(const double* __restrict__ f, long long f_stride, int n, int nh,
                                         double* __restrict__ fe, double* __restrict__ fo, long long h_stride,
                                         long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long b = idx / nh;
    const int h = (int)(idx - b * nh);
    const int xp = n - nh + h, xm = n - 1 - xp;
    const double a = f[b * f_stride + xp];
    const double c = (xm == xp) ? 0.0 : f[b * f_stride + xm];
    fe[b * h_stride + h] = a + c;
    fo[b * h_stride + h] = (xm == xp) ? a * 0.0 : a - c;   // 0 at the node x = 0, NaN if f is not finite there (as 0 * f)
}

__global__ void rows_parity_merge_kernel(const double* __restrict__ ge, const double* __restrict__ go,
                                         long long h_stride, int n, int nh, int n_pad, double* __restrict__ f,
                                         long long f_stride, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long b = idx / nh;
    const int h = (int)(idx - b * nh);
    const int xp = n - nh + h, xm = n - 1 - xp;
    const double e = ge[b * h_stride + h], o = go[b * h_stride + h];
    f[b * f_stride + xp] = e + o;
    if (xm != xp) f[b * f_stride + xm] = e - o;
    if (h == 0)
        for (int c = n; c < n_pad; c++) f[b * f_stride + c] = 0.0;   // pad columns of the row (16-byte pitch)
}

// dst[r][c] = src[(r0 + r*rs)*src_ld + c0 + c*cs] for r < rows, c < cols; pad columns up to dst_ld are zeroed
__global__ void strided_copy_kernel(const double* __restrict__ src, long long src_ld, int r0, int rs, int c0, int cs,
                                    double* __restrict__ dst, long long dst_ld, long long rows, int cols,
                                    long long src_bstride, long long dst_bstride) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * dst_ld) return;
    const long long r = idx / dst_ld;
    const int c = (int)(idx - r * dst_ld);
    const double* sb = src + blockIdx.y * src_bstride;
    dst[blockIdx.y * dst_bstride + idx] = (c < cols) ? sb[(r0 + r * rs) * src_ld + c0 + (long long)c * cs] : 0.0;
}

// The same along a middle axis: in[p][x][c] (x < n, c < cols, batch entries in_stride apart) ->
//   be[p][h][c] = in[p][lo+h][c] + in[p][n-1-lo-h][c],  bo[p][h][c] = in[p][lo+h][c] - in[p][n-1-lo-h][c]
// and back: out[p][lo+h][c] = ge + go, out[p][n-1-lo-h][c] = ge - go (out batch entries out_stride apart).
__global__ void axis_parity_split_kernel(const double* __restrict__ in, long long in_stride, int n, int nh,
                                         long long cols, double* __restrict__ be, double* __restrict__ bo,
                                         long long per_batch) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_batch) return;
    const int h = (int)(idx / cols);
    const long long c = idx - (long long)h * cols;
    const int xp = n - nh + h, xm = n - 1 - xp;
    const double* src = in + blockIdx.y * in_stride;
    const double a = src[(long long)xp * cols + c];
    const double b = (xm == xp) ? 0.0 : src[(long long)xm * cols + c];
    be[blockIdx.y * per_batch + idx] = a + b;
    bo[blockIdx.y * per_batch + idx] = (xm == xp) ? a * 0.0 : a - b;
}

__global__ void axis_parity_merge_kernel(const double* __restrict__ ge, const double* __restrict__ go, int n, int nh,
                                         long long cols, double* __restrict__ out, long long out_stride,
                                         long long per_batch) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_batch) return;
    const int h = (int)(idx / cols);
    const long long c = idx - (long long)h * cols;
    const int xp = n - nh + h, xm = n - 1 - xp;
    const double e = ge[blockIdx.y * per_batch + idx], o = go[blockIdx.y * per_batch + idx];
    double* dst = out + blockIdx.y * out_stride;
    dst[(long long)xp * cols + c] = e + o;
    if (xm != xp) dst[(long long)xm * cols + c] = e - o;
}

int launch_axis_parity_split(const double* in, long long in_stride, int n, int nh, long long cols, int batch,
                             double* be, double* bo, cudaStream_t s) {
    const long long per_batch = (long long)nh * cols;
    for (int b0 = 0; b0 < batch; b0 += kMaxGridY) {
        dim3 grid(cdiv(per_batch, 256), std::min(batch - b0, kMaxGridY));
        axis_parity_split_kernel<<<grid, 256, 0, s>>>(in + b0 * in_stride, in_stride, n, nh, cols, be + b0 * per_batch,
                                                      bo + b0 * per_batch, per_batch);
        HB_TRY(launch_status());
    }
    return HB_OK;
}

int launch_axis_parity_merge(const double* ge, const double* go, int n, int nh, long long cols, int batch, double* out,
                             long long out_stride, cudaStream_t s) {
    const long long per_batch = (long long)nh * cols;
    for (int b0 = 0; b0 < batch; b0 += kMaxGridY) {
        dim3 grid(cdiv(per_batch, 256), std::min(batch - b0, kMaxGridY));
        axis_parity_merge_kernel<<<grid, 256, 0, s>>>(ge + b0 * per_batch, go + b0 * per_batch, n, nh, cols,
                                                      out + b0 * out_stride, out_stride, per_batch);
        HB_TRY(launch_status());
    }
    return HB_OK;
}

int launch_rows_parity_split(const double* f, long long f_stride, int n, int nh, double* fe, double* fo,
                             long long h_stride, int batch, cudaStream_t s) {
    const long long total = (long long)batch * nh;
    rows_parity_split_kernel<<<cdiv(total, 256), 256, 0, s>>>(f, f_stride, n, nh, fe, fo, h_stride, total);
    return launch_status();
}

int launch_rows_parity_merge(const double* ge, const double* go, long long h_stride, int n, int nh, int n_pad,
                             double* f, long long f_stride, int batch, cudaStream_t s) {
    const long long total = (long long)batch * nh;
    rows_parity_merge_kernel<<<cdiv(total, 256), 256, 0, s>>>(ge, go, h_stride, n, nh, n_pad, f, f_stride, total);
    return launch_status();
}

int launch_strided_copy(const double* src, long long src_ld, int r0, int rs, int c0, int cs, double* dst,
                        long long dst_ld, long long rows, int cols, int batch, long long src_bstride,
                        long long dst_bstride, cudaStream_t s) {
    for (int b0 = 0; b0 < batch; b0 += kMaxGridY) {
        dim3 grid(cdiv(rows * dst_ld, 256), std::min(batch - b0, kMaxGridY));
        strided_copy_kernel<<<grid, 256, 0, s>>>(src + b0 * src_bstride, src_ld, r0, rs, c0, cs, dst + b0 * dst_bstride,
                                                 dst_ld, rows, cols, src_bstride, dst_bstride);
        HB_TRY(launch_status());
    }
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// d-dimensional parity split / merge (see ParityNDArgs).  One thread per point (h_a on the split axes, x_a on the
// others) of every batch entry, last axis fastest: it loads the 2^M mirror images of its point, runs the M
// butterflies in registers and stores the 2^M parity combinations -- every element is read once and written once
// whatever the number of split axes (the per-axis passes this replaces moved the array once per axis).
struct ParityNDWork {          // host-derived addressing of the work index (last axis fastest)
    long long div[HB_MAX_DIM];  // product of the work extents of the faster axes
    long long per_batch;        // work items per batch entry
    int sq_axis[3];             // the split axes, outermost first
};

// I: type of the work-index arithmetic (unsigned when the index fits 32 bits: the decomposition costs several
// divisions per thread, and 64-bit ones are emulated)
template <int M, bool MERGE, typename I>
__global__ void nd_parity_kernel(ParityNDArgs a, ParityNDWork wk, const double* __restrict__ in,
                                 double* __restrict__ out, long long total) {
    const long long idx64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx64 >= total) return;
    const I idx = (I)idx64;
    const long long b = (long long)(idx / (I)wk.per_batch);
    long long nat = b * a.nat_bstride, srt = b * a.srt_bstride;
    // natural / sorted offsets of the point, and per split axis: the two natural positions and the parity offset
    long long pos_p[M], pos_m[M], par[M];
    bool degenerate[M];
#pragma unroll
    for (int q = 0; q < M; q++) {
        const int ax = wk.sq_axis[q];
        const int c = (int)((idx / (I)wk.div[ax]) % (I)a.nh[ax]);
        const int xp = a.n[ax] - a.nh[ax] + c, xm = a.n[ax] - 1 - xp;
        pos_p[q] = xp * a.nat_stride[ax];
        pos_m[q] = xm * a.nat_stride[ax];
        degenerate[q] = (xm == xp);
        par[q] = a.srt_par[ax];
        srt += c * a.srt_stride[ax];
    }
#pragma unroll 1
    for (int ax = 0; ax < a.d; ax++) {
        if (a.split[ax]) continue;
        const int c = (int)((idx / (I)wk.div[ax]) % (I)a.n[ax]);
        nat += c * a.nat_stride[ax];
        srt += c * a.srt_stride[ax];
    }
    const int last = a.d - 1;
    const bool last_split = a.split[last] != 0;   // then slot M - 1 is the last axis
    double v[1 << M];
    if (!MERGE) {
#pragma unroll
        for (int s = 0; s < (1 << M); s++) {
            long long off = nat;
            bool absent = false;
#pragma unroll
            for (int q = 0; q < M; q++) {
                const bool mirror = (s >> q) & 1;
                off += mirror ? pos_m[q] : pos_p[q];
                absent |= mirror && degenerate[q];
            }
            v[s] = absent ? 0.0 : in[off];
        }
#pragma unroll
        for (int q = 0; q < M; q++)
#pragma unroll
            for (int s = 0; s < (1 << M); s++)
                if (!((s >> q) & 1)) {
                    const double x = v[s], y = v[s | (1 << q)];
                    v[s] = x + y;
                    v[s | (1 << q)] = degenerate[q] ? x * 0.0 : x - y;
                }
#pragma unroll
        for (int s = 0; s < (1 << M); s++) {
            long long off = srt;
#pragma unroll
            for (int q = 0; q < M; q++) off += ((s >> q) & 1) ? par[q] : 0;
            out[off] = v[s];
        }
    } else {
#pragma unroll
        for (int s = 0; s < (1 << M); s++) {
            long long off = srt;
#pragma unroll
            for (int q = 0; q < M; q++) off += ((s >> q) & 1) ? par[q] : 0;
            v[s] = in[off];
        }
#pragma unroll
        for (int q = 0; q < M; q++)
#pragma unroll
            for (int s = 0; s < (1 << M); s++)
                if (!((s >> q) & 1)) {
                    const double e = v[s], o = v[s | (1 << q)];
                    v[s] = e + o;
                    v[s | (1 << q)] = e - o;
                }
        // the thread at work coordinate 0 of the last axis also zero-fills the pad columns of the rows it touches
        const bool pads = a.nat_pad_last > a.n[last] && (idx % (I)(last_split ? a.nh[last] : a.n[last])) == 0;
#pragma unroll
        for (int s = 0; s < (1 << M); s++) {
            long long off = nat, row = nat;   // row: the same point with the last axis at its origin
            bool absent = false;
#pragma unroll
            for (int q = 0; q < M; q++) {
                const bool mirror = (s >> q) & 1;
                const long long pq = mirror ? pos_m[q] : pos_p[q];
                off += pq;
                if (!(q == M - 1 && last_split)) row += pq;
                absent |= mirror && degenerate[q];
            }
            if (absent) continue;
            out[off] = v[s];
            if (pads && !(last_split && ((s >> (M - 1)) & 1)))
                for (int c = a.n[last]; c < a.nat_pad_last; c++) out[row + c * a.nat_stride[last]] = 0.0;
        }
    }
}

template <bool MERGE>
static int launch_nd_parity(const ParityNDArgs& a, const double* in, double* out, cudaStream_t s) {
    if (a.d < 1 || a.d > HB_MAX_DIM) { set_error("nd parity pass: 1..%d axes supported", HB_MAX_DIM); return HB_ERR_INVALID; }
    ParityNDWork wk = {};
    int m = 0;
    long long total = 1;
    for (int ax = a.d - 1; ax >= 0; ax--) {
        wk.div[ax] = total;
        total *= a.split[ax] ? a.nh[ax] : a.n[ax];
    }
    for (int ax = 0; ax < a.d; ax++)
        if (a.split[ax]) {
            if (m < 3) wk.sq_axis[m] = ax;
            m++;
        }
    wk.per_batch = total;
    total *= a.batch;
    if (total <= 0) return HB_OK;
    if (m < 1 || m > 3) { set_error("nd parity pass: 1..3 split axes supported (got %d)", m); return HB_ERR_INVALID; }
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) { set_error("nd parity pass: grid too large"); return HB_ERR_INVALID; }
    if (total < 0x7fffffffLL) {
        if (m == 1) nd_parity_kernel<1, MERGE, unsigned><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
        else if (m == 2) nd_parity_kernel<2, MERGE, unsigned><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
        else nd_parity_kernel<3, MERGE, unsigned><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
    } else {
        if (m == 1) nd_parity_kernel<1, MERGE, long long><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
        else if (m == 2) nd_parity_kernel<2, MERGE, long long><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
        else nd_parity_kernel<3, MERGE, long long><<<(unsigned)blocks, 256, 0, s>>>(a, wk, in, out, total);
    }
    return launch_status();
}
int launch_nd_parity_split(const ParityNDArgs& a, const double* natural, double* sorted, cudaStream_t s) {
    return launch_nd_parity<false>(a, natural, sorted, s);
}
int launch_nd_parity_merge(const ParityNDArgs& a, const double* sorted, double* natural, cudaStream_t s) {
    return launch_nd_parity<true>(a, sorted, natural, s);
}

// De-interleave by parity in one pass (1-D inverse transform on a symmetric node set): row r of `src` holds the
// coefficients in natural order; even[r][j] = src[r][2j] (j < Pe), odd[r][j] = src[r][2j+1] (j < Po), pad columns
// up to the even pitches lde / ldo zeroed (they are GEMM operands).  Each source sector is read once.
__global__ void deinterleave_rows_kernel(const double* __restrict__ src, long long src_ld, int Pe, int Po,
                                         double* __restrict__ even, long long lde, double* __restrict__ odd,
                                         long long ldo, long long rows, long long width) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * width) return;
    const long long r = idx / width;
    const int j = (int)(idx - r * width);
    const double* row = src + r * src_ld;
    if (j < lde) even[r * lde + j] = (j < Pe) ? row[2 * j] : 0.0;
    if (j < ldo) odd[r * ldo + j] = (j < Po) ? row[2 * j + 1] : 0.0;
}
int launch_deinterleave_rows(const double* src, long long src_ld, int Pe, int Po, double* even, long long lde, double* odd,
                             long long ldo, long long rows, cudaStream_t s) {
    const long long width = lde > ldo ? lde : ldo;
    if (rows <= 0 || width <= 0) return HB_OK;
    deinterleave_rows_kernel<<<cdiv(rows * width, 256), 256, 0, s>>>(src, src_ld, Pe, Po, even, lde, odd, ldo, rows, width);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Pitch repack / gather helpers.
__global__ void repack_rows_kernel(const double* __restrict__ src, long long src_ld, double* __restrict__ dst,
                                   long long dst_ld, long long rows, int cols) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * dst_ld) return;
    const long long r = idx / dst_ld;
    const int c = (int)(idx - r * dst_ld);
    dst[idx] = (c < cols) ? src[r * src_ld + c] : 0.0;
}

int launch_repack_rows(const double* src, long long src_ld, double* dst, long long dst_ld, long long rows, int cols,
                       cudaStream_t s) {
    repack_rows_kernel<<<cdiv(rows * dst_ld, 256), 256, 0, s>>>(src, src_ld, dst, dst_ld, rows, cols);
    return launch_status();
}

// dst[b][j] = (lut[j] >= 0) ? src[b][lut[j]] : 0      (coefficients in set order -> padded lex box)
__global__ void gather_lut_kernel(const double* __restrict__ src, long long src_stride, const int* __restrict__ lut,
                                  double* __restrict__ dst, long long dst_stride, long long n_dst, int batch) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_dst) return;
    const int p = lut[idx];
    for (int b = blockIdx.y; b < batch; b += gridDim.y)
        dst[b * dst_stride + idx] = (p >= 0) ? src[b * src_stride + p] : 0.0;
}

int launch_gather_lut(const double* src, long long src_stride, const int* lut, double* dst, long long dst_stride,
                      long long n_dst, int batch, cudaStream_t s) {
    dim3 grid(cdiv(n_dst, 256), batch < 1024 ? batch : 1024);
    gather_lut_kernel<<<grid, 256, 0, s>>>(src, src_stride, lut, dst, dst_stride, n_dst, batch);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// tensorize (reference: cpp/src/hermite/tensorize.cpp:86-122 vectors, :197-283 matrices):
//   vectors : out[i]      = prod_j in_j[ map_j[i] ]
//   matrices: out[r][c]   = prod_j in_j[ map_j[r] * size_j + map_j[c] ]
// map_j[i] = position of (m_i restricted to dirs_j) in the j-th factor's index set (host-built LUT).
__global__ void tensorize_vec_kernel(TensorizeArgs a, double* __restrict__ out, int n_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    double v = 1.0;
    for (int j = 0; j < a.k; j++) v = __dmul_rn(v, a.in[j][a.map[j][i]]);
    out[i] = v;
}

__global__ void tensorize_mat_kernel(TensorizeArgs a, double* __restrict__ out, int n_out, int r0) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = r0 + blockIdx.y;
    if (c >= n_out) return;
    // the reference folds factors left to right: ((A0*A1)*A2)...
    double v = a.in[0][(long long)a.map[0][r] * a.size[0] + a.map[0][c]];
    for (int j = 1; j < a.k; j++)
        v = __dmul_rn(v, a.in[j][(long long)a.map[j][r] * a.size[j] + a.map[j][c]]);
    out[(long long)r * n_out + c] = v;
}

int launch_tensorize(const TensorizeArgs& a, bool matrices, double* out, int n_out, cudaStream_t s) {
    if (matrices) {
        for (int r0 = 0; r0 < n_out; r0 += kMaxGridY) {   // gridDim.y <= 65535
            dim3 grid(cdiv(n_out, 256), std::min(n_out - r0, kMaxGridY));
            tensorize_mat_kernel<<<grid, 256, 0, s>>>(a, out, n_out, r0);
        }
    } else {
        tensorize_vec_kernel<<<cdiv(n_out, 256), 256, 0, s>>>(a, out, n_out);
    }
    return launch_status();
}

// project (reference: cpp/src/hermite/project.cpp:35-58, :78-109): out[i] = in[sel[i]],
// out[r][c] = in[sel[r]][sel[c]], sel[i] = position in the full set of the sub-index i embedded on dirs.
__global__ void project_vec_kernel(const double* __restrict__ in, const int* __restrict__ sel, double* out, int n_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = sel[i] >= 0 ? in[sel[i]] : 0.0;
}
__global__ void project_mat_kernel(const double* __restrict__ in, int n_in, const int* __restrict__ sel, double* out,
                                   int n_out, int r0) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = r0 + blockIdx.y;
    if (c >= n_out) return;
    const int sr = sel[r], sc = sel[c];
    out[(long long)r * n_out + c] = (sr >= 0 && sc >= 0) ? in[(long long)sr * n_in + sc] : 0.0;
}
int launch_project(const double* in, int n_in, const int* sel, bool matrix, double* out, int n_out, cudaStream_t s) {
    if (matrix) {
        for (int r0 = 0; r0 < n_out; r0 += kMaxGridY) {   // gridDim.y <= 65535
            dim3 grid(cdiv(n_out, 256), std::min(n_out - r0, kMaxGridY));
            project_mat_kernel<<<grid, 256, 0, s>>>(in, n_in, sel, out, n_out, r0);
        }
    } else {
        project_vec_kernel<<<cdiv(n_out, 256), 256, 0, s>>>(in, sel, out, n_out);
    }
    return launch_status();
}

// varfd (reference: cpp/src/hermite/varf.cpp:290-376): out[r][c] = fac[c] * var[r][src[c]] where, for a
// Hermite axis, src[c] = index(m_c - e_dir) and fac[c] = sqrt(m_c[dir]) (0 / unused when m_c[dir]==0).
__global__ void varfd_kernel(const double* __restrict__ var, long long ldv, const int* __restrict__ src,
                             const double* __restrict__ fac, double* out, long long ldo, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= n) return;
    const int sc = src[c];
    out[(long long)r * ldo + c] = sc >= 0 ? __dmul_rn(var[(long long)r * ldv + sc], fac[c]) : 0.0;
}
int launch_varfd(const double* var, long long ldv, const int* src, const double* fac, double* out, long long ldo,
                 int rows, int n, cudaStream_t s) {
    // gridDim.y is limited to 65535: walk taller row blocks in slices
    for (int r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid(cdiv(n, 256), nr);
        varfd_kernel<<<grid, 256, 0, s>>>(var + (long long)r0 * ldv, ldv, src, fac, out + (long long)r0 * ldo, ldo, n);
        const int rc = launch_status();
        if (rc != HB_OK) return rc;
    }
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// varf assembly when only a few alpha survive the threshold (reference: varf.cpp:236-262; this is the
// regime the reference algebra is numerically meaningful in -- polynomial-like f).  One thread per
// output entry; the retained (alpha, Hf[alpha]) list is walked in the reference's iterator order and
// each term is formed as ((T[a_0][m_0][n_0] * T[a_1][m_1][n_1]) * ...) * Hf, added without FMA, so
// the sum reproduces the reference's rounding.  HBM-write-bound: 8 B per entry.
__global__ void varf_sparse_kernel(VarfSparseArgs a, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = a.row_begin + blockIdx.y;
    if (c >= a.n_polys) return;
    int m[HB_MAX_DIM], nn[HB_MAX_DIM];
    for (int d = 0; d < a.dim; d++) {
        m[d] = a.midx[(long long)r * a.dim + d];
        nn[d] = a.midx[(long long)c * a.dim + d];
    }
    double acc = 0.0;
    for (int t = 0; t < a.n_kept; t++) {
        double prod = 0.0;
        bool zero = false;
        for (int d = 0; d < a.dim; d++) {
            const int al = a.kept_alpha[(long long)t * a.dim + d];
            const double* T = a.T[d];
            const double v = T[(long long)al * a.plane[d] + (long long)m[d] * a.Pp[d] + nn[d]];
            if (v == 0.0) { zero = true; break; }   // the reference only stores non-zeros (sparse.cpp:52)
            prod = (d == 0) ? v : __dmul_rn(prod, v);
        }
        if (zero) continue;
        acc = __dadd_rn(acc, __dmul_rn(prod, a.kept_val[t]));
    }
    out[(long long)blockIdx.y * a.ldo + c] = acc;
}

int launch_varf_sparse(const VarfSparseArgs& a, double* out, cudaStream_t s) {
    const int rows = a.row_end - a.row_begin;
    if (rows <= 0) return HB_OK;
    for (int r0 = 0; r0 < rows; r0 += 65535) {
        VarfSparseArgs b = a;
        b.row_begin = a.row_begin + r0;
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid(cdiv(a.n_polys, 128), nr);
        varf_sparse_kernel<<<grid, 128, 0, s>>>(b, out + (long long)r0 * a.ldo);
    }
    return launch_status();
}

// Banded variant of the same assembly: T[k][i][j] vanishes unless |i-j| <= k, so with the retained
// alpha bounded by amax[d] per axis only the entries with |m_d - n_d| <= amax[d] can be non-zero --
// for polynomial f a few dozen per row out of |I|.  The matrix is zero-filled once (a plain memset,
// HBM-write-bound) and only the band is computed: one thread per (row, band offset), same term order
// and rounding as varf_sparse_kernel.
__global__ void varf_band_kernel(VarfSparseArgs a, VarfBandArgs b, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)(a.row_end - a.row_begin) * b.width_total;
    if (idx >= total) return;
    const int r = a.row_begin + (int)(idx / b.width_total);
    int off = (int)(idx % b.width_total);
    int m[HB_MAX_DIM], nn[HB_MAX_DIM];
    long long lex = 0;
    for (int d = a.dim - 1; d >= 0; d--) {          // offsets: last axis fastest
        const int w = 2 * b.amax[d] + 1;
        const int o = off % w;
        off /= w;
        m[d] = a.midx[(long long)r * a.dim + d];
        nn[d] = m[d] + o - b.amax[d];
        if (nn[d] < 0 || nn[d] >= b.extent[d]) return;
    }
    for (int d = 0; d < a.dim; d++) lex = lex * (d == a.dim - 1 ? b.pitch_last : b.extent[d]) + nn[d];
    const int c = b.lut_box[lex];
    if (c < 0) return;
    double acc = 0.0;
    for (int t = 0; t < a.n_kept; t++) {
        double prod = 0.0;
        bool zero = false;
        for (int d = 0; d < a.dim; d++) {
            const int al = a.kept_alpha[(long long)t * a.dim + d];
            const double v = a.T[d][(long long)al * a.plane[d] + (long long)m[d] * a.Pp[d] + nn[d]];
            if (v == 0.0) { zero = true; break; }
            prod = (d == 0) ? v : __dmul_rn(prod, v);
        }
        if (zero) continue;
        acc = __dadd_rn(acc, __dmul_rn(prod, a.kept_val[t]));
    }
    out[(long long)(r - a.row_begin) * a.ldo + c] = acc;
}

int launch_varf_band(const VarfSparseArgs& a, const VarfBandArgs& b, double* out, cudaStream_t s, bool zero_filled) {
    const long long rows = a.row_end - a.row_begin;
    if (rows <= 0) return HB_OK;
    if (!zero_filled &&
        cudaMemset2DAsync(out, (size_t)a.ldo * 8, 0, (size_t)a.n_polys * 8, (size_t)rows, s) != cudaSuccess)
        return HB_ERR_CUDA;
    const long long total = rows * b.width_total;
    if (total == 0) return HB_OK;
    varf_band_kernel<<<cdiv(total, 256), 256, 0, s>>>(a, b, out);
    return launch_status();
}

// inner (reference: cpp/src/hermite/inner.cpp:36-217): result[i] = sum_j s1[ind1[j][i]] * s2[ind2[j][i]],
// j over the shared ("inner") index set in iterator order, pairs outside either set skipped (:193-194).
// One thread per result entry; the maps are stored [j][i] so that warps read them coalesced; products
// are added sequentially without FMA, as the reference does.
__global__ void inner_kernel(const double* __restrict__ s1, const double* __restrict__ s2, const int* __restrict__ ind1,
                             const int* __restrict__ ind2, int n_res, int n_inner, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_res) return;
    double acc = 0.0;
    for (int j = 0; j < n_inner; j++) {
        const int a = ind1[(long long)j * n_res + i], b = ind2[(long long)j * n_res + i];
        if (a < 0 || b < 0) continue;
        acc = __dadd_rn(acc, __dmul_rn(s1[a], s2[b]));
    }
    out[i] = acc;
}
int launch_inner(const double* s1, const double* s2, const int* ind1, const int* ind2, int n_res, int n_inner,
                 double* out, cudaStream_t s) {
    inner_kernel<<<cdiv(n_res, 128), 128, 0, s>>>(s1, s2, ind1, ind2, n_res, n_inner, out);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Last step of the d-dimensional dense varf (d >= 3, hb_plan.cu: varf_dense_nd): the contraction leaves the full
// pair-space array full[(m_0,n_0)]...[(m_{d-1},n_{d-1})] (pair (m_a, n_a) at m_a * Pp_a + n_a); the matrix entry
// A[r][c] is the cell addressed by the multi-indices of row r and column c.
struct VarfGatherArgs {
    int dim, n_polys, row_begin;
    int Pp[HB_MAX_DIM];
    long long stride[HB_MAX_DIM];   // of the pair index of axis a in `full`
    const int* midx;
};
__global__ void varf_gather_nd_kernel(VarfGatherArgs a, const double* __restrict__ full, double* __restrict__ out,
                                      long long ldo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = a.row_begin + blockIdx.y;
    if (c >= a.n_polys) return;
    long long cell = 0;
    for (int d = 0; d < a.dim; d++)
        cell += ((long long)a.midx[(long long)r * a.dim + d] * a.Pp[d] + a.midx[(long long)c * a.dim + d]) * a.stride[d];
    out[(long long)blockIdx.y * ldo + c] = full[cell];
}
int launch_varf_gather_nd(int dim, int n_polys, int row_begin, int row_end, const int* Pp, const long long* stride,
                          const int* midx, const double* full, double* out, long long ldo, cudaStream_t s) {
    VarfGatherArgs a;
    a.dim = dim; a.n_polys = n_polys; a.midx = midx;
    for (int d = 0; d < dim; d++) { a.Pp[d] = Pp[d]; a.stride[d] = stride[d]; }
    for (int r0 = row_begin; r0 < row_end; r0 += kMaxGridY) {
        const int nr = row_end - r0 < kMaxGridY ? row_end - r0 : kMaxGridY;
        a.row_begin = r0;
        dim3 grid(cdiv(n_polys, 128), nr);
        varf_gather_nd_kernel<<<grid, 128, 0, s>>>(a, full, out + (long long)(r0 - row_begin) * ldo, ldo);
    }
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Weighting / scaling steps of the object API (reference: hermipy/quad.py:287-328, NumPy on the host there), batch
// rows of `n` entries against ONE factor row: op 0  out = a / (b == 0 ? 1e-300 : b)   (quad.py:299: the factor of the
// quadrature divides the function before the transform; 1e-300 stands in where the factor underflowed), op 1
// out = a * b (quad.py:328).  Same IEEE operations as the NumPy expressions, so the results are bit-identical.
__global__ void pointwise_kernel(int op, const double* __restrict__ a, long long a_stride, const double* __restrict__ b,
                                 double* __restrict__ out, long long o_stride, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double bv = b[i];
    const double av = a[(long long)blockIdx.y * a_stride + i];
    out[(long long)blockIdx.y * o_stride + i] = op == 0 ? __ddiv_rn(av, bv == 0.0 ? 1e-300 : bv) : __dmul_rn(av, bv);
}
int launch_pointwise(int op, const double* a, long long a_stride, const double* b, double* out, long long o_stride,
                     long long n, int batch, cudaStream_t s) {
    if (n <= 0 || batch <= 0) return HB_OK;
    for (int b0 = 0; b0 < batch; b0 += kMaxGridY) {
        const int nb = batch - b0 < kMaxGridY ? batch - b0 : kMaxGridY;
        dim3 grid((unsigned)cdiv(n, 256), nb);
        pointwise_kernel<<<grid, 256, 0, s>>>(op, a + (long long)b0 * a_stride, a_stride, b, out + (long long)b0 * o_stride,
                                             o_stride, n);
    }
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// Dense -> CSR compaction (exact zeros dropped, as the reference's to_spmat does, sparse.cpp:44-58).
__global__ void count_nnz_kernel(const double* __restrict__ A, long long ld, int n_cols, long long* row_nnz) {
    const int r = blockIdx.x;
    int cnt = 0;
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) cnt += (A[(long long)r * ld + c] != 0.0);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    __shared__ int sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < (blockDim.x >> 5); i++) t += sh[i];
        row_nnz[r] = t;
    }
}
// one warp per row, ordered compaction with ballot/popc
__global__ void fill_csr_kernel(const double* __restrict__ A, long long ld, int n_rows, int n_cols,
                                const long long* __restrict__ indptr, int* indices, double* data) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int lane = threadIdx.x & 31;
    long long pos = indptr[r];
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int c = c0 + lane;
        const double v = c < n_cols ? A[(long long)r * ld + c] : 0.0;
        const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const long long p = pos + __popc(mask & ((1u << lane) - 1));
            indices[p] = c;
            data[p] = v;
        }
        pos += __popc(mask);
    }
}
int launch_count_nnz(const double* A, long long ld, int n_rows, int n_cols, long long* row_nnz, cudaStream_t s) {
    count_nnz_kernel<<<n_rows, 256, 0, s>>>(A, ld, n_cols, row_nnz);
    return launch_status();
}
int launch_fill_csr(const double* A, long long ld, int n_rows, int n_cols, const long long* indptr, int* indices,
                    double* data, cudaStream_t s) {
    fill_csr_kernel<<<cdiv(n_rows, 8), 256, 0, s>>>(A, ld, n_rows, n_cols, indptr, indices, data);
    return launch_status();
}

}  // namespace hb
