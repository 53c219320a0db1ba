// Consumers of an assembled varf matrix that stays resident in HBM.
//
// Reference: hermipy/varf.py:127-134 (`Varf.__call__`: coeffs = matrix.dot(series.coeffs)) and, through
// it, the Krylov solvers of varf.py:170-237 (`solve(use_gmres=True)`, `eigs`) whose only contact with the
// matrix is this product.  At C2 cube the matrix is 13 GB (C3 cube: 207 GB row-sharded), so the product
// runs where the matrix was assembled; only the vectors (|I| doubles) travel.
//
// y[r] = sum_c A[r][c] x[c]: 2 flops per 8 bytes of A -> HBM-bound; algorithmic bytes = 8 M N (+ 8 (M + N)).
// One CTA owns ROWS rows with the SAME 16-byte phase (consecutive rows when lda is even, every second row
// when it is odd), so that one column partition serves all of them: every x value is fetched once per CTA
// and used ROWS times (x is re-read from L2 by each CTA: 1/ROWS of the matrix traffic), and every matrix
// load is a 16-byte streaming load.  The sums are a fixed tree (thread-strided partials, warp shuffles,
// one pass over the warps), so a product is reproducible run to run.
#include "hb_kernels.h"
#include "hb_plan.h"

#include <cstdint>

namespace hb {

namespace {

__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int ROWS, int NT>
__global__ void __launch_bounds__(NT) matvec_kernel(int M, int N, const double* __restrict__ A, long long lda,
                                                    const double* __restrict__ x, double* __restrict__ y) {
    // rows of this CTA: r0 + i * rs.  lda odd: CTA pairs (2g, 2g+1) take the even / odd rows of a 2*ROWS group.
    const bool odd_ld = (lda & 1) != 0;
    const int rs = odd_ld ? 2 : 1;
    const int r0 = odd_ld ? (blockIdx.x >> 1) * (2 * ROWS) + (blockIdx.x & 1) : blockIdx.x * ROWS;
    if (r0 >= M) return;
    const double* row[ROWS];
    bool live[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) {
        const int r = r0 + i * rs;
        live[i] = r < M;
        row[i] = A + (long long)(live[i] ? r : r0) * lda;   // dead rows alias row r0 (loads stay in bounds)
    }
    // leading elements up to the first 16-byte boundary of the rows (same for all rows of the CTA)
    const int head = (int)((reinterpret_cast<uintptr_t>(row[0]) >> 3) & 1);
    const int npair = (N - head) >> 1;
    double acc[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) acc[i] = 0.0;

    int j = threadIdx.x;
    // two column pairs per row in flight per thread: 2 * ROWS * 16 bytes outstanding
    for (; j + NT < npair; j += 2 * NT) {
        const int c0 = head + 2 * j, c1 = c0 + 2 * NT;
        double2 a0[ROWS], a1[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
            a0[i] = ldg_stream2(row[i] + c0);
            a1[i] = ldg_stream2(row[i] + c1);
        }
        const double x00 = __ldg(x + c0), x01 = __ldg(x + c0 + 1);
        const double x10 = __ldg(x + c1), x11 = __ldg(x + c1 + 1);
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
            acc[i] = fma(a0[i].x, x00, acc[i]);
            acc[i] = fma(a0[i].y, x01, acc[i]);
            acc[i] = fma(a1[i].x, x10, acc[i]);
            acc[i] = fma(a1[i].y, x11, acc[i]);
        }
    }
    if (j < npair) {
        const int c0 = head + 2 * j;
        const double x00 = __ldg(x + c0), x01 = __ldg(x + c0 + 1);
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
            const double2 a0 = ldg_stream2(row[i] + c0);
            acc[i] = fma(a0.x, x00, acc[i]);
            acc[i] = fma(a0.y, x01, acc[i]);
        }
    }
    if (threadIdx.x == 0) {   // the unpaired first / last columns
        if (head) {
            const double xv = x[0];
#pragma unroll
            for (int i = 0; i < ROWS; i++) acc[i] = fma(row[i][0], xv, acc[i]);
        }
        const int last = head + 2 * npair;
        if (last < N) {
            const double xv = x[last];
#pragma unroll
            for (int i = 0; i < ROWS; i++) acc[i] = fma(row[i][last], xv, acc[i]);
        }
    }
    __shared__ double part[ROWS][NT / 32];
#pragma unroll
    for (int i = 0; i < ROWS; i++) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) part[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < ROWS && r0 + (int)threadIdx.x * rs < M) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) v += part[threadIdx.x][w];
        y[r0 + threadIdx.x * rs] = v;
    }
}

}  // namespace

int launch_matvec(int M, int N, const double* A, long long lda, const double* x, double* y, cudaStream_t s) {
    if (M <= 0 || N <= 0) return HB_OK;
    constexpr int ROWS = 4, NT = 256;
    const bool odd_ld = (lda & 1) != 0;
    const int groups = odd_ld ? 2 * ((M + 2 * ROWS - 1) / (2 * ROWS)) : ((M + ROWS - 1) / ROWS);
    matvec_kernel<ROWS, NT><<<groups, NT, 0, s>>>(M, N, A, lda, x, y);
    return launch_status();
}

}  // namespace hb
