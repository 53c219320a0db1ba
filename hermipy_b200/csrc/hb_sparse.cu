// Sparse (CSR) results without a dense intermediate.  Reference: the spmat instantiations of varf
// (cpp/src/hermite/varf.cpp:458-460), tensorize (tensorize.cpp:197-283), project (project.cpp:78-109), varfd
// (varf.cpp:290-376) and the exchange format row_col_val / to_spmat (sparse.cpp:28-100).
//
// Every operation here has the same shape: each output entry comes from one cheaply enumerable CANDIDATE (a band
// offset of a row, a pair of stored entries of two factors, one stored entry of the input), most of which exist.
// A kernel per operation writes the candidates as (key = row << 32 | column, value) with dropped ones keyed ~0
// (outside the index set, or exactly zero: the reference stores non-zeros only, sparse.cpp:52); one radix sort
// (CUB, library code beside the path) orders them by row and column, and three small kernels cut the result into
// indptr / indices / data.  Work and memory are O(candidates), never O(n_polys^2).
#include <cub/cub.cuh>

#include "hb_kernels.h"
#include "hb_plan.h"

namespace hb {

static const unsigned long long kDropped = ~0ull;

__global__ void csr_count_kernel(const unsigned long long* __restrict__ keys, long long n, long long* nnz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keys[i] != kDropped && (i + 1 == n || keys[i + 1] == kDropped)) *nnz = i + 1;   // sorted: dropped keys last
}
// indptr[r] = first sorted entry whose row is >= r (binary search), r = 0 .. n_rows
__global__ void csr_indptr_kernel(const unsigned long long* __restrict__ keys, const long long* __restrict__ nnz_p,
                                  long long n_rows, long long* indptr) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    const long long nnz = *nnz_p;
    const unsigned long long want = (unsigned long long)r << 32;
    long long lo = 0, hi = nnz;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    indptr[r] = lo;
}
__global__ void csr_split_kernel(const unsigned long long* __restrict__ keys, const long long* __restrict__ nnz_p,
                                 int* indices) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *nnz_p) indices[i] = (int)(keys[i] & 0xffffffffull);
}

int csr_from_candidates(long long n_rows, long long n_cand, CsrResult& out, cudaStream_t st) {
    out.rows = n_rows;
    out.nnz = 0;
    HB_TRY(out.indptr.ensure((size_t)(n_rows + 1) * 8));
    HB_TRY(out.scalars.ensure(16));
    HB_CUDA(cudaMemsetAsync(out.scalars.p, 0, 16, st));
    if (n_cand > 0) {
        HB_TRY(out.keys_sorted.ensure((size_t)n_cand * 8));
        HB_TRY(out.data.ensure((size_t)n_cand * 8));
        HB_TRY(out.indices.ensure((size_t)n_cand * 4));
        const int bits = 64;   // dropped keys (~0) must sort last: all bits take part
        size_t temp = 0;
        auto* kin = static_cast<unsigned long long*>(out.keys.p);
        auto* kout = static_cast<unsigned long long*>(out.keys_sorted.p);
        auto* vin = static_cast<double*>(out.vals.p);
        auto* vout = static_cast<double*>(out.data.p);
        if (cub::DeviceRadixSort::SortPairs(nullptr, temp, kin, kout, vin, vout, n_cand, 0, bits, st) != cudaSuccess)
            return HB_ERR_CUDA;
        HB_TRY(out.temp.ensure(temp));
        if (cub::DeviceRadixSort::SortPairs(out.temp.p, temp, kin, kout, vin, vout, n_cand, 0, bits, st) != cudaSuccess)
            return HB_ERR_CUDA;
        note_launch(1);
        long long* nnz_p = static_cast<long long*>(out.scalars.p);
        csr_count_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(kout, n_cand, nnz_p);
        csr_indptr_kernel<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(kout, nnz_p, n_rows, static_cast<long long*>(out.indptr.p));
        csr_split_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(kout, nnz_p, out.indices.i());
        note_launch(3);
        HB_TRY(launch_status());
        HB_CUDA(cudaMemcpyAsync(&out.nnz, out.scalars.p, 8, cudaMemcpyDeviceToHost, st));
    } else {
        HB_CUDA(cudaMemsetAsync(out.indptr.p, 0, (size_t)(n_rows + 1) * 8, st));
    }
    HB_CUDA(cudaStreamSynchronize(st));
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// varf, few alpha kept: the band of varf_band_kernel (hb_kernels.cu), same term order and rounding, written as
// candidates instead of into a zero-filled dense matrix.
__global__ void varf_band_coo_kernel(VarfSparseArgs a, VarfBandArgs b, unsigned long long* __restrict__ keys,
                                     double* __restrict__ vals) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)(a.row_end - a.row_begin) * b.width_total;
    if (idx >= total) return;
    const int r = a.row_begin + (int)(idx / b.width_total);
    int off = (int)(idx % b.width_total);
    int m[HB_MAX_DIM], nn[HB_MAX_DIM];
    bool inside = true;
    for (int d = a.dim - 1; d >= 0; d--) {          // offsets: last axis fastest
        const int w = 2 * b.amax[d] + 1;
        const int o = off % w;
        off /= w;
        m[d] = a.midx[(long long)r * a.dim + d];
        nn[d] = m[d] + o - b.amax[d];
        inside = inside && nn[d] >= 0 && nn[d] < b.extent[d];
    }
    int c = -1;
    double acc = 0.0;
    if (inside) {
        long long lex = 0;
        for (int d = 0; d < a.dim; d++) lex = lex * (d == a.dim - 1 ? b.pitch_last : b.extent[d]) + nn[d];
        c = b.lut_box[lex];
    }
    if (c >= 0) {
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
    }
    const bool keep = c >= 0 && acc != 0.0;
    keys[idx] = keep ? ((unsigned long long)(unsigned)(r - a.row_begin) << 32 | (unsigned)c) : kDropped;
    vals[idx] = acc;
}

int varf_band_csr(const VarfSparseArgs& a, const VarfBandArgs& b, CsrResult& out, cudaStream_t st) {
    const long long rows = a.row_end - a.row_begin;
    const long long total = rows * b.width_total;
    if (total > 0) {
        HB_TRY(out.keys.ensure((size_t)total * 8));
        HB_TRY(out.vals.ensure((size_t)total * 8));
        varf_band_coo_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, b, static_cast<unsigned long long*>(out.keys.p),
                                                                              static_cast<double*>(out.vals.p));
        note_launch(1);
        HB_TRY(launch_status());
    }
    return csr_from_candidates(rows, total, out, st);
}

// ------------------------------------------------------------------------------------------------
// tensorize of k CSR factors (tensorize.cpp:197-283 folded over the factors, :286-332).  Output row r of the
// full set restricts to row map[j][r] of factor j; its stored columns are the combinations of one stored column
// per factor whose combined multi-index lies in the set.  cand_off[r] = first candidate of row r.
__global__ void tensorize_count_kernel(SpTensorizeArgs a, long long* __restrict__ count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_out) return;
    long long c = 1;
    for (int j = 0; j < a.k; j++) {
        const int rj = a.map[j][r];
        c *= rj < 0 ? 0 : a.indptr[j][rj + 1] - a.indptr[j][rj];
    }
    count[r] = c;
}
__global__ void tensorize_coo_kernel(SpTensorizeArgs a, const long long* __restrict__ cand_off,
                                     unsigned long long* __restrict__ keys, double* __restrict__ vals) {
    const int r = blockIdx.x;
    const long long begin = cand_off[r], n = cand_off[r + 1] - begin;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) {
        long long rest = t, lex = 0;
        double v = 0.0;
        // mixed radix over the factors, last factor fastest; the value is ((v0 * v1) * v2) ... as the fold forms it
        long long lexpart[HB_MAX_FACTORS];
        double vj[HB_MAX_FACTORS];
        for (int j = a.k - 1; j >= 0; j--) {
            const int rj = a.map[j][r];
            const long long p0 = a.indptr[j][rj], cnt = a.indptr[j][rj + 1] - p0;
            const long long e = p0 + rest % cnt;
            rest /= cnt;
            vj[j] = a.data[j][e];
            lexpart[j] = a.col_lex[j][a.indices[j][e]];   // contribution of this factor's column multi-index to the box offset
        }
        bool in_box = true;
        for (int j = 0; j < a.k; j++) {
            in_box = in_box && lexpart[j] >= 0;
            lex += lexpart[j];
            v = j == 0 ? vj[0] : __dmul_rn(v, vj[j]);
        }
        const int c = in_box ? a.lex_to_pos[lex] : -1;
        const bool keep = c >= 0 && v != 0.0;
        keys[begin + t] = keep ? ((unsigned long long)(unsigned)r << 32 | (unsigned)c) : kDropped;
        vals[begin + t] = v;
    }
}

int tensorize_csr(const SpTensorizeArgs& a, CsrResult& out, cudaStream_t st) {
    HB_TRY(out.offsets.ensure((size_t)(a.n_out + 1) * 8));
    long long* off = static_cast<long long*>(out.offsets.p);
    HB_CUDA(cudaMemsetAsync(off, 0, (size_t)(a.n_out + 1) * 8, st));
    tensorize_count_kernel<<<(a.n_out + 255) / 256, 256, 0, st>>>(a, off);
    note_launch(1);
    size_t temp = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, temp, off, off, a.n_out + 1, st) != cudaSuccess) return HB_ERR_CUDA;
    HB_TRY(out.temp.ensure(temp));
    if (cub::DeviceScan::ExclusiveSum(out.temp.p, temp, off, off, a.n_out + 1, st) != cudaSuccess) return HB_ERR_CUDA;
    long long total = 0;
    HB_CUDA(cudaMemcpyAsync(&total, off + a.n_out, 8, cudaMemcpyDeviceToHost, st));
    HB_CUDA(cudaStreamSynchronize(st));
    if (total > 0) {
        HB_TRY(out.keys.ensure((size_t)total * 8));
        HB_TRY(out.vals.ensure((size_t)total * 8));
        tensorize_coo_kernel<<<a.n_out, 128, 0, st>>>(a, off, static_cast<unsigned long long*>(out.keys.p),
                                                     static_cast<double*>(out.vals.p));
        note_launch(1);
        HB_TRY(launch_status());
    }
    return csr_from_candidates(a.n_out, total, out, st);
}

// ------------------------------------------------------------------------------------------------
// One candidate per stored entry of a CSR input: the row and the column are renamed through maps (-1 drops the
// entry) and the value is scaled.  project (project.cpp:78-109): both maps = position in the sub-set of an aligned
// multi-index, factor 1.  varfd (varf.cpp:343-372): rows keep their index, column c goes to index(m_c +- e_dir)
// with the factor of that column.
__global__ void relabel_coo_kernel(int n_rows, const long long* __restrict__ indptr, const int* __restrict__ indices,
                                   const double* __restrict__ data, const int* __restrict__ row_map,
                                   const int* __restrict__ col_map, const double* __restrict__ col_fac,
                                   unsigned long long* __restrict__ keys, double* __restrict__ vals) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const int nr = row_map ? row_map[r] : r;
    for (long long e = indptr[r] + (threadIdx.x & 31); e < indptr[r + 1]; e += 32) {
        const int nc = col_map[indices[e]];
        const double v = col_fac ? __dmul_rn(data[e], col_fac[indices[e]]) : data[e];
        const bool keep = nr >= 0 && nc >= 0 && v != 0.0;
        keys[e] = keep ? ((unsigned long long)(unsigned)nr << 32 | (unsigned)nc) : kDropped;
        vals[e] = v;
    }
}

int relabel_csr(int n_rows_in, long long nnz_in, const long long* indptr, const int* indices, const double* data,
                const int* row_map, const int* col_map, const double* col_fac, long long n_rows_out, CsrResult& out,
                cudaStream_t st) {
    if (nnz_in > 0) {
        HB_TRY(out.keys.ensure((size_t)nnz_in * 8));
        HB_TRY(out.vals.ensure((size_t)nnz_in * 8));
        relabel_coo_kernel<<<(n_rows_in + 7) / 8, 256, 0, st>>>(n_rows_in, indptr, indices, data, row_map, col_map, col_fac,
                                                               static_cast<unsigned long long*>(out.keys.p),
                                                               static_cast<double*>(out.vals.p));
        note_launch(1);
        HB_TRY(launch_status());
    }
    return csr_from_candidates(n_rows_out, nnz_in, out, st);
}

}  // namespace hb
