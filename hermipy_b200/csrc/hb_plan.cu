// Plans: sum-factorised d-dimensional Hermite transform (one DMMA GEMM per axis) and the varf
// assembly, on device-resident data.  Reference behaviour: cpp/src/hermite/transform.cpp:44-149
// (_transform), cpp/src/hermite/varf.cpp:164-269 (varf).
#include "hb_plan.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "hb_gemm.cuh"
#include "hb_kernels.h"

namespace hb {

// ---------------------------------------------------------------------------------------------
// Error text of the calling thread.  hb_last_error() hands the message out ONCE: a later failure that sets no text
// of its own then reports "(no detail)" instead of repeating the message of an older, unrelated error.
static thread_local char g_err[512] = "";
static thread_local char g_err_out[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() {
    if (g_err[0]) memcpy(g_err_out, g_err, sizeof(g_err));
    else snprintf(g_err_out, sizeof(g_err_out), "(no detail)");
    g_err[0] = 0;
    return g_err_out;
}
int invalid(const char* where) {
    set_error("%s: invalid arguments", where);
    return HB_ERR_INVALID;
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error '%s' in %s", cudaGetErrorString(e), what);
    return HB_ERR_CUDA;
}
void note_launch(int n) { g_launches += n; }
long long launch_count(int reset) {
    long long v = g_launches.load();
    if (reset) g_launches = 0;
    return v;
}
int launch_status() {
    note_launch(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? HB_OK : cuda_fail(e, "kernel launch");
}

DevBuf::~DevBuf() { release(); }
std::atomic<unsigned long long> g_devbuf_generation{0};

void DevBuf::release() {
    if (p) {
        cudaFree(p);
        g_devbuf_generation.fetch_add(1, std::memory_order_relaxed);   // captured graphs hold this address
    }
    p = nullptr;
    bytes = 0;
}
int DevBuf::ensure(size_t need) {
    if (need <= bytes) return HB_OK;
    release();
    need = (need + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc of %zu bytes failed: %s", need, cudaGetErrorString(e));
        cudaGetLastError();
        return HB_ERR_NOMEM;
    }
    bytes = need;
    return HB_OK;
}
int DevBuf::ensure_zero(size_t need, cudaStream_t s) {
    HB_TRY(ensure(need));
    HB_CUDA(cudaMemsetAsync(p, 0, need, s));
    return HB_OK;
}

}  // namespace hb

hb_plan::~hb_plan() {
    delete plan2;
    if (fill_stream) cudaStreamDestroy(fill_stream);
    if (fill_ready) cudaEventDestroy(fill_ready);
    if (fill_done) cudaEventDestroy(fill_done);
    if (h_Hf_pinned) cudaFreeHost(h_Hf_pinned);
    for (Timed& t : timed) {
        if (t.e0) cudaEventDestroy(t.e0);
        if (t.e1) cudaEventDestroy(t.e1);
    }
}

namespace hb {

int timed_gemm(hb_plan* pl, const GemmArgs& g, int tag, cudaStream_t st) {
    if (!pl->timing) return gemm_launch(g, st);
    // developer timeline: HB_TRACE_TAG restricts the per-CTA trace (hb_debug_trace) to the launches with this tag
    long long* trace = gemm_get_trace();
    const char* only = trace ? getenv("HB_TRACE_TAG") : nullptr;
    const bool mute = only && atoi(only) != tag;
    if (mute) gemm_set_trace(nullptr);
    struct Restore { long long* t; bool on; ~Restore() { if (on) gemm_set_trace(t); } } restore{trace, mute};
    if (pl->n_timed >= (int)pl->timed.size()) {
        hb_plan::Timed t;
        HB_CUDA(cudaEventCreate(&t.e0));
        HB_CUDA(cudaEventCreate(&t.e1));
        pl->timed.push_back(t);
    }
    hb_plan::Timed& t = pl->timed[pl->n_timed++];
    t.tag = tag;
    t.flops = 2.0 * g.M * (double)g.N * g.K * g.batch * g.batch_lo;
    HB_CUDA(cudaEventRecord(t.e0, st));
    const int rc = gemm_launch(g, st);
    HB_CUDA(cudaEventRecord(t.e1, st));
    return rc;
}

// ---------------------------------------------------------------------------------------------
int plan_create(hb_plan** out, int dim, const int* n_pts, int degree, const double* nodes, const double* weights,
                const int* do_fourier, int set) {
    *out = nullptr;
    if (dim < 1 || dim > HB_MAX_DIM || degree < 0 || !n_pts || !nodes || !weights) {
        set_error("plan_create: invalid arguments");
        return HB_ERR_INVALID;
    }
    if (set < 0) return HB_ERR_INDEX_SET;
    for (int a = 0; a < dim; a++) {
        if (n_pts[a] < 1) return HB_ERR_INVALID;
        if (do_fourier && do_fourier[a] == 1 && (degree % 2 == 1)) {
            set_error("Use even degree for Fourier.");  // transform.cpp:92-96
            return HB_ERR_FOURIER_DEGREE;
        }
    }
    auto iset = get_index_set(set, dim, degree);
    if (!iset) {
        set_error("plan_create: index set not constructible (set %d, dim %d, degree %d)", set, dim, degree);
        return HB_ERR_INVALID;
    }
    std::unique_ptr<hb_plan> pl(new hb_plan);
    pl->dim = dim; pl->degree = degree; pl->set = set; pl->iset = iset;
    pl->n_polys = iset->size;
    pl->ax.resize(dim);
    cudaStream_t st = 0;
    size_t off = 0;
    for (int a = 0; a < dim; a++) {
        Axis& A = pl->ax[a];
        A.n = n_pts[a];
        A.P = (int)iset->extent[a];
        A.Pp = (int)even_up(A.P);
        A.np = (int)even_up(A.n);
        A.fourier = do_fourier ? do_fourier[a] : 0;
        A.h_x.assign(nodes + off, nodes + off + A.n);
        A.h_w.assign(weights + off, weights + off + A.n);
        off += A.n;
        HB_TRY(A.x.ensure(A.n * 8));
        HB_TRY(A.w.ensure(A.n * 8));
        HB_CUDA(cudaMemcpyAsync(A.x.p, A.h_x.data(), A.n * 8, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaMemcpyAsync(A.w.p, A.h_w.data(), A.n * 8, cudaMemcpyHostToDevice, st));
        HB_TRY(A.plain.ensure_zero((size_t)A.n * A.Pp * 8, st));
        HB_TRY(A.wgt.ensure_zero((size_t)A.n * A.Pp * 8, st));
        HB_TRY(A.plainT.ensure_zero((size_t)A.P * A.np * 8, st));
        HB_TRY(A.wgtT.ensure_zero((size_t)A.P * A.np * 8, st));
        HB_TRY(launch_basis_tables(A.x.d(), A.w.d(), A.n, A.P, A.fourier, A.plain.d(), A.wgt.d(), A.Pp,
                                   A.plainT.d(), A.wgtT.d(), A.np, st));
        // slots of the blocked pair tables (dense varf); parity-sorted on Hermite axes
        auto add_group = [&](int first, int step) {
            int count = 0;
            for (int k = first; k < A.P; k += step) { A.slots.push_back(k); count++; }
            while (A.slots.size() % 8) A.slots.push_back(-1);
            return (count + 7) / 8;
        };
        // Only the axis whose contraction is halved by parity (axis 0 of the 2-D Gram varf) needs its slots sorted by
        // parity; on the other axis natural order keeps consecutive basis indices -- consecutive output positions --
        // next to each other inside a tile (contiguous runs instead of every second entry).
        // Which one: sets enumerated along an axis (cube, rectangle) gain from natural order (C2 cube stage 2: 39 GB ->
        // 31 GB of DRAM traffic, 17.1 -> 16.4 ms); graded sets (triangle, cross) run along anti-diagonals, where both
        // axes must step alike (sorted on both: C3 triangle 97 ms against 109 ms).  HB_VARF_SORT_ALL_AXES=0/1 overrides.
        bool natural_axis1 = false;
        if (dim == 2 && a == 1 && !A.fourier) {
            const int P0 = (int)iset->extent[0], P1 = (int)iset->extent[1];
            auto pos = [&](int i, int j) { return (i < 0 || j < 0 || i >= P0 || j >= P1) ? -1 : iset->lex_to_pos[(size_t)i * P1 + j]; };
            long long along_axis = 0, along_diag = 0;
            for (int i = 0; i < P0; i++)
                for (int j = 0; j < P1; j++) {
                    const int v = pos(i, j);
                    if (v < 0) continue;
                    along_axis += (pos(i, j + 1) == v + 1) + (pos(i + 1, j) == v + 1);
                    along_diag += (pos(i + 1, j - 1) == v + 1) + (pos(i - 1, j + 1) == v + 1);
                }
            natural_axis1 = along_axis > along_diag;
            const char* e = getenv("HB_VARF_SORT_ALL_AXES");
            if (e && e[0]) natural_axis1 = e[0] == '0';
        }
        pl->varf_natural_axis1 = natural_axis1;
        if (A.fourier || natural_axis1) {
            A.nb_even = add_group(0, 1);
        } else {
            A.nb_even = add_group(0, 2);
            add_group(1, 2);
        }
        A.nb = (int)A.slots.size() / 8;
        A.symmetric_nodes = !A.fourier;
        for (int i = 0; i < A.n && A.symmetric_nodes; i++)
            A.symmetric_nodes = A.h_x[i] == -A.h_x[A.n - 1 - i] && A.h_w[i] == A.h_w[A.n - 1 - i];
    }
    pl->grid_rows = 1;
    pl->box_rows = 1;
    for (int a = 0; a + 1 < dim; a++) {
        pl->grid_rows *= pl->ax[a].n;
        pl->box_rows *= pl->ax[a].P;
    }
    pl->grid_size = pl->grid_rows * pl->ax[dim - 1].np;
    pl->box_size = pl->box_rows * pl->ax[dim - 1].Pp;
    // padded lexicographic box -> position
    {
        std::vector<int> lut((size_t)pl->box_size, -1);
        const int Pl = pl->ax[dim - 1].P, Ppl = pl->ax[dim - 1].Pp;
        const size_t n_lex = iset->lex_to_pos.size();
        for (size_t lex = 0; lex < n_lex; lex++) {
            const size_t row = lex / Pl, col = lex % Pl;
            lut[row * Ppl + col] = iset->lex_to_pos[lex];
        }
        HB_TRY(pl->lut_box.ensure(lut.size() * 4));
        HB_CUDA(cudaMemcpyAsync(pl->lut_box.p, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaStreamSynchronize(st));
    }
    *out = pl.release();
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// 1-D transform on a symmetric node set (every Gauss-Hermite rule): phi_k(-x) = (-1)^k phi_k(x) bit for bit, so
// the even coefficients only see f(x) + f(-x) and the odd ones f(x) - f(-x), both on the nh = ceil(n/2)
// non-negative nodes: two GEMMs of a quarter of the size each instead of one -- half the flops.  The even / odd
// results are interleaved by the store epilogue (column stride 2); the inverse mirrors it (de-interleave, two
// GEMMs, butterfly).  The extra passes move 8 (n + P) bytes per function each way, against 2 n P flops saved.
// HB_TRANSFORM_PARITY: 0 = off, 2 = on for every stage, unset / 1 = on where the stage is large enough for the
// extra split / merge passes (one read + one write of the operand) to pay off.  Read on every call (tests toggle it).
static int parity_mode() {
    const char* e = getenv("HB_TRANSFORM_PARITY");
    return e && e[0] ? atoi(e) : 1;
}
static bool parity_stage(const Axis& A, double stage_flops) {
    const int mode = parity_mode();
    if (mode == 0 || !A.symmetric_nodes || A.n < 4 || A.P < 4) return false;
    return mode == 2 || stage_flops >= 4e9;
}

struct Parity1D { int nh, nhp, lo, Pe, Po, Pep, Pop; };
static Parity1D parity_dims(const Axis& A) {
    Parity1D d;
    d.nh = (A.n + 1) / 2; d.nhp = (int)even_up(d.nh); d.lo = A.n - d.nh;
    d.Pe = (A.P + 1) / 2; d.Po = A.P / 2; d.Pep = (int)even_up(d.Pe); d.Pop = (int)even_up(d.Po);
    return d;
}

static int ensure_parity_tables(Axis& A, cudaStream_t st) {
    if (A.par_tables) return HB_OK;
    const Parity1D d = parity_dims(A);
    HB_TRY(A.pWe.ensure((size_t)d.nh * d.Pep * 8));
    HB_TRY(A.pWo.ensure((size_t)d.nh * d.Pop * 8));
    HB_TRY(A.pPhiE.ensure((size_t)d.nh * d.Pep * 8));
    HB_TRY(A.pPhiO.ensure((size_t)d.nh * d.Pop * 8));
    HB_TRY(A.pWeT.ensure((size_t)d.Pe * d.nhp * 8));
    HB_TRY(A.pWoT.ensure((size_t)d.Po * d.nhp * 8));
    HB_TRY(A.pPe.ensure((size_t)d.Pe * d.nhp * 8));
    HB_TRY(A.pPo.ensure((size_t)d.Po * d.nhp * 8));
    // [node h][basis]: rows lo + h, every second column
    HB_TRY(launch_strided_copy(A.wgt.d(), A.Pp, d.lo, 1, 0, 2, A.pWe.d(), d.Pep, d.nh, d.Pe, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.wgt.d(), A.Pp, d.lo, 1, 1, 2, A.pWo.d(), d.Pop, d.nh, d.Po, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.plain.d(), A.Pp, d.lo, 1, 0, 2, A.pPhiE.d(), d.Pep, d.nh, d.Pe, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.plain.d(), A.Pp, d.lo, 1, 1, 2, A.pPhiO.d(), d.Pop, d.nh, d.Po, 1, 0, 0, st));
    // [basis][node h]: every second row, columns lo + h
    HB_TRY(launch_strided_copy(A.wgtT.d(), A.np, 0, 2, d.lo, 1, A.pWeT.d(), d.nhp, d.Pe, d.nh, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.wgtT.d(), A.np, 1, 2, d.lo, 1, A.pWoT.d(), d.nhp, d.Po, d.nh, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.plainT.d(), A.np, 0, 2, d.lo, 1, A.pPe.d(), d.nhp, d.Pe, d.nh, 1, 0, 0, st));
    HB_TRY(launch_strided_copy(A.plainT.d(), A.np, 1, 2, d.lo, 1, A.pPo.d(), d.nhp, d.Po, d.nh, 1, 0, 0, st));
    A.par_tables = true;
    return HB_OK;
}

// The coefficient box in which the axes of `mask` are parity-sorted: per sorted axis the index runs over the even
// basis functions first and then the odd ones, both blocks padded to the same extent (Pe = ceil(P/2) rows; on the
// last axis Pep = Pe rounded up to even columns, so that both blocks are 16-byte aligned).  Returns the look-up
// table box cell -> position in the index set (-1 for padding and for cells outside the set).
static int sorted_box(hb_plan* pl, int mask, hb_plan::SortedBox** out, cudaStream_t st) {
    auto it = pl->sorted_boxes.find(mask);
    if (it != pl->sorted_boxes.end()) { *out = &it->second; return HB_OK; }
    const int d = pl->dim;
    std::vector<std::vector<int>> index(d);   // box coordinate -> basis index (-1 padding)
    for (int a = 0; a < d; a++) {
        const Axis& A = pl->ax[a];
        const Parity1D q = parity_dims(A);
        const bool last = (a == d - 1);
        if (mask >> a & 1) {   // both parity blocks have the same extent: Pe rows, or Pep columns on the last axis
            const int blk = last ? q.Pep : q.Pe;
            for (int k = 0; k < q.Pe; k++) index[a].push_back(2 * k);
            while ((int)index[a].size() < blk) index[a].push_back(-1);
            for (int k = 0; k < q.Po; k++) index[a].push_back(2 * k + 1);
            while ((int)index[a].size() < 2 * blk) index[a].push_back(-1);
        } else {
            for (int k = 0; k < A.P; k++) index[a].push_back(k);
            if (last) while ((int)index[a].size() < A.Pp) index[a].push_back(-1);
        }
    }
    long long total = 1;
    for (int a = 0; a < d; a++) total *= (long long)index[a].size();
    std::vector<int> lut((size_t)total, -1);
    std::vector<int> coord(d, 0);
    for (long long cell = 0; cell < total; cell++) {
        long long lex = 0;
        bool ok = true;
        for (int a = 0; a < d; a++) {
            const int k = index[a][coord[a]];
            if (k < 0) { ok = false; break; }
            lex = lex * pl->ax[a].P + k;
        }
        if (ok) lut[(size_t)cell] = pl->iset->lex_to_pos[(size_t)lex];
        for (int a = d - 1; a >= 0; a--) {        // next cell, last axis fastest
            if (++coord[a] < (int)index[a].size()) break;
            coord[a] = 0;
        }
    }
    hb_plan::SortedBox& box = pl->sorted_boxes[mask];
    box.pitch = (long long)index[d - 1].size();
    box.size = total;
    HB_TRY(box.lut.ensure((size_t)total * 4));
    HB_CUDA(cudaMemcpyAsync(box.lut.p, lut.data(), (size_t)total * 4, cudaMemcpyHostToDevice, st));
    HB_CUDA(cudaStreamSynchronize(st));   // `lut` goes out of scope
    *out = &box;
    return HB_OK;
}

// Contraction of the LAST grid axis of `rows` rows (pitch f_pitch) into interleaved even / odd coefficients
// (out row pitch out_pitch): split + two half-size GEMMs.  This is the whole forward transform when dim == 1.
static int forward_last_axis_parity(hb_plan* pl, Axis& A, const double* d_f, long long f_pitch, long long rows,
                                    double* d_out, long long out_pitch, int tag, cudaStream_t st, bool sorted = false) {
    const Parity1D d = parity_dims(A);
    HB_TRY(ensure_parity_tables(A, st));
    HB_TRY(pl->parA.ensure((size_t)rows * d.nhp * 8));
    HB_TRY(pl->parB.ensure((size_t)rows * d.nhp * 8));
    HB_TRY(launch_rows_parity_split(d_f, f_pitch, A.n, d.nh, pl->parA.d(), pl->parB.d(), d.nhp, (int)rows, st));
    for (int par = 0; par < 2; par++) {
        GemmArgs g;
        g.M = (int)rows; g.K = d.nh; g.N = par ? d.Po : d.Pe;
        g.A = par ? pl->parB.d() : pl->parA.d(); g.lda = d.nhp;
        g.B = par ? A.pWo.d() : A.pWe.d(); g.ldb = par ? d.Pop : d.Pep;
        g.ldc = out_pitch;
        if (sorted) {   // [even block | odd block at Pep]: plain 16-byte stores
            g.C = d_out + (par ? d.Pep : 0);
        } else {        // natural order: even / odd coefficients interleave (1-D transform)
            g.C = d_out; g.col_stride = 2; g.col_offset = par;
        }
        HB_TRY(timed_gemm(pl, g, tag, st));
    }
    return HB_OK;
}

// Inverse of it: rows of interleaved coefficients (pitch c_pitch) -> grid rows (pitch f_pitch).
static int backward_last_axis_parity(hb_plan* pl, Axis& A, const double* d_coef, long long c_pitch, long long rows,
                                     double* d_f, long long f_pitch, int tag, cudaStream_t st, bool sorted = false) {
    const Parity1D d = parity_dims(A);
    HB_TRY(ensure_parity_tables(A, st));
    const size_t cw = (size_t)d.Pep + d.Pop;
    if (!sorted) HB_TRY(pl->parA.ensure((size_t)rows * cw * 8));
    HB_TRY(pl->parB.ensure((size_t)rows * 2 * d.nhp * 8));
    const double* ce = pl->parA.d();
    const double* co = ce + (size_t)rows * d.Pep;
    long long lde = d.Pep, ldo = d.Pop;
    double* ge = pl->parB.d();
    double* go = ge + (size_t)rows * d.nhp;
    if (sorted) {   // the rows already hold [even block | odd block at Pep]
        ce = d_coef; co = d_coef + d.Pep; lde = ldo = c_pitch;
    } else {        // natural order (1-D transform): de-interleave
        HB_TRY(launch_deinterleave_rows(d_coef, c_pitch, d.Pe, d.Po, pl->parA.d(), d.Pep,
                                        pl->parA.d() + (size_t)rows * d.Pep, d.Pop, rows, st));
    }
    for (int par = 0; par < 2; par++) {
        GemmArgs g;
        g.M = (int)rows; g.K = par ? d.Po : d.Pe; g.N = d.nh;
        g.A = par ? co : ce; g.lda = par ? ldo : lde;
        g.B = par ? A.pPo.d() : A.pPe.d(); g.ldb = d.nhp;
        g.C = par ? go : ge; g.ldc = d.nhp;
        HB_TRY(timed_gemm(pl, g, tag, st));
    }
    return launch_rows_parity_merge(ge, go, d.nhp, A.n, d.nh, (int)std::min<long long>(A.np, f_pitch), d_f, f_pitch,
                                    (int)rows, st);
}

// Which axes of a d-dimensional (d >= 2) transform of `batch` functions are split by parity (bit a).  The split
// costs one extra pass over the grid values per direction (all axes at once), so it is taken when the transform is
// large enough for half the flops to pay for it; every symmetric axis is split then (at most three: the pass
// handles 2^3 mirror images per thread).
static int parity_mask(const hb_plan* pl, int batch) {
    const int mode = parity_mode();
    if (mode == 0) return 0;
    const int d = pl->dim;
    double flops = 0, prefix = (double)batch * pl->grid_rows * pl->ax[d - 1].n, cols = 1;
    for (int a = d - 1; a >= 0; a--) {
        const Axis& A = pl->ax[a];
        prefix /= A.n;
        flops += 2.0 * prefix * A.n * A.P * cols;
        cols *= A.P;
    }
    // With the one-pass split of round 2 the split pays at every size measured (tools/parity_threshold.py: 401 x 401,
    // degree 200: 1 function 74 -> 55 us per round trip, 17 functions 158 -> 135 us; 201 x 201, degree 100: 49 -> 45 us),
    // so the default threshold is 0; HB_TRANSFORM_PARITY_GFLOP restores one.
    static const double min_gflop = [] { const char* e = getenv("HB_TRANSFORM_PARITY_GFLOP"); return e && e[0] ? atof(e) : 0.0; }();
    if (mode != 2 && flops < min_gflop * 1e9) return 0;
    int mask = 0, count = 0;
    for (int a = d - 1; a >= 0 && count < 3; a--) {
        const Axis& A = pl->ax[a];
        if (A.symmetric_nodes && A.n >= 4 && A.P >= 4) { mask |= 1 << a; count++; }
    }
    return mask;
}

int ensure_combined_parity_tables(Axis& A, cudaStream_t st) {
    if (A.cpar_tables) return HB_OK;
    const Parity1D d = parity_dims(A);
    const size_t ik = (size_t)d.nh * d.Pep, ki = (size_t)d.Pe * d.nhp;
    HB_TRY(A.cW.ensure_zero(2 * ik * 8, st));
    HB_TRY(A.cPhi.ensure_zero(2 * ik * 8, st));
    HB_TRY(A.cWT.ensure_zero(2 * ki * 8, st));
    HB_TRY(A.cPhiT.ensure_zero(2 * ki * 8, st));
    for (int par = 0; par < 2; par++) {
        const int Pq = par ? d.Po : d.Pe;
        if (Pq == 0) continue;
        // [node h][basis]: rows lo + h, every second column
        HB_TRY(launch_strided_copy(A.wgt.d(), A.Pp, d.lo, 1, par, 2, A.cW.d() + par * ik, d.Pep, d.nh, Pq, 1, 0, 0, st));
        HB_TRY(launch_strided_copy(A.plain.d(), A.Pp, d.lo, 1, par, 2, A.cPhi.d() + par * ik, d.Pep, d.nh, Pq, 1, 0, 0, st));
        // [basis][node h]: every second row, columns lo + h
        HB_TRY(launch_strided_copy(A.wgtT.d(), A.np, par, 2, d.lo, 1, A.cWT.d() + par * ki, d.nhp, Pq, d.nh, 1, 0, 0, st));
        HB_TRY(launch_strided_copy(A.plainT.d(), A.np, par, 2, d.lo, 1, A.cPhiT.d() + par * ki, d.nhp, Pq, d.nh, 1, 0, 0, st));
    }
    A.cpar_tables = true;
    return HB_OK;
}

// Geometry of a d >= 2 transform whose axes in `mask` are split by parity.  A split axis is indexed by
// (parity, h) on the grid side -- 2 x nh entries, 2 x nhp on the last axis -- and by (parity, k') on the
// coefficient side -- 2 x Pe entries, 2 x Pep on the last axis (see sorted_box).
struct ParityGeom {
    int d = 0;
    long long G[HB_MAX_DIM], C[HB_MAX_DIM];   // grid-side / coefficient-side extent of every axis
    long long grid_size = 1, box_size = 1;
};
static ParityGeom parity_geometry(const hb_plan* pl, int mask) {
    ParityGeom g;
    g.d = pl->dim;
    for (int a = 0; a < g.d; a++) {
        const Axis& A = pl->ax[a];
        const bool last = (a == g.d - 1);
        if (mask >> a & 1) {
            const Parity1D q = parity_dims(A);
            g.G[a] = last ? 2 * q.nhp : 2 * q.nh;
            g.C[a] = last ? 2 * q.Pep : 2 * q.Pe;
        } else {
            g.G[a] = last ? A.np : A.n;
            g.C[a] = last ? A.Pp : A.P;
        }
        g.grid_size *= g.G[a];
        g.box_size *= g.C[a];
    }
    return g;
}

// natural grid layout <-> parity-sorted layout of this plan (batch entries f_stride / geom.grid_size apart)
static ParityNDArgs parity_pass_args(const hb_plan* pl, int mask, const ParityGeom& geom, long long f_stride, int batch) {
    ParityNDArgs pa;
    const int d = pl->dim;
    pa.d = d;
    long long nat = 1, srt = 1;
    for (int a = d - 1; a >= 0; a--) {
        const Axis& A = pl->ax[a];
        const Parity1D q = parity_dims(A);
        pa.n[a] = A.n; pa.nh[a] = q.nh; pa.split[a] = (mask >> a) & 1;
        pa.nat_stride[a] = nat;
        pa.srt_stride[a] = srt;
        pa.srt_par[a] = (a == d - 1 ? q.nhp : q.nh) * srt;
        nat *= (a == d - 1) ? A.np : A.n;
        srt *= geom.G[a];
    }
    pa.batch = batch; pa.nat_bstride = f_stride; pa.srt_bstride = geom.grid_size;
    pa.nat_pad_last = pl->ax[d - 1].np;
    return pa;
}

// Forward transform with the axes in `mask` split by parity: ONE pass sorts the grid values by parity along all
// split axes (nd_parity_kernel), then every axis contraction is one GEMM launch whose two low-level batch entries
// are the even and the odd half (half the flops of the plain stage), the last one scattering into set order.
static int forward_parity(hb_plan* pl, int mask, const double* d_f, long long f_stride, int batch, double* d_coef,
                          long long coef_stride, cudaStream_t st) {
    const int d = pl->dim, L = d - 1;
    const ParityGeom geom = parity_geometry(pl, mask);
    hb_plan::SortedBox* box = nullptr;
    HB_TRY(sorted_box(pl, mask, &box, st));
    for (int a = 0; a < d; a++)
        if (mask >> a & 1) HB_TRY(ensure_combined_parity_tables(pl->ax[a], st));
    // workspace: the sorted grid and every intermediate [grid axes < a][coefficient axes >= a]
    size_t need = (size_t)batch * geom.grid_size;
    {
        long long size = (long long)batch * geom.grid_size;
        for (int a = L; a >= 1; a--) {
            size = size / geom.G[a] * geom.C[a];
            need = std::max(need, (size_t)size);
        }
    }
    HB_TRY(pl->ws[0].ensure(need * 8));
    HB_TRY(pl->ws[1].ensure(need * 8));
    pl->n_timed = 0;
    HB_TRY(launch_nd_parity_split(parity_pass_args(pl, mask, geom, f_stride, batch), d_f, pl->ws[0].d(), st));
    long long prefix = (long long)batch * geom.grid_size / geom.G[L];   // rows of the last-axis stage
    {
        Axis& A = pl->ax[L];
        GemmArgs g;
        g.M = (int)prefix;
        g.A = pl->ws[0].d(); g.lda = geom.G[L];
        g.C = pl->ws[1].d(); g.ldc = geom.C[L];
        if (mask >> L & 1) {
            const Parity1D q = parity_dims(A);
            g.K = q.nh; g.N = q.Pe; g.batch_lo = 2;
            g.strideA_lo = q.nhp;
            g.B = A.cW.d(); g.ldb = q.Pep; g.strideB_lo = (long long)q.nh * q.Pep;
            g.strideC_lo = q.Pep;
        } else {
            g.K = A.n; g.N = A.Pp;
            g.B = A.wgt.d(); g.ldb = A.Pp;
        }
        HB_TRY(timed_gemm(pl, g, d, st));
    }
    int cur = 1;
    long long cols = geom.C[L];
    for (int a = L - 1; a >= 0; a--) {
        Axis& A = pl->ax[a];
        prefix /= geom.G[a];
        GemmArgs g;
        g.N = (int)cols; g.batch = (int)prefix;
        g.B = pl->ws[cur].d(); g.ldb = cols; g.strideB = geom.G[a] * cols;
        if (mask >> a & 1) {
            const Parity1D q = parity_dims(A);
            g.M = q.Pe; g.K = q.nh; g.batch_lo = 2;
            g.A = A.cWT.d(); g.lda = q.nhp; g.strideA_lo = (long long)q.Pe * q.nhp;
            g.strideB_lo = (long long)q.nh * cols;
            g.strideC_lo = (long long)q.Pe * cols;
            g.row_offset_lo = q.Pe;
        } else {
            g.M = A.P; g.K = A.n;
            g.A = A.wgtT.d(); g.lda = A.np;
        }
        if (a == 0) {
            g.C = d_coef; g.strideC = coef_stride; g.strideC_lo = 0; g.epi = EPI_LUT; g.lut = box->lut.i();
        } else {
            g.C = pl->ws[cur ^ 1].d(); g.ldc = cols; g.strideC = geom.C[a] * cols; g.row_offset_lo = 0;
        }
        HB_TRY(timed_gemm(pl, g, 1 + a, st));
        cols *= geom.C[a];
        cur ^= 1;
    }
    return HB_OK;
}

// Inverse of it: coefficients gathered into the parity-sorted box, one GEMM launch per axis (even / odd halves as
// low-level batch entries), and ONE butterfly pass f(+-x) = g_e +- g_o over all split axes at the end.
static int backward_parity(hb_plan* pl, int mask, const double* d_coef, long long coef_stride, int batch, double* d_f,
                           long long f_stride, cudaStream_t st) {
    const int d = pl->dim, L = d - 1;
    const ParityGeom geom = parity_geometry(pl, mask);
    hb_plan::SortedBox* box = nullptr;
    HB_TRY(sorted_box(pl, mask, &box, st));
    for (int a = 0; a < d; a++)
        if (mask >> a & 1) HB_TRY(ensure_combined_parity_tables(pl->ax[a], st));
    size_t need = (size_t)batch * geom.box_size;
    {
        long long size = (long long)batch * geom.box_size;
        for (int a = L; a >= 0; a--) {
            size = size / geom.C[a] * geom.G[a];
            need = std::max(need, (size_t)size);
        }
    }
    HB_TRY(pl->ws[0].ensure(need * 8));
    HB_TRY(pl->ws[1].ensure(need * 8));
    HB_TRY(launch_gather_lut(d_coef, coef_stride, box->lut.i(), pl->ws[0].d(), geom.box_size, geom.box_size, batch, st));
    pl->n_timed = 0;
    long long prefix = (long long)batch * geom.box_size / geom.C[L];
    {
        Axis& A = pl->ax[L];
        GemmArgs g;
        g.M = (int)prefix;
        g.A = pl->ws[0].d(); g.lda = geom.C[L];
        g.C = pl->ws[1].d(); g.ldc = geom.G[L];
        if (mask >> L & 1) {
            const Parity1D q = parity_dims(A);
            g.K = q.Pe; g.N = q.nh; g.batch_lo = 2;
            g.strideA_lo = q.Pep;
            g.B = A.cPhiT.d(); g.ldb = q.nhp; g.strideB_lo = (long long)q.Pe * q.nhp;
            g.strideC_lo = q.nhp;
        } else {
            g.K = A.P; g.N = A.np;
            g.B = A.plainT.d(); g.ldb = A.np;
        }
        HB_TRY(timed_gemm(pl, g, 10 + d, st));
    }
    int cur = 1;
    long long cols = geom.G[L];
    for (int a = L - 1; a >= 0; a--) {
        Axis& A = pl->ax[a];
        prefix /= geom.C[a];
        GemmArgs g;
        g.N = (int)cols; g.batch = (int)prefix;
        g.B = pl->ws[cur].d(); g.ldb = cols; g.strideB = geom.C[a] * cols;
        g.C = pl->ws[cur ^ 1].d(); g.ldc = cols; g.strideC = geom.G[a] * cols;
        if (mask >> a & 1) {
            const Parity1D q = parity_dims(A);
            g.M = q.nh; g.K = q.Pe; g.batch_lo = 2;
            g.A = A.cPhi.d(); g.lda = q.Pep; g.strideA_lo = (long long)q.nh * q.Pep;
            g.strideB_lo = (long long)q.Pe * cols;
            g.strideC_lo = (long long)q.nh * cols;
        } else {
            g.M = A.n; g.K = A.P;
            g.A = A.plain.d(); g.lda = A.Pp;
        }
        HB_TRY(timed_gemm(pl, g, 11 + a, st));
        cols *= geom.G[a];
        cur ^= 1;
    }
    return launch_nd_parity_merge(parity_pass_args(pl, mask, geom, f_stride, batch), pl->ws[cur].d(), d_f, st);
}

// ---------------------------------------------------------------------------------------------
// Forward: c[j] = sum_i f[i] prod_d w_d[i_d] phi_{m_j,d}(x_{i_d})   (transform.cpp:115-139), contracted
// one axis at a time, last axis first.
int plan_forward(hb_plan* pl, const double* d_f, long long f_stride, int batch, double* d_coef,
                 long long coef_stride, cudaStream_t st) {
    if (batch <= 0) return HB_OK;
    const int d = pl->dim;
    if (d == 1 && parity_stage(pl->ax[0], 2.0 * batch * pl->ax[0].n * pl->ax[0].P)) {
        pl->n_timed = 0;
        return forward_last_axis_parity(pl, pl->ax[0], d_f, f_stride, batch, d_coef, coef_stride, 1, st);
    }
    if (d == 1) {
        const Axis& A = pl->ax[0];
        GemmArgs g;
        g.M = batch; g.N = A.P; g.K = A.n;
        g.A = d_f; g.lda = f_stride;
        g.B = A.wgt.d(); g.ldb = A.Pp;
        g.C = d_coef; g.ldc = coef_stride;
        g.epi = EPI_STORE;  // every 1-D index set enumerates 0..P-1 in natural order
        pl->n_timed = 0;
        return timed_gemm(pl, g, 1, st);
    }
    if (const int mask = parity_mask(pl, batch)) return forward_parity(pl, mask, d_f, f_stride, batch, d_coef, coef_stride, st);
    Axis& L = pl->ax[d - 1];
    const long long rows1 = pl->grid_rows;
    const bool flat = (f_stride == rows1 * L.np || batch == 1);
    const int* lut = pl->lut_box.i();
    long long cols = L.Pp;
    // workspace: largest intermediate
    size_t need = (size_t)batch * rows1 * cols;
    {
        long long c = cols, prefix = (long long)batch * rows1;
        for (int a = d - 2; a >= 1; a--) {
            prefix /= pl->ax[a].n;
            c *= pl->ax[a].P;
            need = std::max(need, (size_t)(prefix * c));
        }
    }
    HB_TRY(pl->ws[0].ensure(need * 8));
    if (d > 2) HB_TRY(pl->ws[1].ensure(need * 8));
    pl->n_timed = 0;
    {
        GemmArgs g;
        g.K = L.n; g.N = L.Pp;
        g.A = d_f; g.lda = L.np;
        if (flat) {
            g.M = (int)(batch * rows1); g.batch = 1;
        } else {
            g.M = (int)rows1; g.batch = batch; g.strideA = f_stride; g.strideC = rows1 * cols;
        }
        g.B = L.wgt.d(); g.ldb = L.Pp; g.strideB = 0;
        g.C = pl->ws[0].d(); g.ldc = cols;
        HB_TRY(timed_gemm(pl, g, d, st));
    }
    int cur = 0;
    long long prefix = (long long)batch * rows1;
    for (int a = d - 2; a >= 0; a--) {
        Axis& A = pl->ax[a];
        prefix /= A.n;
        GemmArgs g;
        g.M = A.P; g.K = A.n; g.N = (int)cols; g.batch = (int)prefix;
        g.A = A.wgtT.d(); g.lda = A.np; g.strideA = 0;
        g.B = pl->ws[cur].d(); g.ldb = cols; g.strideB = (long long)A.n * cols;
        if (a == 0) {
            g.C = d_coef; g.strideC = coef_stride; g.epi = EPI_LUT; g.lut = lut;
        } else {
            g.C = pl->ws[cur ^ 1].d(); g.ldc = cols; g.strideC = (long long)A.P * cols;
        }
        HB_TRY(timed_gemm(pl, g, 1 + a, st));
        cols *= A.P;
        cur ^= 1;
    }
    return HB_OK;
}

// Backward: f[i] = sum_j c[j] prod_d phi_{m_j,d}(x_{i_d})   (transform.cpp:121-128,143; weight forced to 1).
int plan_backward(hb_plan* pl, const double* d_coef, long long coef_stride, int batch, double* d_f,
                  long long f_stride, cudaStream_t st) {
    if (batch <= 0) return HB_OK;
    const int d = pl->dim;
    if (d == 1 && parity_stage(pl->ax[0], 2.0 * batch * pl->ax[0].n * pl->ax[0].P)) {
        pl->n_timed = 0;
        return backward_last_axis_parity(pl, pl->ax[0], d_coef, coef_stride, batch, d_f, f_stride, 11, st);
    }
    if (d == 1) {
        const Axis& A = pl->ax[0];
        GemmArgs g;
        g.M = batch; g.N = A.n; g.K = A.P;
        g.A = d_coef; g.lda = coef_stride;
        g.B = A.plainT.d(); g.ldb = A.np;
        g.C = d_f; g.ldc = f_stride;
        pl->n_timed = 0;
        return timed_gemm(pl, g, 11, st);
    }
    if (const int mask = parity_mask(pl, batch)) return backward_parity(pl, mask, d_coef, coef_stride, batch, d_f, f_stride, st);
    Axis& L = pl->ax[d - 1];
    const long long box_pitch = L.Pp, box_size = pl->box_size;
    size_t need = (size_t)batch * box_size;
    {
        long long c = L.np, prefix = (long long)batch * pl->box_rows;
        need = std::max(need, (size_t)(prefix * c));
        for (int a = d - 2; a >= 1; a--) {
            prefix /= pl->ax[a].P;
            c *= pl->ax[a].n;
            need = std::max(need, (size_t)(prefix * c));
        }
    }
    HB_TRY(pl->ws[0].ensure(need * 8));
    HB_TRY(pl->ws[1].ensure(need * 8));
    // coefficients in set order -> zero-padded lexicographic box
    HB_TRY(launch_gather_lut(d_coef, coef_stride, pl->lut_box.i(), pl->ws[0].d(), box_size, box_size, batch, st));
    long long cols = L.np;
    pl->n_timed = 0;
    {
        GemmArgs g;
        g.M = (int)(batch * pl->box_rows); g.K = L.P; g.N = L.np;
        g.A = pl->ws[0].d(); g.lda = box_pitch;
        g.B = L.plainT.d(); g.ldb = L.np;
        g.C = pl->ws[1].d(); g.ldc = cols;
        HB_TRY(timed_gemm(pl, g, 10 + d, st));
    }
    int cur = 1;
    long long prefix = (long long)batch * pl->box_rows;
    for (int a = d - 2; a >= 0; a--) {
        Axis& A = pl->ax[a];
        prefix /= A.P;
        GemmArgs g;
        g.M = A.n; g.K = A.P; g.N = (int)cols; g.batch = (int)prefix;
        g.A = A.plain.d(); g.lda = A.Pp; g.strideA = 0;
        g.B = pl->ws[cur].d(); g.ldb = cols; g.strideB = (long long)A.P * cols;
        if (a == 0) {
            g.C = d_f; g.ldc = cols; g.strideC = f_stride;
        } else {
            g.C = pl->ws[cur ^ 1].d(); g.ldc = cols; g.strideC = (long long)A.n * cols;
        }
        HB_TRY(timed_gemm(pl, g, 11 + a, st));
        cols *= A.n;
        cur ^= 1;
    }
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// Fourier triple products (reference: triple_products_fourier, varf.cpp:99-162), from the complex
// exponential expansion of the basis e_0 = 1/sqrt(2pi), e_{2k} = cos(kx)/sqrt(pi), e_{2k-1} = sin(kx)/sqrt(pi):
// int e_a e_b e_c over [-pi,pi] = 2pi * sum of coefficient products whose frequencies add up to 0.
void fourier_triple_products_host(int degree, int Pp, std::vector<double>& out) {
    const int K = 2 * degree + 1, P = degree + 1;
    out.assign((size_t)K * P * Pp, 0.0);
    struct Term { int freq; double re, im; };
    auto expand = [](int idx, Term t[2]) -> int {
        if (idx == 0) { t[0] = {0, 1.0 / std::sqrt(2 * M_PI), 0.0}; return 1; }
        const double s = 1.0 / std::sqrt(M_PI);
        if (idx % 2 == 0) {  // cos(kx) = (e^{ikx} + e^{-ikx})/2
            const int k = idx / 2;
            t[0] = {k, 0.5 * s, 0.0}; t[1] = {-k, 0.5 * s, 0.0};
        } else {             // sin(kx) = (e^{ikx} - e^{-ikx})/(2i)
            const int k = (idx + 1) / 2;
            t[0] = {k, 0.0, -0.5 * s}; t[1] = {-k, 0.0, 0.5 * s};
        }
        return 2;
    };
    for (int a = 0; a < K; a++)
        for (int b = 0; b < P; b++)
            for (int c = 0; c < P; c++) {
                Term ta[2], tb[2], tc[2];
                const int na = expand(a, ta), nb = expand(b, tb), nc = expand(c, tc);
                double acc = 0.0;
                for (int i = 0; i < na; i++)
                    for (int j = 0; j < nb; j++)
                        for (int k = 0; k < nc; k++) {
                            if (ta[i].freq + tb[j].freq + tc[k].freq != 0) continue;
                            // real part of the product of three complex coefficients
                            const double r1 = ta[i].re * tb[j].re - ta[i].im * tb[j].im;
                            const double i1 = ta[i].re * tb[j].im + ta[i].im * tb[j].re;
                            acc += r1 * tc[k].re - i1 * tc[k].im;
                        }
                out[(size_t)a * P * Pp + (size_t)b * Pp + c] = 2 * M_PI * acc;
            }
}

int ensure_triple_products(hb_plan* pl, cudaStream_t st) {
    const int N = pl->degree;
    if (pl->Tdeg == N) return HB_OK;
    pl->TPp = (int)even_up(N + 1);
    const size_t bytes = (size_t)(2 * N + 1) * (N + 1) * pl->TPp * 8;
    bool any_h = false, any_f = false;
    for (const Axis& A : pl->ax) (A.fourier ? any_f : any_h) = true;
    if (any_h) {
        HB_TRY(pl->T.ensure(bytes));
        HB_TRY(launch_triple_products(N, pl->TPp, pl->T.d(), st));
    }
    if (any_f) {
        std::vector<double> h;
        fourier_triple_products_host(N, pl->TPp, h);
        HB_TRY(pl->Tf.ensure(bytes));
        HB_CUDA(cudaMemcpyAsync(pl->Tf.p, h.data(), bytes, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaStreamSynchronize(st));
    }
    pl->Tdeg = N;
    return HB_OK;
}

static int ensure_varf_tables(hb_plan* pl, cudaStream_t st) {
    if (!pl->midx.p) {
        std::vector<int> m(pl->iset->list.begin(), pl->iset->list.end());
        HB_TRY(pl->midx.ensure(m.size() * 4));
        HB_CUDA(cudaMemcpyAsync(pl->midx.p, m.data(), m.size() * 4, cudaMemcpyHostToDevice, st));
        if (pl->dim == 2) {
            // Per-block position tables of the 2-D index set for the scatter epilogue of the dense varf
            // (see GemmParams): positions, the order of the 64 entries of a block by position, and the range.
            const int P1 = pl->ax[1].P, nb0 = pl->ax[0].nb, nb1 = pl->ax[1].nb;
            for (int a = 0; a < 2; a++) {
                Axis& A = pl->ax[a];
                HB_TRY(A.d_slots.ensure(A.slots.size() * 4));
                HB_CUDA(cudaMemcpyAsync(A.d_slots.p, A.slots.data(), A.slots.size() * 4, cudaMemcpyHostToDevice, st));
            }
            std::vector<int> pos((size_t)nb0 * nb1 * 64, -1);
            std::vector<unsigned char> perm((size_t)nb0 * nb1 * 64);
            std::vector<int> range((size_t)nb0 * nb1 * 2);
            for (int ba = 0; ba < nb0; ba++)
                for (int bb = 0; bb < nb1; bb++) {
                    const size_t base = ((size_t)ba * nb1 + bb) * 64;
                    int lo = 0x7fffffff, hi = -1;
                    for (int e = 0; e < 64; e++) {
                        const int a = pl->ax[0].slots[8 * ba + (e >> 3)], b = pl->ax[1].slots[8 * bb + (e & 7)];
                        const int v = (a >= 0 && b >= 0) ? pl->iset->lex_to_pos[(size_t)a * P1 + b] : -1;
                        pos[base + e] = v;
                        if (v >= 0) { lo = std::min(lo, v); hi = std::max(hi, v); }
                        perm[base + e] = (unsigned char)e;
                    }
                    // invalid entries (-1) sort last: compare as unsigned
                    std::sort(perm.begin() + base, perm.begin() + base + 64, [&](unsigned char x, unsigned char y) {
                        return (unsigned)pos[base + x] < (unsigned)pos[base + y];
                    });
                    range[((size_t)ba * nb1 + bb) * 2] = hi >= 0 ? lo : 0;
                    range[((size_t)ba * nb1 + bb) * 2 + 1] = hi;
                }
            // joint order of the 128 column entries of a column tile (two adjacent column blocks) for every bn0:
            // entry = cbl * 64 + slot, sorted by output position, so that runs continue across the two blocks
            // (16-entry runs where the enumeration is consecutive along n1)
            const int tiles_n = (nb1 * nb1 + 1) / 2;
            std::vector<unsigned char> perm2((size_t)nb0 * tiles_n * 128);
            for (int bn0 = 0; bn0 < nb0; bn0++)
                for (int tn = 0; tn < tiles_n; tn++) {
                    int key[128];
                    unsigned char* dst = &perm2[((size_t)bn0 * tiles_n + tn) * 128];
                    for (int e = 0; e < 128; e++) {
                        const int cb = 2 * tn + (e >> 6);
                        int v = -1;
                        if (cb < nb1 * nb1) {
                            const int bn1 = cb % nb1;
                            v = pos[((size_t)bn0 * nb1 + bn1) * 64 + (e & 63)];
                        }
                        key[e] = v;
                        dst[e] = (unsigned char)e;
                    }
                    std::sort(dst, dst + 128, [&](unsigned char x, unsigned char y) { return (unsigned)key[x] < (unsigned)key[y]; });
                }
            HB_TRY(pl->blkperm2.ensure(perm2.size()));
            HB_CUDA(cudaMemcpyAsync(pl->blkperm2.p, perm2.data(), perm2.size(), cudaMemcpyHostToDevice, st));
            HB_TRY(pl->blkpos.ensure(pos.size() * 4));
            HB_TRY(pl->blkperm.ensure(perm.size()));
            HB_TRY(pl->blkrange.ensure(range.size() * 4));
            HB_CUDA(cudaMemcpyAsync(pl->blkpos.p, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaMemcpyAsync(pl->blkperm.p, perm.data(), perm.size(), cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaMemcpyAsync(pl->blkrange.p, range.data(), range.size() * 4, cudaMemcpyHostToDevice, st));
            pl->h_blkrange = range;
        }
        HB_CUDA(cudaStreamSynchronize(st));
    }
    return HB_OK;
}

// two-stage dense contraction for dim == 2 (see DESIGN.md "varf"), on blocked pair tables:
//   G[k0][q] = sum_{k1} coef[k0][k1] * R[k1][q]                       (stage 1, A K-contiguous)
//   A[idx(m0,m1)][idx(n0,n1)] = sum_{k0} L[k0][p] * G[k0][q]           (stage 2, A M-contiguous, scatter epilogue)
// with p = blocked pair (m0,n0) of axis 0 and q = blocked pair (m1,n1) of axis 1.  When the whole matrix
// is requested only the blocks with m0 <= n0 are computed and mirrored (A is symmetric, bit for bit:
// the tables are symmetric in (i,j) and every dot product runs over k in the same order).

// `parity_seg` > 0: the contraction over k0 is halved by parity (symmetric nodes, Gram form): `coef` holds the
// even part of f in rows [0, K0) and the odd part in rows [parity_seg, parity_seg + K0); a row tile whose block
// pair has even m0 + n0 reads the first segment of G, an odd one the second.
static int varf_dense_2d(hb_plan* pl, const double* coef, long long ld_coef, int K0, int K1, const double* Ltab,
                         const double* Rtab, double* d_out, long long ldo, long long row_begin, long long row_end,
                         int parity_seg, cudaStream_t st) {
    const int P0 = pl->ax[0].P, P1 = pl->ax[1].P;
    const int nb0 = pl->ax[0].nb, nb1 = pl->ax[1].nb;
    const int rows1 = parity_seg > 0 ? parity_seg + K0 : K0;   // rows of the stage-1 product
    const long long nrow = (long long)nb0 * nb0 * 64, ncol = (long long)nb1 * nb1 * 64;
    HB_TRY(pl->G.ensure((size_t)rows1 * ncol * 8));
    GemmArgs g1;
    g1.M = rows1; g1.K = K1; g1.N = (int)ncol;
    g1.A = coef; g1.lda = ld_coef;
    g1.B = Rtab; g1.ldb = ncol;
    g1.C = pl->G.d(); g1.ldc = ncol;
    HB_TRY(timed_gemm(pl, g1, 21, st));
    GemmArgs g2;
    g2.M = (int)nrow; g2.K = K0; g2.N = (int)ncol;
    g2.A = Ltab; g2.lda = nrow; g2.a_mmajor = true;
    g2.B = pl->G.d(); g2.ldb = ncol;
    g2.C = d_out; g2.ldc = ldo;
    g2.epi = EPI_VARF2D;
    g2.blkpos = pl->blkpos.i();
    g2.blkperm = static_cast<const unsigned char*>(pl->blkperm.p);
    g2.blkperm2 = static_cast<const unsigned char*>(pl->blkperm2.p);
    g2.blkrange = static_cast<const int2*>(pl->blkrange.p);
    g2.P0 = P0; g2.Pp0 = nb0; g2.P1 = P1; g2.Pp1 = nb1;
    g2.b_koff_odd = parity_seg; g2.nb_even0 = pl->ax[0].nb_even;
    g2.group_m = pl->varf_natural_axis1 ? 3 : 2;   // base-block squares of the tile order (varf_tile_list; measured)
    const bool full = (row_begin == 0 && row_end == pl->n_polys);
    g2.row_begin = (int)row_begin; g2.row_end = full ? 0x7fffffff : (int)row_end;
    g2.symmetric = full ? 1 : 0;
    // persistent ping-pong kernel over the list of needed tiles (cached per row range)
    const long long key[3] = {row_begin, row_end, g2.symmetric};
    if (key[0] != pl->tiles_key[0] || key[1] != pl->tiles_key[1] || key[2] != pl->tiles_key[2]) {
        std::vector<int> list;
        HB_TRY(varf_tile_list(g2, pl->h_blkrange, list));
        pl->n_tiles = (int)(list.size() / 2);
        HB_TRY(pl->tiles.ensure((list.size() + 2) * 4));   // + one int2 slot: the tile counter of the kernel
        if (!list.empty()) {
            HB_CUDA(cudaMemcpyAsync(pl->tiles.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaStreamSynchronize(st));   // `list` goes out of scope
        }
        pl->tiles_key[0] = key[0]; pl->tiles_key[1] = key[1]; pl->tiles_key[2] = key[2];
    }
    if (!pl->timing) return varf_stage2_launch(g2, static_cast<const int2*>(pl->tiles.p), pl->n_tiles, st);
    if (pl->n_timed >= (int)pl->timed.size()) {
        hb_plan::Timed t;
        HB_CUDA(cudaEventCreate(&t.e0));
        HB_CUDA(cudaEventCreate(&t.e1));
        pl->timed.push_back(t);
    }
    hb_plan::Timed& t = pl->timed[pl->n_timed++];
    t.tag = 22;
    t.flops = 2.0 * g2.M * (double)g2.N * g2.K;
    HB_CUDA(cudaEventRecord(t.e0, st));
    const int rc = varf_stage2_launch(g2, static_cast<const int2*>(pl->tiles.p), pl->n_tiles, st);
    HB_CUDA(cudaEventRecord(t.e1, st));
    return rc;
}

// Dense varf in d >= 3 dimensions: the contraction  A = sum_k in[k_0..k_{d-1}] prod_a tab_a[k_a][(m_a, n_a)]  taken one
// axis at a time, last axis first -- d GEMM launches on the FP64 tensor cores -- into the full pair-space array
// full[(m_0,n_0)]..[(m_{d-1},n_{d-1})], from which the rows of the matrix are gathered.  `in` is the grid function
// (Gram form: k = nodes, tables psi_m psi_n) or the box of retained transform coefficients (reference algebra,
// varf.cpp:236-262: k = alpha, tables T[alpha][m][n]); tab[a] is [K_a][C_a] with C_a = P_a * Pp_a.  The pair space
// holds prod P_a^2 cells -- n_polys^2 for the cube, (d!)^2 times that for a triangle -- so this path is for the
// moderate degrees d >= 3 is used at (checked against the memory that is free).
static int varf_dense_nd(hb_plan* pl, const double* in, long long in_ld, const int* K, const double* const* tab,
                         const long long* tab_ld, const int* Pp, double* d_out, long long ldo, long long row_begin, long long row_end,
                         cudaStream_t st) {
    const int d = pl->dim, L = d - 1;
    long long C[HB_MAX_DIM] = {0}, in_rows = 1;
    for (int a = 0; a < d; a++) C[a] = (long long)pl->ax[a].P * Pp[a];
    for (int a = 0; a < L; a++) in_rows *= K[a];
    // largest intermediate: [K_0..K_{a-1}][C_a..C_{d-1}]
    long long need = 0, sz = in_rows * C[L];
    need = sz;
    for (int a = L - 1; a >= 0; a--) { sz = sz / K[a] * C[a]; need = std::max(need, sz); }
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t have = pl->ws[0].bytes + pl->ws[1].bytes;
    if ((double)need * 16.0 > (double)free_b + (double)have - 1e9) {
        set_error("varf (dim %d, degree %d): the pair space of the dense assembly needs 2 x %.1f GB", d, pl->degree, need * 8e-9);
        return HB_ERR_NOMEM;
    }
    HB_TRY(pl->ws[0].ensure((size_t)need * 8));
    HB_TRY(pl->ws[1].ensure((size_t)need * 8));
    {
        GemmArgs g;
        g.M = (int)in_rows; g.K = K[L]; g.N = (int)C[L];
        g.A = in; g.lda = in_ld;
        g.B = tab[L]; g.ldb = tab_ld[L];
        g.C = pl->ws[0].d(); g.ldc = C[L];
        HB_TRY(timed_gemm(pl, g, 21, st));
    }
    int cur = 0;
    long long prefix = in_rows, cols = C[L];
    for (int a = L - 1; a >= 0; a--) {
        prefix /= K[a];
        GemmArgs g;
        g.M = (int)C[a]; g.K = K[a]; g.N = (int)cols; g.batch = (int)prefix;
        g.A = tab[a]; g.lda = tab_ld[a]; g.a_mmajor = true; g.strideA = 0;
        g.B = pl->ws[cur].d(); g.ldb = cols; g.strideB = (long long)K[a] * cols;
        g.C = pl->ws[cur ^ 1].d(); g.ldc = cols; g.strideC = C[a] * cols;
        HB_TRY(timed_gemm(pl, g, 22, st));
        cols *= C[a];
        cur ^= 1;
    }
    long long stride[HB_MAX_DIM];
    stride[L] = 1;
    for (int a = L - 1; a >= 0; a--) stride[a] = stride[a + 1] * C[a + 1];
    return launch_varf_gather_nd(d, (int)pl->n_polys, (int)row_begin, (int)row_end, Pp, stride, pl->midx.i(),
                                 pl->ws[cur].d(), d_out, ldo, st);
}

static int varf_gram(hb_plan* pl, const double* d_f, double* d_out, long long ldo, long long row_begin,
                     long long row_end, cudaStream_t st) {
    if (pl->dim > 2) {
        // psi_m psi_n tables per axis (psi = sqrt(w) phi: overflow-free), then d GEMM stages + gather
        const int d = pl->dim;
        int K[HB_MAX_DIM], Pp[HB_MAX_DIM];
        long long tab_ld[HB_MAX_DIM];
        const double* tab[HB_MAX_DIM];
        if (pl->pairN.size() != (size_t)d) pl->pairN.resize(d);
        for (int a = 0; a < d; a++) {
            Axis& A = pl->ax[a];
            if (!A.psi.p) {
                HB_TRY(A.psi.ensure_zero((size_t)A.n * A.Pp * 8, st));
                HB_TRY(launch_psi_table(A.x.d(), A.w.d(), A.sw.p ? A.sw.d() : nullptr, A.n, A.P, A.fourier, A.psi.d(), A.Pp, st));
            }
            if (!pl->pairN[a].p) {
                HB_TRY(pl->pairN[a].ensure((size_t)A.n * A.P * A.Pp * 8));
                HB_TRY(launch_pair_table(A.psi.d(), A.Pp, nullptr, A.n, A.P, A.Pp, pl->pairN[a].d(), st));
            }
            K[a] = A.n; Pp[a] = A.Pp; tab[a] = pl->pairN[a].d(); tab_ld[a] = (long long)A.P * A.Pp;
        }
        return varf_dense_nd(pl, d_f, pl->ax[d - 1].np, K, tab, tab_ld, Pp, d_out, ldo, row_begin, row_end, st);
    }
    // pair products are formed from psi = sqrt(w) phi (overflow-free), not from w * phi * phi
    for (Axis& A : pl->ax) {
        if (A.psi.p) continue;
        HB_TRY(A.psi.ensure_zero((size_t)A.n * A.Pp * 8, st));
        HB_TRY(launch_psi_table(A.x.d(), A.w.d(), A.sw.p ? A.sw.d() : nullptr, A.n, A.P, A.fourier, A.psi.d(), A.Pp, st));
    }
    if (pl->dim == 1) {
        const Axis& A = pl->ax[0];
        const size_t bytes = (size_t)A.n * A.P * A.Pp * 8;
        if (pl->pair[0].bytes < bytes) {
            HB_TRY(pl->pair[0].ensure(bytes));
            HB_TRY(launch_pair_table(A.psi.d(), A.Pp, nullptr, A.n, A.P, A.Pp, pl->pair[0].d(), st));
        }
        const long long ncol = (long long)A.P * A.Pp;
        HB_TRY(pl->tmp.ensure(ncol * 8));
        GemmArgs g;
        g.M = 1; g.K = A.n; g.N = (int)ncol;
        g.A = d_f; g.lda = A.np;
        g.B = pl->pair[0].d(); g.ldb = ncol;
        g.C = pl->tmp.d(); g.ldc = ncol;
        HB_TRY(timed_gemm(pl, g, 21, st));
        const long long r0 = std::max(0LL, row_begin), r1 = std::min((long long)A.P, row_end);
        if (r1 > r0)
            HB_CUDA(cudaMemcpy2DAsync(d_out, ldo * 8, pl->tmp.d() + r0 * A.Pp, (size_t)A.Pp * 8, (size_t)A.P * 8,
                                      r1 - r0, cudaMemcpyDeviceToDevice, st));
        return HB_OK;
    }
    // Parity: psi_k(-x) = (-1)^k psi_k(x) holds bit for bit when the nodes and weights of axis 0 are symmetric,
    // so L[-x][(m0,n0)] = (-1)^(m0+n0) L[x][(m0,n0)] and the contraction over x0 only needs the non-negative
    // nodes, against the even or the odd part of f (chosen per row tile, whose 8x8 block has one parity):
    // half the flops of the dominant stage.
    static const bool parity_ok = [] { const char* e = getenv("HB_VARF_PARITY"); return !(e && e[0] == '0'); }();
    Axis &A0 = pl->ax[0], &A1 = pl->ax[1];
    const bool parity = parity_ok && A0.symmetric_nodes && A0.n >= 2;
    const int nh = (A0.n + 1) / 2;                       // nodes x >= 0: rows n - nh .. n - 1
    for (int a = 0; a < 2; a++) {
        if (pl->pairB_ok[a]) continue;
        Axis& A = pl->ax[a];
        const bool half = (a == 0 && parity);
        const int rows = half ? nh : A.n, first = half ? A.n - nh : 0;
        HB_TRY(pl->pairB[a].ensure((size_t)rows * A.nb * A.nb * 64 * 8));
        HB_TRY(launch_blocked_pair_table(0, A.psi.d() + (size_t)first * A.Pp, A.Pp, 0, nullptr, rows, A.d_slots.i(),
                                         A.nb, pl->pairB[a].d(), st));
        pl->pairB_ok[a] = true;
    }
    if (!parity)
        return varf_dense_2d(pl, d_f, A1.np, A0.n, A1.n, pl->pairB[0].d(), pl->pairB[1].d(), d_out, ldo, row_begin,
                             row_end, 0, st);
    const int seg = (nh + 15) & ~15;                     // the odd segment starts on a k-tile boundary
    HB_TRY(pl->fsplit.ensure_zero((size_t)(seg + nh) * A1.np * 8, st));
    HB_TRY(launch_parity_split(d_f, A0.n, nh, seg, A1.np, pl->fsplit.d(), st));
    return varf_dense_2d(pl, pl->fsplit.d(), A1.np, nh, A1.n, pl->pairB[0].d(), pl->pairB[1].d(), d_out, ldo, row_begin,
                         row_end, seg, st);
}

// The reference's algebra (varf.cpp:164-269).
// `csr` != nullptr: sparse output.  When the retained alpha leave a narrow band the band is emitted straight as CSR
// (csr->rows = number of owned rows) and d_out is not touched; otherwise csr->rows = -1 and nothing has been
// assembled (the caller falls back to a dense assembly + compaction).
static int varf_reference(hb_plan* pl, const double* d_f, double* d_out, long long ldo, long long row_begin,
                          long long row_end, cudaStream_t st, CsrResult* csr = nullptr) {
    const int d = pl->dim, N = pl->degree;
    HB_TRY(ensure_triple_products(pl, st));
    // Hf = transform at degree 2N on the same kind of index set (varf.cpp:181)
    if (!pl->plan2) {
        std::vector<int> n_pts(d), four(d);
        std::vector<double> nodes, weights;
        for (int a = 0; a < d; a++) {
            n_pts[a] = pl->ax[a].n;
            four[a] = pl->ax[a].fourier;
            nodes.insert(nodes.end(), pl->ax[a].h_x.begin(), pl->ax[a].h_x.end());
            weights.insert(weights.end(), pl->ax[a].h_w.begin(), pl->ax[a].h_w.end());
        }
        HB_TRY(plan_create(&pl->plan2, d, n_pts.data(), 2 * N, nodes.data(), weights.data(), four.data(), pl->set));
    }
    hb_plan* p2 = pl->plan2;
    const long long n2 = p2->n_polys;
    HB_TRY(pl->Hf.ensure(even_up(n2) * 8));
    // The banded assembly (polynomial-like f: the regime the reference algebra is meaningful in) is a zero-fill of
    // the output (HBM-write-bound, 1.8 ms for the 13 GB C2 cube matrix) plus a few entries per row.  What precedes
    // it -- transform of f at degree 2N, fetching Hf, the sequential norm and threshold on the host, ~0.6 ms --
    // does not touch the output, so when the previous call on this plan ended in the banded assembly the zero-fill
    // starts now on a side stream and the band kernel waits for it.  (A wrong guess only costs the wait.)
    const long long fill_rows = row_end - row_begin;
    static const bool early_fill_enabled = [] { const char* e = getenv("HB_VARF_EARLY_FILL"); return !(e && e[0] == '0'); }();
    const bool early_fill = early_fill_enabled && pl->last_path == 2 && fill_rows > 0 && csr == nullptr;
    if (early_fill) {
        if (!pl->fill_stream) {
            HB_CUDA(cudaStreamCreateWithFlags(&pl->fill_stream, cudaStreamNonBlocking));
            HB_CUDA(cudaEventCreateWithFlags(&pl->fill_ready, cudaEventDisableTiming));
            HB_CUDA(cudaEventCreateWithFlags(&pl->fill_done, cudaEventDisableTiming));
        }
    }
    HB_TRY(plan_forward(p2, d_f, pl->grid_size, 1, pl->Hf.d(), even_up(n2), st));
    if (pl->h_Hf_pinned_len < (size_t)n2) {
        if (pl->h_Hf_pinned) cudaFreeHost(pl->h_Hf_pinned);
        pl->h_Hf_pinned = nullptr;
        pl->h_Hf_pinned_len = 0;
        HB_CUDA(cudaMallocHost(&pl->h_Hf_pinned, (size_t)n2 * 8));
        pl->h_Hf_pinned_len = (size_t)n2;
    }
    // Hf goes to the host through a kernel that stores into the page-locked buffer (mapped under unified
    // addressing), not through cudaMemcpyAsync: the copy engine is what executes the zero-fill, and a D2H copy
    // queued behind it waited the whole 1.8 ms (measured: no gain from the early fill until this changed).
    if (early_fill) {
        HB_TRY(launch_repack_rows(pl->Hf.d(), n2, pl->h_Hf_pinned, n2, 1, (int)n2, st));
        // The fill is queued behind the (60 us) transform: launched first, its CTAs would hold every SM slot and the
        // transform kernels of `st` -- same priority -- would run after it.
        HB_CUDA(cudaEventRecord(pl->fill_ready, st));                 // earlier users of the output buffer
        HB_CUDA(cudaStreamWaitEvent(pl->fill_stream, pl->fill_ready, 0));
        HB_CUDA(cudaMemset2DAsync(d_out, (size_t)ldo * 8, 0, (size_t)pl->n_polys * 8, (size_t)fill_rows, pl->fill_stream));
        HB_CUDA(cudaEventRecord(pl->fill_done, pl->fill_stream));
    } else {
        HB_CUDA(cudaMemcpyAsync(pl->h_Hf_pinned, pl->Hf.p, n2 * 8, cudaMemcpyDeviceToHost, st));
    }
    HB_CUDA(cudaStreamSynchronize(st));
    if (early_fill) HB_CUDA(cudaStreamWaitEvent(st, pl->fill_done, 0));   // everything queued on st from here on
    // threshold exactly as the reference: sequential sum of squares, sqrt, max(norm,1) (varf.cpp:190-195),
    // keep alpha unless |Hf| < 1e-12*norm (:238-241; a NaN coefficient is therefore kept).
    const double* Hf = pl->h_Hf_pinned;
    double norm = 0;
    for (long long i = 0; i < n2; i++) norm += std::fabs(Hf[i]) * std::fabs(Hf[i]);
    norm = std::sqrt(norm);
    norm = norm < 1 ? 1 : norm;
    const double thr = 1e-12 * norm;
    std::vector<int> kept_alpha;
    std::vector<double> kept_val;
    for (long long i = 0; i < n2; i++) {
        if (std::fabs(Hf[i]) < thr) continue;
        for (int a = 0; a < d; a++) kept_alpha.push_back((int)p2->iset->list[(size_t)i * d + a]);
        kept_val.push_back(Hf[i]);
    }
    const int n_kept = (int)kept_val.size();
    pl->last_n_kept = n_kept;
    const bool dense = (d == 2 && n_kept > 96);
    if (csr) csr->rows = -1;
    if (csr && dense) return HB_OK;
    // d >= 3 with many alpha kept: the same contraction, one GEMM per axis on the triple-product tables, when its pair
    // space fits (the one-thread-per-entry kernel below walks every kept alpha for every entry)
    const char* nd_gemm = getenv("HB_VARF_ND_GEMM");      // developer switch (tests compare the two assemblies)
    if (d >= 3 && n_kept > 96 && !csr && !(nd_gemm && nd_gemm[0] == '0')) {
        bool hermite_only = true;
        double cells = 1;
        for (int a = 0; a < d; a++) {
            hermite_only = hermite_only && !pl->ax[a].fourier;
            cells *= (double)pl->ax[a].P * pl->TPp;
        }
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (hermite_only && cells * 16.0 < (double)free_b * 0.5) {
            std::vector<long long> ext(d);
            long long box = 1;
            for (int a = 0; a < d; a++) ext[a] = (long long)p2->iset->extent[a];
            const long long ld_last = even_up(ext[d - 1]);
            for (int a = 0; a + 1 < d; a++) box *= ext[a];
            std::vector<double> Hc((size_t)(box * ld_last), 0.0);
            for (int t = 0; t < n_kept; t++) {
                long long cell = 0;
                for (int a = 0; a < d; a++) cell = cell * (a == d - 1 ? ld_last : ext[a]) + kept_alpha[(size_t)t * d + a];
                Hc[(size_t)cell] = kept_val[t];
            }
            HB_TRY(pl->Hc.ensure(Hc.size() * 8));
            HB_CUDA(cudaMemcpyAsync(pl->Hc.p, Hc.data(), Hc.size() * 8, cudaMemcpyHostToDevice, st));
            HB_CUDA(cudaStreamSynchronize(st));
            int K[HB_MAX_DIM], Pp[HB_MAX_DIM];
            long long tab_ld[HB_MAX_DIM];
            const double* tab[HB_MAX_DIM];
            for (int a = 0; a < d; a++) {
                K[a] = (int)ext[a]; Pp[a] = pl->TPp; tab[a] = pl->T.d(); tab_ld[a] = (long long)(N + 1) * pl->TPp;
            }
            pl->last_path = 3;
            return varf_dense_nd(pl, pl->Hc.d(), ld_last, K, tab, tab_ld, Pp, d_out, ldo, row_begin, row_end, st);
        }
    }
    pl->last_path = dense ? 1 : 0;
    if (dense) {
        const int E0 = (int)p2->iset->extent[0], E1 = (int)p2->iset->extent[1];
        const long long ldc = even_up(E1);
        std::vector<double> Hc((size_t)E0 * ldc, 0.0);
        for (int t = 0; t < n_kept; t++) Hc[(size_t)kept_alpha[2 * t] * ldc + kept_alpha[2 * t + 1]] = kept_val[t];
        HB_TRY(pl->Hc.ensure(Hc.size() * 8));
        HB_CUDA(cudaMemcpyAsync(pl->Hc.p, Hc.data(), Hc.size() * 8, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaStreamSynchronize(st));
        const long long plane = (long long)(N + 1) * pl->TPp;
        for (int a = 0; a < 2; a++) {
            if (pl->TB_ok[a]) continue;
            if (a == 1 && pl->ax[1].fourier == pl->ax[0].fourier && pl->ax[1].P == pl->ax[0].P &&
                pl->ax[1].slots == pl->ax[0].slots) continue;  // same basis, same slot order: shares TB[0]
            const int nb = pl->ax[a].nb;
            const double* T = pl->ax[a].fourier ? pl->Tf.d() : pl->T.d();
            HB_TRY(pl->TB[a].ensure((size_t)(2 * N + 1) * nb * nb * 64 * 8));
            HB_TRY(launch_blocked_pair_table(1, T, pl->TPp, plane, nullptr, 2 * N + 1, pl->ax[a].d_slots.i(), nb,
                                             pl->TB[a].d(), st));
            pl->TB_ok[a] = true;
        }
        const double* T0 = pl->TB[0].d();
        const double* T1 = pl->TB_ok[1] ? pl->TB[1].d() : pl->TB[0].d();
        return varf_dense_2d(pl, pl->Hc.d(), ldc, E0, E1, T0, T1, d_out, ldo, row_begin, row_end, 0, st);
    }
    HB_TRY(pl->kept_alpha.ensure(std::max<size_t>(kept_alpha.size(), 1) * 4));
    HB_TRY(pl->kept_val.ensure(std::max<size_t>(kept_val.size(), 1) * 8));
    if (n_kept) {
        HB_CUDA(cudaMemcpyAsync(pl->kept_alpha.p, kept_alpha.data(), kept_alpha.size() * 4, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaMemcpyAsync(pl->kept_val.p, kept_val.data(), kept_val.size() * 8, cudaMemcpyHostToDevice, st));
        HB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
    }
    VarfSparseArgs a;
    a.dim = d; a.n_polys = (int)pl->n_polys; a.n_kept = n_kept;
    a.row_begin = (int)row_begin; a.row_end = (int)row_end; a.ldo = ldo;
    a.midx = pl->midx.i(); a.kept_alpha = pl->kept_alpha.i(); a.kept_val = pl->kept_val.d();
    for (int k = 0; k < d; k++) {
        a.T[k] = pl->ax[k].fourier ? pl->Tf.d() : pl->T.d();
        a.plane[k] = (long long)(N + 1) * pl->TPp;
        a.Pp[k] = pl->TPp;
    }
    // band width from the retained alpha; zero-fill + band when that is much smaller than a full row
    VarfBandArgs b;
    b.width_total = 1;
    for (int k = 0; k < d; k++) {
        // half-width of the band on axis k.  Hermite: T[a][m][n] = 0 unless |m - n| <= a.  Fourier (varf.cpp:99-162):
        // index a has wave number ceil(a/2) and e_a e_m e_n integrates to 0 unless the wave numbers satisfy
        // |k_m - k_n| <= k_a; with COS(k) = 2k, SIN(k) = 2k - 1 that is |m - n| <= 2 ceil(a/2) + 1 (a sine against
        // the cosine two wave numbers up sits a + 2 away for odd a).
        b.amax[k] = 0;
        for (int t = 0; t < n_kept; t++) {
            const int al = kept_alpha[(size_t)t * d + k];
            b.amax[k] = std::max(b.amax[k], pl->ax[k].fourier ? 2 * ((al + 1) / 2) + 1 : al);
        }
        b.extent[k] = pl->ax[k].P;
        b.width_total *= std::min(2 * b.amax[k] + 1, 2 * pl->ax[k].P - 1);
    }
    b.pitch_last = pl->ax[d - 1].Pp;
    b.lut_box = pl->lut_box.i();
    if (b.width_total * 4 <= pl->n_polys) {
        b.width_total = 1;
        for (int k = 0; k < d; k++) b.width_total *= 2 * b.amax[k] + 1;
        if (csr) return varf_band_csr(a, b, *csr, st);
        pl->last_path = 2;
        return launch_varf_band(a, b, d_out, st, /*zero_filled=*/early_fill);
    }
    if (csr) return HB_OK;   // no narrow band: dense assembly + compaction
    return launch_varf_sparse(a, d_out, st);
}

// Extended-range square roots of the weights of one axis for the Gram-form varf (see psi_table_kernel).  The tables
// built from the weights so far are dropped.
int plan_set_sqrt_weights(hb_plan* pl, int axis, const double* sqrt_w) {
    if (axis < 0 || axis >= pl->dim || !sqrt_w) return HB_ERR_INVALID;
    Axis& A = pl->ax[axis];
    for (int i = 0; i < A.n; i++) {
        if (!(sqrt_w[i] >= 0.0)) { set_error("sqrt_weights: entry %d is negative or NaN", i); return HB_ERR_INVALID; }
        if (A.symmetric_nodes && sqrt_w[i] != sqrt_w[A.n - 1 - i]) {
            set_error("sqrt_weights: the axis is bit-symmetric (x_i = -x_{n-1-i}, w_i = w_{n-1-i}); its square roots must be too");
            return HB_ERR_INVALID;
        }
    }
    HB_TRY(A.sw.ensure((size_t)A.n * 8));
    HB_CUDA(cudaMemcpy(A.sw.p, sqrt_w, (size_t)A.n * 8, cudaMemcpyHostToDevice));
    A.psi.release();
    pl->pairB_ok[0] = pl->pairB_ok[1] = false;
    pl->pair[0].release();
    pl->pair[1].release();
    pl->pairN.clear();
    return HB_OK;
}

int plan_varf_csr(hb_plan* pl, const double* d_f, long long row_begin, long long row_end, CsrResult& out, cudaStream_t st) {
    out.rows = -1;
    if (row_begin < 0) row_begin = 0;
    if (row_end > pl->n_polys) row_end = pl->n_polys;
    if (row_end <= row_begin) return HB_OK;
    HB_TRY(ensure_varf_tables(pl, st));
    pl->n_timed = 0;
    return varf_reference(pl, d_f, nullptr, pl->n_polys, row_begin, row_end, st, &out);
}

int plan_varf(hb_plan* pl, const double* d_f, int mode, double* d_out, long long ldo, long long row_begin,
              long long row_end, cudaStream_t st) {
    if (row_begin < 0) row_begin = 0;
    if (row_end > pl->n_polys) row_end = pl->n_polys;
    if (row_end <= row_begin) return HB_OK;
    if (ldo < pl->n_polys) return HB_ERR_INVALID;
    HB_TRY(ensure_varf_tables(pl, st));
    pl->n_timed = 0;
    if (mode == HB_VARF_GRAM) return varf_gram(pl, d_f, d_out, ldo, row_begin, row_end, st);
    if (mode == HB_VARF_REFERENCE) return varf_reference(pl, d_f, d_out, ldo, row_begin, row_end, st);
    set_error("unknown varf mode %d", mode);
    return HB_ERR_INVALID;
}

}  // namespace hb
