// C ABI of libhermite_b200.so (declared in include/hermite_b200.h).  Host-pointer entry points mirror
// the functions of the reference's `hermite_cpp` module (cpp/src/hermite/python/module.cpp:85-167):
// they look up (or build) a cached plan, stage the operands to the device, run the CUDA path and copy
// the result back.  There is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <mutex>
#include <chrono>
#include <vector>

#include "hb_kernels.h"
#include "hb_plan.h"

namespace hb {
const char* last_error();
long long launch_count(int reset);

namespace {

uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

struct PlanKey {
    uint64_t hash;
    int dim, degree, set;
    std::vector<int> n_pts, fourier;
    std::vector<double> nodes, weights;
    bool operator==(const PlanKey& o) const {
        return hash == o.hash && dim == o.dim && degree == o.degree && set == o.set && n_pts == o.n_pts &&
               fourier == o.fourier && nodes == o.nodes && weights == o.weights;
    }
};

struct CacheEntry { PlanKey key; hb_plan* plan; };
std::mutex g_plan_mutex;
std::list<CacheEntry> g_plans;  // most recently used first
const size_t kMaxPlans = 12;

int check_device() {
    static int state = -1;
    if (state < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n < 1) {
            cudaGetLastError();
            set_error("no usable CUDA device (%s); libhermite_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
            return HB_ERR_CUDA;
        }
        state = 1;
    }
    return HB_OK;
}

int cached_plan(hb_plan** out, int dim, const int* n_pts, int degree, const double* nodes, const double* weights,
                const int* do_fourier, const char* index_set) {
    HB_TRY(check_device());
    const int set = index_set_from_name(index_set);
    if (set < 0) { set_error("Invalid index set!"); return HB_ERR_INDEX_SET; }
    if (dim < 1 || dim > HB_MAX_DIM || !n_pts) return invalid(__func__);
    PlanKey key;
    key.dim = dim; key.degree = degree; key.set = set;
    key.n_pts.assign(n_pts, n_pts + dim);
    key.fourier.assign(dim, 0);
    if (do_fourier) key.fourier.assign(do_fourier, do_fourier + dim);
    size_t tot = 0;
    for (int a = 0; a < dim; a++) { if (n_pts[a] < 1) return invalid(__func__); tot += n_pts[a]; }
    key.nodes.assign(nodes, nodes + tot);
    key.weights.assign(weights, weights + tot);
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(key.n_pts.data(), dim * 4, h);
    h = fnv1a(key.fourier.data(), dim * 4, h);
    h = fnv1a(key.nodes.data(), tot * 8, h);
    h = fnv1a(key.weights.data(), tot * 8, h);
    h = fnv1a(&degree, 4, h);
    h = fnv1a(&set, 4, h);
    key.hash = h;
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    for (auto it = g_plans.begin(); it != g_plans.end(); ++it)
        if (it->key == key) {
            g_plans.splice(g_plans.begin(), g_plans, it);
            *out = g_plans.front().plan;
            return HB_OK;
        }
    hb_plan* pl = nullptr;
    HB_TRY(plan_create(&pl, dim, n_pts, degree, nodes, weights, key.fourier.data(), set));
    g_plans.push_front(CacheEntry{std::move(key), pl});
    while (g_plans.size() > kMaxPlans) {
        delete g_plans.back().plan;
        g_plans.pop_back();
    }
    *out = pl;
    return HB_OK;
}

// staging buffers shared by the host-pointer entry points (one call at a time per process)
std::mutex g_call_mutex;
DevBuf g_in, g_out, g_stage, g_stage_coef, g_aux[HB_MAX_FACTORS], g_auxi[HB_MAX_FACTORS];

// last sparse result of the process (the host-pointer calls are serialised by g_call_mutex; a caller that uses
// several threads pairs each *_sp call with its hb_sparse_fetch under its own lock, as hermipy_b200/core.py does)
CsrResult g_sp;
DevBuf g_sp_rownnz, g_sp_in[3 * HB_MAX_FACTORS], g_sp_lex;

int upload_grid(const hb_plan* pl, const double* f, int batch, double* dst) {
    // flat host->device copy, then re-pitch on the device (a pitched H2D copy of short rows is 2.4x slower)
    const Axis& L = pl->ax[pl->dim - 1];
    const size_t n_grid = (size_t)pl->grid_rows * L.n * batch;
    if (L.np == L.n) {
        HB_CUDA(cudaMemcpyAsync(dst, f, n_grid * 8, cudaMemcpyHostToDevice, 0));
        return HB_OK;
    }
    HB_TRY(g_stage.ensure(n_grid * 8));
    HB_CUDA(cudaMemcpyAsync(g_stage.p, f, n_grid * 8, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpy2DAsync(dst, (size_t)L.np * 8, g_stage.p, (size_t)L.n * 8, (size_t)L.n * 8,
                              (size_t)pl->grid_rows * batch, cudaMemcpyDeviceToDevice, 0));
    return HB_OK;
}

// body evaluated at translation + dilation @ node on the plan's grid, into the padded device layout
int plan_discretize(hb_plan* plan, const char* body, const double* translation, const double* dilation, double* d_f,
                    cudaStream_t st) {
    if (!plan->nodes_cat.p) {   // nodes of all axes, concatenated on the device (built on first use)
        size_t tot = 0;
        for (const Axis& A : plan->ax) tot += A.n;
        HB_TRY(plan->nodes_cat.ensure(tot * 8));
        size_t off = 0;
        for (const Axis& A : plan->ax) {
            HB_CUDA(cudaMemcpyAsync(plan->nodes_cat.d() + off, A.x.p, (size_t)A.n * 8, cudaMemcpyDeviceToDevice, st));
            off += A.n;
        }
    }
    std::vector<int> n_pts;
    for (const Axis& A : plan->ax) n_pts.push_back(A.n);
    return discretize_device(body, plan->dim, n_pts.data(), plan->nodes_cat.d(), translation, dilation, d_f,
                             plan->ax[plan->dim - 1].np, st);
}

int set_or_err(const char* index_set) {
    const int set = index_set_from_name(index_set);
    if (set < 0) set_error("Invalid index set!");
    return set;
}

// position LUT: for every multi-index of `full` restricted to `dirs` -> position in `sub` (or -1)
std::vector<int> restriction_map(const IndexSet& full, const std::vector<int>& dirs, const IndexSet& sub) {
    std::vector<int> map(full.size);
    std::vector<uint32_t> m(dirs.size());
    for (uint32_t i = 0; i < full.size; i++) {
        for (size_t j = 0; j < dirs.size(); j++) m[j] = full.list[(size_t)i * full.dim + dirs[j]];
        map[i] = sub.position(m.data());
    }
    return map;
}

}  // namespace
}  // namespace hb

using namespace hb;

extern "C" {

const char* hb_last_error(void) { return hb::last_error(); }
const char* hb_version(void) { return "hermite_b200 0.1 (sm_100a)"; }
long long hb_launch_count(int reset) { return hb::launch_count(reset); }
// developer hook (not in the public header): device buffer of 2 + 5*capacity int64 words, [1] = capacity
void hb_debug_trace(void* dev) { hb::gemm_set_trace(static_cast<long long*>(dev)); }

int hb_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); if (count) *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    if (count) *count = n;
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
int hb_index_size(const char* index_set, int dim, int degree, long long* size) {
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    const long long s = index_size(set, dim, degree);
    if (s < 0) { set_error("index_size: invalid (dim %d, degree %d)", dim, degree); return HB_ERR_INVALID; }
    *size = s;
    return HB_OK;
}

int hb_index_get_degree(const char* index_set, int dim, long long n_polys, int* degree) {
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (index_get_degree(set, dim, n_polys, degree)) {
        set_error("Can't find a degree with %lld polynomials in dimension %d", n_polys, dim);
        return HB_ERR_DEGREE;
    }
    return HB_OK;
}

int hb_index_list(const char* index_set, int dim, int degree, uint32_t* out, size_t cap) {
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    auto s = get_index_set(set, dim, degree);
    if (!s) { set_error("index_list: invalid (dim %d, degree %d)", dim, degree); return HB_ERR_INVALID; }
    if (cap < s->list.size()) return HB_ERR_NOMEM;
    memcpy(out, s->list.data(), s->list.size() * sizeof(uint32_t));
    return HB_OK;
}

int hb_triangle_index(const uint32_t* m, int dim, uint32_t* index) {
    if (!m || dim < 1) return invalid(__func__);
    *index = triangle_index(m, dim);
    return HB_OK;
}
int hb_cube_index(const uint32_t* m, int dim, uint32_t* index) {
    if (!m || dim < 1) return invalid(__func__);
    *index = cube_index(m, dim);
    return HB_OK;
}

int hb_index_positions(const char* index_set, int dim, int degree, const uint32_t* m, size_t count, int32_t* pos) {
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    auto s = get_index_set(set, dim, degree);
    if (!s || (count && (!m || !pos))) { set_error("index_positions: invalid (dim %d, degree %d)", dim, degree); return HB_ERR_INVALID; }
    for (size_t i = 0; i < count; i++) {
        bool inside = true;
        for (int a = 0; a < dim; a++) inside = inside && m[i * dim + a] < s->extent[a];
        pos[i] = inside ? s->position(m + i * dim) : -1;
    }
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
int hb_hermite_eval(const double* x, int n, int degree, double* out) {
    HB_TRY(check_device());
    if (n < 0 || degree < 0) return invalid(__func__);
    if (n == 0) return HB_OK;
    std::lock_guard<std::mutex> lock(g_call_mutex);
    const int P = degree + 1;
    HB_TRY(g_in.ensure((size_t)n * 8));
    HB_TRY(g_out.ensure((size_t)n * P * 8));
    HB_CUDA(cudaMemcpyAsync(g_in.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(launch_basis_tables(g_in.d(), nullptr, n, P, 0, g_out.d(), nullptr, P, nullptr, nullptr, 0, 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, (size_t)n * P * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

// Copy-in / compute / copy-out pipeline of the host-pointer batch transform: the batch is cut into chunks, and
// chunk i+1 is uploaded (H2D engine) while chunk i is transformed (SMs) and chunk i-1 is downloaded (D2H
// engine).  Three non-blocking streams chained by events; chunks run in order on the compute stream, so the
// plan's workspace is never used by two chunks at once.
//
// Issuing that pipeline costs ~100 CUDA API calls (per chunk: copies, events, tensor-map encodes, 3-6 launches),
// a few hundred microseconds of host time that is on the critical path of a 1 ms call.  A caller that keeps
// passing the same page-locked buffers -- a time-stepping loop -- gets the whole pipeline replayed from a CUDA
// graph instead: the first call with a given (plan, batch, direction, buffers) runs eagerly (and lets the GEMM
// autotuner time its candidates), the second is captured, later ones are one cudaGraphLaunch.
namespace hb {
namespace {
struct Pipeline {
    static const int kMaxChunks = 16;
    cudaStream_t in = nullptr, compute = nullptr, out = nullptr;
    cudaEvent_t uploaded[kMaxChunks] = {nullptr}, computed[kMaxChunks] = {nullptr};
    cudaEvent_t fork = nullptr, join_in = nullptr, join_out = nullptr;
    cudaEvent_t downloaded[kMaxChunks] = {nullptr}, start = nullptr;   // HB_PIPE_DEBUG=1: per-chunk timeline on stderr
    bool ok = false, debug = false;
    int last_chunks = 0;
    int init() {
        if (ok) return HB_OK;
        HB_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
        debug = getenv("HB_PIPE_DEBUG") && atoi(getenv("HB_PIPE_DEBUG")) != 0;
        const unsigned flags = debug ? cudaEventDefault : cudaEventDisableTiming;
        for (int i = 0; i < kMaxChunks; i++) {
            HB_CUDA(cudaEventCreateWithFlags(&uploaded[i], flags));
            HB_CUDA(cudaEventCreateWithFlags(&computed[i], flags));
            HB_CUDA(cudaEventCreateWithFlags(&downloaded[i], flags));
        }
        HB_CUDA(cudaEventCreateWithFlags(&start, flags));
        HB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        HB_CUDA(cudaEventCreateWithFlags(&join_in, cudaEventDisableTiming));
        HB_CUDA(cudaEventCreateWithFlags(&join_out, cudaEventDisableTiming));
        ok = true;
        return HB_OK;
    }
};
Pipeline g_pipe;

// one captured pipeline; valid while no device buffer it refers to has been freed (g_devbuf_generation)
struct PipeGraph {
    const hb_plan* plan = nullptr;
    int batch = 0, forward = 0;
    const void* src = nullptr;
    const void* dst = nullptr;
    unsigned long long generation = 0, stamp = 0;
    int calls = 0;               // calls seen with this key (0: slot free)
    bool refused = false;        // capture failed once: stay on the eager path
    long long launches = 0;      // kernels of this library inside the graph (hb_launch_count bookkeeping)
    cudaGraphExec_t exec = nullptr;
    void drop() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
    }
};
const int kMaxPipeGraphs = 8;
PipeGraph g_pipe_graphs[kMaxPipeGraphs];
unsigned long long g_pipe_stamp = 0;

bool page_locked(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// enqueue the whole pipeline on g_pipe's streams (eagerly, or into the capture that g_pipe.compute is in)
int issue_pipeline(hb_plan* pl, int dim, int batch, const double* f, int forward, double* out, bool capturing) {
    const Axis& L = pl->ax[dim - 1];
    const long long n_grid = pl->grid_rows * L.n;
    // Coefficient rows on the device: in d >= 2 dimensions they are only touched through the look-up table (scalar
    // scatter of the last forward stage, gather of the first inverse stage), so they can keep the host's row length
    // and cross PCIe without a re-pitch; in 1-D they are a GEMM operand and need an even pitch.
    const long long cstride = dim >= 2 ? pl->n_polys : even_up(pl->n_polys);
    const long long in_dev = forward ? pl->grid_size : cstride, out_dev = forward ? cstride : pl->grid_size;
    const bool pitched = L.np != L.n, cpitched = cstride != pl->n_polys;
    // Chunk sizes.  A chunk's transform takes about as long for 2 functions as for 17 at the C2 size (one wave of
    // GEMM tiles either way: ~60-90 us), so equal small chunks make the SMs, not PCIe, the bottleneck (6 x 88 us
    // against 400 us of upload).  The chunks therefore HALVE: the big ones keep the compute stream behind the
    // copy engine, and what is left exposed at the end -- transform + download of the last chunk (forward), or
    // upload + transform of the first (inverse) -- belongs to the smallest one (each chunk takes 58 % of what is
    // left).  Forward runs large -> small (the tail is after the last upload), inverse small -> large (the head
    // is before the first download).
    const double total_mb = (double)batch * pl->grid_size * 8e-6;
    static const int max_chunks = [] { const char* e = getenv("HB_PIPE_CHUNKS"); const int v = e ? atoi(e) : 8;
                                       return std::max(1, std::min(v, (int)Pipeline::kMaxChunks)); }();
    int n_chunks = (int)std::min<double>(max_chunks, std::max(1.0, total_mb / 8.0 + 0.5));
    if (total_mb >= 4.0) n_chunks = std::max(n_chunks, std::min(3, max_chunks));
    n_chunks = std::min(n_chunks, batch);
    int sizes[Pipeline::kMaxChunks];
    for (int c = 0, left = batch; c < n_chunks; c++) {
        const int chunks_left = n_chunks - c;
        sizes[c] = chunks_left == 1 ? left : std::max(1, std::min((int)(0.58 * left + 0.99), left - (chunks_left - 1)));
        left -= sizes[c];
    }
    if (!forward) std::reverse(sizes, sizes + n_chunks);
    if (capturing) {   // fork the upload stream off the capture origin
        HB_CUDA(cudaEventRecord(g_pipe.fork, g_pipe.compute));
        HB_CUDA(cudaStreamWaitEvent(g_pipe.in, g_pipe.fork, 0));
    }
    g_pipe.last_chunks = n_chunks;
    if (g_pipe.debug) HB_CUDA(cudaEventRecord(g_pipe.start, g_pipe.in));
    // Host rows of n doubles <-> device rows padded to an even count (16-byte TMA rows).  A pitched copy over
    // PCIe runs at 23 GB/s on this PCIe 5 x16 link against 55 GB/s for a flat one (tools/pcie_probe.cu), so the
    // rows cross the bus contiguously (g_stage) and are re-pitched by a kernel (a 2-D device-to-device memcpy
    // would queue on a copy engine, behind or ahead of the PCIe transfers the pipeline is trying to keep busy).
    // All uploads are queued first.  With both directions busy this link gives each ~30 GB/s instead of 55, so
    // the forward transform (large uploads, small downloads) holds its downloads back until the last upload has
    // landed: the uploads then finish at full rate and the downloads overlap the transform of the last chunk.
    for (int c = 0, b0 = 0; c < n_chunks; b0 += sizes[c], c++) {
        const int nb = sizes[c];
        if (forward) {
            double* dst = pitched ? g_stage.d() + (size_t)b0 * n_grid : g_in.d() + (size_t)b0 * in_dev;
            HB_CUDA(cudaMemcpyAsync(dst, f + (size_t)b0 * n_grid, (size_t)nb * n_grid * 8, cudaMemcpyHostToDevice, g_pipe.in));
        } else {
            double* dst = cpitched ? g_stage_coef.d() + (size_t)b0 * pl->n_polys : g_in.d() + (size_t)b0 * in_dev;
            HB_CUDA(cudaMemcpyAsync(dst, f + (size_t)b0 * pl->n_polys, (size_t)nb * pl->n_polys * 8, cudaMemcpyHostToDevice,
                                    g_pipe.in));
        }
        HB_CUDA(cudaEventRecord(g_pipe.uploaded[c], g_pipe.in));
    }
    if (forward && n_chunks > 1) HB_CUDA(cudaStreamWaitEvent(g_pipe.out, g_pipe.uploaded[n_chunks - 1], 0));
    for (int c = 0, b0 = 0; c < n_chunks; b0 += sizes[c], c++) {
        const int nb = sizes[c];
        double* d_in = g_in.d() + (size_t)b0 * in_dev;
        double* d_out = g_out.d() + (size_t)b0 * out_dev;
        double* d_flat = g_stage.d() + (size_t)b0 * n_grid;
        double* d_cflat = g_stage_coef.d() + (size_t)b0 * pl->n_polys;
        HB_CUDA(cudaStreamWaitEvent(g_pipe.compute, g_pipe.uploaded[c], 0));
        if (forward) {
            if (pitched) HB_TRY(launch_repack_rows(d_flat, L.n, d_in, L.np, pl->grid_rows * nb, L.n, g_pipe.compute));
            HB_TRY(plan_forward(pl, d_in, pl->grid_size, nb, d_out, cstride, g_pipe.compute));
            if (cpitched) HB_TRY(launch_repack_rows(d_out, cstride, d_cflat, pl->n_polys, nb, (int)pl->n_polys, g_pipe.compute));
        } else {
            if (cpitched) HB_TRY(launch_repack_rows(d_cflat, pl->n_polys, d_in, cstride, nb, (int)pl->n_polys, g_pipe.compute));
            HB_TRY(plan_backward(pl, d_in, cstride, nb, d_out, pl->grid_size, g_pipe.compute));
            if (pitched) HB_TRY(launch_repack_rows(d_out, L.np, d_flat, L.n, pl->grid_rows * nb, L.n, g_pipe.compute));
        }
        HB_CUDA(cudaEventRecord(g_pipe.computed[c], g_pipe.compute));
        HB_CUDA(cudaStreamWaitEvent(g_pipe.out, g_pipe.computed[c], 0));
        if (forward) {
            HB_CUDA(cudaMemcpyAsync(out + (size_t)b0 * pl->n_polys, cpitched ? d_cflat : d_out, (size_t)nb * pl->n_polys * 8,
                                    cudaMemcpyDeviceToHost, g_pipe.out));
        } else {
            HB_CUDA(cudaMemcpyAsync(out + (size_t)b0 * n_grid, pitched ? d_flat : d_out, (size_t)nb * n_grid * 8,
                                    cudaMemcpyDeviceToHost, g_pipe.out));
        }
        if (g_pipe.debug) HB_CUDA(cudaEventRecord(g_pipe.downloaded[c], g_pipe.out));
    }
    if (capturing) {   // join both side streams back into the origin
        HB_CUDA(cudaEventRecord(g_pipe.join_in, g_pipe.in));
        HB_CUDA(cudaStreamWaitEvent(g_pipe.compute, g_pipe.join_in, 0));
        HB_CUDA(cudaEventRecord(g_pipe.join_out, g_pipe.out));
        HB_CUDA(cudaStreamWaitEvent(g_pipe.compute, g_pipe.join_out, 0));
    }
    return HB_OK;
}
}  // namespace
}  // namespace hb

int hb_transform_batch(int dim, const int* n_pts, int degree, int batch, const double* f, const double* nodes,
                       const double* weights, const int* do_fourier, int forward, const char* index_set,
                       double* out, size_t out_len_each) {
    // the call mutex is taken BEFORE the plan look-up: the LRU eviction in cached_plan() deletes plans, and a plan
    // may only be deleted while no other host-pointer call is using it
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(cached_plan(&pl, dim, n_pts, degree, nodes, weights, do_fourier, index_set));
    if (batch < 1) return HB_OK;
    HB_TRY(g_pipe.init());
    const Axis& L = pl->ax[dim - 1];
    const long long n_grid = pl->grid_rows * L.n;
    const long long cstride = dim >= 2 ? pl->n_polys : even_up(pl->n_polys);   // see issue_pipeline
    const size_t expect = forward ? (size_t)pl->n_polys : (size_t)n_grid;
    if (out_len_each != expect) {
        set_error("transform: out_len %zu != %s %zu", out_len_each, forward ? "n_polys" : "grid size", expect);
        return HB_ERR_INVALID;
    }
    const long long in_dev = forward ? pl->grid_size : cstride, out_dev = forward ? cstride : pl->grid_size;
    HB_TRY(g_in.ensure((size_t)batch * in_dev * 8));
    HB_TRY(g_out.ensure((size_t)batch * out_dev * 8));
    if (L.np != L.n) HB_TRY(g_stage.ensure((size_t)batch * n_grid * 8));
    if (cstride != pl->n_polys) HB_TRY(g_stage_coef.ensure((size_t)batch * pl->n_polys * 8));

    // ---- graph replay for repeated calls on the same page-locked buffers ----
    static const bool use_graphs = [] { const char* e = getenv("HB_PIPE_GRAPH"); return !(e && atoi(e) == 0); }();
    PipeGraph* slot = nullptr;
    if (use_graphs && !g_pipe.debug && (size_t)batch * n_grid * 8 >= (1u << 20) && page_locked(f) && page_locked(out)) {
        PipeGraph* oldest = &g_pipe_graphs[0];
        for (PipeGraph& g : g_pipe_graphs) {
            if (g.calls && g.plan == pl && g.batch == batch && g.forward == forward && g.src == f && g.dst == out) {
                slot = &g;
                break;
            }
            if (g.stamp < oldest->stamp) oldest = &g;
        }
        if (!slot) {
            slot = oldest;
            slot->drop();
            *slot = PipeGraph();
            slot->plan = pl; slot->batch = batch; slot->forward = forward; slot->src = f; slot->dst = out;
        }
        slot->stamp = ++g_pipe_stamp;
        slot->calls++;
        const unsigned long long gen = g_devbuf_generation.load(std::memory_order_relaxed);
        if (slot->exec && slot->generation != gen) { slot->drop(); slot->calls = 2; }   // a buffer moved: recapture
        if (!slot->exec && slot->calls >= 2 && !slot->refused) {
            cudaGraph_t graph = nullptr;
            bool good = cudaStreamBeginCapture(g_pipe.compute, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (good) {
                const long long launches_before = hb::launch_count(0);
                const int rc = issue_pipeline(pl, dim, batch, f, forward, out, true);
                const cudaError_t e = cudaStreamEndCapture(g_pipe.compute, &graph);
                good = rc == HB_OK && e == cudaSuccess && graph != nullptr;
                if (good) good = cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                slot->launches = hb::launch_count(0) - launches_before;
                hb::note_launch(-slot->launches);   // nothing ran yet: replays are counted when launched
            }
            if (!good) {
                cudaGetLastError();
                slot->drop();
                slot->refused = true;
            }
            // anything allocated while capturing (first use of a table) moved the generation: take it after
            slot->generation = g_devbuf_generation.load(std::memory_order_relaxed);
            if (good && slot->generation != gen) { slot->drop(); slot->calls = 1; }
        }
        if (slot->exec) {
            HB_CUDA(cudaGraphLaunch(slot->exec, g_pipe.compute));
            hb::note_launch(slot->launches);
            HB_CUDA(cudaStreamSynchronize(g_pipe.compute));
            return HB_OK;
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    HB_TRY(issue_pipeline(pl, dim, batch, f, forward, out, false));
    const auto t1 = std::chrono::steady_clock::now();
    HB_CUDA(cudaStreamSynchronize(g_pipe.out));
    HB_CUDA(cudaStreamSynchronize(g_pipe.compute));
    if (g_pipe.debug) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[hb pipe] %s batch %d: issue %.0f us, issue+wait %.0f us; per chunk (us after the first upload starts):",
                forward ? "forward" : "inverse", batch, std::chrono::duration<double, std::micro>(t1 - t0).count(),
                std::chrono::duration<double, std::micro>(t2 - t0).count());
        for (int c = 0; c < g_pipe.last_chunks; c++) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, g_pipe.start, g_pipe.uploaded[c]);
            cudaEventElapsedTime(&b, g_pipe.start, g_pipe.computed[c]);
            cudaEventElapsedTime(&d, g_pipe.start, g_pipe.downloaded[c]);
            fprintf(stderr, " [up %.0f, done %.0f, down %.0f]", a * 1e3, b * 1e3, d * 1e3);
        }
        fprintf(stderr, "\n");
    }
    return HB_OK;
}

int hb_transform(int dim, const int* n_pts, int degree, const double* f, const double* nodes,
                 const double* weights, const int* do_fourier, int forward, const char* index_set, double* out,
                 size_t out_len) {
    return hb_transform_batch(dim, n_pts, degree, 1, f, nodes, weights, do_fourier, forward, index_set, out, out_len);
}

// ---- the object API's transform / eval / varf without host arithmetic or extra PCIe trips ------------------
// (reference: hermipy/quad.py:287-342.  There the function and the factor of the quadrature are discretised, divided
// or multiplied with NumPy on the host, and only then handed to the C++ transform.)
int hb_transform_weighted(int dim, const int* n_pts, int degree, int batch, const double* f, const char* f_body,
                          const char* factor_body, const double* nodes, const double* weights, const double* translation,
                          const double* dilation, const int* do_fourier, const char* index_set, double* out,
                          size_t out_len_each) {
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(cached_plan(&pl, dim, n_pts, degree, nodes, weights, do_fourier, index_set));
    if (batch < 1) return HB_OK;
    if ((f == nullptr) == (f_body == nullptr) || (f_body && batch != 1)) {
        set_error("transform_weighted: give grid values OR one expression");
        return HB_ERR_INVALID;
    }
    if (out_len_each != (size_t)pl->n_polys) { set_error("transform_weighted: out_len %zu != n_polys %lld", out_len_each, pl->n_polys); return HB_ERR_INVALID; }
    const long long gs = pl->grid_size, cs = even_up(pl->n_polys);
    HB_TRY(g_in.ensure((size_t)batch * gs * 8));
    HB_TRY(g_out.ensure((size_t)batch * cs * 8));
    if (f) {
        HB_TRY(upload_grid(pl, f, batch, g_in.d()));
    } else {
        HB_CUDA(cudaMemsetAsync(g_in.p, 0, (size_t)gs * 8, 0));
        HB_TRY(plan_discretize(pl, f_body, translation, dilation, g_in.d(), 0));
    }
    if (factor_body) {
        HB_TRY(g_aux[0].ensure((size_t)gs * 8));
        HB_CUDA(cudaMemsetAsync(g_aux[0].p, 0, (size_t)gs * 8, 0));
        HB_TRY(plan_discretize(pl, factor_body, translation, dilation, g_aux[0].d(), 0));
        HB_TRY(launch_pointwise(0, g_in.d(), gs, g_aux[0].d(), g_in.d(), gs, gs, batch, 0));
    }
    HB_TRY(plan_forward(pl, g_in.d(), gs, batch, g_out.d(), cs, 0));
    HB_CUDA(cudaMemcpy2DAsync(out, (size_t)pl->n_polys * 8, g_out.p, (size_t)cs * 8, (size_t)pl->n_polys * 8, batch,
                              cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

int hb_eval_weighted(int dim, const int* n_pts, int degree, const double* coeffs, size_t n_polys, const double* eval_nodes,
                     const double* nodes, const double* weights, const char* factor_body, const double* translation,
                     const double* dilation, const int* do_fourier, const char* index_set, double* out, size_t out_len) {
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(cached_plan(&pl, dim, n_pts, degree, eval_nodes, weights, do_fourier, index_set));
    const Axis& L = pl->ax[dim - 1];
    const long long gs = pl->grid_size, n_grid = pl->grid_rows * L.n;
    if (n_polys != (size_t)pl->n_polys || out_len != (size_t)n_grid) {
        set_error("eval_weighted: %zu coefficients / %zu values, expected %lld / %lld", n_polys, out_len, pl->n_polys, n_grid);
        return HB_ERR_INVALID;
    }
    HB_TRY(g_in.ensure((size_t)even_up(pl->n_polys) * 8));
    HB_TRY(g_out.ensure((size_t)gs * 8));
    HB_CUDA(cudaMemcpyAsync(g_in.p, coeffs, n_polys * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(plan_backward(pl, g_in.d(), even_up(pl->n_polys), 1, g_out.d(), gs, 0));
    if (factor_body) {
        // the factor lives on the quadrature's OWN nodes (the plan above sits on the mapped nodes of the series)
        size_t tot = 0;
        for (int a = 0; a < dim; a++) tot += n_pts[a];
        HB_TRY(g_aux[1].ensure(tot * 8));
        HB_TRY(g_aux[0].ensure((size_t)gs * 8));
        HB_CUDA(cudaMemcpyAsync(g_aux[1].p, nodes, tot * 8, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaMemsetAsync(g_aux[0].p, 0, (size_t)gs * 8, 0));
        HB_TRY(discretize_device(factor_body, dim, n_pts, g_aux[1].d(), translation, dilation, g_aux[0].d(), L.np, 0));
        HB_TRY(launch_pointwise(1, g_out.d(), gs, g_aux[0].d(), g_out.d(), gs, gs, 1, 0));
    }
    HB_CUDA(cudaMemcpy2DAsync(out, (size_t)L.n * 8, g_out.p, (size_t)L.np * 8, (size_t)L.n * 8, (size_t)pl->grid_rows,
                              cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

int hb_integrate(int dim, const int* n_pts, const double* f, const double* nodes, const double* weights,
                 const int* do_fourier, double* result) {
    // transform.cpp:170-181: degree-0 triangle transform, times sqrt(2 pi) per Fourier axis
    double c0 = 0;
    HB_TRY(hb_transform(dim, n_pts, 0, f, nodes, weights, do_fourier, 1, "triangle", &c0, 1));
    double factor = 1;
    for (int a = 0; a < dim; a++)
        if (do_fourier && do_fourier[a] == 1) factor *= std::sqrt(2 * M_PI);
    *result = factor * c0;
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// discretize (module.cpp:88 -> discretize.cpp:112-126)
int hb_discretize(const char* body, int dim, const int* n_pts, const double* nodes, const double* translation,
                  const double* dilation, double* out, size_t out_len) {
    HB_TRY(check_device());
    if (!body || dim < 1 || dim > HB_MAX_DIM || !n_pts || !nodes) return invalid(__func__);
    size_t tot = 0, total = 1;
    for (int a = 0; a < dim; a++) { if (n_pts[a] < 1) return invalid(__func__); tot += n_pts[a]; total *= n_pts[a]; }
    if (out_len != total) { set_error("discretize: out_len %zu != grid size %zu", out_len, total); return HB_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(g_in.ensure(tot * 8));
    HB_TRY(g_out.ensure(total * 8));
    HB_CUDA(cudaMemcpyAsync(g_in.p, nodes, tot * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(discretize_device(body, dim, n_pts, g_in.d(), translation, dilation, g_out.d(), n_pts[dim - 1], 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, total * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
static int triple_products_impl(int degree, bool fourier, double* out) {
    HB_TRY(check_device());
    if (degree < 0) return invalid(__func__);
    if (fourier && (degree % 2 == 1)) { set_error("Use even degree for Fourier"); return HB_ERR_FOURIER_DEGREE; }
    const int P = degree + 1, Pp = (int)even_up(P), K = 2 * degree + 1;
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(g_out.ensure((size_t)K * P * Pp * 8));
    if (fourier) {
        std::vector<double> h;
        fourier_triple_products_host(degree, Pp, h);
        for (int k = 0; k < K; k++)
            for (int i = 0; i < P; i++) memcpy(out + ((size_t)k * P + i) * P, &h[((size_t)k * P + i) * Pp], P * 8);
        return HB_OK;
    }
    HB_TRY(launch_triple_products(degree, Pp, g_out.d(), 0));
    HB_CUDA(cudaMemcpy2DAsync(out, (size_t)P * 8, g_out.p, (size_t)Pp * 8, (size_t)P * 8, (size_t)K * P,
                              cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}
int hb_triple_products(int degree, double* out) { return triple_products_impl(degree, false, out); }
int hb_triple_products_fourier(int degree, double* out) { return triple_products_impl(degree, true, out); }

// grid values of the varf argument on the device: uploaded, or evaluated from an expression (no host round trip)
static int varf_input(hb_plan* pl, const double* f, const char* f_body, const double* translation, const double* dilation) {
    HB_TRY(g_in.ensure((size_t)pl->grid_size * 8));
    if (f) return upload_grid(pl, f, 1, g_in.d());
    HB_CUDA(cudaMemsetAsync(g_in.p, 0, (size_t)pl->grid_size * 8, 0));
    return plan_discretize(pl, f_body, translation, dilation, g_in.d(), 0);
}

static int varf_to_device(hb_plan** plp, int dim, const int* n_pts, int degree, const double* f, const double* nodes,
                          const double* weights, const int* do_fourier, const char* index_set, int mode,
                          const char* f_body = nullptr, const double* translation = nullptr, const double* dilation = nullptr) {
    hb_plan* pl = nullptr;
    HB_TRY(cached_plan(&pl, dim, n_pts, degree, nodes, weights, do_fourier, index_set));
    *plp = pl;
    HB_TRY(g_out.ensure((size_t)pl->n_polys * pl->n_polys * 8));
    HB_TRY(varf_input(pl, f, f_body, translation, dilation));
    return plan_varf(pl, g_in.d(), mode, g_out.d(), pl->n_polys, 0, pl->n_polys, 0);
}

int hb_varf(int dim, const int* n_pts, int degree, const double* f, const double* nodes, const double* weights,
            const int* do_fourier, const char* index_set, int mode, double* out, size_t n_polys) {
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(varf_to_device(&pl, dim, n_pts, degree, f, nodes, weights, do_fourier, index_set, mode));
    if (n_polys != (size_t)pl->n_polys) { set_error("varf: n_polys %zu != %lld", n_polys, pl->n_polys); return HB_ERR_INVALID; }
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, n_polys * n_polys * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

// dense device matrix (n x n, pitch ld) -> g_sp, exact zeros dropped (sparse.cpp:44-58)
static int compact_dense_to_csr(const double* A, int n, long long ld) {
    HB_TRY(g_sp_rownnz.ensure((size_t)n * 8));
    HB_TRY(launch_count_nnz(A, ld, n, n, static_cast<long long*>(g_sp_rownnz.p), 0));
    std::vector<long long> rn(n), ip(n + 1, 0);
    HB_CUDA(cudaMemcpyAsync(rn.data(), g_sp_rownnz.p, (size_t)n * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    for (int i = 0; i < n; i++) ip[i + 1] = ip[i] + rn[i];
    g_sp.rows = n;
    g_sp.nnz = ip[n];
    HB_TRY(g_sp.indptr.ensure((size_t)(n + 1) * 8));
    HB_TRY(g_sp.indices.ensure(std::max<size_t>(g_sp.nnz, 1) * 4));
    HB_TRY(g_sp.data.ensure(std::max<size_t>(g_sp.nnz, 1) * 8));
    HB_CUDA(cudaMemcpyAsync(g_sp.indptr.p, ip.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(launch_fill_csr(A, ld, n, n, static_cast<long long*>(g_sp.indptr.p), g_sp.indices.i(), g_sp.data.d(), 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

static int varf_sp_impl(int dim, const int* n_pts, int degree, const double* f, const char* f_body,
                        const double* translation, const double* dilation, const double* nodes, const double* weights,
                        const int* do_fourier, const char* index_set, int mode, long long* nnz);

int hb_varf_sp(int dim, const int* n_pts, int degree, const double* f, const double* nodes, const double* weights,
               const int* do_fourier, const char* index_set, int mode, long long* nnz) {
    return varf_sp_impl(dim, n_pts, degree, f, nullptr, nullptr, nullptr, nodes, weights, do_fourier, index_set, mode, nnz);
}

/* varf of a function given as an expression in v[i] (evaluated on the device at translation + dilation @ node):
 * what Quad.varf('expr') needs, without discretising on the host first (quad.py:308-328) */
int hb_varf_function(int dim, const int* n_pts, int degree, const char* f_body, const double* translation,
                     const double* dilation, const double* nodes, const double* weights, const int* do_fourier,
                     const char* index_set, int mode, double* out, size_t n_polys) {
    if (!f_body) return invalid(__func__);
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(varf_to_device(&pl, dim, n_pts, degree, nullptr, nodes, weights, do_fourier, index_set, mode, f_body, translation, dilation));
    if (n_polys != (size_t)pl->n_polys) { set_error("varf: n_polys %zu != %lld", n_polys, pl->n_polys); return HB_ERR_INVALID; }
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, n_polys * n_polys * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}
int hb_varf_function_sp(int dim, const int* n_pts, int degree, const char* f_body, const double* translation,
                        const double* dilation, const double* nodes, const double* weights, const int* do_fourier,
                        const char* index_set, int mode, long long* nnz) {
    if (!f_body) return invalid(__func__);
    return varf_sp_impl(dim, n_pts, degree, nullptr, f_body, translation, dilation, nodes, weights, do_fourier, index_set, mode, nnz);
}

static int varf_sp_impl(int dim, const int* n_pts, int degree, const double* f, const char* f_body,
                        const double* translation, const double* dilation, const double* nodes, const double* weights,
                        const int* do_fourier, const char* index_set, int mode, long long* nnz) {
    std::lock_guard<std::mutex> lock(g_call_mutex);
    hb_plan* pl = nullptr;
    HB_TRY(cached_plan(&pl, dim, n_pts, degree, nodes, weights, do_fourier, index_set));
    HB_TRY(varf_input(pl, f, f_body, translation, dilation));
    g_sp.rows = -1;
    // polynomial-like f: the band of retained triple products goes straight to CSR, no n_polys^2 buffer
    if (mode == HB_VARF_REFERENCE) HB_TRY(plan_varf_csr(pl, g_in.d(), 0, pl->n_polys, g_sp, 0));
    if (g_sp.rows < 0) {
        HB_TRY(g_out.ensure((size_t)pl->n_polys * pl->n_polys * 8));
        HB_TRY(plan_varf(pl, g_in.d(), mode, g_out.d(), pl->n_polys, 0, pl->n_polys, 0));
        HB_TRY(compact_dense_to_csr(g_out.d(), (int)pl->n_polys, pl->n_polys));
    }
    *nnz = g_sp.nnz;
    return HB_OK;
}

int hb_plan_varf_sp(hb_plan* plan, const double* d_f, int mode, long long row_begin, long long row_end, long long* nnz,
                    void* stream) {
    if (!plan || !nnz) return invalid(__func__);
    std::lock_guard<std::mutex> lock(g_call_mutex);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (row_begin < 0) row_begin = 0;
    if (row_end > plan->n_polys) row_end = plan->n_polys;
    g_sp.rows = -1;
    if (mode == HB_VARF_REFERENCE) HB_TRY(plan_varf_csr(plan, d_f, row_begin, row_end, g_sp, st));
    if (g_sp.rows < 0) {
        set_error("hb_plan_varf_sp: f is not polynomial-like (no narrow band of retained triple products); assemble the "
                  "dense row block with hb_plan_varf");
        return HB_ERR_INVALID;
    }
    *nnz = g_sp.nnz;
    return HB_OK;
}

int hb_sparse_fetch(long long* indptr, int* indices, double* data) {
    std::lock_guard<std::mutex> lock(g_call_mutex);
    if (g_sp.rows < 0) { set_error("hb_sparse_fetch: no sparse result"); return HB_ERR_INVALID; }
    HB_CUDA(cudaMemcpy(indptr, g_sp.indptr.p, (size_t)(g_sp.rows + 1) * 8, cudaMemcpyDeviceToHost));
    if (g_sp.nnz) {
        HB_CUDA(cudaMemcpy(indices, g_sp.indices.p, (size_t)g_sp.nnz * 4, cudaMemcpyDeviceToHost));
        HB_CUDA(cudaMemcpy(data, g_sp.data.p, (size_t)g_sp.nnz * 8, cudaMemcpyDeviceToHost));
    }
    return HB_OK;
}

// column map of varfd (varf.cpp:423-453): entry (i, j) = fac[j] * var[i][src[j]], src[j] = index(m_j -/+ e_dir)
static int varfd_column_map(int dim, int degree, int direction, size_t n_polys, int do_fourier, const char* index_set,
                            std::vector<int>& src, std::vector<double>& fac) {
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    auto s = get_index_set(set, dim, degree);
    if (!s || direction < 0 || direction >= dim) return invalid(__func__);
    if (n_polys != s->size) { set_error("varfd: matrix size %zu != n_polys %u", n_polys, s->size); return HB_ERR_INVALID; }
    const int n = (int)n_polys;
    src.assign(n, -1);
    fac.assign(n, 0.0);
    std::vector<uint32_t> m(dim);
    for (int j = 0; j < n; j++) {
        for (int a = 0; a < dim; a++) m[a] = s->list[(size_t)j * dim + a];
        const uint32_t md = m[direction];
        if (md == 0) continue;
        if (do_fourier == 0) {
            m[direction] = md - 1;
            fac[j] = std::sqrt((double)md);
        } else {
            const bool is_cos = (md % 2) == 0;
            m[direction] = is_cos ? md - 1 : md + 1;
            const double wave = (double)((md + (is_cos ? 0 : 1)) / 2);
            fac[j] = is_cos ? -wave : wave;
        }
        src[j] = s->position(m.data());
    }
    return HB_OK;
}

int hb_varfd(int dim, int degree, int direction, const double* var, size_t n_polys, int do_fourier,
             const char* index_set, double* out) {
    HB_TRY(check_device());
    std::vector<int> src;
    std::vector<double> fac;
    HB_TRY(varfd_column_map(dim, degree, direction, n_polys, do_fourier, index_set, src, fac));
    const int n = (int)n_polys;
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(g_in.ensure(n_polys * n_polys * 8));
    HB_TRY(g_out.ensure(n_polys * n_polys * 8));
    HB_TRY(g_auxi[0].ensure((size_t)n * 4));
    HB_TRY(g_aux[0].ensure((size_t)n * 8));
    HB_CUDA(cudaMemcpyAsync(g_in.p, var, n_polys * n_polys * 8, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_auxi[0].p, src.data(), (size_t)n * 4, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_aux[0].p, fac.data(), (size_t)n * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(launch_varfd(g_in.d(), n, g_auxi[0].i(), g_aux[0].d(), g_out.d(), n, n, n, 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, n_polys * n_polys * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
static int tensorize_impl(int k, const double* const* inputs, const int* lens, const int* dirs_flat,
                          const int* dir_lens, const char* index_set, bool matrices, double* out, size_t out_n) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (k < 1 || k > HB_MAX_FACTORS) { set_error("tensorize: between 1 and %d factors supported", HB_MAX_FACTORS); return HB_ERR_INVALID; }
    int dim = 0;
    std::vector<std::vector<int>> dirs(k);
    for (int j = 0, off = 0; j < k; j++) {
        dirs[j].assign(dirs_flat + off, dirs_flat + off + dir_lens[j]);
        off += dir_lens[j];
        dim += dir_lens[j];
    }
    std::vector<char> seen(dim, 0);
    for (auto& dj : dirs)
        for (int v : dj) {
            if (v < 0 || v >= dim) { set_error("Missing direction, exiting ..."); return HB_ERR_INVALID; }
            if (seen[v]) { set_error("The same direction appears twice, exiting ..."); return HB_ERR_INVALID; }
            seen[v] = 1;
        }
    int degree = 0;
    if (index_get_degree(set, dir_lens[0], lens[0], &degree)) {
        set_error("Can't find a degree with %d polynomials in dimension %d", lens[0], dir_lens[0]);
        return HB_ERR_DEGREE;
    }
    auto full = get_index_set(set, dim, degree);
    if (!full) return invalid(__func__);
    if (out_n != full->size) { set_error("tensorize: output size %zu != %u", out_n, full->size); return HB_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    TensorizeArgs a;
    a.k = k;
    for (int j = 0; j < k; j++) {
        auto sub = get_index_set(set, dir_lens[j], degree);
        if (!sub || (int)sub->size != lens[j]) {
            set_error("Size of input does not match dimension! Expected dimension: %u Actual dimension: %d",
                      sub ? sub->size : 0u, lens[j]);
            return HB_ERR_INVALID;
        }
        std::vector<int> map = restriction_map(*full, dirs[j], *sub);
        const size_t in_elems = matrices ? (size_t)lens[j] * lens[j] : (size_t)lens[j];
        HB_TRY(g_aux[j].ensure(in_elems * 8));
        HB_TRY(g_auxi[j].ensure(map.size() * 4));
        HB_CUDA(cudaMemcpyAsync(g_aux[j].p, inputs[j], in_elems * 8, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaMemcpyAsync(g_auxi[j].p, map.data(), map.size() * 4, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaStreamSynchronize(0));  // `map` is a temporary
        a.in[j] = g_aux[j].d();
        a.map[j] = g_auxi[j].i();
        a.size[j] = lens[j];
    }
    const size_t out_elems = matrices ? out_n * out_n : out_n;
    HB_TRY(g_out.ensure(out_elems * 8));
    HB_TRY(launch_tensorize(a, matrices, g_out.d(), (int)out_n, 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, out_elems * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

int hb_tensorize_vecs(int k, const double* const* inputs, const int* lens, const int* dirs_flat,
                      const int* dir_lens, const char* index_set, double* out, size_t out_len) {
    return tensorize_impl(k, inputs, lens, dirs_flat, dir_lens, index_set, false, out, out_len);
}
int hb_tensorize_mats(int k, const double* const* inputs, const int* sizes, const int* dirs_flat,
                      const int* dir_lens, const char* index_set, double* out, size_t n_out) {
    return tensorize_impl(k, inputs, sizes, dirs_flat, dir_lens, index_set, true, out, n_out);
}

static int project_impl(const double* input, size_t n_in, int dim, const int* dirs, int n_dirs, const char* index_set,
                        bool matrix, double* out, size_t n_out) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (n_dirs < 1 || n_dirs > dim) return invalid(__func__);
    int degree = 0;
    if (index_get_degree(set, dim, (long long)n_in, &degree)) {
        set_error("Can't find a degree with %zu polynomials in dimension %d", n_in, dim);
        return HB_ERR_DEGREE;
    }
    auto full = get_index_set(set, dim, degree);
    auto sub = get_index_set(set, n_dirs, degree);
    if (!full || !sub) return invalid(__func__);
    if (n_out != sub->size) { set_error("project: output size %zu != %u", n_out, sub->size); return HB_ERR_INVALID; }
    // sel[i] = position in the full set of sub-index i embedded on dirs (zeros elsewhere); project.cpp:48-56
    // scans the full set and writes aligned entries, which visits each sub-index at most once.
    std::vector<int> sel(n_out, -1);
    std::vector<uint32_t> m(dim);
    for (uint32_t i = 0; i < sub->size; i++) {
        std::fill(m.begin(), m.end(), 0u);
        for (int j = 0; j < n_dirs; j++) {
            if (dirs[j] < 0 || dirs[j] >= dim) return invalid(__func__);
            m[dirs[j]] = sub->list[(size_t)i * n_dirs + j];
        }
        sel[i] = full->position(m.data());
    }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    const size_t in_elems = matrix ? n_in * n_in : n_in, out_elems = matrix ? n_out * n_out : n_out;
    HB_TRY(g_in.ensure(in_elems * 8));
    HB_TRY(g_out.ensure(out_elems * 8));
    HB_TRY(g_auxi[0].ensure(n_out * 4));
    HB_CUDA(cudaMemcpyAsync(g_in.p, input, in_elems * 8, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_auxi[0].p, sel.data(), n_out * 4, cudaMemcpyHostToDevice, 0));
    HB_TRY(launch_project(g_in.d(), (int)n_in, g_auxi[0].i(), matrix, g_out.d(), (int)n_out, 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, out_elems * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

int hb_project_vec(const double* input, size_t len, int dim, const int* dirs, int n_dirs, const char* index_set,
                   double* out, size_t out_len) {
    return project_impl(input, len, dim, dirs, n_dirs, index_set, false, out, out_len);
}
int hb_project_mat(const double* input, size_t n, int dim, const int* dirs, int n_dirs, const char* index_set,
                   double* out, size_t n_out) {
    return project_impl(input, n, dim, dirs, n_dirs, index_set, true, out, n_out);
}

// ---- CSR inputs, CSR result (no dense intermediate) -----------------------------------------------------
static int upload_csr(int slot, int n, const long long* indptr, const int* indices, const double* data) {
    const long long nnz = indptr[n];
    HB_TRY(g_sp_in[3 * slot].ensure((size_t)(n + 1) * 8));
    HB_TRY(g_sp_in[3 * slot + 1].ensure(std::max<size_t>(nnz, 1) * 4));
    HB_TRY(g_sp_in[3 * slot + 2].ensure(std::max<size_t>(nnz, 1) * 8));
    HB_CUDA(cudaMemcpyAsync(g_sp_in[3 * slot].p, indptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (nnz) {
        HB_CUDA(cudaMemcpyAsync(g_sp_in[3 * slot + 1].p, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaMemcpyAsync(g_sp_in[3 * slot + 2].p, data, (size_t)nnz * 8, cudaMemcpyHostToDevice, 0));
    }
    return HB_OK;
}

int hb_tensorize_mats_sp(int k, const long long* const* indptr, const int* const* indices, const double* const* data,
                         const int* sizes, const int* dirs_flat, const int* dir_lens, const char* index_set,
                         long long* n_out, long long* nnz) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (k < 1 || k > HB_MAX_FACTORS) { set_error("tensorize: between 1 and %d factors supported", HB_MAX_FACTORS); return HB_ERR_INVALID; }
    int dim = 0;
    std::vector<std::vector<int>> dirs(k);
    for (int j = 0, off = 0; j < k; j++) {
        dirs[j].assign(dirs_flat + off, dirs_flat + off + dir_lens[j]);
        off += dir_lens[j];
        dim += dir_lens[j];
    }
    std::vector<char> seen(dim, 0);
    for (auto& dj : dirs)
        for (int v : dj) {
            if (v < 0 || v >= dim) { set_error("Missing direction, exiting ..."); return HB_ERR_INVALID; }
            if (seen[v]) { set_error("The same direction appears twice, exiting ..."); return HB_ERR_INVALID; }
            seen[v] = 1;
        }
    int degree = 0;
    if (index_get_degree(set, dir_lens[0], sizes[0], &degree)) {
        set_error("Can't find a degree with %d polynomials in dimension %d", sizes[0], dir_lens[0]);
        return HB_ERR_DEGREE;
    }
    auto full = get_index_set(set, dim, degree);
    if (!full) return invalid(__func__);
    // strides of the lexicographic box of the full set (axis 0 slowest)
    std::vector<long long> stride(dim, 1);
    for (int a = dim - 2; a >= 0; a--) stride[a] = stride[a + 1] * full->extent[a + 1];
    std::lock_guard<std::mutex> lock(g_call_mutex);
    SpTensorizeArgs a;
    a.k = k; a.n_out = (int)full->size;
    for (int j = 0; j < k; j++) {
        auto sub = get_index_set(set, dir_lens[j], degree);
        if (!sub || (int)sub->size != sizes[j]) {
            set_error("Size of input does not match dimension! Expected dimension: %u Actual dimension: %d",
                      sub ? sub->size : 0u, sizes[j]);
            return HB_ERR_INVALID;
        }
        HB_TRY(upload_csr(j, sizes[j], indptr[j], indices[j], data[j]));
        std::vector<int> map = restriction_map(*full, dirs[j], *sub);
        std::vector<long long> col_lex(sub->size, 0);
        for (uint32_t c = 0; c < sub->size; c++)
            for (int i = 0; i < dir_lens[j] && col_lex[c] >= 0; i++) {
                const uint32_t v = sub->list[(size_t)c * dir_lens[j] + i];
                // a component beyond the bounding box of the full set (the rectangle sets differ per axis): no such column
                if (v >= full->extent[dirs[j][i]]) col_lex[c] = -1;
                else col_lex[c] += (long long)v * stride[dirs[j][i]];
            }
        HB_TRY(g_auxi[j].ensure(map.size() * 4));
        HB_TRY(g_aux[j].ensure(col_lex.size() * 8));
        HB_CUDA(cudaMemcpyAsync(g_auxi[j].p, map.data(), map.size() * 4, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaMemcpyAsync(g_aux[j].p, col_lex.data(), col_lex.size() * 8, cudaMemcpyHostToDevice, 0));
        HB_CUDA(cudaStreamSynchronize(0));   // temporaries
        a.indptr[j] = static_cast<const long long*>(g_sp_in[3 * j].p);
        a.indices[j] = g_sp_in[3 * j + 1].i();
        a.data[j] = g_sp_in[3 * j + 2].d();
        a.map[j] = g_auxi[j].i();
        a.col_lex[j] = static_cast<const long long*>(g_aux[j].p);
    }
    HB_TRY(g_sp_lex.ensure(full->lex_to_pos.size() * 4));
    HB_CUDA(cudaMemcpyAsync(g_sp_lex.p, full->lex_to_pos.data(), full->lex_to_pos.size() * 4, cudaMemcpyHostToDevice, 0));
    a.lex_to_pos = g_sp_lex.i();
    HB_TRY(tensorize_csr(a, g_sp, 0));
    *n_out = g_sp.rows;
    *nnz = g_sp.nnz;
    return HB_OK;
}

int hb_project_mat_sp(const long long* indptr, const int* indices, const double* data, size_t n_in, int dim,
                      const int* dirs, int n_dirs, const char* index_set, long long* n_out, long long* nnz) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (n_dirs < 1 || n_dirs > dim) return invalid(__func__);
    int degree = 0;
    if (index_get_degree(set, dim, (long long)n_in, &degree)) {
        set_error("Can't find a degree with %zu polynomials in dimension %d", n_in, dim);
        return HB_ERR_DEGREE;
    }
    auto full = get_index_set(set, dim, degree);
    auto sub = get_index_set(set, n_dirs, degree);
    if (!full || !sub) return invalid(__func__);
    // position in the sub-set of every ALIGNED multi-index of the full set (zero outside dirs), -1 otherwise
    std::vector<int> to_sub(full->size, -1);
    std::vector<uint32_t> m(dim);
    for (uint32_t i = 0; i < sub->size; i++) {
        std::fill(m.begin(), m.end(), 0u);
        for (int j = 0; j < n_dirs; j++) {
            if (dirs[j] < 0 || dirs[j] >= dim) return invalid(__func__);
            m[dirs[j]] = sub->list[(size_t)i * n_dirs + j];
        }
        const int p = full->position(m.data());
        if (p >= 0) to_sub[p] = (int)i;
    }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(upload_csr(0, (int)n_in, indptr, indices, data));
    HB_TRY(g_auxi[0].ensure(to_sub.size() * 4));
    HB_CUDA(cudaMemcpyAsync(g_auxi[0].p, to_sub.data(), to_sub.size() * 4, cudaMemcpyHostToDevice, 0));
    HB_TRY(relabel_csr((int)n_in, indptr[n_in], static_cast<const long long*>(g_sp_in[0].p), g_sp_in[1].i(), g_sp_in[2].d(),
                       g_auxi[0].i(), g_auxi[0].i(), nullptr, sub->size, g_sp, 0));
    *n_out = g_sp.rows;
    *nnz = g_sp.nnz;
    return HB_OK;
}

int hb_varfd_sp(int dim, int degree, int direction, const long long* indptr, const int* indices, const double* data,
                size_t n_polys, int do_fourier, const char* index_set, long long* nnz) {
    HB_TRY(check_device());
    std::vector<int> src;
    std::vector<double> fac;
    HB_TRY(varfd_column_map(dim, degree, direction, n_polys, do_fourier, index_set, src, fac));
    // forward form of the map: stored column c of the input lands in column dst[c] with the factor of that column
    const int n = (int)n_polys;
    std::vector<int> dst(n, -1);
    std::vector<double> dfac(n, 0.0);
    for (int j = 0; j < n; j++)
        if (src[j] >= 0) { dst[src[j]] = j; dfac[src[j]] = fac[j]; }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(upload_csr(0, n, indptr, indices, data));
    HB_TRY(g_auxi[0].ensure((size_t)n * 4));
    HB_TRY(g_aux[0].ensure((size_t)n * 8));
    HB_CUDA(cudaMemcpyAsync(g_auxi[0].p, dst.data(), (size_t)n * 4, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_aux[0].p, dfac.data(), (size_t)n * 8, cudaMemcpyHostToDevice, 0));
    HB_TRY(relabel_csr(n, indptr[n], static_cast<const long long*>(g_sp_in[0].p), g_sp_in[1].i(), g_sp_in[2].d(), nullptr,
                       g_auxi[0].i(), g_aux[0].d(), n, g_sp, 0));
    *nnz = g_sp.nnz;
    return HB_OK;
}

int hb_inner(const double* s1, size_t n1, const double* s2, size_t n2, const int* dirs1, int nd1, const int* dirs2,
             int nd2, const char* index_set, double* out, size_t out_len) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    if (nd1 < 1 || nd2 < 1) return invalid(__func__);
    // merge the (increasing) direction lists: shared directions are contracted, the others survive
    // in increasing order (inner.cpp:61-84)
    std::vector<int> from1, slot, inner1, inner2;   // for every result axis: comes from s1? which axis there
    int i = 0, j = 0;
    while (i < nd1 || j < nd2) {
        if (i == nd1 || (j < nd2 && dirs1[i] > dirs2[j])) { from1.push_back(0); slot.push_back(j++); }
        else if (j == nd2 || (i < nd1 && dirs1[i] < dirs2[j])) { from1.push_back(1); slot.push_back(i++); }
        else { inner1.push_back(i++); inner2.push_back(j++); }
    }
    const int dim_res = (int)slot.size(), dim_in = (int)inner1.size();
    int degree = 0, degree2 = 0;
    if (index_get_degree(set, nd1, (long long)n1, &degree) || index_get_degree(set, nd2, (long long)n2, &degree2) ||
        degree != degree2) {
        set_error("inner: the two series do not have the same degree");
        return HB_ERR_DEGREE;
    }
    auto it1 = get_index_set(set, nd1, degree), it2 = get_index_set(set, nd2, degree);
    std::shared_ptr<const IndexSet> it_res = dim_res ? get_index_set(set, dim_res, degree) : nullptr;
    std::shared_ptr<const IndexSet> it_in = dim_in ? get_index_set(set, dim_in, degree) : nullptr;
    if (!it1 || !it2 || (dim_res && !it_res) || (dim_in && !it_in)) return invalid(__func__);
    const int n_res = dim_res ? (int)it_res->size : 1, n_in = dim_in ? (int)it_in->size : 1;
    if (out_len != (size_t)n_res) { set_error("inner: output size %zu != %d", out_len, n_res); return HB_ERR_INVALID; }
    std::vector<int> ind1((size_t)n_res * n_in), ind2((size_t)n_res * n_in);
    std::vector<uint32_t> m1(nd1), m2(nd2);
    for (int r = 0; r < n_res; r++) {
        for (int k = 0; k < dim_res; k++) {
            const uint32_t v = it_res->list[(size_t)r * dim_res + k];
            if (from1[k]) m1[slot[k]] = v; else m2[slot[k]] = v;
        }
        for (int q = 0; q < n_in; q++) {
            for (int k = 0; k < dim_in; k++) {
                const uint32_t v = it_in->list[(size_t)q * dim_in + k];
                m1[inner1[k]] = v;
                m2[inner2[k]] = v;
            }
            ind1[(size_t)q * n_res + r] = it1->position(m1.data());
            ind2[(size_t)q * n_res + r] = it2->position(m2.data());
        }
    }
    std::lock_guard<std::mutex> lock(g_call_mutex);
    HB_TRY(g_aux[0].ensure(n1 * 8));
    HB_TRY(g_aux[1].ensure(n2 * 8));
    HB_TRY(g_auxi[0].ensure(ind1.size() * 4));
    HB_TRY(g_auxi[1].ensure(ind2.size() * 4));
    HB_TRY(g_out.ensure((size_t)n_res * 8));
    HB_CUDA(cudaMemcpyAsync(g_aux[0].p, s1, n1 * 8, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_aux[1].p, s2, n2 * 8, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_auxi[0].p, ind1.data(), ind1.size() * 4, cudaMemcpyHostToDevice, 0));
    HB_CUDA(cudaMemcpyAsync(g_auxi[1].p, ind2.data(), ind2.size() * 4, cudaMemcpyHostToDevice, 0));
    HB_TRY(launch_inner(g_aux[0].d(), g_aux[1].d(), g_auxi[0].i(), g_auxi[1].i(), n_res, n_in, g_out.d(), 0));
    HB_CUDA(cudaMemcpyAsync(out, g_out.p, (size_t)n_res * 8, cudaMemcpyDeviceToHost, 0));
    HB_CUDA(cudaStreamSynchronize(0));
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
int hb_plan_create(hb_plan** plan, int dim, const int* n_pts, int degree, const double* nodes,
                   const double* weights, const int* do_fourier, const char* index_set) {
    HB_TRY(check_device());
    const int set = set_or_err(index_set);
    if (set < 0) return HB_ERR_INDEX_SET;
    return plan_create(plan, dim, n_pts, degree, nodes, weights, do_fourier, set);
}
void hb_plan_destroy(hb_plan* plan) { delete plan; }
long long hb_plan_n_polys(const hb_plan* plan) { return plan->n_polys; }
long long hb_plan_grid_pitch(const hb_plan* plan) { return plan->ax[plan->dim - 1].np; }
long long hb_plan_grid_size(const hb_plan* plan) { return plan->grid_size; }
int hb_plan_forward(hb_plan* plan, const double* d_f, long long f_stride, int batch, double* d_coef,
                    long long coef_stride, void* stream) {
    return plan_forward(plan, d_f, f_stride, batch, d_coef, coef_stride, static_cast<cudaStream_t>(stream));
}
int hb_plan_backward(hb_plan* plan, const double* d_coef, long long coef_stride, int batch, double* d_f,
                     long long f_stride, void* stream) {
    return plan_backward(plan, d_coef, coef_stride, batch, d_f, f_stride, static_cast<cudaStream_t>(stream));
}
int hb_plan_varf(hb_plan* plan, const double* d_f, int mode, double* d_out, long long ldo, long long row_begin,
                 long long row_end, void* stream) {
    return plan_varf(plan, d_f, mode, d_out, ldo, row_begin, row_end, static_cast<cudaStream_t>(stream));
}

int hb_plan_discretize(hb_plan* plan, const char* body, const double* translation, const double* dilation,
                       double* d_f, void* stream) {
    return plan_discretize(plan, body, translation, dilation, d_f, static_cast<cudaStream_t>(stream));
}
int hb_plan_set_sqrt_weights(hb_plan* plan, int axis, const double* sqrt_w) {
    if (!plan) return invalid(__func__);
    std::lock_guard<std::mutex> lock(g_call_mutex);
    return plan_set_sqrt_weights(plan, axis, sqrt_w);
}
int hb_plan_set_timing(hb_plan* plan, int on) { plan->timing = on != 0; plan->n_timed = 0; return HB_OK; }
int hb_plan_timing_count(const hb_plan* plan) { return plan->n_timed; }
int hb_plan_timing_get(hb_plan* plan, int i, int* tag, double* ms, double* padded_flops) {
    if (i < 0 || i >= plan->n_timed) return invalid(__func__);
    hb_plan::Timed& t = plan->timed[i];
    HB_CUDA(cudaEventSynchronize(t.e1));
    float el = 0;
    HB_CUDA(cudaEventElapsedTime(&el, t.e0, t.e1));
    *tag = t.tag; *ms = el; *padded_flops = t.flops;
    return HB_OK;
}
// Generic device GEMM (row-major FP64, see hb_gemm.cuh) and access to a plan's basis tables, used by
// the multi-GPU composition in hermipy_b200/dist.py.
int hb_dev_dgemm(int M, int N, int K, int batch, const double* A, long long lda, long long strideA, int a_mmajor,
                 const double* B, long long ldb, long long strideB, double* C, long long ldc, long long strideC,
                 void* stream) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.batch = batch;
    g.A = A; g.lda = lda; g.strideA = strideA; g.a_mmajor = a_mmajor != 0;
    g.B = B; g.ldb = ldb; g.strideB = strideB;
    g.C = C; g.ldc = ldc; g.strideC = strideC;
    const int rc = gemm_launch(g, static_cast<cudaStream_t>(stream));
    if (rc == HB_ERR_ALIGN) set_error("hb_dev_dgemm: operands need 16-byte aligned bases and even leading dimensions");
    return rc;
}
int hb_dev_dgemm_scatter(int M, int N, int K, int batch, const double* A, long long lda, long long strideA,
                         int a_mmajor, const double* B, long long ldb, long long strideB, double* const* blocks,
                         int n_blocks, int block_rows, long long ldc, long long strideC, void* stream) {
    if (n_blocks < 1 || n_blocks > 16 || block_rows < 1 || (long long)n_blocks * block_rows < M) {
        set_error("hb_dev_dgemm_scatter: need 1..16 row blocks covering M");
        return HB_ERR_INVALID;
    }
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.batch = batch;
    g.A = A; g.lda = lda; g.strideA = strideA; g.a_mmajor = a_mmajor != 0;
    g.B = B; g.ldb = ldb; g.strideB = strideB;
    g.C = blocks[0]; g.ldc = ldc; g.strideC = strideC;
    g.block_rows = block_rows;
    for (int i = 0; i < n_blocks; i++) g.blocks[i] = blocks[i];
    return gemm_launch(g, static_cast<cudaStream_t>(stream));
}
int hb_dev_dgemm_ex(const hb_gemm_desc* q, void* stream) {
    if (!q) return invalid(__func__);
    GemmArgs g;
    g.M = q->M; g.N = q->N; g.K = q->K; g.batch = q->batch; g.batch_lo = q->batch_lo > 0 ? q->batch_lo : 1;
    g.A = q->A; g.lda = q->lda; g.strideA = q->strideA; g.strideA_lo = q->strideA_lo; g.a_mmajor = q->a_mmajor != 0;
    g.B = q->B; g.ldb = q->ldb; g.strideB = q->strideB; g.strideB_lo = q->strideB_lo;
    g.C = q->C; g.ldc = q->ldc; g.strideC = q->strideC; g.strideC_lo = q->strideC_lo;
    if (q->n_blocks > 0) {
        const long long last = (long long)q->row_offset + (long long)(g.batch_lo - 1) * q->row_offset_lo + q->M - 1;
        if (q->n_blocks > 16 || q->block_rows < 1 || !q->blocks || q->row_offset < 0 || q->row_offset_lo < 0 ||
            last >= (long long)q->n_blocks * q->block_rows) {
            set_error("hb_dev_dgemm_ex: need 1..16 row blocks covering destination rows 0..%lld", last);
            return HB_ERR_INVALID;
        }
        g.block_rows = q->block_rows; g.row_offset = q->row_offset; g.row_offset_lo = q->row_offset_lo;
        g.C = q->blocks[0];
        for (int i = 0; i < q->n_blocks; i++) g.blocks[i] = q->blocks[i];
    }
    if (q->signal_state) {
        if (q->n_signal < 0 || q->n_signal > 16 || (q->n_signal > 0 && !q->signal_flags)) {
            set_error("hb_dev_dgemm_ex: 0..16 signal flags");
            return HB_ERR_INVALID;
        }
        g.signal_state = q->signal_state; g.n_signal = q->n_signal;
        for (int i = 0; i < q->n_signal; i++) g.signal_flags[i] = q->signal_flags[i];
    }
    if (q->wait_flags) {
        if (q->n_wait < 0 || !q->wait_epoch) { set_error("hb_dev_dgemm_ex: wait_flags needs wait_epoch"); return HB_ERR_INVALID; }
        g.wait_flags = q->wait_flags; g.n_wait = q->n_wait; g.wait_epoch = q->wait_epoch;
    }
    return gemm_launch(g, static_cast<cudaStream_t>(stream));
}
int hb_dev_parity_pass(int merge, int d, const int* n, const int* split, const long long* nat_stride,
                       const long long* srt_stride, const long long* srt_par, long long batch, long long nat_bstride,
                       long long srt_bstride, int nat_pad_last, const double* in, double* out, void* stream) {
    HB_TRY(check_device());
    if (d < 1 || d > HB_MAX_DIM || !n || !split || !nat_stride || !srt_stride || !srt_par) return invalid(__func__);
    ParityNDArgs a;
    a.d = d;
    for (int k = 0; k < d; k++) {
        a.n[k] = n[k]; a.nh[k] = (n[k] + 1) / 2; a.split[k] = split[k] != 0;
        a.nat_stride[k] = nat_stride[k]; a.srt_stride[k] = srt_stride[k]; a.srt_par[k] = srt_par[k];
    }
    a.batch = batch; a.nat_bstride = nat_bstride; a.srt_bstride = srt_bstride; a.nat_pad_last = nat_pad_last;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return merge ? launch_nd_parity_merge(a, in, out, st) : launch_nd_parity_split(a, in, out, st);
}
int hb_plan_parity_table(hb_plan* plan, int axis, int kind, const double** ptr, long long* rows, long long* cols,
                         long long* ld, long long* stride, void* stream) {
    if (axis < 0 || axis >= plan->dim) return invalid(__func__);
    Axis& A = plan->ax[axis];
    if (!A.symmetric_nodes) { set_error("hb_plan_parity_table: axis %d is not a bit-symmetric Hermite axis", axis); return HB_ERR_INVALID; }
    HB_TRY(ensure_combined_parity_tables(A, static_cast<cudaStream_t>(stream)));
    const int nh = (A.n + 1) / 2, Pe = (A.P + 1) / 2;
    const long long nhp = even_up(nh), Pep = even_up(Pe);
    switch (kind) {
        case 0: *ptr = A.cW.d(); *rows = nh; *cols = Pe; *ld = Pep; *stride = nh * Pep; break;
        case 1: *ptr = A.cWT.d(); *rows = Pe; *cols = nh; *ld = nhp; *stride = Pe * nhp; break;
        case 2: *ptr = A.cPhi.d(); *rows = nh; *cols = Pe; *ld = Pep; *stride = nh * Pep; break;
        case 3: *ptr = A.cPhiT.d(); *rows = Pe; *cols = nh; *ld = nhp; *stride = Pe * nhp; break;
        default: return invalid(__func__);
    }
    return HB_OK;
}
int hb_dev_matvec(int M, int N, const double* A, long long lda, const double* x, double* y, void* stream) {
    HB_TRY(check_device());
    if (M < 0 || N < 0 || lda < N) { set_error("hb_dev_matvec: need lda >= N >= 0"); return HB_ERR_INVALID; }
    return launch_matvec(M, N, A, lda, x, y, static_cast<cudaStream_t>(stream));
}
int hb_dev_varfd(int dim, int degree, int direction, const double* var, long long ldv, long long n_rows,
                 size_t n_polys, int do_fourier, const char* index_set, double* out, long long ldo, void* stream) {
    HB_TRY(check_device());
    if (var == out) { set_error("hb_dev_varfd: in-place operation is not supported"); return HB_ERR_INVALID; }
    std::vector<int> src;
    std::vector<double> fac;
    HB_TRY(varfd_column_map(dim, degree, direction, n_polys, do_fourier, index_set, src, fac));
    const int n = (int)n_polys;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the column map is small (12 bytes per column): stream-ordered scratch, freed behind the kernel
    int* d_src = nullptr;
    double* d_fac = nullptr;
    HB_CUDA(cudaMallocAsync(&d_src, (size_t)n * 4, st));
    HB_CUDA(cudaMallocAsync(&d_fac, (size_t)n * 8, st));
    HB_CUDA(cudaMemcpyAsync(d_src, src.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    HB_CUDA(cudaMemcpyAsync(d_fac, fac.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    const int rc = launch_varfd(var, ldv, d_src, d_fac, out, ldo, (int)n_rows, n, st);
    cudaFreeAsync(d_src, st);
    cudaFreeAsync(d_fac, st);
    HB_CUDA(cudaStreamSynchronize(st));   // src / fac are host temporaries of this call
    return rc;
}
int hb_plan_table(const hb_plan* plan, int axis, int kind, const double** ptr, long long* rows, long long* cols,
                  long long* ld) {
    if (axis < 0 || axis >= plan->dim) return invalid(__func__);
    const Axis& A = plan->ax[axis];
    switch (kind) {
        case 0: *ptr = A.plain.d(); *rows = A.n; *cols = A.P; *ld = A.Pp; break;   // phi_k(x_i)        [i][k]
        case 1: *ptr = A.wgt.d(); *rows = A.n; *cols = A.P; *ld = A.Pp; break;     // w_i phi_k(x_i)    [i][k]
        case 2: *ptr = A.plainT.d(); *rows = A.P; *cols = A.n; *ld = A.np; break;  // phi_k(x_i)        [k][i]
        case 3: *ptr = A.wgtT.d(); *rows = A.P; *cols = A.n; *ld = A.np; break;    // w_i phi_k(x_i)    [k][i]
        default: return invalid(__func__);
    }
    return HB_OK;
}
int hb_plan_varf_info(const hb_plan* plan, int* n_kept, int* path) {
    *n_kept = plan->last_n_kept;
    *path = plan->last_path;
    return HB_OK;
}

}  // extern "C"
