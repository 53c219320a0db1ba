// Host launcher for the DMMA/TMA GEMM: builds the TMA tensor maps (cuTensorMapEncodeTiled obtained
// through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda) and picks
// a tile configuration.
#include "hb_gemm_launch.cuh"

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

namespace hb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// rank-3 FP64 map: dims {d0 (contiguous), d1 (pitch ld elements), d2 (batch, stride elements)}
int make_map(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2,
                    uint64_t ld, uint64_t stride, uint32_t b0, uint32_t b1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return HB_ERR_CUDA;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {ld * 8, stride * 8};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[hermite_b200] cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu ld %llu\n",
                (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                (unsigned long long)ld);
        return HB_ERR_CUDA;
    }
    return HB_OK;
}

// rank-4 FP64 map: dims {d0 (contiguous), d1 (pitch ld), n_lo (stride_lo), n_hi (stride_hi)}, strides in elements
int make_map4(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t n_lo, uint64_t n_hi, uint64_t ld,
              uint64_t stride_lo, uint64_t stride_hi, uint32_t b0, uint32_t b1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return HB_ERR_CUDA;
    cuuint64_t dims[4] = {d0, d1, n_lo, n_hi};
    cuuint64_t strides[3] = {ld * 8, stride_lo * 8, stride_hi * 8};
    cuuint32_t box[4] = {b0, b1, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu, ld %llu, strides %llu %llu", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)n_lo, (unsigned long long)n_hi,
                  (unsigned long long)ld, (unsigned long long)stride_lo, (unsigned long long)stride_hi);
        return HB_ERR_CUDA;
    }
    return HB_OK;
}

// Tile configurations: {BM, BN, resident CTAs per SM}.  The launcher picks the one with the smallest
// modelled time: small or oddly shaped problems (P = 201, n = 401, batch 17 ...) otherwise lose most of
// their time to tile quantisation and partial waves over the 148 SMs.
struct TileCfg { int bm, bn, ctas; };
static const int kNumCfgs = 12;
static const TileCfg kCfgs[] = {{128, 128, 1}, {64, 64, 2}, {64, 128, 2}, {64, 208, 1}, {32, 208, 2},
                                {48, 208, 1},  {96, 128, 1}, {32, 128, 2}, {48, 208, 1}, {32, 208, 1},
                                {112, 128, 1}, {128, 112, 1}};

// Modelled time (microseconds) of configuration i on the problem `a` (used to rank candidates; fitted before the
// CTAs became persistent, so the per-wave fixed cost is now pessimistic for multi-wave problems -- the first-call
// tuner, which measures, has the last word for everything below 150 us): full waves over 148 SMs x resident CTAs,
// a partial last wave, a per-configuration steady-state efficiency and a per-configuration fixed cost
// (prologue + epilogue, largely hidden between co-resident CTAs).  Efficiencies and fixed costs are a
// least-squares fit to tools/cfg_sweep.py on B200 (nine shapes from 0.5 GFLOP to 1.3 TFLOP, fit error <= 10 %
// except for partial second waves of the two-CTA configurations).
static double model_us(const GemmArgs& a, int i) {
    static const double kEff[kNumCfgs] = {1.005, 1.036, 0.956, 0.957, 0.926, 0.881, 0.994, 0.979, 0.96, 0.90, 1.0, 1.0};
    static const double kFixedUs[kNumCfgs] = {7.8, 11.7, 10.0, 6.9, 12.0, 7.9, 5.8, 10.8, 7.9, 7.0, 7.8, 7.8};
    const double sms = 148.0, clk_per_us = 1920.0;
    const double ksteps = (a.K + 15) / 16;
    const TileCfg& c = kCfgs[i];
    const long long tiles = (long long)((a.M + c.bm - 1) / c.bm) * ((a.N + c.bn - 1) / c.bn) * a.batch * a.batch_lo;
    // DMMA-bound time of one tile in SM clocks (one DMMA.8x8x4 per 4 clocks per SM)
    const double tile_t = (double)c.bm * c.bn * ksteps * 16.0 / 256.0 * 4.0;
    const long long slots = (long long)sms * c.ctas;
    const long long full = tiles / slots, rem = tiles % slots;
    double clk = full * c.ctas * tile_t;
    double n_fixed = full * (c.ctas > 1 ? 0.3 : 1.0);
    if (rem) { clk += std::ceil((double)rem / sms) * tile_t; n_fixed += 1.0; }
    return clk / clk_per_us / kEff[i] + n_fixed * kFixedUs[i];
}

static int dispatch(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream);

// Small problems (modelled below ~150 us) are dominated by tile quantisation over 148 SMs and by fixed costs,
// which the model gets wrong by 10-15 % in either direction; for those, the four best candidates of the model
// are timed once on the actual operands (first call per shape, outside graph capture) and the winner is cached.
// Every configuration accumulates each output over k in the same order, so the result does not depend on the
// choice.  Launches whose epilogue writes to peer memory are never replayed.
struct TuneKey {
    int M, N, K, batch, flags;
    bool operator<(const TuneKey& o) const {
        return std::tie(M, N, K, batch, flags) < std::tie(o.M, o.N, o.K, o.batch, o.flags);
    }
};
static std::map<TuneKey, int> g_tuned;
static std::mutex g_tune_mutex;

static int tuned_cfg(const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
    double t[kNumCfgs];
    int order[kNumCfgs];
    for (int i = 0; i < kNumCfgs; i++) { t[i] = model_us(a, i); order[i] = i; }
    std::sort(order, order + kNumCfgs, [&](int x, int y) { return t[x] < t[y]; });
    static const bool enabled = [] { const char* e = getenv("HB_GEMM_AUTOTUNE"); return !(e && e[0] == '0'); }();
    if (!enabled || t[order[0]] > 150.0 || a.block_rows > 0 || a.epi == EPI_VARF2D || a.signal_state) return order[0];
    // Timing replays the caller's GEMM ~19 times on the caller's operands: only safe when C aliases neither input
    // (an in-place product would be fed its own output on the second launch).  Overlapping calls take the model's pick.
    {
        const int nlo = a.batch_lo;
        auto extent = [nlo](const double* p, long long rows, long long ld, long long stride, int batch, long long stride_lo) {
            const long long n = (batch > 1 ? (long long)(batch - 1) * (stride > 0 ? stride : 0) : 0) +
                                (nlo > 1 ? (long long)(nlo - 1) * (stride_lo > 0 ? stride_lo : 0) : 0) + rows * ld;
            return std::make_pair(reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p + (n > 0 ? n : 0)));
        };
        const auto ra = extent(a.A, a.a_mmajor ? a.K : a.M, a.lda, a.strideA, a.batch, a.strideA_lo);
        const auto rb = extent(a.B, a.K, a.ldb, a.strideB, a.batch, a.strideB_lo);
        auto rc = extent(a.C, a.M, a.ldc > 0 ? a.ldc : a.N, a.strideC, a.batch, a.strideC_lo);
        if (a.epi == EPI_LUT)
            rc.second = rc.first + (uintptr_t)8 * ((a.batch - 1) * (a.strideC > 0 ? a.strideC : 0) +
                                                   (long long)a.M * a.N * a.batch_lo * 2);
        auto overlap = [](std::pair<uintptr_t, uintptr_t> x, std::pair<uintptr_t, uintptr_t> y) {
            return x.first < y.second && y.first < x.second;
        };
        if (overlap(rc, ra) || overlap(rc, rb)) return order[0];
    }
    const TuneKey key{a.M, a.N, a.K, a.batch,
                      (a.a_mmajor ? 1 : 0) | (a.epi << 1) | ((a.strideA == 0) << 4) | ((a.strideB == 0) << 5) |
                          ((a.col_stride > 1) << 6) | (a.batch_lo << 7)};
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    auto it = g_tuned.find(key);
    if (it != g_tuned.end()) return it->second;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return order[0];
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return order[0];
    int best = order[0];
    float best_ms = 1e30f;
    static const bool verbose = [] { const char* e = getenv("HB_GEMM_TUNE_DEBUG"); return e && e[0] == '1'; }();
    for (int c = 0; c < 4; c++) {   // four: configurations 5 and 8 share a tile and sit next to each other in the model
        const int cfg = order[c];
        if (dispatch(cfg, a, p, stream) != HB_OK) continue;      // warm-up (and first-use attribute setting)
        // events resolve ~1 us, the candidates are often closer than that: time runs of 6 launches
        float ms_min = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0, stream);
            for (int i = 0; i < 6; i++) dispatch(cfg, a, p, stream);
            cudaEventRecord(e1, stream);
            if (cudaEventSynchronize(e1) != cudaSuccess) break;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            ms_min = std::min(ms_min, ms / 6);
        }
        // a later candidate must beat the model's favourite by more than 1 %: the run of identical launches timed
        // here cannot separate closer ones (C2 forward stage 2: 35.7 us for configurations 1 and 4 here, 39.3 /
        // 37.3 us inside the transform), and a pick that flips with timing noise would make run times irreproducible
        if (ms_min < 0.99f * best_ms) { best_ms = ms_min; best = cfg; }
        if (verbose)
            fprintf(stderr, "[hb tune] M %d N %d K %d batch %d epi %d: cfg %d modelled %.1f us, measured %.1f us\n", a.M, a.N,
                    a.K, a.batch, a.epi, cfg, t[cfg], ms_min * 1e3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    g_tuned[key] = best;
    return best;
}

static long long* g_trace = nullptr;
void gemm_set_trace(long long* dev) { g_trace = dev; }
long long* gemm_get_trace() { return g_trace; }

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0 || a.batch_lo <= 0) return HB_OK;
    if (a.K <= 0) { set_error("gemm: K = %d (need K >= 1)", a.K); return HB_ERR_INVALID; }
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || (a.batch > 1 && ((a.strideA & 1) || (a.strideB & 1))) ||
        (a.batch_lo > 1 && ((a.strideA_lo & 1) || (a.strideB_lo & 1)))) {
        set_error("gemm: operands need 16-byte aligned bases and even leading dimensions / batch strides (lda %lld, "
                  "ldb %lld, strideA %lld, strideB %lld)", a.lda, a.ldb, a.strideA, a.strideB);
        return HB_ERR_ALIGN;
    }
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch;
    p.C = a.C; p.ldc = a.ldc; p.strideC = a.strideC;
    p.epi = a.epi; p.lut = a.lut;
    p.P0 = a.P0; p.Pp0 = a.Pp0; p.P1 = a.P1; p.Pp1 = a.Pp1;
    p.row_begin = a.row_begin; p.row_end = a.row_end; p.symmetric = a.symmetric;
    p.blkpos = a.blkpos; p.blkperm = a.blkperm; p.blkperm2 = a.blkperm2; p.blkrange = a.blkrange;
    p.b_koff_odd = 0; p.nb_even0 = 0;
    p.trace = g_trace;
    { const char* dbg = getenv("HB_VARF_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    p.alpha = a.alpha;
    p.block_rows = a.block_rows;
    p.col_stride = a.col_stride > 0 ? a.col_stride : 1; p.col_offset = a.col_offset;
    p.row_stride = a.row_stride > 0 ? a.row_stride : 1; p.row_offset = a.row_offset;
    p.batch_lo = a.batch_lo; p.strideC_lo = a.strideC_lo;
    p.row_offset_lo = a.row_offset_lo; p.col_offset_lo = a.col_offset_lo;
    for (int i = 0; i < GEMM_MAX_BLOCKS; i++) p.blocks[i] = a.blocks[i];
    p.wait_flags = a.wait_flags; p.wait_epoch = a.wait_epoch; p.n_wait = a.wait_flags ? a.n_wait : 0;
    p.signal_state = a.signal_state; p.n_signal = a.signal_state ? a.n_signal : 0;
    for (int i = 0; i < GEMM_MAX_BLOCKS; i++) p.signal_flags[i] = a.signal_flags[i];
    if (a.block_rows > 0 && (a.M + a.block_rows - 1) / a.block_rows > GEMM_MAX_BLOCKS) {
        set_error("gemm: at most %d destination row blocks", GEMM_MAX_BLOCKS);
        return HB_ERR_INVALID;
    }
    int cfg = a.cfg;
    if (cfg < 0) {
        static const char* env = getenv("HB_GEMM_CFG_FIXED");      // developer override, read once
        const char* dyn = getenv("HB_GEMM_CFG");                   // developer sweep (tools/cfg_sweep.py)
        if (dyn && dyn[0] && dyn[0] != '-') cfg = atoi(dyn);
        else if (env && env[0]) cfg = atoi(env);
    }
    // Row-block scatter (the destinations may be peer GPUs): staged bulk-store epilogue when the layout allows
    // 16-byte runs (HB_GEMM_BULK=0 keeps the register epilogue).  Small tiles win here -- several rounds per SM, so
    // that one tile's copies drain over NVLink under the next tile's main loop (tools/dist_variants.py).
    if (a.block_rows > 0 && a.epi == EPI_STORE && !a.a_mmajor && p.col_stride == 1 && (a.N & 1) == 0 &&
        (a.ldc & 1) == 0 && (a.strideC & 1) == 0 && (a.strideC_lo & 1) == 0) {
        const char* off = getenv("HB_GEMM_BULK");             // read per call: tests compare the two epilogues
        bool aligned = !(off && off[0] == '0');
        for (int i = 0; i < GEMM_MAX_BLOCKS && aligned; i++)
            aligned = (reinterpret_cast<uintptr_t>(a.blocks[i]) & 15) == 0;
        if (aligned) {
            static const int kBulkOf[kNumCfgs] = {6, 1, 2, 2, 7, 2, 6, 7, 2, 7, 6, 6};
            const char* pick = getenv("HB_GEMM_BULK_CFG");
            int bc = (cfg >= 0 && cfg < kNumCfgs) ? kBulkOf[cfg] : 1;
            if (cfg < 0 && pick && pick[0]) bc = kBulkOf[std::max(0, std::min(kNumCfgs - 1, atoi(pick)))];
            p.epi = EPI_BULK;
            return gemm_dispatch_bulk_k(bc, a, p, stream);
        }
    }
    if (cfg < 0 || cfg >= kNumCfgs) cfg = tuned_cfg(a, p, stream);
    return dispatch(cfg, a, p, stream);
}

static int dispatch(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
    if (a.epi == EPI_VARF2D) {
        set_error("gemm: the varf scatter epilogue belongs to the persistent kernel (varf_stage2_launch)");
        return HB_ERR_INVALID;
    }
    if (a.epi == EPI_LUT) {
        if (a.a_mmajor) { set_error("gemm: the look-up-table epilogue needs a K-contiguous A operand"); return HB_ERR_INVALID; }
        return gemm_dispatch_lut_k(cfg, a, p, stream);
    }
    return a.a_mmajor ? gemm_dispatch_store_m(cfg, a, p, stream) : gemm_dispatch_store_k(cfg, a, p, stream);
}

}  // namespace hb
