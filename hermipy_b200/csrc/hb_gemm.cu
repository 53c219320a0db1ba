// Host launcher for the DMMA/TMA GEMM: builds the TMA tensor maps (cuTensorMapEncodeTiled obtained
// through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda) and picks
// a tile configuration.
#include "hb_gemm.cuh"
#include "hb_internal.h"
#include "hb_plan.h"

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <mutex>

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
static int make_map(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2,
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

template <int BM, int BN, int WMc, int WNc, int ST, bool AM, int MINB>
static int launch_cfg(const GemmArgs& a, const GemmParams& p_in, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    // A batch stride of 0 means "shared operand": the map gets a batch extent of 1 and the kernel
    // multiplies the batch coordinate by a_bmul/b_bmul = 0.
    const bool a_batched = a.batch > 1 && a.strideA != 0, b_batched = a.batch > 1 && a.strideB != 0;
    const uint64_t strideA = a_batched ? (uint64_t)a.strideA : (uint64_t)a.lda * (AM ? a.K : a.M);
    const uint64_t strideB = b_batched ? (uint64_t)a.strideB : (uint64_t)a.ldb * a.K;
    int rc;
    if (AM)
        rc = make_map(&tmA, a.A, a.M, a.K, a_batched ? a.batch : 1, a.lda, strideA, 16, GEMM_BK);
    else
        rc = make_map(&tmA, a.A, a.K, a.M, a_batched ? a.batch : 1, a.lda, strideA, GEMM_BK, BM);
    if (rc) return rc;
    rc = make_map(&tmB, a.B, a.N, a.K, b_batched ? a.batch : 1, a.ldb, strideB, 16, GEMM_BK);
    if (rc) return rc;
    GemmParams p = p_in;
    p.tiles_m = (a.M + BM - 1) / BM;
    p.tiles_n = (a.N + BN - 1) / BN;
    p.a_bmul = a_batched ? 1 : 0;
    p.b_bmul = b_batched ? 1 : 0;
    const long long n_cta = (long long)p.tiles_m * p.tiles_n * a.batch;
    if (n_cta > 0x7fffffffLL) return HB_ERR_INVALID;
    auto kern = dgemm_dmma_tma_kernel<BM, BN, WMc, WNc, ST, AM, MINB>;
    constexpr int smem = gemm_smem_bytes<BM, BN, ST>();
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return HB_ERR_CUDA;
        attr_set = true;
    }
    kern<<<(unsigned)n_cta, (WMc * WNc + 1) * 32, smem, stream>>>(tmA, tmB, p);
    return launch_status();
}

// Tile configurations: {BM, BN, resident CTAs per SM}.  The launcher picks the one with the smallest
// modelled time: small or oddly shaped problems (P = 201, n = 401, batch 17 ...) otherwise lose most of
// their time to tile quantisation and partial waves over the 148 SMs.
struct TileCfg { int bm, bn, ctas; };
static const TileCfg kCfgs[] = {{128, 128, 1}, {64, 64, 2}, {64, 128, 2}, {64, 208, 1}, {32, 208, 2},
                                {48, 208, 1},  {96, 128, 1}, {32, 128, 2}};
static const int kNumCfgs = 8;

static int pick_cfg(const GemmArgs& a) {
    const double sms = 148.0;
    const double ksteps = (a.K + 15) / 16;
    double best = 1e300;
    int best_i = 0;
    for (int i = 0; i < kNumCfgs; i++) {
        const TileCfg& c = kCfgs[i];
        const long long tiles = (long long)((a.M + c.bm - 1) / c.bm) * ((a.N + c.bn - 1) / c.bn) * a.batch;
        // DMMA-bound time of one tile in SM clocks (one DMMA.8x8x4 per 4 clocks per SM), plus a fixed
        // prologue/epilogue cost that co-resident CTAs overlap with each other's main loop.
        const double tile_t = (double)c.bm * c.bn * ksteps * 16.0 / 256.0 * 4.0;
        const double fixed = 5000.0 + 0.02 * c.bm * c.bn;
        const long long slots = (long long)sms * c.ctas;
        const long long full = tiles / slots, rem = tiles % slots;
        double t = full * (c.ctas * tile_t + (c.ctas > 1 ? 0.3 : 1.0) * fixed);
        if (rem) t += std::ceil((double)rem / sms) * tile_t + fixed;
        // small tiles re-read their operands more often (shared-memory and L2 traffic per flop)
        t *= 1.0 + 6.0 / c.bm + 6.0 / c.bn;
        if (t < best) { best = t; best_i = i; }
    }
    return best_i;
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return HB_OK;
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
        (reinterpret_cast<uintptr_t>(a.B) & 15) || (a.batch > 1 && ((a.strideA & 1) || (a.strideB & 1))) || a.K <= 0)
        return HB_ERR_ALIGN;
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch;
    p.C = a.C; p.ldc = a.ldc; p.strideC = a.strideC;
    p.epi = a.epi; p.lut = a.lut;
    p.P0 = a.P0; p.Pp0 = a.Pp0; p.P1 = a.P1; p.Pp1 = a.Pp1;
    p.row_begin = a.row_begin; p.row_end = a.row_end; p.tri_degree = a.tri_degree; p.symmetric = a.symmetric; p.lut_mode = a.lut_mode;
    { const char* dbg = getenv("HB_VARF_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    p.alpha = a.alpha;
    p.block_rows = a.block_rows;
    for (int i = 0; i < GEMM_MAX_BLOCKS; i++) p.blocks[i] = a.blocks[i];
    if (a.block_rows > 0 && (a.M + a.block_rows - 1) / a.block_rows > GEMM_MAX_BLOCKS) return HB_ERR_INVALID;
    int cfg = a.cfg;
    if (cfg < 0) {
        static const char* env = getenv("HB_GEMM_CFG_FIXED");      // developer override, read once
        const char* dyn = getenv("HB_GEMM_CFG");                   // developer sweep (tools/cfg_sweep.py)
        if (dyn && dyn[0] && dyn[0] != '-') cfg = atoi(dyn);
        else if (env && env[0]) cfg = atoi(env);
    }
    if (cfg < 0 || cfg >= kNumCfgs) cfg = pick_cfg(a);
    if (a.a_mmajor) {
        switch (cfg) {
            case 0: return launch_cfg<128, 128, 4, 2, 4, true, 1>(a, p, stream);
            case 1: return launch_cfg<64, 64, 2, 2, 6, true, 2>(a, p, stream);
            case 2: return launch_cfg<64, 128, 2, 2, 4, true, 2>(a, p, stream);
            case 3: return launch_cfg<64, 208, 4, 2, 4, true, 1>(a, p, stream);
            case 4: return launch_cfg<32, 208, 2, 2, 3, true, 2>(a, p, stream);
            case 5: return launch_cfg<48, 208, 3, 2, 4, true, 1>(a, p, stream);
            case 6: return launch_cfg<96, 128, 3, 2, 4, true, 1>(a, p, stream);
            default: return launch_cfg<32, 128, 2, 2, 4, true, 2>(a, p, stream);
        }
    } else {
        switch (cfg) {
            case 0: return launch_cfg<128, 128, 4, 2, 4, false, 1>(a, p, stream);
            case 1: return launch_cfg<64, 64, 2, 2, 6, false, 2>(a, p, stream);
            case 2: return launch_cfg<64, 128, 2, 2, 4, false, 2>(a, p, stream);
            case 3: return launch_cfg<64, 208, 4, 2, 4, false, 1>(a, p, stream);
            case 4: return launch_cfg<32, 208, 2, 2, 3, false, 2>(a, p, stream);
            case 5: return launch_cfg<48, 208, 3, 2, 4, false, 1>(a, p, stream);
            case 6: return launch_cfg<96, 128, 3, 2, 4, false, 1>(a, p, stream);
            default: return launch_cfg<32, 128, 2, 2, 4, false, 2>(a, p, stream);
        }
    }
}

}  // namespace hb
