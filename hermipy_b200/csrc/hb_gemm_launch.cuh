// Host-side launch template of the DMMA/TMA GEMM, shared by the translation units that instantiate the
// kernel (hb_gemm_inst_*.cu: one per (epilogue, A layout) so that they compile in parallel).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "hb_gemm.cuh"
#include "hb_internal.h"
#include "hb_plan.h"

namespace hb {

// current device (index clamped to 63 for the per-device bit masks) and its SM count, cached per device
struct DeviceInfo { int index, sms; };
inline DeviceInfo current_device_info() {
    static std::atomic<int> sms[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    if (dev > 63) dev = 63;
    int n = sms[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        sms[dev].store(n, std::memory_order_relaxed);
    }
    return DeviceInfo{dev, n};
}

// rank-3 FP64 tensor map: dims {d0 (contiguous), d1 (pitch ld elements), d2 (batch, stride elements)}
int make_map(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld, uint64_t stride,
             uint32_t b0, uint32_t b1);
// rank-4: dims {d0, d1, n_lo, n_hi} with strides {ld, stride_lo, stride_hi} (the two-level batch of the GEMM)
int make_map4(CUtensorMap* m, const double* base, uint64_t d0, uint64_t d1, uint64_t n_lo, uint64_t n_hi, uint64_t ld,
              uint64_t stride_lo, uint64_t stride_hi, uint32_t b0, uint32_t b1);

// Tile configurations {BM, BN, warps along M, warps along N, pipeline stages, resident CTAs per SM}.  The number
// of consumer warps is a multiple of 4 everywhere: FP64 DMMA issues at one instruction per 16 clocks per SM
// sub-partition, so 6 consumer warps (2+2+1+1 over the four sub-partitions) cap the tile at 75 % of the rate.
// One warp per sub-partition cannot keep that cadence alone either (19.3 clocks per DMMA measured in the main loop
// of configuration 5: 83 %), two can (96 % for configuration 1): configuration 8 is configuration 5's tile on
// 2 x 4 consumer warps, the 26 column slots dealt 7,7,6,6 so that each sub-partition owns 39 of the 156 MMAs;
// configuration 9 does the same for the 32 x 208 tile when its tiles fit one wave at one CTA per SM.
// Configurations 10 / 11 (112 x 128, 128 x 112) exist for the per-rank GEMMs of the 8-GPU sharded transform, whose
// 8192 rows (or columns) x 2 parities make 128 tiles of 128 x 128 -- 20 of the 148 SMs idle -- but 148 tiles of 112.
#define HB_GEMM_CONFIGS(X, AM, EPI)   \
    X(0, 128, 128, 4, 2, 4, AM, 1, EPI) \
    X(1, 64, 64, 2, 2, 6, AM, 2, EPI)   \
    X(2, 64, 128, 2, 2, 4, AM, 2, EPI)  \
    X(3, 64, 208, 4, 2, 4, AM, 1, EPI)  \
    X(4, 32, 208, 2, 2, 3, AM, 2, EPI)  \
    X(5, 48, 208, 2, 2, 4, AM, 1, EPI)  \
    X(6, 96, 128, 4, 2, 4, AM, 1, EPI)  \
    X(7, 32, 128, 2, 2, 4, AM, 2, EPI)  \
    X(8, 48, 208, 2, 4, 4, AM, 1, EPI)  \
    X(9, 32, 208, 2, 4, 4, AM, 1, EPI)  \
    X(10, 112, 128, 2, 4, 4, AM, 1, EPI) \
    X(11, 128, 112, 4, 2, 4, AM, 1, EPI)

// The bulk-store epilogue (EPI_BULK) carries a staging sub-tile per consumer warp in shared memory: same tiles, fewer
// pipeline stages / resident CTAs where the two would not fit 227 KB.  Configuration ids as above (the others map to
// the nearest of these in gemm_launch).
#define HB_GEMM_BULK_CONFIGS(X, AM, EPI) \
    X(1, 64, 64, 2, 2, 4, AM, 2, EPI)    \
    X(2, 64, 128, 2, 2, 4, AM, 1, EPI)   \
    X(6, 96, 128, 4, 2, 4, AM, 1, EPI)   \
    X(7, 32, 128, 2, 2, 3, AM, 2, EPI)

template <int BM, int BN, int WMc, int WNc, int ST, bool AM, int MINB, int EPI>
static int launch_cfg(const GemmArgs& a, const GemmParams& p_in, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    // An operand that does not vary with a batch level (stride 0, or a level with one entry) gets an extent of 1 at
    // that level of its map and the kernel multiplies the coordinate by 0.
    const bool a_hi = a.batch > 1 && a.strideA != 0, b_hi = a.batch > 1 && a.strideB != 0;
    const bool a_lo = a.batch_lo > 1 && a.strideA_lo != 0, b_lo = a.batch_lo > 1 && a.strideB_lo != 0;
    const uint64_t dummyA = (uint64_t)a.lda * (AM ? a.K : a.M), dummyB = (uint64_t)a.ldb * a.K;
    int rc;
    if (AM)
        rc = make_map4(&tmA, a.A, a.M, a.K, a_lo ? a.batch_lo : 1, a_hi ? a.batch : 1, a.lda,
                       a_lo ? (uint64_t)a.strideA_lo : dummyA, a_hi ? (uint64_t)a.strideA : dummyA, 16, GEMM_BK);
    else
        rc = make_map4(&tmA, a.A, a.K, a.M, a_lo ? a.batch_lo : 1, a_hi ? a.batch : 1, a.lda,
                       a_lo ? (uint64_t)a.strideA_lo : dummyA, a_hi ? (uint64_t)a.strideA : dummyA, GEMM_BK, BM);
    if (rc) return rc;
    rc = make_map4(&tmB, a.B, a.N, a.K, b_lo ? a.batch_lo : 1, b_hi ? a.batch : 1, a.ldb,
                   b_lo ? (uint64_t)a.strideB_lo : dummyB, b_hi ? (uint64_t)a.strideB : dummyB, 16, GEMM_BK);
    if (rc) return rc;
    GemmParams p = p_in;
    p.tiles_m = (a.M + BM - 1) / BM;
    p.tiles_n = (a.N + BN - 1) / BN;
    p.group_m = a.group_m > 0 ? a.group_m : 1;
    p.a_lo_mul = a_lo; p.a_hi_mul = a_hi; p.b_lo_mul = b_lo; p.b_hi_mul = b_hi;
    const long long n_tiles = (long long)p.tiles_m * p.tiles_n * a.batch * a.batch_lo;
    if (n_tiles > 0x7fffffffLL - 65536) return HB_ERR_INVALID;   // the kernel's tile counter steps by the grid size in int
    p.n_tiles = (int)n_tiles;
    const DeviceInfo dev = current_device_info();
    const long long n_cta = std::min<long long>(n_tiles, (long long)dev.sms * MINB);   // persistent: one wave
    auto kern = dgemm_dmma_tma_kernel<BM, BN, WMc, WNc, ST, AM, MINB, EPI>;
    constexpr int smem = EPI == EPI_BULK ? gemm_smem_bytes_bulk<BM, BN, ST, WMc, WNc>() : gemm_smem_bytes<BM, BN, ST>();
    // the > 48 KB dynamic shared memory opt-in is a per-device property of the function: one bit per device
    static std::atomic<unsigned long long> attr_set{0};
    if (!(attr_set.load(std::memory_order_relaxed) >> dev.index & 1)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return HB_ERR_CUDA;
        attr_set.fetch_or(1ull << dev.index, std::memory_order_relaxed);
    }
    // Programmatic dependent launch (HB_GEMM_PDL=0 turns it off): the CTA launch and barrier setup of this kernel overlap
    // the tail of its predecessor in the stream (every warp blocks in griddepcontrol.wait before touching operands; the
    // predecessor triggers after the main loop of its last tile; a predecessor that never triggers -- any non-GEMM
    // kernel -- releases its dependents when it completes, i.e. the launch is an ordinary one).  Measured on B200 with
    // the persistent CTAs of round 2, graph replay: C2 step 159.5 -> 156.2 us, C4 1.1385 -> 1.1368 ms, C5 and the
    // 8-GPU sharded C4 unchanged (round 1, before the CTAs were persistent, it made the C2 step slower).
    static const bool pdl = [] { const char* e = getenv("HB_GEMM_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_cta);
    cfg.blockDim = dim3((WMc * WNc + 4) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p) != cudaSuccess) return launch_status();
    return launch_status();
}

#define HB_GEMM_CASE(ID, BM, BN, WM, WN, ST, AM, MINB, EPI) \
    case ID: return launch_cfg<BM, BN, WM, WN, ST, AM, MINB, EPI>(a, p, stream);

// one dispatcher per (epilogue, A layout); each lives in its own translation unit
int gemm_dispatch_store_k(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream);  // A K-contiguous
int gemm_dispatch_store_m(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream);  // A M-contiguous
int gemm_dispatch_lut_k(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream);
int gemm_dispatch_bulk_k(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream);   // cfg in {1, 2, 6, 7}

}  // namespace hb
