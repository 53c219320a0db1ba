// FP64 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring
// (mbarrier full/empty pipeline, one producer warp) -> register fragments -> mma.sync m8n8k4 (DMMA).
//
// This single kernel is the contraction engine behind every dense step of the hot path:
//   * per-axis Hermite transform stages  (reference: the naive loop cpp/src/hermite/transform.cpp:115-146,
//     here sum-factorised into one GEMM per axis),
//   * the two-stage varf / triple-product contraction (reference: cpp/src/hermite/varf.cpp:236-262).
//
// C[b] (M x N) = A[b] (M x K) * B[b] (K x N), FP64, row-major C.
//   A layout:  A_MMAJOR = false : A stored [M][K] (K contiguous)   -- grid values / coefficients
//              A_MMAJOR = true  : A stored [K][M] (M contiguous)   -- triple-product / Gram tables
//   B layout:  always [K][N] (N contiguous).
// Leading dimensions must be even (16-byte TMA stride rule); the host pads or repacks otherwise.
//
// Shared-memory fragment addressing.  DMMA 8x8x4 wants 32-byte runs per row from 4 rows per half-warp,
// which hits 2-way bank conflicts under the plain 128B TMA swizzle.  Row/column order inside an MMA is
// free (the epilogue knows the permutation), so fragment row g of a K-contiguous tile maps to tile row
// rowmap(g) = 2*(g&3) + (g>>2), and fragment column g of an N-contiguous tile maps to column
// colmap(g) = (g&1) + 8*((g>>1)&1) + 2*(g>>2) inside a 16-wide group: both are conflict-free with the
// hardware swizzle (chunk' = chunk ^ (row & 7)).
#pragma once
#include "hb_ptx.cuh"

namespace hb {

enum EpilogueMode : int {
    EPI_STORE = 0,    // C[b][row*ldc + col] = v
    EPI_LUT = 1,      // out[b*strideC + lut[row*N + col]] = v   (lut < 0: dropped)
    EPI_VARF2D = 2,   // row=(m0,n0), col=(m1,n1): out[(lut[m0*P+m1]-row_begin)*ldc + lut[n0*P+n1]] = v
};

constexpr int GEMM_MAX_BLOCKS = 16;

struct GemmParams {
    int M, N, K, batch;
    double* C;
    long long ldc;      // row pitch of C (EPI_STORE) / of the output matrix (EPI_VARF2D)
    long long strideC;  // batch stride of C / of the scattered output (EPI_LUT)
    int epi;
    const int* lut;     // EPI_LUT: M*N entries; EPI_VARF2D: P0*P1 entries (lex (m0,m1) -> position or -1)
    int P0, Pp0;        // EPI_VARF2D: GEMM row = m0*Pp0 + n0 (axis-0 pair, n0 < P0 valid)
    int P1, Pp1;        // EPI_VARF2D: GEMM col = m1*Pp1 + n1 (axis-1 pair, n1 < P1 valid)
    int row_begin, row_end;  // EPI_VARF2D: owned output rows [row_begin,row_end) (row sharding)
    int tri_degree;     // EPI_VARF2D: >=0 -> skip tiles with m0+m1 > tri_degree (triangle-like sets)
    double alpha;       // result scale
    int tiles_m, tiles_n;    // grid is linear: blockIdx.x = (b*tiles_m + tile_m)*tiles_n + tile_n
    int a_bmul, b_bmul;      // 0: operand shared by all batch entries, 1: batched
    // EPI_STORE with row blocks: output row r goes to blocks[r / block_rows] + (r % block_rows)*ldc.
    // The block pointers may be peer-GPU addresses (NVLink P2P): this is how a GEMM's epilogue performs
    // the all-to-all of the sharded transform.
    int block_rows;          // 0 = off
    double* blocks[GEMM_MAX_BLOCKS];
};

constexpr int GEMM_BK = 16;  // doubles per k-tile = one 128-byte swizzle span

template <int BM, int BN, int STAGES>
constexpr int gemm_smem_bytes() {
    return STAGES * (BM + BN) * GEMM_BK * 8 + 1024 /*alignment slack*/ + 2 * STAGES * 8;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool A_MMAJOR, int MIN_CTAS>
__global__ void __launch_bounds__((WARPS_M* WARPS_N + 1) * 32, MIN_CTAS)
    dgemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int NCW = WARPS_M * WARPS_N;  // consumer warps
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MI = WM / 8, NJ = WN / 8;
    static_assert(WN % 8 == 0 && WM % 8 == 0 && BN % 16 == 0 && (!A_MMAJOR || BM % 16 == 0), "tile granularity");
    constexpr int A_BYTES = BM * GEMM_BK * 8, B_BYTES = BN * GEMM_BK * 8;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES]
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int tile_n = blockIdx.x % p.tiles_n;
    const int tile_m = (blockIdx.x / p.tiles_n) % p.tiles_m;
    const int b = blockIdx.x / (p.tiles_n * p.tiles_m);
    const int m_base = tile_m * BM, n_base = tile_n * BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (p.epi == EPI_VARF2D) {
        // Tile-level skipping.  GEMM row = (m0, n0), column = (m1, n1); the entry lands at output row
        // lut[m0,m1], column lut[n0,n1].
        const int m_last = min(m_base + BM, p.M) - 1, n_last = min(n_base + BN, p.N) - 1;
        const int m0_lo = m_base / p.Pp0, m0_hi = m_last / p.Pp0;
        const int m1_lo = n_base / p.Pp1, m1_hi = n_last / p.Pp1;
        if (p.tri_degree >= 0) {
            // total-degree sets keep only m0+m1 <= degree and n0+n1 <= degree: the smallest values decide.
            if (m0_lo + m1_lo > p.tri_degree) return;
            const int n0_lo = (m0_lo == m0_hi) ? m_base - m0_lo * p.Pp0 : 0;
            const int n1_lo = (m1_lo == m1_hi) ? n_base - m1_lo * p.Pp1 : 0;
            if (n0_lo + n1_lo > p.tri_degree) return;
        }
        // row sharding: skip the tile unless some (m0, m1) in it maps to an owned output row
        bool owned = false;
        for (int a = m0_lo; a <= m0_hi && !owned; a++)
            for (int c = m1_lo; c <= m1_hi; c++) {
                const int r = (a < p.P0 && c < p.P1) ? p.lut[a * p.P1 + c] : -1;
                if (r >= p.row_begin && r < p.row_end) { owned = true; break; }
            }
        if (!owned) return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (STAGES + s), NCW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int KT = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == NCW) {
        // ---------------- TMA producer (one elected lane) ----------------
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            for (int kt = 0; kt < KT; kt++) {
                const int s = kt % STAGES;
                if (kt >= STAGES) mbar_wait(bar_base + 8 * (STAGES + s), ((kt / STAGES) - 1) & 1);
                const uint32_t full = bar_base + 8 * s;
                const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + A_BYTES;
                mbar_arrive_expect_tx(full, STAGE_BYTES);
                const int k0 = kt * GEMM_BK;
                if (A_MMAJOR) {
#pragma unroll
                    for (int g = 0; g < BM / 16; g++)
                        tma_load_3d(sA + g * 2048, &tmA, full, m_base + 16 * g, k0, b * p.a_bmul);
                } else {
                    tma_load_3d(sA, &tmA, full, k0, m_base, b * p.a_bmul);
                }
#pragma unroll
                for (int g = 0; g < BN / 16; g++)
                    tma_load_3d(sB + g * 2048, &tmB, full, n_base + 16 * g, k0, b * p.b_bmul);
            }
        }
        return;
    }

    // ---------------- consumers: DMMA ----------------
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int g = lane >> 2, t = lane & 3;
    const int rm = 2 * (g & 3) + (g >> 2);                          // rowmap(g)
    const int cm = (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2);     // colmap(g)

    // Byte offsets of this lane's fragment elements inside a stage, for k-step s = 0:
    //   K-contiguous tile : row*128 + (((2s + (t>>1)) ^ (row&7)) << 4) + (t&1)*8
    //   N-contiguous tile : grp*2048 + (4s+t)*128 + (((n_in>>1) ^ ((4s+t)&7)) << 4) + (n_in&1)*8
    uint32_t a_off[MI], b_off[NJ];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
        if (A_MMAJOR) {
            const int im = wm * MI + mi;  // MMA slot along M: 16-wide group im>>1, half im&1
            const int m_in = 4 * (im & 1) + cm;
            a_off[mi] = (im >> 1) * 2048 + t * 128 + ((((m_in >> 1) ^ t)) << 4) + (m_in & 1) * 8;
        } else {
            const int row = wm * WM + 8 * mi + rm;
            a_off[mi] = row * 128 + (((t >> 1) ^ rm) << 4) + (t & 1) * 8;
        }
    }
#pragma unroll
    for (int nj = 0; nj < NJ; nj++) {
        const int jn = wn * NJ + nj;  // MMA slot along N: 16-wide group jn>>1, half jn&1
        const int n_in = 4 * (jn & 1) + cm;
        b_off[nj] = A_BYTES + (jn >> 1) * 2048 + t * 128 + ((((n_in >> 1) ^ t)) << 4) + (n_in & 1) * 8;
    }

    double acc[MI][NJ][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int nj = 0; nj < NJ; nj++) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

    for (int kt = 0; kt < KT; kt++) {
        const int s = kt % STAGES;
        mbar_wait(bar_base + 8 * s, (kt / STAGES) & 1);
        const uint8_t* st = smem_gen + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            double af[MI], bf[NJ];
            // k-step ks: K-contiguous: chunk ^= 2*ks -> byte ^= ks<<5 ; N-contiguous: row += 4*ks
            // (+512*ks bytes) and (row&7) ^= 4*(ks&1) -> byte ^= (ks&1)<<6.
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                const uint32_t o = A_MMAJOR ? ((a_off[mi] + 512 * ks) ^ ((ks & 1) << 6))
                                            : (a_off[mi] ^ (ks << 5));
                af[mi] = *reinterpret_cast<const double*>(st + o);
            }
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const uint32_t o = (b_off[nj] + 512 * ks) ^ ((ks & 1) << 6);
                bf[nj] = *reinterpret_cast<const double*>(st + o);
            }
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int nj = 0; nj < NJ; nj++) dmma_8x8x4(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + s));
    }

    // ---------------- epilogue ----------------
#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
        const int im = wm * MI + mi;
        const int row = m_base + (A_MMAJOR ? (16 * (im >> 1) + 4 * (im & 1) + cm) : (wm * WM + 8 * mi + rm));
        if (row >= p.M) continue;
#pragma unroll
        for (int nj = 0; nj < NJ; nj++) {
            const int jn = wn * NJ + nj;
            const int col = n_base + 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
            if (col >= p.N) continue;
            const double v0 = acc[mi][nj][0] * p.alpha, v1 = acc[mi][nj][1] * p.alpha;
            const bool has1 = (col + 1 < p.N);
            if (p.epi == EPI_STORE) {
                double* dst = (p.block_rows ? p.blocks[row / p.block_rows] + (long long)(row % p.block_rows) * p.ldc
                                            : p.C + (long long)row * p.ldc) + (long long)b * p.strideC + col;
                if (has1 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    dst[0] = v0;
                    if (has1) dst[1] = v1;
                }
            } else if (p.epi == EPI_LUT) {
                double* out = p.C + (long long)b * p.strideC;
                const long long li = (long long)row * p.N + col;
                const int i0 = p.lut[li];
                if (i0 >= 0) out[i0] = v0;
                if (has1) {
                    const int i1 = p.lut[li + 1];
                    if (i1 >= 0) out[i1] = v1;
                }
            } else {  // EPI_VARF2D
                const int m0 = row / p.Pp0, n0 = row - m0 * p.Pp0;
                if (n0 >= p.P0) continue;
                const int m1 = col / p.Pp1, n1 = col - m1 * p.Pp1;   // col is even and Pp1 is even:
                if (n1 >= p.P1) continue;                            // col and col+1 share m1
                const int r_out = p.lut[m0 * p.P1 + m1];
                if (r_out < p.row_begin || r_out >= p.row_end) continue;
                double* orow = p.C + (long long)(r_out - p.row_begin) * p.ldc;
                const int c0 = p.lut[n0 * p.P1 + n1];
                const int c1 = (has1 && n1 + 1 < p.P1) ? p.lut[n0 * p.P1 + n1 + 1] : -1;
                if (c1 == c0 + 1 && c0 >= 0 && ((reinterpret_cast<uintptr_t>(orow + c0) & 15) == 0)) {
                    *reinterpret_cast<double2*>(orow + c0) = make_double2(v0, v1);   // one 16-byte store
                } else {
                    if (c0 >= 0) orow[c0] = v0;
                    if (c1 >= 0) orow[c1] = v1;
                }
            }
        }
    }
}

}  // namespace hb
