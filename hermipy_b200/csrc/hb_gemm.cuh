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
                      // pairs are indexed in 8x8 blocks: p(i,j) = ((i>>3)*nb + (j>>3))*64 + (i&7)*8 + (j&7)
};

constexpr int GEMM_MAX_BLOCKS = 16;

struct GemmParams {
    int M, N, K, batch;
    double* C;
    long long ldc;      // row pitch of C (EPI_STORE) / of the output matrix (EPI_VARF2D)
    long long strideC;  // batch stride of C / of the scattered output (EPI_LUT)
    int epi;
    const int* lut;     // EPI_LUT: M*N entries; EPI_VARF2D: P0*P1 entries (lex (m0,m1) -> position or -1)
    int P0, Pp0;        // EPI_VARF2D: basis count of axis 0 and its number of 8-blocks (GEMM row = blocked pair (m0,n0))
    int P1, Pp1;        // EPI_VARF2D: same for axis 1 (GEMM col = blocked pair (m1,n1))
    int symmetric;      // EPI_VARF2D: compute only pairs with m0 <= n0 and mirror them (A is symmetric)
    int debug;          // developer switch (HB_VARF_DEBUG): bit0 drops direct stores, bit1 drops mirror stores
    int lut_mode;       // EPI_VARF2D: 0 = positions from lut, 1 = cube closed form, 2 = triangle closed form
    int row_begin, row_end;  // EPI_VARF2D: owned output rows [row_begin,row_end) (row sharding)
    int tri_degree;     // EPI_VARF2D: >=0 -> skip tiles with m0+m1 > tri_degree (triangle-like sets)
    double alpha;       // result scale
    int tiles_m, tiles_n;    // grid is linear: blockIdx.x = b*tiles_m*tiles_n + (grouped tile order, see the kernel)
    int group_m;             // tile rows per rasterisation group (>= 1)
    int a_bmul, b_bmul;      // 0: operand shared by all batch entries, 1: batched
    // EPI_STORE with row blocks: output row r goes to blocks[r / block_rows] + (r % block_rows)*ldc.
    // The block pointers may be peer-GPU addresses (NVLink P2P): this is how a GEMM's epilogue performs
    // the all-to-all of the sharded transform.
    int block_rows;          // 0 = off
    double* blocks[GEMM_MAX_BLOCKS];
};

// Position of the multi-index (a, b) in the 2-D index set (reference order, iterators.cpp:106-321): closed
// forms for the max-norm shell order (cube) and the graded order (triangle) keep dependent loads out of
// the scatter epilogue; the other sets go through the look-up table.
__device__ __forceinline__ int pair_position(const GemmParams& p, int a, int b) {
    if (p.lut_mode == 1) return b >= a ? b * b + b + a : a * a + b;
    if (p.lut_mode == 2) {
        const int g = a + b;
        return g <= p.tri_degree ? g * (g + 1) / 2 + a : -1;
    }
    return p.lut[a * p.P1 + b];
}

constexpr int GEMM_BK = 16;  // doubles per k-tile = one 128-byte swizzle span

template <int BM, int BN, int STAGES>
constexpr int gemm_smem_bytes() {
    return STAGES * (BM + BN) * GEMM_BK * 8 + 1024 /*alignment slack*/ + 2 * STAGES * 8;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool A_MMAJOR, int MIN_CTAS, int EPI>
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

    // Tile order inside one batch entry: groups of `group_m` tile rows, tile row fastest inside a group, so
    // that the CTAs resident at the same time share a few A row-tiles and a few B column-tiles in L2
    // (group_m = 1 is the plain row-major order).
    const int per_batch = p.tiles_n * p.tiles_m;
    const int b = blockIdx.x / per_batch;
    const int t_in = blockIdx.x - b * per_batch;
    const int per_group = p.group_m * p.tiles_n;
    const int first_m = (t_in / per_group) * p.group_m;
    const int g_rows = min(p.group_m, p.tiles_m - first_m);
    const int t_grp = t_in - (t_in / per_group) * per_group;
    const int tile_m = first_m + t_grp % g_rows;
    const int tile_n = t_grp / g_rows;
    const int m_base = tile_m * BM, n_base = tile_n * BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if constexpr (EPI == EPI_VARF2D) {
        // Tile-level skipping.  GEMM row = blocked pair (m0, n0), column = blocked pair (m1, n1); the entry
        // lands at output row lut[m0,m1], column lut[n0,n1].  A tile is a few 8x8 blocks on each side:
        // it is needed if some (row block, column block) combination is.
        const int rb_lo = m_base >> 6, rb_hi = (min(m_base + BM, p.M) - 1) >> 6;
        const int cb_lo = n_base >> 6, cb_hi = (min(n_base + BN, p.N) - 1) >> 6;
        bool needed = false;
        for (int rb = rb_lo; rb <= rb_hi && !needed; rb++) {
            const int bm0 = rb / p.Pp0, bn0 = rb - bm0 * p.Pp0;
            if (p.symmetric && bm0 > bn0) continue;              // mirrored from the (bn0, bm0) block
            for (int cb = cb_lo; cb <= cb_hi && !needed; cb++) {
                const int bm1 = cb / p.Pp1, bn1 = cb - bm1 * p.Pp1;
                // total-degree sets keep only m0+m1 <= degree and n0+n1 <= degree: the smallest values decide
                if (p.tri_degree >= 0 && (8 * (bm0 + bm1) > p.tri_degree || 8 * (bn0 + bn1) > p.tri_degree)) continue;
                if (p.row_begin <= 0 && p.row_end >= 0x7fffffff) { needed = true; break; }
                // row sharding: some (m0, m1) of the block pair must map to an owned output row
                for (int i = 0; i < 8 && !needed; i++)
                    for (int j = 0; j < 8; j++) {
                        const int a = 8 * bm0 + i, c = 8 * bm1 + j;
                        const int r = (a < p.P0 && c < p.P1) ? pair_position(p, a, c) : -1;
                        if (r >= p.row_begin && r < p.row_end) { needed = true; break; }
                    }
            }
        }
        if (!needed) return;
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
    if constexpr (BM == 64 && BN == 128 && A_MMAJOR) {
        if constexpr (EPI == EPI_VARF2D) {
            // Coalesced scatter through shared memory.  The tile is one 8x8 block of (m0,n0) pairs by two
            // 8x8 blocks of (m1,n1) pairs.  Output runs are contiguous along n1 or n0 (direct entry
            // A[idx(m0,m1)][idx(n0,n1)]) and along m1 or m0 (mirrored entry), never along a register
            // fragment, so the accumulators are staged in the (now idle) pipeline buffers and written
            // out 8 consecutive entries at a time.
            constexpr int SP = BN + 2;   // staging pitch in doubles
            double* S = reinterpret_cast<double*>(smem_gen);
            asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");   // all warps done with the stages
#pragma unroll
            for (int mi = 0; mi < MI; mi++) {
                const int im = wm * MI + mi;
                const int rl = 16 * (im >> 1) + 4 * (im & 1) + cm;
#pragma unroll
                for (int nj = 0; nj < NJ; nj++) {
                    const int jn = wn * NJ + nj;
                    const int cl = 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
                    *reinterpret_cast<double2*>(S + rl * SP + cl) =
                        make_double2(acc[mi][nj][0] * p.alpha, acc[mi][nj][1] * p.alpha);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
            const int rb = m_base >> 6, bm0 = rb / p.Pp0, bn0 = rb - bm0 * p.Pp0;
            const int f8 = lane & 7, s4 = lane >> 3;     // fast index (8 consecutive outputs), slow index (4)
            const bool row_interior = 8 * bm0 + 8 <= p.P0 && 8 * bn0 + 8 <= p.P0;
            if (p.debug & 4) return;
            // pass 0: direct entries A[idx(m0,m1)][idx(n0,n1)]; pass 1: mirrored entries (symmetric mode,
            // off-diagonal part).  Everything that does not depend on the lane is hoisted per column block.
            const int n_pass = (p.symmetric && bm0 <= bn0) ? 2 : 1;
            for (int pass = 0; pass < n_pass; pass++) {
                for (int cbl = 0; cbl < 2; cbl++) {
                    const int cb = (n_base >> 6) + cbl;
                    if (cb * 64 >= p.N) break;
                    const int bm1 = cb / p.Pp1, bn1 = cb - bm1 * p.Pp1;
                    const bool interior = row_interior && 8 * bm1 + 8 <= p.P1 && 8 * bn1 + 8 <= p.P1;
                    // which of the two patch indices runs along consecutive output entries
                    const bool swap = pass == 0 ? (bn1 > bn0) : (bm1 > bm0);
                    const double* Sc = S + cbl * 64;
#pragma unroll 4
                    for (int v = warp; v < 128; v += NCW) {
                        const int l0 = v >> 4, l1 = (v >> 1) & 7;
                        const int fa = swap ? s4 + 4 * (v & 1) : f8;     // index paired with l1's axis
                        const int fb = swap ? f8 : s4 + 4 * (v & 1);     // index paired with l0's axis
                        // pass 0: (m0l, m1l) = (l0, l1) fixed, (n0l, n1l) = (fb, fa); pass 1: the roles swap
                        const int m0l = pass == 0 ? l0 : fb, m1l = pass == 0 ? l1 : fa;
                        const int n0l = pass == 0 ? fb : l0, n1l = pass == 0 ? fa : l1;
                        const int m0 = 8 * bm0 + m0l, n0 = 8 * bn0 + n0l, m1 = 8 * bm1 + m1l, n1 = 8 * bn1 + n1l;
                        if (!interior && (m0 >= p.P0 || n0 >= p.P0 || m1 >= p.P1 || n1 >= p.P1)) continue;
                        if (p.symmetric && (pass ? m0 >= n0 : m0 > n0)) continue;
                        const int r_out = pair_position(p, m0, m1), c_out = pair_position(p, n0, n1);
                        if (r_out < 0 || c_out < 0) continue;
                        const double val = Sc[(m0l * 8 + n0l) * SP + m1l * 8 + n1l];
                        if (pass == 0) {
                            if (r_out >= p.row_begin && r_out < p.row_end && !(p.debug & 1))
                                p.C[(long long)(r_out - p.row_begin) * p.ldc + c_out] = val;
                        } else if (!(p.debug & 2)) {
                            p.C[(long long)c_out * p.ldc + r_out] = val;
                        }
                    }
                }
            }
            return;
        }
    }
    if constexpr (EPI == EPI_LUT) {
        // Scatter through the look-up table (last stage of a d-dim forward transform).  The table entries
        // of one accumulator row are fetched first (independent loads in flight together), then stored.
        double* out = p.C + (long long)b * p.strideC;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int im = wm * MI + mi;
            const int row = m_base + (A_MMAJOR ? (16 * (im >> 1) + 4 * (im & 1) + cm) : (wm * WM + 8 * mi + rm));
            if (row >= p.M) continue;
            const int* __restrict__ lrow = p.lut + (long long)row * p.N;
            int2 idx[NJ];
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const int jn = wn * NJ + nj;
                const int col = n_base + 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
                idx[nj] = make_int2(-1, -1);
                if (col + 1 < p.N && (p.N & 1) == 0) idx[nj] = __ldg(reinterpret_cast<const int2*>(lrow + col));
                else {
                    if (col < p.N) idx[nj].x = __ldg(lrow + col);
                    if (col + 1 < p.N) idx[nj].y = __ldg(lrow + col + 1);
                }
            }
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                if (idx[nj].x >= 0) out[idx[nj].x] = acc[mi][nj][0] * p.alpha;
                if (idx[nj].y >= 0) out[idx[nj].y] = acc[mi][nj][1] * p.alpha;
            }
        }
        return;
    }
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
            if constexpr (EPI == EPI_STORE) {
                double* dst = (p.block_rows ? p.blocks[row / p.block_rows] + (long long)(row % p.block_rows) * p.ldc
                                            : p.C + (long long)row * p.ldc) + (long long)b * p.strideC + col;
                if (has1 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    dst[0] = v0;
                    if (has1) dst[1] = v1;
                }
            } else if constexpr (EPI == EPI_VARF2D) {
                const int rb = row >> 6, bm0 = rb / p.Pp0, bn0 = rb - bm0 * p.Pp0;
                const int m0 = 8 * bm0 + ((row >> 3) & 7), n0 = 8 * bn0 + (row & 7);
                if (m0 >= p.P0 || n0 >= p.P0 || (p.symmetric && m0 > n0)) continue;
                const int cb = col >> 6, bm1 = cb / p.Pp1, bn1 = cb - bm1 * p.Pp1;
                const int m1 = 8 * bm1 + ((col >> 3) & 7), n1 = 8 * bn1 + (col & 7);   // col is even: col+1 -> n1+1
                if (m1 >= p.P1 || n1 >= p.P1) continue;
                const int r_out = pair_position(p, m0, m1);
                if (r_out < 0) continue;
                const int c0 = pair_position(p, n0, n1);
                const int c1 = (has1 && n1 + 1 < p.P1) ? pair_position(p, n0, n1 + 1) : -1;
                if (r_out >= p.row_begin && r_out < p.row_end) {
                    double* orow = p.C + (long long)(r_out - p.row_begin) * p.ldc;
                    if (c1 == c0 + 1 && c0 >= 0 && ((reinterpret_cast<uintptr_t>(orow + c0) & 15) == 0)) {
                        *reinterpret_cast<double2*>(orow + c0) = make_double2(v0, v1);   // one 16-byte store
                    } else {
                        if (c0 >= 0) orow[c0] = v0;
                        if (c1 >= 0) orow[c1] = v1;
                    }
                }
                if (p.symmetric && m0 < n0) {   // mirror: A[idx(n0,n1)][idx(m0,m1)] (full row range only)
                    if (c0 >= 0) p.C[(long long)c0 * p.ldc + r_out] = v0;
                    if (c1 >= 0) p.C[(long long)c1 * p.ldc + r_out] = v1;
                }
            }
        }
    }
}

}  // namespace hb
