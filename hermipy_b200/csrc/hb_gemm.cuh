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
    EPI_VARF2D = 2,   // stage 2 of the dense 2-D varf: scatter A[idx(m0,m1)][idx(n0,n1)]; handled by the persistent
                      // kernel of hb_varf_gemm.cu (this generic kernel has no instantiation for it)
    EPI_BULK = 3,     // EPI_STORE through a shared-memory staging tile and bulk asynchronous copies: every row of a
                      // warp's sub-tile leaves as ONE contiguous run of full 128-byte lines.  Chosen by the launcher
                      // for the row-block scatter (peer-GPU destinations): NVLink moves full lines at several times
                      // the rate of the 32-byte pieces the register epilogue emits, and the copies run under the
                      // next tile's main loop.
};

constexpr int GEMM_MAX_BLOCKS = 16;

struct GemmParams {
    int M, N, K, batch;
    double* C;
    long long ldc;      // row pitch of C (EPI_STORE) / of the output matrix (EPI_VARF2D)
    long long strideC;  // batch stride of C / of the scattered output (EPI_LUT)
    int epi;
    const int* lut;     // EPI_LUT: M*N entries (padded lexicographic box -> position in the index set, or -1)
    // EPI_VARF2D (persistent kernel, hb_varf_gemm.cu).  GEMM row = blocked pair (m0,n0) of axis-0 indices, GEMM column = blocked pair (m1,n1) of axis-1
    // indices, pairs laid out in 8x8 blocks: p(i,j) = ((i>>3)*nb + (j>>3))*64 + (i&7)*8 + (j&7).  The entry lands at
    // A[idx(m0,m1)][idx(n0,n1)].  Positions come from per-block tables of the 2-D index set (built once per plan):
    //   blkpos  [(ba*Pp1 + bb)*64 + (a&7)*8 + (b&7)] = idx(a, b) for a in block ba of axis 0, b in block bb of axis 1
    //                                                  (-1 outside the set or beyond the basis);
    //   blkperm [(ba*Pp1 + bb)*64 + s]               = the 64 entries of the block sorted by position (invalid last);
    //   blkrange[ba*Pp1 + bb]                        = {smallest, largest} valid position of the block, {0,-1} if none.
    const int* blkpos;
    const unsigned char* blkperm;
    const unsigned char* blkperm2;   // persistent kernel: joint order of the 128 column entries of a column tile, per bn0
    const int2* blkrange;
    int P0, Pp0;        // basis count of axis 0 and its number of 8-blocks
    int P1, Pp1;        // same for axis 1
    int symmetric;      // compute only pairs with m0 <= n0 and mirror them (A is symmetric); full row range only
    int b_koff_odd;     // > 0: row tiles whose block pair (bm0,bn0) has odd parity read B from k = b_koff_odd on
    int nb_even0;       //      (parity of a block of axis 0: its index >= nb_even0), persistent kernel only
    int debug;          // developer switch (HB_VARF_DEBUG): bit0 drops direct stores, bit1 mirror stores, bit2 both loops
    int row_begin, row_end;  // owned output rows [row_begin,row_end) (row sharding)
    long long* trace;   // developer timeline (hb_debug_trace): [0] = record counter, [1] = capacity, then 5 words per CTA
    double alpha;       // result scale
    int tiles_m, tiles_n;    // tiles are numbered b*tiles_m*tiles_n + (grouped tile order, see the kernel)
    int n_tiles;             // tiles_m * tiles_n * batch; CTA c of the (at most one-wave) grid takes c, c + grid, ...
    int group_m;             // tile rows per rasterisation group (>= 1)
    // Two-level batch index b = b_hi * batch_lo + b_lo (batch_lo = 1: plain batch).  The low level is how a parity-
    // split transform stage runs its even and odd halves in ONE launch: b_lo selects the per-parity table and the
    // even / odd block of the data operand and of the result.  Per level and operand a multiplier 0 / 1 says whether
    // the operand varies with that index (tensor-map coordinates 2 and 3).
    int batch_lo;
    int a_lo_mul, a_hi_mul, b_lo_mul, b_hi_mul;
    long long strideC_lo;    // C offset per b_lo (strideC is per b_hi)
    int row_offset_lo;       // added to row_offset per b_lo (EPI_LUT rows, destination row of the block scatter)
    int col_offset_lo;       // added to col_offset per b_lo (interleaved columns)
    // EPI_STORE with row blocks: output row r goes to blocks[r / block_rows] + (r % block_rows)*ldc.
    // The block pointers may be peer-GPU addresses (NVLink P2P): this is how a GEMM's epilogue performs
    // the all-to-all of the sharded transform.
    int block_rows;          // 0 = off
    int col_stride, col_offset;   // EPI_STORE: column c is stored at col_offset + c * col_stride (1, 0 = plain)
    int row_stride, row_offset;   // EPI_LUT: GEMM row r is row row_offset + r * row_stride of the look-up table
    double* blocks[GEMM_MAX_BLOCKS];
    // Exchange flags (hb_gemm_desc): consumer side -- the TMA producer waits for wait_flags[i] >= *wait_epoch before its
    // first load; producer side -- the last CTA to finish bumps signal_state[0] and publishes it to the peers.
    const unsigned long long* wait_flags;
    const unsigned long long* wait_epoch;
    int n_wait, n_signal;
    unsigned long long* signal_state;
    unsigned long long* signal_flags[GEMM_MAX_BLOCKS];
};

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int GEMM_BK = 16;  // doubles per k-tile = one 128-byte swizzle span

// developer timeline (hb_debug_trace): word `w` of this CTA's 5-word record = now; record 0 also holds the SM id.
// Called by one lane of consumer warp 0; nothing is kept in registers between calls.
// Compiled in only with -DHB_GEMM_TRACE (make NVCC="nvcc -DHB_GEMM_TRACE"): even predicated off, the hook in the
// main loop cost the 128-register two-CTA configurations 2 us of a 37 us kernel.
__device__ __forceinline__ void gemm_trace(long long* trace, int w) {
#ifndef HB_GEMM_TRACE
    (void)trace; (void)w;
    return;
#endif
    if (trace == nullptr || (long long)blockIdx.x >= trace[1]) return;
    long long* rec = trace + 2 + 5 * (long long)blockIdx.x;
    if (w == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rec[0] = smid;
        atomicAdd(reinterpret_cast<unsigned long long*>(trace), 1ull);
    }
    rec[w == 0 ? 1 : w + 1] = global_ns();
}

template <int BM, int BN, int STAGES>
constexpr int gemm_smem_bytes() {
    return STAGES * (BM + BN) * GEMM_BK * 8 + 1024 /*alignment slack*/ + 2 * STAGES * 8;
}
// EPI_BULK: one staging sub-tile per consumer warp, rows padded by 16 bytes (conflict-free 16-byte fragment stores)
__host__ __device__ constexpr int gemm_stage_pitch(int WN) { return WN * 8 + 16; }
template <int BM, int BN, int STAGES, int WARPS_M, int WARPS_N>
constexpr int gemm_smem_bytes_bulk() {
    return gemm_smem_bytes<BM, BN, STAGES>() + 16 + WARPS_M * WARPS_N * (BM / WARPS_M) * gemm_stage_pitch(BN / WARPS_N);
}

// Register budget.  A CTA is NCW consumer warps plus one producer WARP GROUP (4 warps, one of them active), so
// that `setmaxnreg` can move registers between the roles: the register file is handed out per warp group, and a
// (NCW + 1)-warp CTA is charged for NCW + 4 warps anyway.  With 8 consumer warps (1 CTA/SM) every thread starts
// with 168 registers, with 4 consumer warps and 2 CTAs/SM with 128; the producer group drops to 40 and the
// consumers grow to 232 / 216, enough for the 128 accumulator registers of a 32 x 64 warp tile plus fragments
// (at 168 ptxas spills and shuffles accumulators through extra moves).  4 consumer warps at 1 CTA/SM already
// start with 255 registers.
template <int NCW, int MIN_CTAS>
struct GemmRegs {
    static constexpr bool kRealloc = (NCW == 8 && MIN_CTAS == 1) || (NCW == 4 && MIN_CTAS == 2);
    static constexpr int kConsumer = (NCW == 8) ? 232 : 216;
    static constexpr int kProducer = 40;
};

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool A_MMAJOR, int MIN_CTAS, int EPI>
__global__ void __launch_bounds__((WARPS_M* WARPS_N + 4) * 32, MIN_CTAS)
    dgemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                          const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int NCW = WARPS_M * WARPS_N;  // consumer warps
    constexpr int WM = BM / WARPS_M;
    // The tile is NU = BN/8 eight-column MMA slots wide.  When WARPS_N does not divide NU the first NU % WARPS_N
    // warp columns take one slot more (RAGGED): NJ is the per-warp maximum, `nj_full` says whether this warp uses
    // its last slot.  (48 x 208 over 2 x 4 warps: 7,7,6,6 slots -- with warp w on sub-partition w % 4 and
    // wn = w / 2, every sub-partition gets 13 slots x 3 row blocks.)
    constexpr int NU = BN / 8;
    constexpr int NJ = (NU + WARPS_N - 1) / WARPS_N, MI = WM / 8;
    constexpr bool RAGGED = (NU % WARPS_N) != 0;
    static_assert(BM % WARPS_M == 0 && WM % 8 == 0 && BN % 16 == 0 && (!A_MMAJOR || BM % 16 == 0), "tile granularity");
    constexpr int A_BYTES = BM * GEMM_BK * 8, B_BYTES = BN * GEMM_BK * 8;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES]
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    // PERSISTENT CTAs: the grid is at most one wave (SMs x resident CTAs), CTA c works on tiles c, c + grid, ...
    // The producer keeps the ring full across tile boundaries (the k-tile counter `it` runs on), so the first
    // operands of the next tile land while the consumers are still storing the current one: a tile after the first
    // costs its main loop plus the part of the epilogue that the prefetch depth cannot hide, not a CTA launch, a
    // barrier setup and a cold pipeline (1.1-1.9 us + 1.2-2.6 us per tile measured before, tools/gemm_timeline.py).
    // Tile order inside one batch entry: groups of `group_m` tile rows, tile row fastest inside a group, so
    // that the CTAs resident at the same time share a few A row-tiles and a few B column-tiles in L2
    // (group_m = 1 is the plain row-major order).
    const int per_batch = p.tiles_n * p.tiles_m;
    const int per_group = p.group_m * p.tiles_n;
    auto tile_coords = [&](int tile, int& b_lo, int& b_hi, int& m_base, int& n_base) {
        const int b = tile / per_batch;
        b_hi = b / p.batch_lo;
        b_lo = b - b_hi * p.batch_lo;
        const int t_in = tile - b * per_batch;
        const int first_m = (t_in / per_group) * p.group_m;
        const int g_rows = min(p.group_m, p.tiles_m - first_m);
        const int t_grp = t_in - (t_in / per_group) * per_group;
        m_base = (first_m + t_grp % g_rows) * BM;
        n_base = (t_grp / g_rows) * BN;
    };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) gemm_trace(p.trace, 0);   // [1] CTA start

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

    // Everything above (CTA launch, barrier setup) overlaps the tail of the previous kernel of the stream when
    // this one was launched programmatically; operands are only read from here on.  Every warp blocks here (a
    // hardware wait): consumer warps spinning on their mbarriers instead would steal issue slots from the
    // kernel that is still running.
    grid_dependency_wait();
    using Regs = GemmRegs<NCW, MIN_CTAS>;
    if (warp >= NCW) {
        if constexpr (Regs::kRealloc) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Regs::kProducer));
        // ---------------- TMA producer (one elected lane of the first warp of the producer group) ----------------
        if (warp == NCW && lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            if (p.wait_flags != nullptr) {
                // an operand is being delivered by peer GPUs: wait for their completion flags (acquire, system scope),
                // then order the TMA loads (async proxy) behind the acquire
                const unsigned long long epoch = *reinterpret_cast<const volatile unsigned long long*>(p.wait_epoch);
                for (int i = 0; i < p.n_wait; i++)
                    while (ld_relaxed_sys(p.wait_flags + i) < epoch) {
                    }
                __threadfence_system();
                fence_proxy_async_global();
            }

            int it = 0;   // k-tiles issued so far by this CTA, over all of its tiles
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                int b_lo, b_hi, m_base, n_base;
                tile_coords(tile, b_lo, b_hi, m_base, n_base);
                const int a2 = b_lo * p.a_lo_mul, a3 = b_hi * p.a_hi_mul, b2 = b_lo * p.b_lo_mul, b3 = b_hi * p.b_hi_mul;
                for (int kt = 0; kt < KT; kt++, it++) {
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait(bar_base + 8 * (STAGES + s), ((it / STAGES) - 1) & 1);
                    const uint32_t full = bar_base + 8 * s;
                    const uint32_t sA = smem_base + s * STAGE_BYTES, sB = sA + A_BYTES;
                    mbar_arrive_expect_tx(full, STAGE_BYTES);
                    const int k0 = kt * GEMM_BK;
                    if (A_MMAJOR) {
#pragma unroll
                        for (int g = 0; g < BM / 16; g++)
                            tma_load_4d(sA + g * 2048, &tmA, full, m_base + 16 * g, k0, a2, a3);
                    } else {
                        tma_load_4d(sA, &tmA, full, k0, m_base, a2, a3);
                    }
#pragma unroll
                    for (int g = 0; g < BN / 16; g++)
                        tma_load_4d(sB + g * 2048, &tmB, full, n_base + 16 * g, k0, b2, b3);
                }
            }
        }
        return;
    }

    // ---------------- consumers: DMMA ----------------
    if constexpr (Regs::kRealloc) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Regs::kConsumer));
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int jn0 = RAGGED ? wn * (NU / WARPS_N) + min(wn, NU % WARPS_N) : wn * NJ;   // first MMA slot along N
    const bool nj_full = !RAGGED || wn < NU % WARPS_N;
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
        const int jn = jn0 + nj;  // MMA slot along N: 16-wide group jn>>1, half jn&1
        const int n_in = 4 * (jn & 1) + cm;
        b_off[nj] = A_BYTES + (jn >> 1) * 2048 + t * 128 + ((((n_in >> 1) ^ t)) << 4) + (n_in & 1) * 8;
    }

    // Releasing a stage.  The fragment loads (generic proxy, LDS) of a k-tile must have COMPLETED before the producer
    // may let the TMA engine (async proxy) overwrite the stage.  An mbarrier.arrive placed right behind the k-tile
    // does not guarantee that: nothing ties it to the DMMAs that consume the loaded registers, and ptxas schedules it
    // above them -- measured as a few wrong operand elements per ~1e8 in the last k-steps of a tile whenever the
    // producer was fast enough (the 4-consumer-warp configurations with 2 CTAs per SM).  The release of stage
    // it - 1 is therefore issued behind the full-barrier WAIT of k-tile `it`: the spin loop is a scheduling boundary,
    // every DMMA of k-tile it - 1 has been issued before it, and a DMMA only issues once its operand loads have
    // returned.  In steady state the data of k-tile `it` is already there, so the stage is handed back as early as
    // before; the last stage of a tile is released when the next tile starts (the producer has STAGES - 1 stages to
    // prefetch into during the epilogue), and never after the last tile (nobody waits for it).
    int pending_release = -1;
    int it = 0;   // k-tiles consumed so far by this CTA, over all of its tiles
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int b_lo, b_hi, m_base, n_base;
    tile_coords(tile, b_lo, b_hi, m_base, n_base);
    const long long c_batch_off = (long long)b_hi * p.strideC + (long long)b_lo * p.strideC_lo;
    const bool first_tile = tile == (int)blockIdx.x, last_tile = tile + (int)gridDim.x >= p.n_tiles;
    double acc[MI][NJ][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int nj = 0; nj < NJ; nj++) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

    for (int kt = 0; kt < KT; kt++, it++) {
        const int s = it % STAGES;
        mbar_wait(bar_base + 8 * s, (it / STAGES) & 1);
        if (pending_release >= 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 8 * (STAGES + pending_release));
        }
        pending_release = s;
        if (kt == 0 && first_tile && threadIdx.x == 0) gemm_trace(p.trace, 1);   // [2] first operands landed
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
                bf[nj] = 0.0;
                if (!RAGGED || nj < NJ - 1 || nj_full) bf[nj] = *reinterpret_cast<const double*>(st + o);
            }
#pragma unroll
            for (int mi = 0; mi < MI; mi++)
#pragma unroll
                for (int nj = 0; nj < NJ; nj++)
                    if (!RAGGED || nj < NJ - 1 || nj_full) dmma_8x8x4(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
        }
    }

    // ---------------- epilogue ----------------
    // Programmatic dependent launch: once every CTA is past its main loop the successor's CTAs may take the
    // slots this grid frees and run their prologue under our epilogues (they block in grid_dependency_wait
    // until this grid has completed).  Triggering earlier lets them steal slots from our own later waves.
    if (last_tile && warp == 0 && lane == 0) grid_launch_dependents();
    if (first_tile && threadIdx.x == 0) gemm_trace(p.trace, 2);   // [3] main loop of the first tile done
    if constexpr (EPI == EPI_LUT) {
        // Scatter through the look-up table (last stage of a d-dim forward transform).  The table entries
        // of one accumulator row are fetched first (independent loads in flight together), then stored.
        double* out = p.C + c_batch_off;
        const int lut_row0 = p.row_offset + b_lo * p.row_offset_lo;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int im = wm * MI + mi;
            const int row = m_base + (A_MMAJOR ? (16 * (im >> 1) + 4 * (im & 1) + cm) : (wm * WM + 8 * mi + rm));
            if (row >= p.M) continue;
            const int* __restrict__ lrow = p.lut + (long long)(lut_row0 + row * p.row_stride) * p.N;
            int2 idx[NJ];
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const int jn = jn0 + nj;
                const int col = n_base + 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
                idx[nj] = make_int2(-1, -1);
                if (RAGGED && nj == NJ - 1 && !nj_full) continue;   // slot not owned by this warp
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
        if (first_tile && threadIdx.x == 0) gemm_trace(p.trace, 3);   // [4] this warp's epilogue issued
        continue;   // next tile
    }
    if constexpr (EPI == EPI_BULK) {
        // Staged store (launcher guarantees: K-contiguous A, unit column stride, even N, 16-byte aligned rows).
        static_assert(!RAGGED && !A_MMAJOR && (NJ % 2) == 0, "bulk epilogue: whole 16-column groups per warp");
        constexpr int WN = NJ * 8, PITCH = gemm_stage_pitch(WN);
        const uint32_t stg = ((bar_base + 2 * STAGES * 8 + 15u) & ~15u) + warp * (WM * PITCH);
        // the copies of this warp's previous tile must have read the staging buffer
        if (!first_tile) {
            if (lane < WM) bulk_wait_read();
            __syncwarp();
        }
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const uint32_t o = stg + (8 * mi + rm) * PITCH + (16 * (nj >> 1) + 4 * (nj & 1) + 8 * (t & 1) + 2 * (t >> 1)) * 8;
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(o), "d"(acc[mi][nj][0] * p.alpha),
                             "d"(acc[mi][nj][1] * p.alpha) : "memory");
            }
        fence_proxy_async();      // generic-proxy writes above -> visible to the bulk copies (async proxy) below
        __syncwarp();
        for (int r = lane; r < WM; r += 32) {
            const int row = m_base + wm * WM + r;
            const int col0 = n_base + 8 * jn0;
            if (row < p.M && col0 < p.N) {
                double* rowp;
                if (p.block_rows) {
                    const int drow = p.row_offset + b_lo * p.row_offset_lo + row;
                    rowp = p.blocks[drow / p.block_rows] + (long long)(drow % p.block_rows) * p.ldc;
                } else {
                    rowp = p.C + (long long)row * p.ldc;
                }
                bulk_store_s2g(rowp + c_batch_off + col0, stg + r * PITCH, (uint32_t)min(WN, p.N - col0) * 8u);
            }
        }
        if (lane < WM) {
            bulk_commit();
            if (last_tile) bulk_wait_all();
        }
        if (first_tile && threadIdx.x == 0) gemm_trace(p.trace, 3);
        continue;   // next tile
    }
    // Plain store.  Everything that depends only on the row (destination block, pitch, alignment) is resolved once
    // per accumulator row; the common case -- one destination, unit column stride, interior tile -- is a run of
    // 16-byte stores with no per-element address arithmetic (the fully unrolled per-element form of this
    // epilogue was large enough to miss the instruction cache: 22 % of the samples of a 38 us kernel).
    const bool interior_n = n_base + BN <= p.N;
#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
        const int im = wm * MI + mi;
        const int row = m_base + (A_MMAJOR ? (16 * (im >> 1) + 4 * (im & 1) + cm) : (wm * WM + 8 * mi + rm));
        if (row >= p.M) continue;
        double* rowp;
        if (p.block_rows) {   // destination row = row_offset + b_lo * row_offset_lo + row, dealt over the row blocks
            const int drow = p.row_offset + b_lo * p.row_offset_lo + row;
            rowp = p.blocks[drow / p.block_rows] + (long long)(drow % p.block_rows) * p.ldc;
        } else {
            rowp = p.C + (long long)row * p.ldc;
        }
        rowp += c_batch_off;
        const int col_offset = p.col_offset + b_lo * p.col_offset_lo;
        if (p.col_stride != 1) {   // interleaved output (even / odd coefficients of the 1-D parity transform)
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const int jn = jn0 + nj;
                const int col = n_base + 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
                if (col >= p.N) continue;
                if (RAGGED && nj == NJ - 1 && !nj_full) continue;   // slot not owned by this warp
                double* dst = rowp + col_offset + (long long)col * p.col_stride;
                dst[0] = acc[mi][nj][0] * p.alpha;
                if (col + 1 < p.N) dst[p.col_stride] = acc[mi][nj][1] * p.alpha;
            }
            continue;
        }
        const bool vec = (reinterpret_cast<uintptr_t>(rowp) & 15) == 0;   // columns are even: rows decide alignment
        double* base = rowp + n_base + 8 * (t & 1) + 2 * (t >> 1);
        if (vec && interior_n) {
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const int jn = jn0 + nj;
                if (RAGGED && nj == NJ - 1 && !nj_full) continue;   // slot not owned by this warp
                *reinterpret_cast<double2*>(base + 16 * (jn >> 1) + 4 * (jn & 1)) =
                    make_double2(acc[mi][nj][0] * p.alpha, acc[mi][nj][1] * p.alpha);
            }
            continue;
        }
#pragma unroll
        for (int nj = 0; nj < NJ; nj++) {
            const int jn = jn0 + nj;
            const int off = 16 * (jn >> 1) + 4 * (jn & 1);
            const int col = n_base + off + 8 * (t & 1) + 2 * (t >> 1);
            if (col >= p.N) continue;
            if (RAGGED && nj == NJ - 1 && !nj_full) continue;   // slot not owned by this warp
            const double v0 = acc[mi][nj][0] * p.alpha, v1 = acc[mi][nj][1] * p.alpha;
            if (vec && col + 1 < p.N) {
                *reinterpret_cast<double2*>(base + off) = make_double2(v0, v1);
            } else {
                base[off] = v0;
                if (col + 1 < p.N) base[off + 1] = v1;
            }
        }
    }
    if (first_tile && threadIdx.x == 0) gemm_trace(p.trace, 3);   // [4] this warp's epilogue issued
    }   // tile loop
    if (p.signal_state != nullptr) {
        // Publish "this rank's stores have landed" to the peers.  The consumer warps meet at a named barrier (the
        // producer group has exited; bulk copies were waited for by their issuing lanes), one thread fences system-wide
        // -- cumulative over everything the barrier ordered before it -- and counts the CTA in; the last CTA of the grid
        // bumps the epoch, fences once more and writes the epoch to this rank's slot on every peer.
        asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
        if (threadIdx.x == 0) {
            unsigned int* done = reinterpret_cast<unsigned int*>(p.signal_state + 1);
            __threadfence_system();
            if (atomicAdd(done, 1u) == gridDim.x - 1) {
                const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(p.signal_state) + 1;
                *reinterpret_cast<volatile unsigned long long*>(p.signal_state) = epoch;
                *reinterpret_cast<volatile unsigned int*>(done) = 0u;
                __threadfence_system();
                for (int i = 0; i < p.n_signal; i++) st_relaxed_sys(p.signal_flags[i], epoch);
            }
        }
    }
}

}  // namespace hb
