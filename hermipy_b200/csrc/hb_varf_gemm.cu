// Stage 2 of the dense 2-D varf, A[idx(m0,m1)][idx(n0,n1)] = sum_k L[k][(m0,n0)] * G[k][(m1,n1)]
// (reference: the accumulation loop cpp/src/hermite/varf.cpp:236-262), as a persistent, warp-specialised
// "ping-pong" kernel.
//
// Why not the generic GEMM kernel with two CTAs per SM: the FP64 tensor pipe is the shared bottleneck of two
// co-resident CTAs, so the CTA that is behind catches up whenever the other one leaves its main loop, and the two
// end up in lock step -- both in the main loop (each at half rate), then both in the scatter epilogue with the
// pipe idle (measured on B200: 10.5 % of the time no main loop at all, tools/varf_timeline.py).  Here ONE CTA per
// SM holds two consumer groups (4 warps each, one per SM sub-partition) that are kept HALF A TILE APART by a
// progress flag: their main loops overlap (one group alone reaches only ~2/3 of the DMMA rate, two keep the pipe
// full), but they never start or end together, so the scatter of one group always runs under the main loop of the
// other, and its producer warp already streams the operands of its next tile.
//
//   warps 0-3  consumer group 0        warp 8  TMA producer of group 0     (warps 10, 11: idle, they only complete
//   warps 4-7  consumer group 1        warp 9  TMA producer of group 1      the producer warp group for setmaxnreg)
//
// Tiles are 64 x 128 (one 8x8 block of (m0,n0) pairs by two 8x8 blocks of (m1,n1) pairs) with K = number of
// nodes (Gram form) or 2N+1 (reference algebra); only needed tiles (owned rows, upper half in symmetric mode,
// non-empty blocks of the index set) are listed by the host, in an L2-friendly order, and drawn dynamically by
// the 2 x gridDim groups from a global counter.  Shared memory: 3-stage TMA ring per group (2 x 72 KB), one staging buffer for the
// epilogue (65 KB) guarded by a flag, position tables of the current tile.
#include <cstdlib>

#include <algorithm>
#include <vector>

#include "hb_gemm_launch.cuh"

namespace hb {

namespace {

constexpr int PP_BM = 64, PP_BN = 128, PP_STAGES = 3, PP_GROUP_WARPS = 4;
constexpr int PP_A_BYTES = PP_BM * GEMM_BK * 8, PP_B_BYTES = PP_BN * GEMM_BK * 8;
constexpr int PP_STAGE_BYTES = PP_A_BYTES + PP_B_BYTES;            // 24 KB
constexpr int PP_RING_BYTES = PP_STAGES * PP_STAGE_BYTES;          // 72 KB per group
constexpr int PP_SP = PP_BN + 2;                                   // staging pitch (doubles)
constexpr int PP_STAGING_BYTES = PP_BM * PP_SP * 8;                // 65 KB
constexpr int PP_TABLE_BYTES = 2 * (2 * 64 * 4 + 2 * 64) + 128;    // per group: row/col positions, permutations, joint column order
constexpr int PP_SMEM = 1024 + 2 * PP_RING_BYTES + PP_STAGING_BYTES + 2 * PP_TABLE_BYTES + 2 * 2 * PP_STAGES * 8 + 64;

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
__device__ __forceinline__ void group_sync(int g) {   // the 128 consumer threads of group g
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
}

__global__ void __launch_bounds__(384, 1)
    varf2d_pingpong_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const GemmParams p, const int2* __restrict__ tiles, int n_tiles,
                           int* __restrict__ counter) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout: ring[0] | ring[1] | staging | tables[0] | tables[1] | barriers | flags
    const uint32_t staging_off = 2 * PP_RING_BYTES;
    const uint32_t table_off = staging_off + PP_STAGING_BYTES;
    const uint32_t bar_off = table_off + 2 * PP_TABLE_BYTES;
    const uint32_t flag_off = bar_off + 2 * 2 * PP_STAGES * 8;
    int* flags = reinterpret_cast<int*>(smem_gen + flag_off);   // [1], [2] progress of group 0, 1; [3] staging busy;
                                                                // [4+2g], [5+2g] tile queue of group g; [8+g] tiles published

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KT = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (threadIdx.x == 0) {
        for (int g = 0; g < 2; g++)
            for (int s = 0; s < PP_STAGES; s++) {
                mbar_init(smem_base + bar_off + 8 * (g * 2 * PP_STAGES + s), 1);                          // full
                mbar_init(smem_base + bar_off + 8 * (g * 2 * PP_STAGES + PP_STAGES + s), PP_GROUP_WARPS);  // empty
            }
        for (int i = 0; i < 16; i++) flags[i] = 0;
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int g = warp >= 8 ? warp - 8 : warp >> 2;            // group of this warp
    const uint32_t ring = smem_base + g * PP_RING_BYTES;
    const uint32_t full0 = smem_base + bar_off + 8 * (g * 2 * PP_STAGES), empty0 = full0 + 8 * PP_STAGES;

    // Register budget: 384 threads start with 168 registers each; the producer warp group (warps 8-11, two of
    // them idle) gives most of its share back and the two consumer groups grow to 232, which holds the 128
    // accumulator registers plus fragments without the register shuffling ptxas needs at 168.
    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp >= 10) return;
        // ---------------- TMA producer of group g ----------------
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            int it = 0;
            for (int n = 0;; n++) {
                // tiles are handed out dynamically: the consumer group publishes the index of its next tile one
                // tile ahead (see below), so the producer runs on across tile boundaries
                while (ld_volatile(flags + 8 + g) < n + 1) {
                }
                const int ti = ld_volatile(flags + 4 + 2 * g + (n & 1));
                if (ti >= n_tiles) break;
                const int2 tile = tiles[ti];
                const int m_base = tile.x * PP_BM, n_base = tile.y * PP_BN;
                // parity-halved contraction: odd (m0 + n0) row blocks read the odd segment of B
                const int pb0 = tile.x / p.Pp0, pb1 = tile.x - pb0 * p.Pp0;
                const int kb = ((pb0 >= p.nb_even0) != (pb1 >= p.nb_even0)) ? p.b_koff_odd : 0;
                for (int kt = 0; kt < KT; kt++, it++) {
                    const int s = it % PP_STAGES;
                    if (it >= PP_STAGES) mbar_wait(empty0 + 8 * s, ((it / PP_STAGES) - 1) & 1);
                    const uint32_t full = full0 + 8 * s;
                    const uint32_t sA = ring + s * PP_STAGE_BYTES, sB = sA + PP_A_BYTES;
                    mbar_arrive_expect_tx(full, PP_STAGE_BYTES);
                    const int k0 = kt * GEMM_BK;
#pragma unroll
                    for (int q = 0; q < PP_BM / 16; q++) tma_load_3d(sA + q * 2048, &tmA, full, m_base + 16 * q, k0, 0);
#pragma unroll
                    for (int q = 0; q < PP_BN / 16; q++) tma_load_3d(sB + q * 2048, &tmB, full, n_base + 16 * q, kb + k0, 0);
                }
            }
        }
        return;
    }

    // ---------------- consumer group g ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    constexpr int MI = 4, NJ = 8;                               // warp tile 32 x 64 (2 x 2 warps over 64 x 128)
    const int w = warp & 3, wm = w & 1, wn = w >> 1;
    const int gq = lane >> 2, t = lane & 3;
    const int cm = (gq & 1) + 8 * ((gq >> 1) & 1) + 2 * (gq >> 2);   // conflict-free column map (see hb_gemm.cuh)
    uint32_t a_off[MI], b_off[NJ];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) {
        const int im = wm * MI + mi, m_in = 4 * (im & 1) + cm;
        a_off[mi] = (im >> 1) * 2048 + t * 128 + ((((m_in >> 1) ^ t)) << 4) + (m_in & 1) * 8;
    }
#pragma unroll
    for (int nj = 0; nj < NJ; nj++) {
        const int jn = wn * NJ + nj, n_in = 4 * (jn & 1) + cm;
        b_off[nj] = PP_A_BYTES + (jn >> 1) * 2048 + t * 128 + ((((n_in >> 1) ^ t)) << 4) + (n_in & 1) * 8;
    }
    const uint8_t* ring_gen = smem_gen + g * PP_RING_BYTES;
    double* S = reinterpret_cast<double*>(smem_gen + staging_off);
    int* s_rowpos = reinterpret_cast<int*>(smem_gen + table_off + g * PP_TABLE_BYTES);   // [2][64]
    int* s_colpos = s_rowpos + 128;
    unsigned char* s_rowperm = reinterpret_cast<unsigned char*>(s_colpos + 128);
    unsigned char* s_colperm = s_rowperm + 128;
    unsigned char* s_colperm2 = s_colperm + 128;   // joint order of both column blocks
    const int gt = threadIdx.x & 127;                           // thread index inside the group
    const int tiles_n = (p.N + PP_BN - 1) / PP_BN;

    // Dynamic tile scheduling (the tiles differ in scatter cost and the memory system is the bottleneck, so a
    // static deal leaves some SMs 20 % behind): thread 0 of the group draws tile indices from a global counter,
    // one tile ahead, and publishes them to the group and to its producer warp through a two-slot queue.
    if (gt == 0) {
        flags[4 + 2 * g] = atomicAdd(counter, 1);
        __threadfence_block();
        st_volatile(flags + 8 + g, 1);
    }
    // a stage is released behind the full-barrier wait of the NEXT k-tile, never right behind its own DMMAs: see the
    // comment on `pending_release` in hb_gemm.cuh (LDS of the stage must have completed before the TMA overwrites it)
    int it = 0, n_done = 0, pending_release = -1;
    for (int n = 0;; n++) {
        group_sync(g);                                          // slot n & 1 is published; slot (n+1) & 1 is free
        const int ti = ld_volatile(flags + 4 + 2 * g + (n & 1));
        if (ti >= n_tiles) break;
        if (gt == 0) {
            flags[4 + 2 * g + ((n + 1) & 1)] = atomicAdd(counter, 1);
            __threadfence_block();
            st_volatile(flags + 8 + g, n + 2);
        }
        const int2 tile = tiles[ti];
        const int n_base = tile.y * PP_BN;
        const int bm0 = tile.x / p.Pp0, bn0 = tile.x - bm0 * p.Pp0;

        // position tables of this tile (overlaps the other group's main loop)
        for (int e = gt; e < 128; e += 128) {
            const int cbl = e >> 6, cb = (n_base >> 6) + cbl;
            int rp = -1, cp = -1, rperm = e & 63, cperm = e & 63;
            if (cb * 64 < p.N) {
                const int bm1 = cb / p.Pp1, bn1 = cb - bm1 * p.Pp1;
                const int rblk = (bm0 * p.Pp1 + bm1) * 64 + (e & 63), cblk = (bn0 * p.Pp1 + bn1) * 64 + (e & 63);
                rp = __ldg(p.blkpos + rblk);
                cp = __ldg(p.blkpos + cblk);
                rperm = __ldg(p.blkperm + rblk);
                cperm = __ldg(p.blkperm + cblk);
            }
            s_rowpos[e] = rp; s_colpos[e] = cp;
            s_rowperm[e] = (unsigned char)rperm; s_colperm[e] = (unsigned char)cperm;
            s_colperm2[e] = __ldg(p.blkperm2 + ((size_t)bn0 * tiles_n + tile.y) * 128 + e);
        }

        long long t_tile = 0, t_main0 = 0, t_main1 = 0, t_stage = 0, t_staged = 0, t_direct = 0;
        if (p.trace && gt == 0) t_tile = global_ns();
        // ---- keep the two groups half a tile apart ----
        // prog[q] = 2 * (main loops finished by group q) + (1 once the current one is past its midpoint).  Tile n
        // of group 1 starts after the midpoint of tile n of group 0, tile n of group 0 after the midpoint of tile
        // n-1 of group 1: the main loops overlap (two warps per sub-partition keep the DMMA pipe full) but never
        // start or end together, so one group's scatter always runs under the other group's main loop.
        {
            const int need = g == 0 ? 2 * n_done - 1 : 2 * n_done + 1;
            if (lane == 0)
                while (ld_volatile(flags + 1 + (g ^ 1)) < need) {
                }
            __syncwarp();
        }

        if (p.trace && gt == 0) t_main0 = global_ns();
        double acc[MI][NJ][2];
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

        for (int kt = 0; kt < KT; kt++, it++) {
            const int s = it % PP_STAGES;
            if (kt == KT / 2 && gt == 0) st_volatile(flags + 1 + g, 2 * n_done + 1);
            mbar_wait(full0 + 8 * s, (it / PP_STAGES) & 1);
            if (pending_release >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * pending_release);
            }
            pending_release = s;
            const uint8_t* st = ring_gen + s * PP_STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                double af[MI], bf[NJ];
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
                    af[mi] = *reinterpret_cast<const double*>(st + ((a_off[mi] + 512 * ks) ^ ((ks & 1) << 6)));
#pragma unroll
                for (int nj = 0; nj < NJ; nj++)
                    bf[nj] = *reinterpret_cast<const double*>(st + ((b_off[nj] + 512 * ks) ^ ((ks & 1) << 6)));
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int nj = 0; nj < NJ; nj++) dmma_8x8x4(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
            }
        }

        // ---- hand the pipe over, then scatter ----
        group_sync(g);                                          // all four warps left the main loop
        n_done++;
        if (p.trace && gt == 0) t_main1 = global_ns();
        if (gt == 0) {
            st_volatile(flags + 1 + g, 2 * n_done);
            while (atomicCAS(flags + 3, 0, 1) != 0) {           // staging buffer (the other group may still scatter)
            }
        }
        group_sync(g);
        if (p.trace && gt == 0) t_stage = global_ns();
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
            const int im = wm * MI + mi;
            const int rl = 16 * (im >> 1) + 4 * (im & 1) + cm;
#pragma unroll
            for (int nj = 0; nj < NJ; nj++) {
                const int jn = wn * NJ + nj;
                const int cl = 16 * (jn >> 1) + 4 * (jn & 1) + 8 * (t & 1) + 2 * (t >> 1);
                *reinterpret_cast<double2*>(S + rl * PP_SP + cl) =
                    make_double2(acc[mi][nj][0] * p.alpha, acc[mi][nj][1] * p.alpha);
            }
        }
        group_sync(g);
        if (p.trace && gt == 0) t_staged = global_ns();
        const bool diag = (bm0 == bn0);
        // direct entries A[idx(m0,m1)][idx(n0,n1)] = S[(m0l,n0l)][(m1l,n1l)]: one warp per output row, lanes walk the
        // 128 columns of BOTH column blocks in position order (coalesced whatever the enumeration of the set is;
        // where it is consecutive along n1 the runs continue across the two blocks: 16 entries = 128 bytes).  What
        // a lane needs about its four columns does not depend on the row and is read once per tile.
        {
            // per lane: its four column entries (of 128, both column blocks, in position order)
            int c[4], soff[4], jj[4], cb_[4];
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int e = s_colperm2[lane + 32 * h];
                const int ec = e & 63;
                cb_[h] = e >> 6;
                c[h] = s_colpos[e];
                jj[h] = ec >> 3;
                soff[h] = (ec >> 3) * PP_SP + (e >> 6) * 64 + (ec & 7);
            }
#pragma unroll 4
            for (int q = 0; q < 64 / PP_GROUP_WARPS; q++) {
                const int er = w + PP_GROUP_WARPS * q;
                const int i = er >> 3, k = er & 7;
                const int r0 = s_rowpos[er], r1 = s_rowpos[64 + er];    // output row for column block 0 / 1
                const bool ok0 = r0 >= p.row_begin && r0 < p.row_end, ok1 = r1 >= p.row_begin && r1 < p.row_end;
                if (!ok0 && !ok1) continue;
                double* row0 = p.C + (long long)(r0 - p.row_begin) * p.ldc;
                double* row1 = p.C + (long long)(r1 - p.row_begin) * p.ldc;
                const double* Srow = S + i * 8 * PP_SP + k * 8;
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const bool ok = cb_[h] ? ok1 : ok0;
                    if (ok && c[h] >= 0 && !(p.symmetric && diag && i > jj[h]) && !(p.debug & 1))
                        (cb_[h] ? row1 : row0)[c[h]] = Srow[soff[h]];
                }
            }
        }
        if (p.trace && gt == 0) t_direct = global_ns();
        // mirrored entries A[idx(n0,n1)][idx(m0,m1)] for m0 < n0 (symmetric mode: full row range)
        if (p.symmetric) {
            for (int cbl = 0; cbl < 2; cbl++) {
                const double* Sc = S + cbl * 64;
                int r[2], soff[2], ii[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int er = s_rowperm[cbl * 64 + lane + 32 * h];
                    r[h] = s_rowpos[cbl * 64 + er];
                    ii[h] = er >> 3;
                    soff[h] = (er >> 3) * 8 * PP_SP + (er & 7) * 8;
                }
#pragma unroll 4
                for (int q = 0; q < 64 / PP_GROUP_WARPS; q++) {
                    const int ec = w + PP_GROUP_WARPS * q;
                    const int c = s_colpos[cbl * 64 + ec];
                    if (c < 0) continue;
                    const int j = ec >> 3, l = ec & 7;
                    double* orow = p.C + (long long)c * p.ldc;
                    const double* Scol = Sc + j * PP_SP + l;
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        if (r[h] >= 0 && !(diag && ii[h] >= j) && !(p.debug & 2)) orow[r[h]] = Scol[soff[h]];
                }
            }
        }
        group_sync(g);                                          // staging buffer and tables are free again
        if (gt == 0) {
            __threadfence_block();
            st_volatile(flags + 3, 0);
            if (p.trace) {   // developer timeline: 10 words per tile
                const long long slot = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull);
                if (slot < p.trace[1]) {
                    long long* rec = p.trace + 2 + 10 * slot;
                    rec[0] = ti; rec[1] = blockIdx.x * 2 + g; rec[2] = t_tile; rec[3] = t_main0; rec[4] = t_main1;
                    rec[5] = t_stage; rec[6] = global_ns(); rec[7] = t_staged; rec[8] = t_direct; rec[9] = 0;
                }
            }
        }
    }
    if (gt == 0) st_volatile(flags + 1 + g, 0x7fffffff);        // the other group no longer waits for this one
}

}  // namespace

// List of needed tiles in the order they are handed out.
// Axis 0 is parity-sorted (even basis indices first), so the 8 indices of a block are 2 apart: the entries a tile
// writes along m0 / n0 are every SECOND double of the output -- half of each 32-byte sector; the other half belongs
// to the partner block of the other parity (block b + nb_even).  A sector whose halves reach L2 far apart in time is
// evicted half-written and completed by a DRAM read-modify-write: ncu showed 26 GB read + 26 GB written for the 13 GB
// matrix of C2/cube.  Tiles are therefore listed in FAMILIES that complete each other's sectors -- for base blocks
// (i, j): (i,j), (i,j'), (i',j), (i',j') and, in symmetric mode, their transposed family, whose mirrored entries land
// in the same rows -- back to back, so that the dynamic scheduler runs them at the same time on neighbouring groups
// and the halves merge in L2.  Families are walked in G x G squares of base blocks, and for every column tile the whole
// square back to back, so that a square keeps its part of the left table in L2 while the right table streams.
int varf_tile_list(const GemmArgs& a, const std::vector<int>& blkrange, std::vector<int>& out) {
    const int tiles_n = (a.N + PP_BN - 1) / PP_BN;
    static const int G_env = [] { const char* e = getenv("HB_VARF_GROUP"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
    const int G = G_env > 0 ? G_env : (a.group_m > 1 ? a.group_m : 4);
    static const bool families = [] { const char* e = getenv("HB_VARF_FAMILIES"); return !(e && e[0] == '0'); }();
    out.clear();
    auto needed = [&](int bm0, int bn0, int tn) {
        if (a.symmetric && bm0 > bn0) return false;
        for (int cbl = 0; cbl < PP_BN / 64; cbl++) {
            const int cb = tn * (PP_BN / 64) + cbl;
            if (cb * 64 >= a.N) break;
            const int bm1 = cb / a.Pp1, bn1 = cb - bm1 * a.Pp1;
            const int* rr = &blkrange[2 * (bm0 * a.Pp1 + bm1)];
            const int* cr = &blkrange[2 * (bn0 * a.Pp1 + bn1)];
            if (rr[1] >= rr[0] && cr[1] >= cr[0] && rr[1] >= a.row_begin && rr[0] < a.row_end) return true;
        }
        return false;
    };
    // base blocks: the even-parity blocks [0, nbe); partner(b) = the odd-parity block over the same index range
    const int nbe = (families && a.nb_even0 > 0 && a.nb_even0 < a.Pp0) ? a.nb_even0 : a.Pp0;
    auto partner = [&](int b) { return b + nbe < a.Pp0 ? b + nbe : -1; };
    auto emit = [&](int x, int y, int tn) {
        if (x >= 0 && y >= 0 && needed(x, y, tn)) { out.push_back(x * a.Pp0 + y); out.push_back(tn); }
    };
    for (int m0 = 0; m0 < nbe; m0 += G)
        for (int n0 = a.symmetric ? m0 : 0; n0 < nbe; n0 += G)
            for (int tn = 0; tn < tiles_n; tn++)
                for (int i = m0; i < std::min(m0 + G, nbe); i++)
                    for (int j = (a.symmetric && n0 == m0) ? i : n0; j < std::min(n0 + G, nbe); j++) {
                        const int ip = partner(i), jp = partner(j);
                        emit(i, j, tn); emit(i, jp, tn); emit(ip, j, tn); emit(ip, jp, tn);
                        if (a.symmetric && i != j) { emit(j, i, tn); emit(j, ip, tn); emit(jp, i, tn); emit(jp, ip, tn); }
                    }
    return HB_OK;
}

int varf_stage2_launch(const GemmArgs& a, const int2* d_tiles, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return HB_OK;
    if ((a.lda & 1) || (a.ldb & 1) || (reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15))
        return HB_ERR_ALIGN;
    CUtensorMap tmA, tmB;
    HB_TRY(make_map(&tmA, a.A, a.M, a.K, 1, a.lda, (uint64_t)a.lda * a.K, 16, GEMM_BK));
    const int b_rows = a.b_koff_odd > 0 ? a.b_koff_odd + a.K : a.K;   // B holds two k-segments when the parity split is on
    HB_TRY(make_map(&tmB, a.B, a.N, b_rows, 1, a.ldb, (uint64_t)a.ldb * b_rows, 16, GEMM_BK));
    GemmParams p = {};
    p.M = a.M; p.N = a.N; p.K = a.K; p.batch = 1;
    p.C = a.C; p.ldc = a.ldc;
    p.epi = EPI_VARF2D;
    p.blkpos = a.blkpos; p.blkperm = a.blkperm; p.blkperm2 = a.blkperm2; p.blkrange = a.blkrange;
    p.P0 = a.P0; p.Pp0 = a.Pp0; p.P1 = a.P1; p.Pp1 = a.Pp1;
    p.symmetric = a.symmetric; p.row_begin = a.row_begin; p.row_end = a.row_end;
    p.b_koff_odd = a.b_koff_odd; p.nb_even0 = a.nb_even0;
    p.alpha = a.alpha;
    p.col_stride = 1;
    p.trace = gemm_get_trace();
    { const char* dbg = getenv("HB_VARF_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }   // developer switch: bit0 / bit1 drop the direct / mirrored stores
    const DeviceInfo dev = current_device_info();
    static std::atomic<unsigned long long> attr_set{0};   // one bit per device (the opt-in is per device)
    if (!(attr_set.load(std::memory_order_relaxed) >> dev.index & 1)) {
        if (cudaFuncSetAttribute(varf2d_pingpong_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM) != cudaSuccess)
            return HB_ERR_CUDA;
        attr_set.fetch_or(1ull << dev.index, std::memory_order_relaxed);
    }
    const int grid = std::min(dev.sms, (n_tiles + 1) / 2);
    int* counter = reinterpret_cast<int*>(const_cast<int2*>(d_tiles) + n_tiles);   // one spare entry after the list
    HB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    varf2d_pingpong_kernel<<<grid, 384, PP_SMEM, stream>>>(tmA, tmB, p, d_tiles, n_tiles, counter);
    return launch_status();
}

}  // namespace hb
