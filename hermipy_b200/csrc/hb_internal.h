// Internal (non-ABI) declarations shared by the translation units of libhermite_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <vector>

#include "../../include/hermite_b200.h"

namespace hb {

struct GemmArgs {
    int M = 0, N = 0, K = 0, batch = 1;   // batch = number of b_hi entries; total GEMMs = batch * batch_lo
    const double* A = nullptr;  // a_mmajor ? [K][M] : [M][K]
    long long lda = 0, strideA = 0;  // stride 0 with batch > 1: operand shared by the batch
    bool a_mmajor = false;
    const double* B = nullptr;  // [K][N]
    long long ldb = 0, strideB = 0;
    double* C = nullptr;
    long long ldc = 0, strideC = 0;
    // two-level batch index b = b_hi * batch_lo + b_lo (see GemmParams): strideX is per b_hi, strideX_lo per b_lo;
    // a stride of 0 at a level with more than one entry shares the operand across that level
    int batch_lo = 1;
    long long strideA_lo = 0, strideB_lo = 0, strideC_lo = 0;
    int row_offset_lo = 0, col_offset_lo = 0;
    int epi = 0;
    const int* lut = nullptr;
    // varf scatter epilogue (see GemmParams in hb_gemm.cuh)
    const int* blkpos = nullptr;
    const unsigned char* blkperm = nullptr;
    const unsigned char* blkperm2 = nullptr;   // [(bn0 * tiles_n + tile_n) * 128 + s]: joint order of a column tile's entries
    const int2* blkrange = nullptr;
    int P0 = 0, Pp0 = 0, P1 = 0, Pp1 = 0, row_begin = 0, row_end = 0, symmetric = 0;
    int b_koff_odd = 0, nb_even0 = 0;   // parity-halved contraction (0 = off): see varf_dense_2d in hb_plan.cu
    double alpha = 1.0;
    int block_rows = 0;              // EPI_STORE: rows [i*block_rows, (i+1)*block_rows) -> blocks[i] (C unused)
    double* blocks[16] = {nullptr};
    int col_stride = 1, col_offset = 0;   // EPI_STORE: column c goes to col_offset + c*col_stride
    int row_stride = 1, row_offset = 0;   // EPI_LUT: GEMM row r uses look-up-table row row_offset + r*row_stride
    int group_m = 1;                 // tile rows per rasterisation group (L2 locality of big GEMMs)
    // exchange flags (see hb_gemm_desc in include/hermite_b200.h)
    unsigned long long* signal_flags[16] = {nullptr};
    int n_signal = 0;
    unsigned long long* signal_state = nullptr;
    const unsigned long long* wait_flags = nullptr;
    int n_wait = 0;
    const unsigned long long* wait_epoch = nullptr;
    int cfg = -1;  // -1 auto (cost model), else index into the tile-configuration table of hb_gemm.cu
};

int gemm_launch(const GemmArgs& a, cudaStream_t stream);
void gemm_set_trace(long long* dev);   // developer timeline of the varf stage-2 kernel (nullptr = off)
long long* gemm_get_trace();

// Stage 2 of the dense 2-D varf as a persistent ping-pong kernel (hb_varf_gemm.cu): `a` as for gemm_launch with
// epi = EPI_VARF2D; the tile list (pairs tile_m, tile_n) comes from varf_tile_list.
int varf_tile_list(const GemmArgs& a, const std::vector<int>& blkrange, std::vector<int>& out);
int varf_stage2_launch(const GemmArgs& a, const int2* d_tiles, int n_tiles, cudaStream_t stream);   // developer timeline of the varf stage-2 kernel (nullptr = off)

}  // namespace hb
