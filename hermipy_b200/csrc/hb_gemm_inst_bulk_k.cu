// Instantiations of the DMMA/TMA GEMM kernel: staged bulk-store epilogue (EPI_BULK), A operand K-contiguous.
#include "hb_gemm_launch.cuh"

namespace hb {
int gemm_dispatch_bulk_k(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
    switch (cfg) {
        HB_GEMM_BULK_CONFIGS(HB_GEMM_CASE, false, EPI_BULK)
        default: return HB_ERR_INVALID;
    }
}
}  // namespace hb
