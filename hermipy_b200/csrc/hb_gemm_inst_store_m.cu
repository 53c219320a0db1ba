// Instantiations of the DMMA/TMA GEMM kernel: epilogue EPI_STORE, A operand M-contiguous ([K][M]).
#include "hb_gemm_launch.cuh"

namespace hb {
int gemm_dispatch_store_m(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
    switch (cfg) {
        HB_GEMM_CONFIGS(HB_GEMM_CASE, true, EPI_STORE)
        default: return HB_ERR_INVALID;
    }
}
}  // namespace hb
