// Instantiation of the DMMA/TMA GEMM kernel with the varf scatter epilogue (stage 2 of the dense 2-D varf):
// 64 x 128 tiles, A M-contiguous (blocked pair tables), two CTAs per SM.
#include "hb_gemm_launch.cuh"

namespace hb {
int gemm_dispatch_varf_m(int cfg, const GemmArgs& a, const GemmParams& p, cudaStream_t stream) {
    if (cfg != 2) return HB_ERR_INVALID;
    return launch_cfg<64, 128, 2, 2, 4, true, 2, EPI_VARF2D>(a, p, stream);
}
}  // namespace hb
