// tcgen05 implicit-GEMM path of fdg_conv2d (placeholder until the kernel lands: reports "unsupported").
#include "common.cuh"

namespace fdg {
int conv2d_umma_supported(const FdgConv* p) { (void)p; return 0; }
int conv2d_umma(const FdgConv* p, cudaStream_t st) { (void)p; (void)st; set_error("tcgen05 conv path not built"); return FDG_ENOSUPPORT; }
}  // namespace fdg
