// fdg_conv2d dispatch: tcgen05 implicit GEMM (conv_umma.cu) where the shape qualifies, fp32 SIMT otherwise.
#include "common.cuh"

namespace fdg {
int conv2d_validate(const FdgConv* p);
int conv2d_simt(const FdgConv* p, cudaStream_t st);
int conv2d_umma_supported(const FdgConv* p);
int conv2d_umma(const FdgConv* p, cudaStream_t st);
int conv2d_thin_supported(const FdgConv* p);
int conv2d_thin(const FdgConv* p, cudaStream_t st);
int conv2d_cin1_supported(const FdgConv* p);
int conv2d_cin1(const FdgConv* p, cudaStream_t st);
int conv2d_k1_supported(const FdgConv* p);
int conv2d_k1(const FdgConv* p, cudaStream_t st);
}  // namespace fdg

extern "C" int fdg_conv2d(const FdgConv* p, fdg_stream_t stream) {
  int rc = fdg::conv2d_validate(p);
  if (rc != FDG_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->impl == 1) return fdg::conv2d_simt(p, st);
  if (p->impl == 0) {   // thin layers: direct HBM-bound kernels
    if (fdg::conv2d_cin1_supported(p)) return fdg::conv2d_cin1(p, st);
    if (fdg::conv2d_thin_supported(p)) return fdg::conv2d_thin(p, st);
  }
  if (!p->e_scale && fdg::conv2d_k1_supported(p)) return fdg::conv2d_k1(p, st);
  const int ok = fdg::conv2d_umma_supported(p);
  if (p->e_scale && !ok) { fdg::set_error("fdg_conv2d: BatchNorm-backward epilogue requested but the shape is not tcgen05-eligible"); return FDG_ENOSUPPORT; }
  if (p->impl == 2) {
    if (!ok) { fdg::set_error("fdg_conv2d: impl=tcgen05 requested but shape/layout unsupported"); return FDG_ENOSUPPORT; }
    return fdg::conv2d_umma(p, st);
  }
  return ok ? fdg::conv2d_umma(p, st) : fdg::conv2d_simt(p, st);
}
