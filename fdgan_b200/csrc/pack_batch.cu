// fdg_pack_batch: every weight-operand repack of one network pass in ONE launch.
//
// The tensor-core kernels read weights as packed operand images (fp32 GEMM layout -> bf16 hi/lo, pre-swizzled); parameters
// stay in the PyTorch layout (state-dict compatible) and change every optimiser step, so the images are rebuilt per pass:
// 355 launches of ~4 us per training step (1.4 ms at batch 16; a third of all launches at batch 1).  Here the host builds a
// job table once per network (ops.PackPlan) and each pass refreshes all images with one launch per dependency level
// (fp32 GEMM operands first, the tcgen05 images made from them second).
#include "pack.cuh"

namespace fdg {

int umma_ntile(int taps, int Cout);
int umma_tap_pair(int taps, int Cin);
int umma_nchunks(int taps, int Cin);

__global__ void __launch_bounds__(256) pack_batch_kernel(const FdgPackJob* __restrict__ jobs, int njobs) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  // job of this block: last job whose first_block <= blockIdx.x (first_block is ascending)
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const FdgPackJob j = jobs[lo];
  const int64_t i0 = (int64_t)(blockIdx.x - j.first_block) * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)j.nblocks * blockDim.x;
  const float* __restrict__ src = j.src;
  if (j.kind <= 2) {
    float* __restrict__ dst = static_cast<float*>(j.dst);
    for (int64_t i = i0; i < j.total; i += step) pack_w_item(src, j.cout, j.cin, j.r, j.s, j.kind, dst, j.ld, i);
  } else if (j.kind == FDG_PACK_UMMA) {
    uint8_t* __restrict__ dst = static_cast<uint8_t*>(j.dst);
    const int cch = (j.cin + 63) / 64;
    for (int64_t i = i0; i < j.total; i += step) pack_umma_item(src, j.ld, j.r, j.cin, j.cout, j.s & 0xffff, cch, dst, i, j.s >> 16);
  } else {
    uint8_t* __restrict__ dst = static_cast<uint8_t*>(j.dst);
    for (int64_t i = i0; i < j.total; i += step) pack_k1_item(src, j.ld, j.cin, j.cout, dst, i);
  }
}

}  // namespace fdg

using namespace fdg;

// Work items of a job (the host sizes first_block / nblocks from it); -1 for a malformed job.
extern "C" int64_t fdg_pack_job_items(const FdgPackJob* j) {
  if (!j || j->cout <= 0 || j->cin <= 0) return -1;
  if (j->kind >= 0 && j->kind <= 2) {
    if (j->r <= 0 || j->s <= 0 || j->ld <= 0) return -1;
    const int64_t K = j->kind == 0 ? (int64_t)j->r * j->s * j->cin : (j->kind == 1 ? (int64_t)j->r * j->s * j->cout : j->cout);
    return K * j->ld;
  }
  if (j->kind == FDG_PACK_UMMA) {   // r = taps, s = fdg_umma_tile_code (must equal the kernel's tile / chunk choice for this filter)
    const int code = umma_ntile(j->r, j->cout) | (umma_tap_pair(j->r, j->cin) << 16);
    if (j->r <= 0 || j->s != code || j->ld < j->cout) return -1;
    const int nt = code & 0xffff;
    return (int64_t)cdiv(j->cout, nt) * umma_nchunks(j->r, j->cin) * nt * 8;
  }
  if (j->kind == FDG_PACK_K1) {
    if (j->cout > 32 || j->ld < j->cout) return -1;
    return (int64_t)cdiv(j->cin, 64) * 3 * 96 * 8;
  }
  return -1;
}

extern "C" int fdg_pack_batch(const FdgPackJob* jobs_dev, int njobs, int total_blocks, fdg_stream_t stream) {
  FDG_REQUIRE(jobs_dev && njobs > 0 && total_blocks > 0, "fdg_pack_batch: bad arguments");
  launch_k(pack_batch_kernel, dim3((unsigned)total_blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, jobs_dev, njobs);
  return check_launch("fdg_pack_batch");
}
