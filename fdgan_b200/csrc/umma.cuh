// tcgen05 / mbarrier / bulk-copy PTX wrappers and operand helpers shared by the sm_100a tensor-core kernels.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace fdg {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// multicast form: the bytes land at the same shared-memory offset in every CTA of `mask`, each one's mbarrier (same offset) counts them
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 |
// SBO(1024 B >> 4)<<32 | version 1 <<46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- thread-block clusters (pairs of CTAs sharing an operand through multicast bulk copies)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit that arrives on the barrier at the same shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// same without the wait: several loads can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 pair -> packed bf16x2 (lo half = first element) and the residual pair
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups (halo tiles: the pitch of one image row)
__device__ __forceinline__ uint64_t umma_desc_k128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// MN-major SWIZZLE_128B descriptor: rows of 64 MN-contiguous bf16 (128 B), 8 K-rows per 1024-byte swizzle atom;
// LBO = byte stride between 64-element MN blocks, SBO = byte stride between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// instruction descriptor with both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// Low / high 32-bit words of a SWIZZLE_128B descriptor (K-major: LBO field 1, SBO = bytes between 8-row groups;
// MN-major: LBO = bytes between 64-element blocks).  The 14-bit start-address field lives in the low word, so
// stepping through K slices or filter taps is a 32-bit add on the low word only.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFF) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }

// Issues the 12 tcgen05.mma of one 64-deep chunk (bf16x3 split: hi*hi, hi*lo, lo*hi; four K=16 slices each) with the
// minimum of scalar work per instruction -- a single thread feeds the tensor core, so its instruction latency is the
// issue-rate ceiling.  kstep = encoded descriptor units between K slices (2 = 32 B for K-major, 128 = 2048 B for MN-major).
__device__ __forceinline__ void umma_chunk12(uint32_t d_tmem, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t a_hw, uint32_t b_hi_lo,
                                             uint32_t b_lo_lo, uint32_t b_hw, uint32_t idesc, uint32_t accumulate_first, uint32_t kstep) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .pred p, pt;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%4, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, p;\n\t"
      "mad.lo.u32 ta, %9, 1, %1;\n\t"
      "mad.lo.u32 tb, %9, 1, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %1;\n\t"
      "mad.lo.u32 tb, %9, 2, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %1;\n\t"
      "mad.lo.u32 tb, %9, 3, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%5, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %1;\n\t"
      "mad.lo.u32 tb, %9, 1, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %1;\n\t"
      "mad.lo.u32 tb, %9, 2, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %1;\n\t"
      "mad.lo.u32 tb, %9, 3, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mov.b64 da, {%2, %3};\n\t"
      "mov.b64 db, {%4, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %2;\n\t"
      "mad.lo.u32 tb, %9, 1, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %2;\n\t"
      "mad.lo.u32 tb, %9, 2, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %2;\n\t"
      "mad.lo.u32 tb, %9, 3, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi_lo), "r"(a_lo_lo), "r"(a_hw), "r"(b_hi_lo), "r"(b_lo_lo), "r"(b_hw), "r"(idesc), "r"(accumulate_first), "r"(kstep)
      : "memory");
}

// same with different descriptor steps for the A and B operands (halo tiles: A steps by image rows)
__device__ __forceinline__ void umma_chunk12_ab(uint32_t d_tmem, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t a_hw, uint32_t b_hi_lo,
                                             uint32_t b_lo_lo, uint32_t b_hw, uint32_t idesc, uint32_t accumulate_first, uint32_t kstep, uint32_t kstep_b) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb;\n\t"
      ".reg .pred p, pt;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%4, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, p;\n\t"
      "mad.lo.u32 ta, %9, 1, %1;\n\t"
      "mad.lo.u32 tb, %10, 1, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %1;\n\t"
      "mad.lo.u32 tb, %10, 2, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %1;\n\t"
      "mad.lo.u32 tb, %10, 3, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%5, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %1;\n\t"
      "mad.lo.u32 tb, %10, 1, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %1;\n\t"
      "mad.lo.u32 tb, %10, 2, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %1;\n\t"
      "mad.lo.u32 tb, %10, 3, %5;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mov.b64 da, {%2, %3};\n\t"
      "mov.b64 db, {%4, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %2;\n\t"
      "mad.lo.u32 tb, %10, 1, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %2;\n\t"
      "mad.lo.u32 tb, %10, 2, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %2;\n\t"
      "mad.lo.u32 tb, %10, 3, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %7, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi_lo), "r"(a_lo_lo), "r"(a_hw), "r"(b_hi_lo), "r"(b_lo_lo), "r"(b_hw), "r"(idesc), "r"(accumulate_first), "r"(kstep), "r"(kstep_b)
      : "memory");
}

// Concatenated-B form of the bf16x3 split: the hi and lo images of the B operand are adjacent in shared memory, so
// ONE MMA of width 2*NT multiplies A_hi with [B_hi | B_lo] (columns [0, NT) += hi*hi, [NT, 2NT) += hi*lo) and a
// second one of width NT adds A_lo * B_hi into columns [NT, 2NT): 8 instead of 12 MMAs per 64-deep chunk and one
// third fewer shared-memory reads of the A operand (the binding resource for small N).  The epilogue adds the two
// column halves.  idesc2 / idesc1 = instruction descriptors for N = 2NT / NT; kstep_a / kstep_b as in umma_chunk12_ab.
__device__ __forceinline__ void umma_chunk8(uint32_t d_tmem, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t a_hw, uint32_t b_hi_lo,
                                            uint32_t b_hw, uint32_t idesc2, uint32_t idesc1, uint32_t accumulate_first, uint32_t kstep_a,
                                            uint32_t kstep_b, uint32_t nt_cols) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb, d2;\n\t"
      ".reg .pred p, pt;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.u32 d2, %0, %11;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
      "mov.b64 da, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %1;\n\t"
      "mad.lo.u32 tb, %10, 1, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, pt;\n\t"
      "mad.lo.u32 ta, %9, 1, %2;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %1;\n\t"
      "mad.lo.u32 tb, %10, 2, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, pt;\n\t"
      "mad.lo.u32 ta, %9, 2, %2;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %7, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %1;\n\t"
      "mad.lo.u32 tb, %10, 3, %4;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "mov.b64 db, {tb, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, pt;\n\t"
      "mad.lo.u32 ta, %9, 3, %2;\n\t"
      "mov.b64 da, {ta, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %7, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi_lo), "r"(a_lo_lo), "r"(a_hw), "r"(b_hi_lo), "r"(b_hw), "r"(idesc2), "r"(idesc1), "r"(accumulate_first),
        "r"(kstep_a), "r"(kstep_b), "r"(nt_cols)
      : "memory");
}

// one K = 16 slice of the concatenated-B scheme (see umma_chunk8): D[0, 2NT) (+)= A_hi [B_hi | B_lo], D[NT, 2NT) += A_lo B_hi
__device__ __forceinline__ void umma_concat_slice(uint32_t d_tmem, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t a_hw, uint32_t b_lo,
                                                  uint32_t b_hw, uint32_t idesc2, uint32_t idesc1, uint32_t accumulate, uint32_t nt_cols) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 d2;\n\t"
      ".reg .pred p, pt;\n\t"
      "setp.ne.b32 p, %8, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.u32 d2, %0, %9;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
      "mov.b64 da, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %7, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi_lo), "r"(a_lo_lo), "r"(a_hw), "r"(b_lo), "r"(b_hw), "r"(idesc2), "r"(idesc1), "r"(accumulate), "r"(nt_cols)
      : "memory");
}

// Two filter taps x two K = 16 slices of the concatenated-B scheme in one block (conv_halo pair mode, 32 input channels): the B
// descriptor walks the four 32-byte K slices of the stage, the A descriptors are the halo tile at tap 0's / tap 1's row shift.
__device__ __forceinline__ void umma_pair8(uint32_t d_tmem, uint32_t a0_hi, uint32_t a0_lo, uint32_t a1_hi, uint32_t a1_lo, uint32_t a_hw,
                                           uint32_t b_lo, uint32_t b_hw, uint32_t idesc2, uint32_t idesc1, uint32_t accumulate_first,
                                           uint32_t nt_cols) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 ta, tb, d2;\n\t"
      ".reg .pred p, pt;\n\t"
      "setp.ne.b32 p, %10, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.u32 d2, %0, %11;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%6, %7};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %8, p;\n\t"
      "mov.b64 da, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %9, pt;\n\t"
      "add.u32 ta, %1, 2;\n\t"
      "add.u32 tb, %6, 2;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %7};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %8, pt;\n\t"
      "add.u32 ta, %2, 2;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %9, pt;\n\t"
      "add.u32 tb, %6, 4;\n\t"
      "mov.b64 da, {%3, %5};\n\t"
      "mov.b64 db, {tb, %7};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %8, pt;\n\t"
      "mov.b64 da, {%4, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %9, pt;\n\t"
      "add.u32 ta, %3, 2;\n\t"
      "add.u32 tb, %6, 6;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "mov.b64 db, {tb, %7};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %8, pt;\n\t"
      "add.u32 ta, %4, 2;\n\t"
      "mov.b64 da, {ta, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [d2], da, db, %9, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a0_hi), "r"(a0_lo), "r"(a1_hi), "r"(a1_lo), "r"(a_hw), "r"(b_lo), "r"(b_hw), "r"(idesc2), "r"(idesc1),
        "r"(accumulate_first), "r"(nt_cols)
      : "memory");
}

// one tcgen05.mma from 32-bit descriptor halves
__device__ __forceinline__ void umma_single(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hw, uint32_t b_lo, uint32_t b_hw, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hw), "r"(b_lo), "r"(b_hw), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace fdg
