// TMA (bulk tensor copy) helpers: host-side tensor-map construction through the driver entry point (no libcuda link
// dependency) and the device-side store / group instructions used by the tcgen05 epilogues.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace fdg {

// fp32 tensor map of rank 2..4 with SWIZZLE_128B boxes whose innermost extent is 32 floats (128 bytes).
// dims[0] is the contiguous (channel) dimension; strides_bytes[i] is the stride of dims[i + 1].
// Returns false when the view cannot be described (alignment / stride rules of cuTensorMapEncodeTiled).
bool make_tmap_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
// same for a bf16 tensor (boxes whose innermost extent is 64 elements = 128 bytes)
bool make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

// bulk tensor load global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t smem, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

// multicast form: the box lands at the same shared-memory offset in every CTA of `mask`, each one's mbarrier (same offset) counts the bytes
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t smem, const void* tmap, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no completion tracking): issued a few chunks ahead of the real load, it
// turns that load's DRAM latency into an L2 hit, so fewer bytes have to be in flight per SM to keep HBM busy
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(smem), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace fdg
