// Hardware probe: does a K-major SWIZZLE_128B UMMA descriptor whose start address is shifted by whole 128-byte rows
// (not 1024-byte aligned) and whose SBO is not 1024 read the rows one expects when the data was stored with the
// swizzle computed from ABSOLUTE shared-memory address bits?  Variants: base_offset = 0 or (start >> 7) & 7.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../../fdgan_b200/csrc/umma.cuh"
using namespace fdg;

__device__ __forceinline__ uint64_t desc_k128_ex(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}

// smem: A region of NROWS pixel rows x 128 B (row r holds 64 bf16 = value f(r,k)), B = identity-like [32 x 64]
// D[m][n] = sum_k A[row(m)][k] * B[n][k], with B[n][k] = (k == n) -> D[m][n] = A[row(m)][n], n < 32
__global__ void probe(int shift, int pitch_rows, int use_base_off, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int t = threadIdx.x;
  const int NROWS = 320;
  // A: row r, 16-byte chunk j stored at absolute-address swizzle
  for (int i = t; i < NROWS * 8; i += blockDim.x) {
    const int r = i >> 3, j = i & 7;
    const uint32_t rowaddr = base + r * 128;
    const uint32_t addr = rowaddr + ((j ^ ((rowaddr >> 7) & 7)) << 4);
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)(r % 251) + 0.0f * (j * 8 + e) + ((j * 8 + e) == 0 ? 0.f : 0.f) + (float)((j * 8 + e) * 256 % 7 == 0 ? 0 : 0));
    // value encodes row and column: r + 1000*(col%2)?  keep exact in bf16: use small ints: row%128 and col separately
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)((r * 3 + (j * 8 + e)) % 250));
    uint4 pk = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w));
  }
  const uint32_t bbase = base + NROWS * 128;   // 1024-aligned since NROWS*128 = 40960
  for (int i = t; i < 32 * 8; i += blockDim.x) {
    const int n = i >> 3, j = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((j * 8 + e) == n ? 1.f : 0.f);
    const uint32_t addr = bbase + n * 128 + ((j ^ (n & 7)) << 4);
    uint4 pk = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w));
  }
  if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tbase;
  if (t == 0) {
    const uint32_t astart = base + shift * 128;
    const uint32_t sbo = pitch_rows * 128;
    const uint32_t bo = use_base_off ? ((astart >> 7) & 7) : 0;
    const uint32_t idesc = umma_idesc_bf16(128, 32);
    for (int k4 = 0; k4 < 4; ++k4)
      umma_bf16(tm, desc_k128_ex(astart + k4 * 32, sbo, bo), umma_desc_k128(bbase + k4 * 32), idesc, k4 > 0);
    umma_commit(smem_u32(&bar));
  }
  if (t < 128) {
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tm + ((uint32_t)((t >> 5) * 32) << 16), v);
    for (int n = 0; n < 32; ++n) out[t * 32 + n] = v[n];
    tc_fence_before();
  }
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  float* h = (float*)malloc(128 * 32 * 4);
  const int smem = 320 * 128 + 32 * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pitch = 8; pitch <= 10; pitch += 2)
    for (int bo = 0; bo < 2; ++bo)
      for (int shift = 0; shift <= 11; ++shift) {
        probe<<<1, 128, smem>>>(shift, pitch, bo, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pitch %d bo %d shift %d: CUDA error %s\n", pitch, bo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int r = shift + (m / 8) * pitch + (m % 8);
          for (int n = 0; n < 32; ++n) {
            const float want = (float)((r * 3 + n) % 250);
            if (h[m * 32 + n] != want) ++bad;
          }
        }
        printf("pitch_rows %2d base_off_mode %d shift %2d: %s (%d mismatches)\n", pitch, bo, shift, bad ? "MISMATCH" : "ok", bad);
      }
  return 0;
}
