// Hardware probe for the wgrad_k1 design: an MN-major SWIZZLE_64B B operand whose N blocks (32 elements = 64 B) OVERLAP --
// LBO = 64 B, i.e. N block v is the same [pixel][32 channel] tile shifted by v pixel rows -- and whose start address is
// shifted by whole 64-byte rows (not aligned to the 512-byte swizzle repeat).  Data is stored with the swizzle computed from
// ABSOLUTE shared-memory address bits (Swizzle<2,4,3>: bits [4,5] ^= bits [7,8]).
//   D[m][n] = sum_k A[m][k] * B[n][k],  A (K-major SWIZZLE_128B) = selector A[m][k] = (k == m % 16)
//   -> D[m][(v, co)] = G[pixel = shift + m % 16 + v][co]
// nvcc -gencode arch=compute_100a,code=sm_100a -o probe_mn64 probe_mn64.cu && ./probe_mn64
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../../fdgan_b200/csrc/umma.cuh"
using namespace fdg;

__device__ __forceinline__ uint64_t desc_mn64(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)4 << 61);
}
// D fp32, A/B bf16, A K-major, B MN-major (bit 16)
__host__ __device__ constexpr uint32_t idesc_bmn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int NPIX = 256;

__global__ void probe(int shift, int nblocks, int use_base_off, float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int t = threadIdx.x;
  // A: 128 rows x 64 k (K-major SWIZZLE_128B), selector on k < 16
  for (int i = t; i < 128 * 8; i += blockDim.x) {
    const int m = i >> 3, j = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((j * 8 + e) == (m % 16) ? 1.f : 0.f);
    const uint32_t addr = base + m * 128 + ((j ^ (m & 7)) << 4);
    uint4 pk = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w));
  }
  // G: NPIX pixel rows x 32 channels (64 B rows), 16-byte chunk j of row p at the absolute-address swizzle
  const uint32_t gbase = base + 128 * 128;
  for (int i = t; i < NPIX * 4; i += blockDim.x) {
    const int p = i >> 2, j = i & 3;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16((float)((p * 3 + j * 8 + e) % 250));
    const uint32_t lin = gbase + p * 64 + j * 16;
    const uint32_t addr = lin ^ (((lin >> 7) & 3u) << 4);
    uint4 pk = *reinterpret_cast<uint4*>(v);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w));
  }
  if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tbase;
  const int N = nblocks * 32;
  if (t == 0) {
    const uint32_t gstart = gbase + shift * 64;
    const uint32_t bo = use_base_off ? ((gstart >> 7) & 3) : 0;
    const uint32_t idesc = idesc_bmn(128, N);
    umma_bf16(tm, umma_desc_k128(base), desc_mn64(gstart, 64, 512, bo), idesc, 0);
    umma_commit(smem_u32(&bar));
  }
  if (t < 128) {
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int g = 0; g < nblocks; ++g) {
      float v[32];
      tmem_ld32(tm + ((uint32_t)((t >> 5) * 32) << 16) + g * 32, v);
      for (int n = 0; n < 32; ++n) out[t * 128 + g * 32 + n] = v[n];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 128 * 4);
  float* h = (float*)malloc(128 * 128 * 4);
  const int smem = 128 * 128 + NPIX * 64 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int total_bad = 0;
  for (int nb = 1; nb <= 3; nb += 2)
    for (int bo = 0; bo < 2; ++bo)
      for (int shift = 0; shift <= 40; ++shift) {
        probe<<<1, 128, smem>>>(shift, nb, bo, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("nblocks %d bo %d shift %d: CUDA error %s\n", nb, bo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int v = 0; v < nb; ++v)
            for (int co = 0; co < 32; ++co) {
              const int p = shift + (m % 16) + v;
              const float want = (float)((p * 3 + co) % 250);
              if (h[m * 128 + v * 32 + co] != want) ++bad;
            }
        if (bad || shift % 8 == 0 || shift == 18 || shift == 19)
          printf("N blocks %d base_off_mode %d shift %2d: %s (%d mismatches)\n", nb, bo, shift, bad ? "MISMATCH" : "ok", bad);
        if (bo == 0) total_bad += bad;
      }
  printf("base_off_mode 0 total mismatches: %d\n", total_bad);
  return 0;
}
