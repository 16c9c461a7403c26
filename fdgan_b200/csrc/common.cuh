// Shared helpers for the fdgan_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fdgan_b200.h"

namespace fdg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);
int current_device();         // cudaGetDevice clamped to [0, 63]: index of per-device one-time state
int device_sm_count();        // SM count of the current device (cached per device)
bool epi_warp_stores();      // tcgen05 epilogues: one bulk tensor store per WARP and group (FDG_EPI_WARP=0: one per group, 128-thread barriers)
int dbg_flags();             // ablation switches for the tcgen05 kernels (fdg_set_option("dbg", v)); 0 in production
void set_dbg_flags(int v);  // cudaGetLastError -> FDG_ECUDA

// optional per-launch CUDA-event timing (bench.py's live roofline measurement); families of kernels
enum ProfFamily { PF_CONV_SIMT = 0, PF_CONV_UMMA = 1, PF_WGRAD = 2, PF_EW = 3, PF_FREQ = 4, PF_OTHER = 5, PF_COUNT = 6 };
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int family, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
};

// Programmatic dependent launch (PDL).  Kernels on the hot path are launched with programmatic stream serialization: the
// next kernel's CTAs may be scheduled -- and run their set-up (barrier init, TMEM allocation, index arithmetic) -- while this
// kernel drains, instead of after it.  Contract: a kernel launched through launch_k() executes pdl_wait() before it touches any
// global memory (it returns once every earlier kernel in the stream has completed and flushed), then pdl_trigger() so that its own
// successor may be scheduled.  At batch 1 the step is ~800 launches of 10-20 us: the hidden launch latency is ~15 % of it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();          // FDG_PDL (default 1)

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface through check_launch (cudaGetLastError)
}

// same, as thread-block clusters of `cluster` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define FDG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      fdg::set_error(__VA_ARGS__);        \
      return FDG_EINVAL;                  \
    }                                     \
  } while (0)

__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float prologue_act(float v, float slope) { return v > 0.f ? v : slope * v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// vectorisable view: unit channel stride, 16-byte aligned base and every other stride a multiple of 4 elements
inline bool vec4_ok(const FdgTensor& t) {
  return t.p != nullptr && t.sc == 1 && aligned16(t.p) && (t.sn % 4 == 0) && (t.sh % 4 == 0) && (t.sw % 4 == 0);
}

}  // namespace fdg
