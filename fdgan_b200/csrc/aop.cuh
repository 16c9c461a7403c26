// The conv A-operand: gather (direct / 2x2 average / nearest x2) + prologue (affine, leaky slope),
// shared by the forward/dgrad kernels and the weight-gradient kernels.
#pragma once
#include "common.cuh"

namespace fdg {

struct AOp {
  FdgTensor x;
  int H, W;  // logical extent
  int gather;
  int has_affine;
  const float* scale;
  const float* shift;
  float slope;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 pro4(const AOp& p, float4 v, int c) {
  if (p.has_affine) {
    const float4 sc = ld4(p.scale + c), sh = ld4(p.shift + c);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
    v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
  }
  const float s = p.slope;
  v.x = prologue_act(v.x, s); v.y = prologue_act(v.y, s);
  v.z = prologue_act(v.z, s); v.w = prologue_act(v.w, s);
  return v;
}

__device__ __forceinline__ float pro1(const AOp& p, float v, int c) {
  if (p.has_affine) v = fmaf(v, __ldg(p.scale + c), __ldg(p.shift + c));
  return prologue_act(v, p.slope);
}

// A-operand element(s) at logical input position (n, iy, ix), channel(s) c.. ; caller guarantees in-range.
__device__ __forceinline__ float4 fetch4(const AOp& p, int n, int iy, int ix, int c) {
  const FdgTensor& x = p.x;
  if (p.gather == FDG_GATHER_AVGPOOL2) {
    const float* b = x.p + n * x.sn + (int64_t)(2 * iy) * x.sh + (int64_t)(2 * ix) * x.sw + c;
    float4 v0 = pro4(p, ld4(b), c), v1 = pro4(p, ld4(b + x.sw), c);
    float4 v2 = pro4(p, ld4(b + x.sh), c), v3 = pro4(p, ld4(b + x.sh + x.sw), c);
    return make_float4(0.25f * ((v0.x + v1.x) + (v2.x + v3.x)), 0.25f * ((v0.y + v1.y) + (v2.y + v3.y)),
                       0.25f * ((v0.z + v1.z) + (v2.z + v3.z)), 0.25f * ((v0.w + v1.w) + (v2.w + v3.w)));
  }
  if (p.gather == FDG_GATHER_UP2) { iy >>= 1; ix >>= 1; }
  return pro4(p, ld4(x.p + n * x.sn + (int64_t)iy * x.sh + (int64_t)ix * x.sw + c), c);
}

__device__ __forceinline__ float fetch1(const AOp& p, int n, int iy, int ix, int c) {
  const FdgTensor& x = p.x;
  if (p.gather == FDG_GATHER_AVGPOOL2) {
    const float* b = x.p + n * x.sn + (int64_t)(2 * iy) * x.sh + (int64_t)(2 * ix) * x.sw + (int64_t)c * x.sc;
    return 0.25f * ((pro1(p, __ldg(b), c) + pro1(p, __ldg(b + x.sw), c)) +
                    (pro1(p, __ldg(b + x.sh), c) + pro1(p, __ldg(b + x.sh + x.sw), c)));
  }
  if (p.gather == FDG_GATHER_UP2) { iy >>= 1; ix >>= 1; }
  return pro1(p, __ldg(x.p + n * x.sn + (int64_t)iy * x.sh + (int64_t)ix * x.sw + (int64_t)c * x.sc), c);
}

inline bool aop_vec_ok(const AOp& a, int Cin) {
  return vec4_ok(a.x) && (Cin % 4 == 0) && (!a.has_affine || (aligned16(a.scale) && aligned16(a.shift)));
}

}  // namespace fdg
