// HBM-bound kernels of the hot path: BatchNorm bookkeeping, the element-wise halves of the
// BatchNorm/ReLU/LeakyReLU backward, max pooling, strided copies, weight repacking and fused Adam.
// All of them stream NHWC rows with 128-bit accesses when the views allow it and reduce per-channel
// partial sums with warp shuffles before touching global memory.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "pack.cuh"

namespace fdg {

// ------------------------------------------------------------------ BatchNorm finalize (forward)
__global__ void bn_finalize_kernel(FdgBnFinalize p) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  float scale, shift;
  if (p.training) {
    const double mean = p.stats[c] / p.count;
    double var = p.stats[p.stats_ld + c] / p.count - mean * mean;  // biased
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)p.eps);
    scale = (float)((double)p.gamma[c] * invstd);
    shift = (float)((double)p.beta[c] - mean * (double)p.gamma[c] * invstd);
    if (p.mean) p.mean[c] = (float)mean;
    if (p.invstd) p.invstd[c] = (float)invstd;
    if (p.running_mean) {
      const double unbiased = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
      p.running_mean[c] = (float)((1.0 - p.momentum) * (double)p.running_mean[c] + p.momentum * mean);
      p.running_var[c] = (float)((1.0 - p.momentum) * (double)p.running_var[c] + p.momentum * unbiased);
    }
  } else {
    const float invstd = 1.0f / sqrtf(p.running_var[c] + p.eps);
    scale = p.gamma[c] * invstd;
    shift = p.beta[c] - p.running_mean[c] * scale;
    if (p.mean) p.mean[c] = p.running_mean[c];
    if (p.invstd) p.invstd[c] = invstd;
  }
  p.scale[c] = scale;
  p.shift[c] = shift;
}

// ------------------------------------------------------------------ BatchNorm finalize (backward)
// dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)),  xhat = (x-mu)*invstd
//    = alpha*dz + beta*x + delta
__global__ void bn_bwd_finalize_kernel(FdgBnBwdFinalize p) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const double s1 = p.stats[c], s2 = p.stats[p.C + c];  // sum dz, sum dz*x
  const double mu = p.mean[c], is = p.invstd[c], g = p.gamma[c];
  const double sum_dz_xhat = (s2 - mu * s1) * is;
  const double m1 = s1 / p.count, m2 = sum_dz_xhat / p.count;
  const double alpha = g * is;
  const double beta = -alpha * m2 * is;
  const double delta = -alpha * m1 + alpha * m2 * is * mu;
  p.coef[c] = p.unit_alpha ? 1.f : (float)alpha;
  p.coef[p.C + c] = (float)beta;
  p.coef[2 * p.C + c] = (float)delta;
  if (p.acc_beta) p.acc_beta[c] += (float)beta;
  if (p.acc_delta) p.acc_delta[c] += (float)delta;
  if (p.dgamma) p.dgamma[c] = (p.accumulate ? p.dgamma[c] : 0.f) + (float)sum_dz_xhat;
  if (p.dbeta) p.dbeta[c] = (p.accumulate ? p.dbeta[c] : 0.f) + (float)s1;
}

// ------------------------------------------------------------------ element-wise backward
// One warp-wide row of channels per pixel group: thread handles channel group cg (VW channels) for
// pixels strided by the number of pixel lanes.  Stats mode keeps per-thread partial sums over its pixels
// and reduces across the pixel lanes of the CTA in shared memory before the fp64 atomics.
template <int VW, int UNR>
__global__ void __launch_bounds__(256) ew_bwd_kernel(FdgEwBwd p, int64_t M, int cgroups, int pix_lanes) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ float sm[];  // stats: [2][pix_lanes][cgroups*VW] partials
  const int cg = threadIdx.x % cgroups;
  const int pl = threadIdx.x / cgroups;
  const int c = (blockIdx.y * cgroups + cg) * VW;
  const bool cv = c < p.C && pl < pix_lanes;
  const int HW = p.H * p.W;
  float sc[VW], sh[VW], ca[VW], cb[VW], cd[VW];
#pragma unroll
  for (int u = 0; u < VW; ++u) {
    const int cc = min(c + u, p.C - 1);
    sc[u] = p.has_affine ? p.scale[cc] : 1.f;
    sh[u] = p.has_affine ? p.shift[cc] : 0.f;
    ca[u] = p.coef ? p.coef[cc] : 1.f;
    cb[u] = p.coef ? p.coef[p.C + cc] : 0.f;
    cd[u] = p.coef ? p.coef[2 * p.C + cc] : 0.f;
  }
  float s1[VW], s2[VW];
#pragma unroll
  for (int u = 0; u < VW; ++u) { s1[u] = 0.f; s2[u] = 0.f; }

  if (cv) {
    // UNR pixels per thread in flight per iteration (loads first, then arithmetic in the same pixel order as a plain loop)
    const bool small = M <= 0x7fffffffLL;
    const int64_t step = (int64_t)gridDim.x * pix_lanes;
    for (int64_t m0 = (int64_t)blockIdx.x * pix_lanes + pl; m0 < M; m0 += UNR * step) {
      float gv[UNR][VW], xv[UNR][VW];
      int64_t oo[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int64_t m = m0 + k * step;
        if (m < M) {
          // 64-bit division costs ~100 instructions; M < 2^31 for every tensor of the path
          const int n = small ? (int)m / HW : (int)(m / HW);
          const int rem = small ? (int)m - n * HW : (int)(m - (int64_t)n * HW);
          const int h = rem / p.W, w = rem - h * p.W;
          const int gh = p.g_gather == FDG_GATHER_UP2 ? h >> 1 : h, gw = p.g_gather == FDG_GATHER_UP2 ? w >> 1 : w;
          const float* gp = p.g.p + n * p.g.sn + (int64_t)gh * p.g.sh + (int64_t)gw * p.g.sw + (int64_t)c * p.g.sc;
          const float* xp = p.x.p + n * p.x.sn + (int64_t)h * p.x.sh + (int64_t)w * p.x.sw + (int64_t)c * p.x.sc;
          oo[k] = n * p.out.sn + (int64_t)h * p.out.sh + (int64_t)w * p.out.sw + (int64_t)c * p.out.sc;
          if (VW == 4) {
            const float4 a = *reinterpret_cast<const float4*>(gp);  // plain load: out may alias g (in-place)
            const float4 b = __ldg(reinterpret_cast<const float4*>(xp));
            gv[k][0] = a.x; gv[k][1] = a.y; gv[k][2] = a.z; gv[k][3] = a.w;
            xv[k][0] = b.x; xv[k][1] = b.y; xv[k][2] = b.z; xv[k][3] = b.w;
          } else {
            gv[k][0] = *gp;
            xv[k][0] = __ldg(xp);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int64_t m = m0 + k * step;
        if (m < M) {
          float o[VW];
#pragma unroll
          for (int u = 0; u < VW; ++u) {
            const float v = fmaf(xv[k][u], sc[u], sh[u]);
            const float dz = p.gscale * gv[k][u] * (v > 0.f ? 1.f : p.slope);
            s1[u] += dz;
            s2[u] += dz * xv[k][u];
            o[u] = fmaf(ca[u], dz, fmaf(cb[u], xv[k][u], cd[u]));
          }
          if (!p.stats) {
            float* op = p.out.p + oo[k];
            if (VW == 4) {
              float4 r = make_float4(o[0], o[1], o[2], o[3]);
              if (p.accumulate) { const float4 old = *reinterpret_cast<const float4*>(op); r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
              *reinterpret_cast<float4*>(op) = r;
            } else {
              *op = p.accumulate ? *op + o[0] : o[0];
            }
          }
        }
      }
    }
  }
  if (p.stats) {
    const int CW = cgroups * VW;
    float* r1 = sm;
    float* r2 = sm + pix_lanes * CW;
    if (pl < pix_lanes) {
#pragma unroll
      for (int u = 0; u < VW; ++u) {
        r1[pl * CW + cg * VW + u] = s1[u];
        r2[pl * CW + cg * VW + u] = s2[u];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CW; i += blockDim.x) {
      const int cc = blockIdx.y * CW + i;
      if (cc < p.C) {
        float a = 0.f, b = 0.f;
        for (int l = 0; l < pix_lanes; ++l) { a += r1[l * CW + i]; b += r2[l * CW + i]; }
        atomicAdd(p.stats + cc, (double)a);
        atomicAdd(p.stats + p.C + cc, (double)b);
      }
    }
  }
}

// Transition backward (gradient at HALF resolution, FDG_GATHER_UP2: the adjoint of the 2x2 average pool folded into the gather):
// a thread owns 4 channels of one pooled pixel, i.e. a 2x2 block of x / out that shares ONE gradient vector -- one gradient load
// and four (stats) or eight (accumulating apply) independent 128-bit loads in flight per thread, no per-pixel index division.
// The generic kernel above ran these six launches at 3.0 TB/s (2.8 ms of the step).
template <bool STATS>
__global__ void __launch_bounds__(256) ew_bwd_pool_kernel(FdgEwBwd p, int64_t MP, int cgroups, int pix_lanes) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ float sm[];
  const int cg = threadIdx.x % cgroups;
  const int pl = threadIdx.x / cgroups;
  const int c = (blockIdx.y * cgroups + cg) * 4;
  const bool cv = c < p.C && pl < pix_lanes;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ca = sc, cb = sh, cd = sh;
  if (cv) {
    if (p.has_affine) { sc = *reinterpret_cast<const float4*>(p.scale + c); sh = *reinterpret_cast<const float4*>(p.shift + c); }
    if (p.coef) {
      ca = *reinterpret_cast<const float4*>(p.coef + c);
      cb = *reinterpret_cast<const float4*>(p.coef + p.C + c);
      cd = *reinterpret_cast<const float4*>(p.coef + 2 * p.C + c);
    }
  }
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float gs = p.gscale, sl = p.slope;
  const int PH = p.H >> 1, PW = p.W >> 1, PHW = PH * PW;
  if (cv) {
    const int64_t step = (int64_t)gridDim.x * pix_lanes;
    for (int64_t m = (int64_t)blockIdx.x * pix_lanes + pl; m < MP; m += step) {
      const int n = (int)(m / PHW);
      const int rem = (int)(m - (int64_t)n * PHW);
      const int ph = rem / PW, pw = rem - ph * PW;
      const float4 g = *reinterpret_cast<const float4*>(p.g.p + n * p.g.sn + (int64_t)ph * p.g.sh + (int64_t)pw * p.g.sw + c);
      const float* xb = p.x.p + n * p.x.sn + (int64_t)(2 * ph) * p.x.sh + (int64_t)(2 * pw) * p.x.sw + c;
      float4 xv[4];
      xv[0] = __ldg(reinterpret_cast<const float4*>(xb));
      xv[1] = __ldg(reinterpret_cast<const float4*>(xb + p.x.sw));
      xv[2] = __ldg(reinterpret_cast<const float4*>(xb + p.x.sh));
      xv[3] = __ldg(reinterpret_cast<const float4*>(xb + p.x.sh + p.x.sw));
      float* ob = nullptr;
      float4 old[4];
      if (!STATS) {
        ob = p.out.p + n * p.out.sn + (int64_t)(2 * ph) * p.out.sh + (int64_t)(2 * pw) * p.out.sw + c;
        if (p.accumulate) {
          old[0] = *reinterpret_cast<const float4*>(ob);
          old[1] = *reinterpret_cast<const float4*>(ob + p.out.sw);
          old[2] = *reinterpret_cast<const float4*>(ob + p.out.sh);
          old[3] = *reinterpret_cast<const float4*>(ob + p.out.sh + p.out.sw);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 dz;
        dz.x = gs * g.x * (fmaf(xv[k].x, sc.x, sh.x) > 0.f ? 1.f : sl);
        dz.y = gs * g.y * (fmaf(xv[k].y, sc.y, sh.y) > 0.f ? 1.f : sl);
        dz.z = gs * g.z * (fmaf(xv[k].z, sc.z, sh.z) > 0.f ? 1.f : sl);
        dz.w = gs * g.w * (fmaf(xv[k].w, sc.w, sh.w) > 0.f ? 1.f : sl);
        if (STATS) {
          s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
          s2.x = fmaf(dz.x, xv[k].x, s2.x); s2.y = fmaf(dz.y, xv[k].y, s2.y);
          s2.z = fmaf(dz.z, xv[k].z, s2.z); s2.w = fmaf(dz.w, xv[k].w, s2.w);
        } else {
          float4 o;
          o.x = fmaf(ca.x, dz.x, fmaf(cb.x, xv[k].x, cd.x)); o.y = fmaf(ca.y, dz.y, fmaf(cb.y, xv[k].y, cd.y));
          o.z = fmaf(ca.z, dz.z, fmaf(cb.z, xv[k].z, cd.z)); o.w = fmaf(ca.w, dz.w, fmaf(cb.w, xv[k].w, cd.w));
          if (p.accumulate) { o.x += old[k].x; o.y += old[k].y; o.z += old[k].z; o.w += old[k].w; }
          *reinterpret_cast<float4*>(ob + (int64_t)(k >> 1) * p.out.sh + (int64_t)(k & 1) * p.out.sw) = o;
        }
      }
    }
  }
  if (STATS) {
    const int CW = cgroups * 4;
    float* r1 = sm;
    float* r2 = sm + pix_lanes * CW;
    if (pl < pix_lanes) {
      *reinterpret_cast<float4*>(&r1[pl * CW + cg * 4]) = s1;
      *reinterpret_cast<float4*>(&r2[pl * CW + cg * 4]) = s2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CW; i += blockDim.x) {
      const int cc = blockIdx.y * CW + i;
      if (cc < p.C) {
        float a = 0.f, b = 0.f;
        for (int l = 0; l < pix_lanes; ++l) { a += r1[l * CW + i]; b += r2[l * CW + i]; }
        atomicAdd(p.stats + cc, (double)a);
        atomicAdd(p.stats + p.C + cc, (double)b);
      }
    }
  }
}

// Fast path of ew_bwd: all views are pixel-linear (address = base + pixel * row stride + channel), 4 channels per
// thread, four pixels per thread in flight per iteration (8 independent 128-bit loads).
template <bool STATS>
__global__ void __launch_bounds__(256) ew_bwd_linear_kernel(FdgEwBwd p, int64_t M, int cgroups, int pix_lanes) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ float sm[];
  const int cg = threadIdx.x % cgroups;
  const int pl = threadIdx.x / cgroups;
  const int c = (blockIdx.y * cgroups + cg) * 4;
  const bool cv = c < p.C && pl < pix_lanes;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ca = sc, cb = sh, cd = sh;
  if (cv) {
    if (p.has_affine) { sc = *reinterpret_cast<const float4*>(p.scale + c); sh = *reinterpret_cast<const float4*>(p.shift + c); }
    if (p.coef) {
      ca = *reinterpret_cast<const float4*>(p.coef + c);
      cb = *reinterpret_cast<const float4*>(p.coef + p.C + c);
      cd = *reinterpret_cast<const float4*>(p.coef + 2 * p.C + c);
    }
  }
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  const float gs = p.gscale, sl = p.slope;
  if (cv) {
    const int64_t step = (int64_t)gridDim.x * pix_lanes;
    const float* gbase = p.g.p + c;
    const float* xbase = p.x.p + c;
    float* obase = p.out.p ? p.out.p + c : nullptr;
    for (int64_t m0 = (int64_t)blockIdx.x * pix_lanes + pl; m0 < M; m0 += 4 * step) {
      float4 gv[4], xv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t m = m0 + k * step;
        if (m < M) {
          gv[k] = *reinterpret_cast<const float4*>(gbase + m * p.g.sw);
          xv[k] = __ldg(reinterpret_cast<const float4*>(xbase + m * p.x.sw));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t m = m0 + k * step;
        if (m < M) {
          float4 dz;
          dz.x = gs * gv[k].x * (fmaf(xv[k].x, sc.x, sh.x) > 0.f ? 1.f : sl);
          dz.y = gs * gv[k].y * (fmaf(xv[k].y, sc.y, sh.y) > 0.f ? 1.f : sl);
          dz.z = gs * gv[k].z * (fmaf(xv[k].z, sc.z, sh.z) > 0.f ? 1.f : sl);
          dz.w = gs * gv[k].w * (fmaf(xv[k].w, sc.w, sh.w) > 0.f ? 1.f : sl);
          if (STATS) {
            s1.x += dz.x; s1.y += dz.y; s1.z += dz.z; s1.w += dz.w;
            s2.x = fmaf(dz.x, xv[k].x, s2.x); s2.y = fmaf(dz.y, xv[k].y, s2.y);
            s2.z = fmaf(dz.z, xv[k].z, s2.z); s2.w = fmaf(dz.w, xv[k].w, s2.w);
          } else {
            float4 o;
            o.x = fmaf(ca.x, dz.x, fmaf(cb.x, xv[k].x, cd.x)); o.y = fmaf(ca.y, dz.y, fmaf(cb.y, xv[k].y, cd.y));
            o.z = fmaf(ca.z, dz.z, fmaf(cb.z, xv[k].z, cd.z)); o.w = fmaf(ca.w, dz.w, fmaf(cb.w, xv[k].w, cd.w));
            if (p.out_split) {      // split-bf16 planes [M][C]: hi, then lo (the tensor-core kernels' operand format)
              __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(p.out_split) + m * p.C + c;
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
              const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
              const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
              uint2 hv, lv;
              hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
              lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
              *reinterpret_cast<uint2*>(hp) = hv;
              *reinterpret_cast<uint2*>(hp + M * p.C) = lv;
              if (!obase) continue;      // planes only; with `out` given as well the fp32 tensor is written too (a second consumer reads fp32)
            }
            float* op = obase + m * p.out.sw;
            if (p.accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(op);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(op) = o;
          }
        }
      }
    }
  }
  if (STATS) {
    const int CW = cgroups * 4;
    float* r1 = sm;
    float* r2 = sm + pix_lanes * CW;
    if (pl < pix_lanes) {
      *reinterpret_cast<float4*>(&r1[pl * CW + cg * 4]) = s1;
      *reinterpret_cast<float4*>(&r2[pl * CW + cg * 4]) = s2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CW; i += blockDim.x) {
      const int cc = blockIdx.y * CW + i;
      if (cc < p.C) {
        float a = 0.f, b = 0.f;
        for (int l = 0; l < pix_lanes; ++l) { a += r1[l * CW + i]; b += r2[l * CW + i]; }
        atomicAdd(p.stats + cc, (double)a);
        atomicAdd(p.stats + p.C + cc, (double)b);
      }
    }
  }
}

// ------------------------------------------------------------------ deferred affine part of the BatchNorm backward
__global__ void __launch_bounds__(256) affine_accum_kernel(FdgTensor x, FdgTensor out, int64_t total4, int H, int W, int C4, const float* cb,
                                                           const float* cd) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t r = i / C4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x.p + n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(cb + c)), d = __ldg(reinterpret_cast<const float4*>(cd + c));
    float4* op = reinterpret_cast<float4*>(out.p + n * out.sn + (int64_t)h * out.sh + (int64_t)w * out.sw + c);
    float4 o = *op;
    o.x += fmaf(b.x, xv.x, d.x); o.y += fmaf(b.y, xv.y, d.y); o.z += fmaf(b.z, xv.z, d.z); o.w += fmaf(b.w, xv.w, d.w);
    *op = o;
  }
}

// ------------------------------------------------------------------ max pool 2x2
__global__ void maxpool2_fwd_kernel(FdgTensor x, FdgTensor y, int64_t total, int OH, int OW, int C) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    const float* b = x.p + n * x.sn + (int64_t)(2 * oh) * x.sh + (int64_t)(2 * ow) * x.sw + (int64_t)c * x.sc;
    const float v = fmaxf(fmaxf(__ldg(b), __ldg(b + x.sw)), fmaxf(__ldg(b + x.sh), __ldg(b + x.sh + x.sw)));
    y.p[n * y.sn + (int64_t)oh * y.sh + (int64_t)ow * y.sw + (int64_t)c * y.sc] = v;
  }
}

__global__ void maxpool2_bwd_kernel(FdgTensor x, FdgTensor gy, FdgTensor gx, int64_t total, int OH, int OW, int C,
                                    int accumulate) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    const float* b = x.p + n * x.sn + (int64_t)(2 * oh) * x.sh + (int64_t)(2 * ow) * x.sw + (int64_t)c * x.sc;
    const float v[4] = {__ldg(b), __ldg(b + x.sw), __ldg(b + x.sh), __ldg(b + x.sh + x.sw)};
    int arg = 0;
    float best = v[0];
#pragma unroll
    for (int d = 1; d < 4; ++d) if (v[d] > best) { best = v[d]; arg = d; }  // first maximum, scan order
    const float g = __ldg(gy.p + n * gy.sn + (int64_t)oh * gy.sh + (int64_t)ow * gy.sw + (int64_t)c * gy.sc);
    float* o = gx.p + n * gx.sn + (int64_t)(2 * oh) * gx.sh + (int64_t)(2 * ow) * gx.sw + (int64_t)c * gx.sc;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float* q = o + (d >> 1) * gx.sh + (d & 1) * gx.sw;
      const float val = (d == arg && (!(accumulate & 2) || best > 0.f)) ? g : 0.f;
      *q = (accumulate & 1) ? *q + val : val;
    }
  }
}

// 128-bit variants (unit channel stride, C % 4 == 0): thread = four channels of one pooled pixel, 32-bit index arithmetic.
// Backward flags: bit 0 accumulate, bit 1 multiply by [max > 0] (the ReLU mask of the pooled tensor: Vgg16 pools post-ReLU stages,
// so the separate mask pass over the un-pooled gradient disappears).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(256) maxpool2_fwd_vec4_kernel(FdgTensor x, FdgTensor y, int total4, int OH, int OW, int C4) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const int c = (i % C4) * 4;
    int r = i / C4;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH, n = r / OH;
    const float* b = x.p + n * x.sn + (int64_t)(2 * oh) * x.sh + (int64_t)(2 * ow) * x.sw + c;
    const float4 v0 = ldg4(b), v1 = ldg4(b + x.sw), v2 = ldg4(b + x.sh), v3 = ldg4(b + x.sh + x.sw);
    float4 m;
    m.x = fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x)); m.y = fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y));
    m.z = fmaxf(fmaxf(v0.z, v1.z), fmaxf(v2.z, v3.z)); m.w = fmaxf(fmaxf(v0.w, v1.w), fmaxf(v2.w, v3.w));
    *reinterpret_cast<float4*>(y.p + n * y.sn + (int64_t)oh * y.sh + (int64_t)ow * y.sw + c) = m;
  }
}

__device__ __forceinline__ void maxpool_route(float a, float b, float c, float d, float g, bool mask, float& o0, float& o1, float& o2, float& o3) {
  int arg = 0;
  float best = a;
  if (b > best) { best = b; arg = 1; }      // first maximum, scan order
  if (c > best) { best = c; arg = 2; }
  if (d > best) { best = d; arg = 3; }
  if (mask && !(best > 0.f)) g = 0.f;
  o0 = arg == 0 ? g : 0.f; o1 = arg == 1 ? g : 0.f; o2 = arg == 2 ? g : 0.f; o3 = arg == 3 ? g : 0.f;
}

__global__ void __launch_bounds__(256) maxpool2_bwd_vec4_kernel(FdgTensor x, FdgTensor gy, FdgTensor gx, int total4, int OH, int OW, int C4,
                                                                int flags) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const bool acc = flags & 1, mask = flags & 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const int c = (i % C4) * 4;
    int r = i / C4;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH, n = r / OH;
    const float* b = x.p + n * x.sn + (int64_t)(2 * oh) * x.sh + (int64_t)(2 * ow) * x.sw + c;
    const float4 v0 = ldg4(b), v1 = ldg4(b + x.sw), v2 = ldg4(b + x.sh), v3 = ldg4(b + x.sh + x.sw);
    const float4 g = ldg4(gy.p + n * gy.sn + (int64_t)oh * gy.sh + (int64_t)ow * gy.sw + c);
    float4 o0, o1, o2, o3;
    maxpool_route(v0.x, v1.x, v2.x, v3.x, g.x, mask, o0.x, o1.x, o2.x, o3.x);
    maxpool_route(v0.y, v1.y, v2.y, v3.y, g.y, mask, o0.y, o1.y, o2.y, o3.y);
    maxpool_route(v0.z, v1.z, v2.z, v3.z, g.z, mask, o0.z, o1.z, o2.z, o3.z);
    maxpool_route(v0.w, v1.w, v2.w, v3.w, g.w, mask, o0.w, o1.w, o2.w, o3.w);
    float* o = gx.p + n * gx.sn + (int64_t)(2 * oh) * gx.sh + (int64_t)(2 * ow) * gx.sw + c;
    float4* q0 = reinterpret_cast<float4*>(o), *q1 = reinterpret_cast<float4*>(o + gx.sw);
    float4* q2 = reinterpret_cast<float4*>(o + gx.sh), *q3 = reinterpret_cast<float4*>(o + gx.sh + gx.sw);
    if (acc) {
      const float4 p0 = *q0, p1 = *q1, p2 = *q2, p3 = *q3;
      o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w; o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
      o2.x += p2.x; o2.y += p2.y; o2.z += p2.z; o2.w += p2.w; o3.x += p3.x; o3.y += p3.y; o3.z += p3.z; o3.w += p3.w;
    }
    *q0 = o0; *q1 = o1; *q2 = o2; *q3 = o3;
  }
}

// ------------------------------------------------------------------ strided gather-copy with leaky slope
__global__ void copy4d_kernel(FdgTensor x, FdgTensor y, int64_t total, int H, int W, int C, int gather, float slope,
                              float scale, int accumulate) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float v;
    if (gather == FDG_GATHER_AVGPOOL2) {
      const float* b = x.p + n * x.sn + (int64_t)(2 * h) * x.sh + (int64_t)(2 * w) * x.sw + (int64_t)c * x.sc;
      const float v0 = __ldg(b), v1 = __ldg(b + x.sw), v2 = __ldg(b + x.sh), v3 = __ldg(b + x.sh + x.sw);
      v = 0.25f * ((prologue_act(v0, slope) + prologue_act(v1, slope)) + (prologue_act(v2, slope) + prologue_act(v3, slope)));
    } else {
      const int hh = gather == FDG_GATHER_UP2 ? h >> 1 : h, ww = gather == FDG_GATHER_UP2 ? w >> 1 : w;
      v = prologue_act(__ldg(x.p + n * x.sn + (int64_t)hh * x.sh + (int64_t)ww * x.sw + (int64_t)c * x.sc), slope);
    }
    v *= scale;
    float* o = y.p + n * y.sn + (int64_t)h * y.sh + (int64_t)w * y.sw + (int64_t)c * y.sc;
    *o = accumulate ? *o + v : v;
  }
}

// same arithmetic, four channels per thread: both views have unit channel stride, 16-byte aligned rows and C % 4 == 0
// (channel slices of the NHWC concat buffers: every copy of the generator's forward / backward walk)
__global__ void __launch_bounds__(256) copy4d_vec4_kernel(FdgTensor x, FdgTensor y, int64_t total4, int H, int W, int C4, int gather,
                                                          float slope, float scale, int accumulate) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    int c, w, h, n;
    if (total4 <= 0x7fffffffLL) {     // 32-bit index arithmetic (a 64-bit division costs ~100 instructions)
      const int i32 = (int)i, r0 = i32 / C4, r1 = r0 / W;
      c = (i32 - r0 * C4) * 4; w = r0 - r1 * W; n = r1 / H; h = r1 - n * H;
    } else {
      c = (int)(i % C4) * 4;
      int64_t r = i / C4;
      w = (int)(r % W); r /= W;
      h = (int)(r % H);
      n = (int)(r / H);
    }
    float4 v;
    if (gather == FDG_GATHER_AVGPOOL2) {
      const float* b = x.p + n * x.sn + (int64_t)(2 * h) * x.sh + (int64_t)(2 * w) * x.sw + c;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(b)), v1 = __ldg(reinterpret_cast<const float4*>(b + x.sw));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(b + x.sh)), v3 = __ldg(reinterpret_cast<const float4*>(b + x.sh + x.sw));
      v.x = 0.25f * ((prologue_act(v0.x, slope) + prologue_act(v1.x, slope)) + (prologue_act(v2.x, slope) + prologue_act(v3.x, slope)));
      v.y = 0.25f * ((prologue_act(v0.y, slope) + prologue_act(v1.y, slope)) + (prologue_act(v2.y, slope) + prologue_act(v3.y, slope)));
      v.z = 0.25f * ((prologue_act(v0.z, slope) + prologue_act(v1.z, slope)) + (prologue_act(v2.z, slope) + prologue_act(v3.z, slope)));
      v.w = 0.25f * ((prologue_act(v0.w, slope) + prologue_act(v1.w, slope)) + (prologue_act(v2.w, slope) + prologue_act(v3.w, slope)));
    } else {
      const int hh = gather == FDG_GATHER_UP2 ? h >> 1 : h, ww = gather == FDG_GATHER_UP2 ? w >> 1 : w;
      v = __ldg(reinterpret_cast<const float4*>(x.p + n * x.sn + (int64_t)hh * x.sh + (int64_t)ww * x.sw + c));
      v.x = prologue_act(v.x, slope); v.y = prologue_act(v.y, slope); v.z = prologue_act(v.z, slope); v.w = prologue_act(v.w, slope);
    }
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    float4* o = reinterpret_cast<float4*>(y.p + n * y.sn + (int64_t)h * y.sh + (int64_t)w * y.sw + c);
    if (accumulate) {
      const float4 q = *o;
      v.x = q.x + v.x; v.y = q.y + v.y; v.z = q.z + v.z; v.w = q.w + v.w;
    }
    *o = v;
  }
}

__global__ void act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ out, int64_t n,
                               int act) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float yy = y[i];
    out[i] = g[i] * (act == FDG_ACT_TANH ? (1.f - yy * yy) : yy * (1.f - yy));
  }
}

// transposed convolution as a gather (data gradient of a strided conv); one thread per (pixel, ci)
__global__ void dgrad_strided_kernel(FdgDgradStrided p, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % p.Cin);
    int64_t r = i / p.Cin;
    const int x = (int)(r % p.W); r /= p.W;
    const int y = (int)(r % p.H);
    const int n = (int)(r / p.H);
    float acc = 0.f;
    for (int kr = 0; kr < p.R; ++kr) {
      const int ty = y + p.pad - kr;
      if (ty < 0 || ty % p.stride != 0) continue;
      const int oy = ty / p.stride;
      if (oy >= p.OH) continue;
      for (int ks = 0; ks < p.S; ++ks) {
        const int tx = x + p.pad - ks;
        if (tx < 0 || tx % p.stride != 0) continue;
        const int ox = tx / p.stride;
        if (ox >= p.OW) continue;
        const float* gp = p.g.p + n * p.g.sn + (int64_t)oy * p.g.sh + (int64_t)ox * p.g.sw;
        const float* wp = p.w + ((int64_t)ci * p.R + kr) * p.S + ks;
        for (int co = 0; co < p.Cout; ++co)
          acc = fmaf(__ldg(gp + (int64_t)co * p.g.sc), __ldg(wp + (int64_t)co * p.Cin * p.R * p.S), acc);
      }
    }
    float* o = p.dx.p + n * p.dx.sn + (int64_t)y * p.dx.sh + (int64_t)x * p.dx.sw + (int64_t)ci * p.dx.sc;
    *o = p.accumulate ? *o + acc : acc;
  }
}

// ------------------------------------------------------------------ pooled BatchNorm + activation (transition blocks)
// y(n, h, w, c) = mean over the 2x2 block of act(x * scale[c] + shift[c]): the input of a torchvision transition's 1x1 convolution
// with the average pool commuted in front of it (models/densenet.py:214-221), materialised ONCE.  Forward convolution and weight
// gradient then run as plain 1x1 kernels on a quarter of the pixels instead of gathering 2x2 blocks through their loaders
// (four loads + prologue per operand element).  128-bit accesses; x, y: unit channel stride, 16-byte aligned, C % 4 == 0.
__global__ void __launch_bounds__(256) pool2_bn_act_kernel(FdgTensor x, FdgTensor y, int64_t total4, int OH, int OW, int C4, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float slope) {
  pdl_wait();      // PDL contract (common.cuh)
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t r = i / C4;
    const int w = (int)(r % OW); r /= OW;
    const int h = (int)(r % OH);
    const int n = (int)(r / OH);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) { sc = __ldg(reinterpret_cast<const float4*>(scale + c)); sh = __ldg(reinterpret_cast<const float4*>(shift + c)); }
    const float* b = x.p + n * x.sn + (int64_t)(2 * h) * x.sh + (int64_t)(2 * w) * x.sw + c;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(b)), v1 = __ldg(reinterpret_cast<const float4*>(b + x.sw));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(b + x.sh)), v3 = __ldg(reinterpret_cast<const float4*>(b + x.sh + x.sw));
    float4 o;   // same association as the gather of the convolution loaders (aop.cuh:fetch4): (v0 + v1) + (v2 + v3)
    o.x = 0.25f * ((prologue_act(fmaf(v0.x, sc.x, sh.x), slope) + prologue_act(fmaf(v1.x, sc.x, sh.x), slope)) +
                   (prologue_act(fmaf(v2.x, sc.x, sh.x), slope) + prologue_act(fmaf(v3.x, sc.x, sh.x), slope)));
    o.y = 0.25f * ((prologue_act(fmaf(v0.y, sc.y, sh.y), slope) + prologue_act(fmaf(v1.y, sc.y, sh.y), slope)) +
                   (prologue_act(fmaf(v2.y, sc.y, sh.y), slope) + prologue_act(fmaf(v3.y, sc.y, sh.y), slope)));
    o.z = 0.25f * ((prologue_act(fmaf(v0.z, sc.z, sh.z), slope) + prologue_act(fmaf(v1.z, sc.z, sh.z), slope)) +
                   (prologue_act(fmaf(v2.z, sc.z, sh.z), slope) + prologue_act(fmaf(v3.z, sc.z, sh.z), slope)));
    o.w = 0.25f * ((prologue_act(fmaf(v0.w, sc.w, sh.w), slope) + prologue_act(fmaf(v1.w, sc.w, sh.w), slope)) +
                   (prologue_act(fmaf(v2.w, sc.w, sh.w), slope) + prologue_act(fmaf(v3.w, sc.w, sh.w), slope)));
    *reinterpret_cast<float4*>(y.p + n * y.sn + (int64_t)h * y.sh + (int64_t)w * y.sw + c) = o;
  }
}

// ------------------------------------------------------------------ single-output-channel convolutions by taps
// A stride-1 RxS convolution with ONE output channel (Fusion-D layer 5: 288 -> 1, 4x4) is a GEMV per pixel; on the implicit-GEMM
// kernels it pads N to 32 and streams R*S weight tiles.  Re-associated, it is a 1x1 convolution Cin -> R*S (one column per tap:
// s[p][t] = sum_c a[p][c] w[c][t], a dense GEMM that reads the input once) followed by a sum of the taps over shifted pixels:
//   fwd:   out(n, oy, ox) = act( sum_t s(n, oy + ky - pad, ox + kx - pad)[t] )                       (fdg_tap_sum)
//   wgrad: dW[c][t] = sum_p a[p][c] G(p)[t],  G(n, y, x)[t] = g(n, y - ky + pad, x - kx + pad)        (fdg_tap_spread + 1x1 wgrad)
__global__ void __launch_bounds__(256) tap_sum_kernel(FdgTensor s, FdgTensor out, int64_t total, int H, int W, int R, int S, int pad, int OH,
                                                      int OW, int act) {
  pdl_wait();      // PDL contract (common.cuh)
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    int64_t q = i / OW;
    const int oy = (int)(q % OH);
    const int n = (int)(q / OH);
    float acc = 0.f;
    for (int ky = 0; ky < R; ++ky) {
      const int y = oy + ky - pad;
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < S; ++kx) {
        const int x = ox + kx - pad;
        if (x < 0 || x >= W) continue;
        acc += __ldg(s.p + n * s.sn + (int64_t)y * s.sh + (int64_t)x * s.sw + (int64_t)(ky * S + kx) * s.sc);
      }
    }
    if (act == FDG_ACT_RELU) acc = fmaxf(acc, 0.f);
    else if (act == FDG_ACT_TANH) acc = tanhf(acc);
    else if (act == FDG_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    out.p[n * out.sn + (int64_t)oy * out.sh + (int64_t)ox * out.sw] = acc;
  }
}

__global__ void __launch_bounds__(256) tap_spread_kernel(FdgTensor g, FdgTensor gs, int64_t total, int H, int W, int R, int S, int pad, int OH,
                                                         int OW) {
  pdl_wait();      // PDL contract (common.cuh)
  pdl_trigger();
  const int T = R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    int64_t q = i / T;
    const int x = (int)(q % W); q /= W;
    const int y = (int)(q % H);
    const int n = (int)(q / H);
    const int ky = t / S, kx = t - ky * S;
    const int oy = y - ky + pad, ox = x - kx + pad;
    float v = 0.f;
    if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) v = __ldg(g.p + n * g.sn + (int64_t)oy * g.sh + (int64_t)ox * g.sw);
    gs.p[n * gs.sn + (int64_t)y * gs.sh + (int64_t)x * gs.sw + (int64_t)t * gs.sc] = v;
  }
}

// ------------------------------------------------------------------ weight repack
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S, int mode,
                                   float* __restrict__ out, int out_ld, int64_t total) {
  // total = K * out_ld over the destination; gather from the source layout
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    pack_w_item(w, Cout, Cin, R, S, mode, out, out_ld, i);
}

// ------------------------------------------------------------------ fused Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float gscale) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

// Device-side step counter (CUDA-graph replay: no host value may change between replays).  state = {step, bc1, sqrt(bc2)}
__global__ void adam_prepare_kernel(float* state, float b1, float b2) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const float step = state[0] + 1.f;
  state[0] = step;
  state[1] = 1.f - powf(b1, step);
  state[2] = sqrtf(1.f - powf(b2, step));
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                                float lr, float b1, float b2, float eps, const float* __restrict__ state, float gscale) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const float bc1 = state[1], bc2_sqrt = state[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

// ------------------------------------------------------------------ fused loss value + gradient
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, float target,
                                                        int kind, int64_t n, float scale, float* __restrict__ grad,
                                                        int accumulate, double* loss) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  float part = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float av = a[i];
    float l, g;
    if (kind == FDG_LOSS_L1) {
      const float d = av - b[i];
      l = fabsf(d);
      g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else if (kind == FDG_LOSS_MSE) {
      const float d = av - b[i];
      l = d * d;
      g = 2.f * d;
    } else {
      const float la = fmaxf(logf(av), -100.f), l1a = fmaxf(log1pf(-av), -100.f);
      l = -(target * la + (1.f - target) * l1a);
      g = (av - target) / fmaxf(av * (1.f - av), 1e-12f);
    }
    part += l;
    if (grad) grad[i] = (accumulate ? grad[i] : 0.f) + scale * g;
  }
  part = warp_sum(part);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(loss, (double)t * (double)scale);
  }
}

static inline unsigned grid_for(int64_t total, int block) {
  int64_t g = cdiv64(total, block);
  const int64_t cap = 148 * 16;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fdg

using namespace fdg;

namespace fdg { int dgrad_strided_small(const FdgDgradStrided* p, cudaStream_t st); }

extern "C" {

int fdg_bn_finalize(const FdgBnFinalize* p, fdg_stream_t stream) {
  FDG_REQUIRE(p && p->C > 0 && p->gamma && p->beta && p->scale && p->shift, "fdg_bn_finalize: bad arguments");
  FDG_REQUIRE(!p->training || (p->stats && p->count > 0), "fdg_bn_finalize: training mode needs stats and count");
  FDG_REQUIRE(p->training || (p->running_mean && p->running_var), "fdg_bn_finalize: eval mode needs running stats");
  launch_k(bn_finalize_kernel, dim3(cdiv(p->C, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, *p);
  return check_launch("fdg_bn_finalize");
}

int fdg_bn_bwd_finalize(const FdgBnBwdFinalize* p, fdg_stream_t stream) {
  FDG_REQUIRE(p && p->C > 0 && p->stats && p->gamma && p->mean && p->invstd && p->coef && p->count > 0,
              "fdg_bn_bwd_finalize: bad arguments");
  launch_k(bn_bwd_finalize_kernel, dim3(cdiv(p->C, 128)), dim3(128), (size_t)(0), (cudaStream_t)stream, *p);
  return check_launch("fdg_bn_bwd_finalize");
}

int fdg_ew_bwd(const FdgEwBwd* p, fdg_stream_t stream) {
  FDG_REQUIRE(p && p->g.p && p->x.p && p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0, "fdg_ew_bwd: bad arguments");
  FDG_REQUIRE(p->stats || p->out.p || p->out_split, "fdg_ew_bwd: neither stats nor out given");
  FDG_REQUIRE(!p->has_affine || (p->scale && p->shift), "fdg_ew_bwd: affine without scale/shift");
  FDG_REQUIRE(p->g_gather == FDG_GATHER_DIRECT || p->g_gather == FDG_GATHER_UP2, "fdg_ew_bwd: bad gather");
  const bool vec = vec4_ok(p->g) && vec4_ok(p->x) && (p->stats || p->out_split || vec4_ok(p->out)) && (p->C % 4 == 0);
  const int VW = vec ? 4 : 1;
  const int groups_total = cdiv(p->C, VW);
  const int cgroups = groups_total < 64 ? groups_total : 64;   // channel groups per CTA
  const int pix_lanes = 256 / cgroups;
  const int64_t M = (int64_t)p->N * p->H * p->W;
  int64_t gx = cdiv64(M, (int64_t)pix_lanes * 8);
  const int gy = cdiv(groups_total, cgroups);
  const int64_t cap = (148 * 8) / gy + 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const size_t smem = p->stats ? sizeof(float) * 2 * pix_lanes * cgroups * VW : 0;
  dim3 grid((unsigned)gx, gy);
  ProfScope prof(PF_EW, 4.0 * (double)M * p->C, 4.0 * (double)M * p->C * (p->stats ? 2.0 : (p->accumulate ? 4.0 : 3.0)),
                 (cudaStream_t)stream);
  auto linear = [&](const FdgTensor& t) { return t.sh == (int64_t)p->W * t.sw && t.sn == (int64_t)p->H * t.sh; };
  const bool fast = vec && p->g_gather == FDG_GATHER_DIRECT && linear(p->g) && linear(p->x) && (p->stats || p->out_split || linear(p->out)) &&
                    (!p->has_affine || (aligned16(p->scale) && aligned16(p->shift))) && (!p->coef || (aligned16(p->coef) && p->C % 4 == 0));
  if (p->out_split && !p->stats) {
    FDG_REQUIRE(fast && !p->accumulate && aligned16(p->out_split),
                "fdg_ew_bwd: out_split needs the pixel-linear 128-bit path (unit channel stride, C %% 4 == 0, direct gather) and no accumulate");
  }
  if (fast) {
    if (p->stats) launch_k(ew_bwd_linear_kernel<true>, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, *p, M, cgroups, pix_lanes);
    else launch_k(ew_bwd_linear_kernel<false>, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, *p, M, cgroups, pix_lanes);
  } else if (vec && p->g_gather == FDG_GATHER_UP2 && p->H % 2 == 0 && p->W % 2 == 0 && !p->out_split &&
             (!p->has_affine || (aligned16(p->scale) && aligned16(p->shift))) && (!p->coef || aligned16(p->coef))) {
    // transition backward: one thread per pooled pixel and 4 channels (2x2 block of x / out)
    const int64_t MP = M / 4;
    int64_t gxp = cdiv64(MP, (int64_t)pix_lanes * 4);
    if (gxp > cap) gxp = cap;
    if (gxp < 1) gxp = 1;
    dim3 gridp((unsigned)gxp, gy);
    if (p->stats) launch_k(ew_bwd_pool_kernel<true>, gridp, dim3(256), smem, (cudaStream_t)stream, *p, MP, cgroups, pix_lanes);
    else launch_k(ew_bwd_pool_kernel<false>, gridp, dim3(256), (size_t)0, (cudaStream_t)stream, *p, MP, cgroups, pix_lanes);
  } else if (vec) {
    // measured on the three transition backward passes (gradient gathered at half resolution): four pixels in flight cost
    // 104 registers and occupancy, 3.16 ms / step against 2.79 ms for the plain loop (profiles/r01h_launches_final.md)
    static const int unr = [] { const char* e = getenv("FDG_EW_UNR"); return e ? atoi(e) : 1; }();
    if (unr == 4) launch_k(ew_bwd_kernel<4, 4>, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, *p, M, cgroups, pix_lanes);
    else launch_k(ew_bwd_kernel<4, 1>, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, *p, M, cgroups, pix_lanes);
  } else launch_k(ew_bwd_kernel<1, 4>, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, *p, M, cgroups, pix_lanes);
  return check_launch("fdg_ew_bwd");
}

int fdg_affine_accum(const FdgTensor* x, const FdgTensor* out, int N, int H, int W, int C, const float* cb, const float* cd,
                     fdg_stream_t stream) {
  FDG_REQUIRE(x && out && x->p && out->p && cb && cd && N > 0 && H > 0 && W > 0 && C > 0, "fdg_affine_accum: bad arguments");
  FDG_REQUIRE(C % 4 == 0 && vec4_ok(*x) && vec4_ok(*out) && aligned16(cb) && aligned16(cd),
              "fdg_affine_accum: needs unit-stride channels, C % 4 == 0 and 16-byte aligned views / vectors");
  const int64_t total4 = (int64_t)N * H * W * (C / 4);
  ProfScope prof(PF_EW, 2.0 * (double)total4 * 4, 4.0 * (double)total4 * 4 * 3.0, (cudaStream_t)stream);
  launch_k(affine_accum_kernel, dim3(grid_for(total4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *out, total4, H, W, C / 4, cb, cd);
  return check_launch("fdg_affine_accum");
}

int fdg_maxpool2_fwd(const FdgTensor* x, const FdgTensor* y, int N, int OH, int OW, int C, fdg_stream_t stream) {
  FDG_REQUIRE(x && y && x->p && y->p && N > 0 && OH > 0 && OW > 0 && C > 0, "fdg_maxpool2_fwd: bad arguments");
  const int64_t total = (int64_t)N * OH * OW * C;
  if (C % 4 == 0 && fdg::vec4_ok(*x) && fdg::vec4_ok(*y) && total / 4 < (1ll << 31)) {
    launch_k(maxpool2_fwd_vec4_kernel, dim3(grid_for(total / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *y, (int)(total / 4), OH, OW, C / 4);
    return check_launch("fdg_maxpool2_fwd");
  }
  launch_k(maxpool2_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *y, total, OH, OW, C);
  return check_launch("fdg_maxpool2_fwd");
}

int fdg_maxpool2_bwd(const FdgTensor* x, const FdgTensor* gy, const FdgTensor* gx, int N, int OH, int OW, int C,
                     int accumulate, fdg_stream_t stream) {
  FDG_REQUIRE(x && gy && gx && x->p && gy->p && gx->p && N > 0 && OH > 0 && OW > 0 && C > 0,
              "fdg_maxpool2_bwd: bad arguments");
  const int64_t total = (int64_t)N * OH * OW * C;
  if (C % 4 == 0 && fdg::vec4_ok(*x) && fdg::vec4_ok(*gy) && fdg::vec4_ok(*gx) && total / 4 < (1ll << 31)) {
    launch_k(maxpool2_bwd_vec4_kernel, dim3(grid_for(total / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *gy, *gx, (int)(total / 4), OH, OW, C / 4,
             accumulate);
    return check_launch("fdg_maxpool2_bwd");
  }
  launch_k(maxpool2_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *gy, *gx, total, OH, OW, C, accumulate);
  return check_launch("fdg_maxpool2_bwd");
}

int fdg_copy4d(const FdgTensor* x, const FdgTensor* y, int N, int H, int W, int C, int gather, float slope, float scale,
               int accumulate, fdg_stream_t stream) {
  FDG_REQUIRE(x && y && x->p && y->p && N > 0 && H > 0 && W > 0 && C > 0, "fdg_copy4d: bad arguments");
  FDG_REQUIRE(gather >= 0 && gather <= 2, "fdg_copy4d: bad gather mode");
  const int64_t total = (int64_t)N * H * W * C;
  if (C % 4 == 0 && fdg::vec4_ok(*x) && fdg::vec4_ok(*y)) {
    launch_k(copy4d_vec4_kernel, dim3(grid_for(total / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *y, total / 4, H, W, C / 4, gather, slope, scale,
                                                                                   accumulate);
    return check_launch("fdg_copy4d");
  }
  launch_k(copy4d_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *y, total, H, W, C, gather, slope, scale, accumulate);
  return check_launch("fdg_copy4d");
}

int fdg_pool2_bn_act(const FdgTensor* x, const FdgTensor* y, int N, int OH, int OW, int C, const float* scale, const float* shift, float slope,
                     fdg_stream_t stream) {
  FDG_REQUIRE(x && y && x->p && y->p && N > 0 && OH > 0 && OW > 0 && C > 0, "fdg_pool2_bn_act: bad arguments");
  FDG_REQUIRE((scale == nullptr) == (shift == nullptr), "fdg_pool2_bn_act: scale and shift come together");
  if (!(C % 4 == 0 && fdg::vec4_ok(*x) && fdg::vec4_ok(*y) && (!scale || (fdg::aligned16(scale) && fdg::aligned16(shift))))) {
    fdg::set_error("fdg_pool2_bn_act: needs unit channel stride, 16-byte aligned views and C %% 4 == 0");
    return FDG_ENOSUPPORT;
  }
  const int64_t total4 = (int64_t)N * OH * OW * (C / 4);
  ProfScope prof(PF_OTHER, 0.0, 4.0 * 5.0 * (double)N * OH * OW * C, (cudaStream_t)stream);
  launch_k(pool2_bn_act_kernel, dim3(grid_for(total4, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, *x, *y, total4, OH, OW, C / 4, scale, shift, slope);
  return check_launch("fdg_pool2_bn_act");
}

int fdg_tap_sum(const FdgTensor* s, const FdgTensor* out, int N, int H, int W, int R, int S, int pad, int act, fdg_stream_t stream) {
  FDG_REQUIRE(s && out && s->p && out->p && N > 0 && H > 0 && W > 0 && R > 0 && S > 0 && pad >= 0, "fdg_tap_sum: bad arguments");
  const int OH = H + 2 * pad - R + 1, OW = W + 2 * pad - S + 1;
  FDG_REQUIRE(OH > 0 && OW > 0, "fdg_tap_sum: empty output");
  const int64_t total = (int64_t)N * OH * OW;
  launch_k(tap_sum_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, *s, *out, total, H, W, R, S, pad, OH, OW, act);
  return check_launch("fdg_tap_sum");
}

int fdg_tap_spread(const FdgTensor* g, const FdgTensor* gs, int N, int H, int W, int R, int S, int pad, fdg_stream_t stream) {
  FDG_REQUIRE(g && gs && g->p && gs->p && N > 0 && H > 0 && W > 0 && R > 0 && S > 0 && pad >= 0, "fdg_tap_spread: bad arguments");
  const int OH = H + 2 * pad - R + 1, OW = W + 2 * pad - S + 1;
  FDG_REQUIRE(OH > 0 && OW > 0, "fdg_tap_spread: empty gradient");
  const int64_t total = (int64_t)N * H * W * R * S;
  launch_k(tap_spread_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, *g, *gs, total, H, W, R, S, pad, OH, OW);
  return check_launch("fdg_tap_spread");
}

int fdg_act_bwd(const float* g, const float* y, float* out, int64_t n, int act, fdg_stream_t stream) {
  FDG_REQUIRE(g && y && out && n > 0 && (act == FDG_ACT_TANH || act == FDG_ACT_SIGMOID), "fdg_act_bwd: bad arguments");
  launch_k(act_bwd_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, g, y, out, n, act);
  return check_launch("fdg_act_bwd");
}

int fdg_conv2d_dgrad_strided(const FdgDgradStrided* p, fdg_stream_t stream) {
  FDG_REQUIRE(p && p->g.p && p->w && p->dx.p, "fdg_conv2d_dgrad_strided: null pointer");
  FDG_REQUIRE(p->N > 0 && p->OH > 0 && p->OW > 0 && p->Cout > 0 && p->Cin > 0 && p->R > 0 && p->S > 0 && p->stride > 0 &&
                  p->H > 0 && p->W > 0, "fdg_conv2d_dgrad_strided: bad extents");
  FDG_REQUIRE(p->OH == (p->H + 2 * p->pad - p->R) / p->stride + 1 && p->OW == (p->W + 2 * p->pad - p->S) / p->stride + 1,
              "fdg_conv2d_dgrad_strided: OH/OW inconsistent");
  {
    const int rc = fdg::dgrad_strided_small(p, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  const int64_t total = (int64_t)p->N * p->H * p->W * p->Cin;
  dgrad_strided_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(*p, total);
  return check_launch("fdg_conv2d_dgrad_strided");
}

int fdg_pack_weight(const float* w, int Cout, int Cin, int R, int S, int mode, float* out, int out_ld,
                    fdg_stream_t stream) {
  FDG_REQUIRE(w && out && Cout > 0 && Cin > 0 && R > 0 && S > 0 && mode >= 0 && mode <= 2, "fdg_pack_weight: bad arguments");
  const int K = mode == 0 ? R * S * Cin : (mode == 1 ? R * S * Cout : Cout);
  const int ncols = mode == 0 ? Cout : Cin;
  FDG_REQUIRE(out_ld >= ncols, "fdg_pack_weight: out_ld too small");
  FDG_REQUIRE(mode != 2 || (R == 1 && S == 1), "fdg_pack_weight: mode 2 is 1x1 only");
  const int64_t total = (int64_t)K * out_ld;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, R, S, mode, out, out_ld, total);
  return check_launch("fdg_pack_weight");
}

int fdg_loss_grad(const float* a, const float* b, float target, int kind, int64_t n, float scale, float* grad,
                  int accumulate, double* loss, fdg_stream_t stream) {
  FDG_REQUIRE(a && loss && n > 0 && kind >= 0 && kind <= 2, "fdg_loss_grad: bad arguments");
  FDG_REQUIRE(kind == FDG_LOSS_BCE || b, "fdg_loss_grad: L1/MSE need a second operand");
  launch_k(loss_grad_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, b, target, kind, n, scale, grad, accumulate, loss);
  return check_launch("fdg_loss_grad");
}

int fdg_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, fdg_stream_t stream) {
  FDG_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "fdg_adam_flat: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  launch_k(adam_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                  bc1, sqrtf(bc2), grad_scale);
  return check_launch("fdg_adam_flat");
}

// Same update with the step counter and the bias corrections kept on the device (state: 3 floats, zero-initialised), so that
// a captured CUDA graph of the training step can be replayed without any host-side value changing between replays.
int fdg_adam_flat_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float* state, float grad_scale, fdg_stream_t stream) {
  FDG_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n > 0, "fdg_adam_flat_dev: bad arguments");
  launch_k(adam_prepare_kernel, dim3(1), dim3(1), (size_t)(0), (cudaStream_t)stream, state, beta1, beta2);
  int rc = check_launch("fdg_adam_flat_dev[prepare]");
  if (rc != FDG_OK) return rc;
  launch_k(adam_dev_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, state,
                                                                      grad_scale);
  return check_launch("fdg_adam_flat_dev");
}

}  // extern "C"
