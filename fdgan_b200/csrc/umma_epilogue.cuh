// Epilogue of the tcgen05 convolution kernels: one 32-channel group of one output pixel, held in registers after
// tcgen05.ld, goes through alpha / bias / activation / backward mask, is stored (normal, 2x2 nearest-upsample or
// accumulate) and contributes to the per-channel BatchNorm statistics (padded shared-memory transpose: lane l ends up
// with the column sums of channel c0 + l over the warp's 32 pixels).
#pragma once
#include "aop.cuh"
#include "umma.cuh"

namespace fdg {

__device__ __forceinline__ void umma_epilogue_group(const FdgConv& p, bool yvec, bool evec, float (&v)[32], bool mv, int n, int oy,
                                                    int ox, int c0, int lane, float (*tile)[33], float& acc1, float& acc2) {
    const int nvalid = p.Cout - c0 < 32 ? p.Cout - c0 : 32;
    const bool full = nvalid == 32;
    // ---- alpha, bias
    if (p.bias) {
      if (full) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = ld4(p.bias + c0 + 4 * q);
          v[4 * q] = fmaf(v[4 * q], p.alpha, bb.x); v[4 * q + 1] = fmaf(v[4 * q + 1], p.alpha, bb.y);
          v[4 * q + 2] = fmaf(v[4 * q + 2], p.alpha, bb.z); v[4 * q + 3] = fmaf(v[4 * q + 3], p.alpha, bb.w);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] = fmaf(v[u], p.alpha, u < nvalid ? __ldg(p.bias + c0 + u) : 0.f);
      }
    } else if (p.alpha != 1.f) {
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] *= p.alpha;
    }
    // ---- activation (uniform switch hoisted out of the element loop)
    if (p.act == FDG_ACT_RELU) {
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] = fmaxf(v[u], 0.f);
    } else if (p.act == FDG_ACT_TANH) {
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] = tanhf(v[u]);
    } else if (p.act == FDG_ACT_SIGMOID) {
#pragma unroll
      for (int u = 0; u < 32; ++u) v[u] = 1.f / (1.f + expf(-v[u]));
    }
    // ---- ReLU / LeakyReLU backward mask from a second tensor
    if (p.e.p && mv) {
      const float* ep = p.e.p + n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw + (int64_t)c0 * p.e.sc;
      if (evec && full) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ev = ld4(ep + 4 * q);
          v[4 * q] *= ev.x > 0.f ? 1.f : p.eslope; v[4 * q + 1] *= ev.y > 0.f ? 1.f : p.eslope;
          v[4 * q + 2] *= ev.z > 0.f ? 1.f : p.eslope; v[4 * q + 3] *= ev.w > 0.f ? 1.f : p.eslope;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (u < nvalid) v[u] *= __ldg(ep + (int64_t)u * p.e.sc) > 0.f ? 1.f : p.eslope;
      }
    }
    if (!mv || !full) {
#pragma unroll
      for (int u = 0; u < 32; ++u) if (!mv || u >= nvalid) v[u] = 0.f;
    }
    if (mv) {
      const int reps = p.store == FDG_STORE_UP2 ? 4 : 1;
      for (int d = 0; d < reps; ++d) {
        const int yy = p.store == FDG_STORE_UP2 ? 2 * oy + (d >> 1) : oy, xx = p.store == FDG_STORE_UP2 ? 2 * ox + (d & 1) : ox;
        float* yp = p.y.p + n * p.y.sn + (int64_t)yy * p.y.sh + (int64_t)xx * p.y.sw + (int64_t)c0 * p.y.sc;
        if (yvec && full) {
          if (p.store == FDG_STORE_ACCUM) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 old = *reinterpret_cast<const float4*>(yp + 4 * q);
              *reinterpret_cast<float4*>(yp + 4 * q) = make_float4(v[4 * q] + old.x, v[4 * q + 1] + old.y, v[4 * q + 2] + old.z, v[4 * q + 3] + old.w);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(yp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (u < nvalid) {
              float* q1 = yp + (int64_t)u * p.y.sc;
              *q1 = p.store == FDG_STORE_ACCUM ? *q1 + v[u] : v[u];
            }
        }
      }
    }
    if (p.stats) {
      // column sums through a padded shared-memory transpose: lane l ends up with the sums of column l
#pragma unroll
      for (int u = 0; u < 32; ++u) tile[u][lane] = v[u];
      __syncwarp();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const float xv = tile[lane][rr];
        s1 += xv;
        s2 = fmaf(xv, xv, s2);
      }
      __syncwarp();
      acc1 += s1;
      acc2 += s2;
    }
}

}  // namespace fdg
