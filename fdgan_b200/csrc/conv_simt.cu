// fdg_conv2d, fp32 SIMT implicit-GEMM path.
//
// The general-purpose (any shape / stride / layout) convolution of the hot path: exact fp32 FMA
// arithmetic, used for the HBM-bound odd shapes (K = 27, N = 3, N = 1, NCHW image I/O, 4x4 stride 2)
// and as the fallback of the tcgen05 path (conv_umma.cu) for shapes it does not cover.
// Tile: 128 output pixels x BN output channels x 16-deep K steps, 256 threads, 8 x (BN/16) outputs
// per thread, register-prefetch double buffering, XOR-swizzled transposed A tile in shared memory.
#include "aop.cuh"

namespace fdg {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;

struct ConvArgs {
  FdgConv c;
  AOp ao;
  int64_t M;     // N*OH*OW
  int Ktot;      // R*S*Cin
  int ksteps;    // number of BK-deep steps
  int cchunks;   // VEC: ceil(Cin/BK)
  int wvec;      // weights loadable as float4
  int yvec;      // outputs storable as vectors
};

__device__ __forceinline__ float epi_act(float v, int act) {
  switch (act) {
    case FDG_ACT_RELU: return fmaxf(v, 0.f);
    case FDG_ACT_TANH: return tanhf(v);
    case FDG_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(NT, 2) conv_simt_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int TN = BN / 16;
  constexpr int BV = (BK * BN / 4 + NT - 1) / NT;  // float4 B loads per thread
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ float sred[2][BN];

  const FdgConv& p = a.c;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int OHW = p.OH * p.OW;

  // ---- A-load role: this thread stages channels [a_chunk*4, +4) of pixels a_pix0 and a_pix0+64
  const int a_chunk = t & 3, a_pix0 = t >> 2;
  int pn[2], piy[2], pix[2];
  bool pv[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t m = m0 + a_pix0 + 64 * j;
    pv[j] = m < a.M;
    const int64_t mm = pv[j] ? m : 0;
    pn[j] = (int)(mm / OHW);
    const int rem = (int)(mm - (int64_t)pn[j] * OHW);
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
    piy[j] = oy * p.stride - p.pad;
    pix[j] = ox * p.stride - p.pad;
  }

  float ra[2][4];
  float4 rb[BV];

  auto load_a = [&](int step) {
    if (VEC) {
      const int tap = step / a.cchunks;
      const int c = (step - tap * a.cchunks) * BK + a_chunk * 4;
      const int r = tap / p.S, s = tap - r * p.S;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int iy = piy[j] + r, ix = pix[j] + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pv[j] && c < p.Cin && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = fetch4(a.ao, pn[j], iy, ix, c);
        ra[j][0] = v.x; ra[j][1] = v.y; ra[j][2] = v.z; ra[j][3] = v.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = step * BK + a_chunk * 4 + e;
        const bool kv = k < a.Ktot;
        const int tap = kv ? k / p.Cin : 0;
        const int c = kv ? k - tap * p.Cin : 0;
        const int r = tap / p.S, s = tap - r * p.S;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int iy = piy[j] + r, ix = pix[j] + s;
          float v = 0.f;
          if (kv && pv[j] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = fetch1(a.ao, pn[j], iy, ix, c);
          ra[j][e] = v;
        }
      }
    }
  };

  auto load_b = [&](int step) {
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int idx = t + NT * i;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < BK) {
        int kb;
        bool kv;
        if (VEC) {
          const int tap = step / a.cchunks;
          const int c = (step - tap * a.cchunks) * BK + row;
          kv = c < p.Cin;
          kb = tap * p.Cin + c;
        } else {
          kb = step * BK + row;
          kv = kb < a.Ktot;
        }
        const int n = n0 + col;
        if (kv && n < p.Cout) {
          const float* wp = p.w + (int64_t)kb * p.w_ld + n;
          if (a.wvec && n + 3 < p.w_ld) {
            v = ld4(wp);
          } else {
            v.x = __ldg(wp);
            if (n + 1 < p.Cout) v.y = __ldg(wp + 1);
            if (n + 2 < p.Cout) v.z = __ldg(wp + 2);
            if (n + 3 < p.Cout) v.w = __ldg(wp + 3);
          }
          // columns >= Cout inside a vector load come from the zero padding of the packed buffer, or are
          // never stored; keep them finite
          if (n + 1 >= p.Cout) v.y = 0.f;
          if (n + 2 >= p.Cout) v.z = 0.f;
          if (n + 3 >= p.Cout) v.w = 0.f;
        }
      }
      rb[i] = v;
    }
  };

  auto store_ab = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = (a_pix0 + 64 * j) ^ (a_chunk << 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) As[buf][a_chunk * 4 + e][m] = ra[j][e];
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int idx = t + NT * i;
      const int row = idx / (BN / 4), col = (idx - row * (BN / 4)) * 4;
      if (row < BK) *reinterpret_cast<float4*>(&Bs[buf][row][col]) = rb[i];
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (t < BN) { sred[0][t] = 0.f; sred[1][t] = 0.f; }

  load_a(0);
  load_b(0);
  store_ab(0);
  __syncthreads();

  for (int step = 0; step < a.ksteps; ++step) {
    const int cur = step & 1;
    const bool more = step + 1 < a.ksteps;
    if (more) { load_a(step + 1); load_b(step + 1); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const int mb = (ty * 8) ^ (((k >> 2) & 3) << 3);
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][mb]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][mb + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      if (TN == 2) {
        const float2 b = *reinterpret_cast<const float2*>(&Bs[cur][k][tx * 2]);
        bv[0] = b.x; bv[1] = b.y;
      } else {
#pragma unroll
        for (int q = 0; q < TN / 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][q * 64 + tx * 4]);
          bv[q * 4 + 0] = b.x; bv[q * 4 + 1] = b.y; bv[q * 4 + 2] = b.z; bv[q * 4 + 3] = b.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_ab(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  constexpr int NQ = (TN == 2) ? 1 : TN / 4;   // column groups
  constexpr int QW = (TN == 2) ? 2 : 4;        // columns per group
  float ssum[TN], ssq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= a.M) continue;
    const int n = (int)(m / OHW);
    const int rem = (int)(m - (int64_t)n * OHW);
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int cb = n0 + (TN == 2 ? tx * 2 : q * 64 + tx * 4);
      if (cb >= p.Cout) continue;
      float v[QW];
#pragma unroll
      for (int u = 0; u < QW; ++u) {
        const int c = cb + u;
        float r = acc[i][q * 4 + u] * p.alpha;
        if (c < p.Cout) {
          if (p.bias) r += __ldg(p.bias + c);
          r = epi_act(r, p.act);
          if (p.e.p) {
            const float ev = __ldg(p.e.p + n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw + (int64_t)c * p.e.sc);
            r *= (ev > 0.f ? 1.f : p.eslope);
          }
          ssum[q * 4 + u] += r;
          ssq[q * 4 + u] += r * r;
        }
        v[u] = r;
      }
      const bool full = cb + QW <= p.Cout;
      if (p.store == FDG_STORE_UP2) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          float* yp = p.y.p + n * p.y.sn + (int64_t)(2 * oy + (d >> 1)) * p.y.sh + (int64_t)(2 * ox + (d & 1)) * p.y.sw + (int64_t)cb * p.y.sc;
          if (a.yvec && full) {
            if (QW == 4) *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
            else *reinterpret_cast<float2*>(yp) = make_float2(v[0], v[1]);
          } else {
#pragma unroll
            for (int u = 0; u < QW; ++u) if (cb + u < p.Cout) yp[(int64_t)u * p.y.sc] = v[u];
          }
        }
      } else {
        float* yp = p.y.p + n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox * p.y.sw + (int64_t)cb * p.y.sc;
        const bool accum = p.store == FDG_STORE_ACCUM;
        if (a.yvec && full) {
          if (QW == 4) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if (accum) { const float4 old = *reinterpret_cast<const float4*>(yp); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *reinterpret_cast<float4*>(yp) = o;
          } else {
            float2 o = make_float2(v[0], v[1]);
            if (accum) { const float2 old = *reinterpret_cast<const float2*>(yp); o.x += old.x; o.y += old.y; }
            *reinterpret_cast<float2*>(yp) = o;
          }
        } else {
#pragma unroll
          for (int u = 0; u < QW; ++u)
            if (cb + u < p.Cout) {
              float* q1 = yp + (int64_t)u * p.y.sc;
              *q1 = accum ? *q1 + v[u] : v[u];
            }
        }
      }
    }
  }

  if (p.stats) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int u = 0; u < QW; ++u) {
        const int cl = (TN == 2 ? tx * 2 : q * 64 + tx * 4) + u;
        atomicAdd(&sred[0][cl], ssum[q * 4 + u]);
        atomicAdd(&sred[1][cl], ssq[q * 4 + u]);
      }
    __syncthreads();
    if (t < BN && n0 + t < p.Cout) {
      atomicAdd(p.stats + n0 + t, (double)sred[0][t]);
      atomicAdd(p.stats + p.stats_ld + n0 + t, (double)sred[1][t]);
    }
  }
}

template <int BN>
static int launch_simt(const ConvArgs& a, bool vec, cudaStream_t st) {
  dim3 grid((unsigned)cdiv64(a.M, BM), (unsigned)cdiv(a.c.Cout, BN));
  const double gmul = a.c.gather == FDG_GATHER_AVGPOOL2 ? 4.0 : 1.0;
  ProfScope prof(PF_CONV_SIMT, 2.0 * (double)a.M * a.Ktot * a.c.Cout,
                 4.0 * ((double)a.M * a.c.Cout + gmul * (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  if (vec) conv_simt_kernel<BN, true><<<grid, NT, 0, st>>>(a);
  else conv_simt_kernel<BN, false><<<grid, NT, 0, st>>>(a);
  return check_launch("fdg_conv2d[simt]");
}

int conv2d_validate(const FdgConv* p) {
  FDG_REQUIRE(p != nullptr, "fdg_conv2d: null descriptor");
  FDG_REQUIRE(p->x.p && p->y.p && p->w, "fdg_conv2d: null x/y/w pointer");
  FDG_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0, "fdg_conv2d: non-positive extent");
  FDG_REQUIRE(p->R > 0 && p->S > 0 && p->stride > 0 && p->pad >= 0, "fdg_conv2d: bad filter geometry");
  FDG_REQUIRE(p->OH == (p->H + 2 * p->pad - p->R) / p->stride + 1 && p->OW == (p->W + 2 * p->pad - p->S) / p->stride + 1,
              "fdg_conv2d: OH/OW (%d,%d) inconsistent with H,W,R,S,stride,pad", p->OH, p->OW);
  FDG_REQUIRE(p->OH > 0 && p->OW > 0, "fdg_conv2d: empty output");
  FDG_REQUIRE(p->gather >= 0 && p->gather <= 2, "fdg_conv2d: bad gather mode %d", p->gather);
  FDG_REQUIRE(!p->has_affine || (p->scale && p->shift), "fdg_conv2d: affine prologue without scale/shift");
  FDG_REQUIRE(p->w_ld >= p->Cout, "fdg_conv2d: w_ld < Cout");
  FDG_REQUIRE(p->store >= 0 && p->store <= 2, "fdg_conv2d: bad store mode %d", p->store);
  FDG_REQUIRE(!(p->stats && p->store != FDG_STORE_NORMAL && !p->e_scale), "fdg_conv2d: stats need the normal store mode");
  FDG_REQUIRE(!p->stats || p->stats_ld >= p->Cout, "fdg_conv2d: stats_ld < Cout");
  FDG_REQUIRE((int64_t)p->R * p->S * p->Cin < (1ll << 31), "fdg_conv2d: K too large");
  if (p->e_scale) {
    FDG_REQUIRE(p->e_shift && p->e.p && p->store != FDG_STORE_UP2 && p->Cout % 4 == 0 && vec4_ok(p->y) && vec4_ok(p->e) &&
                    aligned16(p->e_scale) && aligned16(p->e_shift) && !p->bias && p->act == FDG_ACT_NONE,
                "fdg_conv2d: the BatchNorm-backward epilogue needs e, e_shift, 128-bit y/e views, Cout %% 4 == 0, no bias/activation");
    FDG_REQUIRE(p->impl != 1, "fdg_conv2d: the BatchNorm-backward epilogue exists on the tcgen05 path only");
  }
  return FDG_OK;
}

int conv2d_simt(const FdgConv* p, cudaStream_t st) {
  ConvArgs a;
  a.c = *p;
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.Ktot = p->R * p->S * p->Cin;
  a.ao = AOp{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  const bool vec = aop_vec_ok(a.ao, p->Cin);
  a.cchunks = cdiv(p->Cin, BK);
  a.ksteps = vec ? p->R * p->S * a.cchunks : cdiv(a.Ktot, BK);
  a.wvec = aligned16(p->w) && (p->w_ld % 4 == 0);
  a.yvec = vec4_ok(p->y);
  if (p->Cout <= 32) return launch_simt<32>(a, vec, st);
  if (p->Cout <= 64) return launch_simt<64>(a, vec, st);
  return launch_simt<128>(a, vec, st);
}

}  // namespace fdg
