// fdg_conv2d_wgrad, fp32 SIMT path: dW[k][co] += sum_pixels a[pixel][k] * g[pixel][co].
//
// GEMM with the pixel index as the contraction dimension (up to 1 M pixels at B=16, 256x256), so the
// work is split over pixels (grid.z) as well as over (k, co) tiles; every CTA reduces its pixel range
// in registers (8 x 4 outputs per thread, tile 128 k x 64 co, 16 pixels per shared-memory stage) and
// adds its partial tile atomically, straight into the PyTorch parameter layout (OIHW or [Cin][Cout]).
#include "aop.cuh"

namespace fdg {

constexpr int WTK = 128;  // k rows per tile
constexpr int WTC = 64;   // output channels per tile
constexpr int WPM = 16;   // pixels per stage
constexpr int WNT = 256;

struct WgradArgs {
  FdgWgrad c;
  AOp ao;
  int64_t M;
  int Ktot;
  int64_t m_per_split;
  int avec, gvec;
};

template <bool AVEC, bool GVEC>
__global__ void __launch_bounds__(WNT, 2) wgrad_simt_kernel(const __grid_constant__ WgradArgs a) {
  __shared__ __align__(16) float As[2][WPM][WTK];
  __shared__ __align__(16) float Gs[2][WPM][WTC];
  const FdgWgrad& p = a.c;
  const int t = threadIdx.x;
  const int k0 = blockIdx.x * WTK, c0 = blockIdx.y * WTC;
  const int64_t mbeg = (int64_t)blockIdx.z * a.m_per_split;
  const int64_t mend = min(a.M, mbeg + a.m_per_split);
  const int OHW = p.OH * p.OW;

  // A-load role: fixed k group (4 consecutive k = 4 channels of one tap), pixels a_p and a_p+8 of each stage
  const int a_kg = t & 31, a_p = t >> 5;
  const int ak = k0 + a_kg * 4;
  int a_r[4], a_s[4], a_c[4];
  bool a_kv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = ak + e;
    a_kv[e] = k < a.Ktot;
    const int tap = a_kv[e] ? k / p.Cin : 0;
    a_c[e] = a_kv[e] ? k - tap * p.Cin : 0;
    a_r[e] = tap / p.S;
    a_s[e] = tap - a_r[e] * p.S;
  }
  // G-load role: pixel g_p, channel group g_cg
  const int g_p = t >> 4, g_cg = t & 15;

  float ra[2][4];
  float rg[4];

  auto load = [&](int64_t mb) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t m = mb + a_p + 8 * j;
      ra[j][0] = ra[j][1] = ra[j][2] = ra[j][3] = 0.f;
      if (m < mend) {
        const int n = (int)(m / OHW);
        const int rem = (int)(m - (int64_t)n * OHW);
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        if (AVEC) {
          const int iy = oy * p.stride - p.pad + a_r[0], ix = ox * p.stride - p.pad + a_s[0];
          if (a_kv[0] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            const float4 v = fetch4(a.ao, n, iy, ix, a_c[0]);
            ra[j][0] = v.x; ra[j][1] = v.y; ra[j][2] = v.z; ra[j][3] = v.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int iy = oy * p.stride - p.pad + a_r[e], ix = ox * p.stride - p.pad + a_s[e];
            if (a_kv[e] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) ra[j][e] = fetch1(a.ao, n, iy, ix, a_c[e]);
          }
        }
      }
    }
    {
      const int64_t m = mb + g_p;
      rg[0] = rg[1] = rg[2] = rg[3] = 0.f;
      const int c = c0 + g_cg * 4;
      if (m < mend && c < p.Cout) {
        const int n = (int)(m / OHW);
        const int rem = (int)(m - (int64_t)n * OHW);
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        const float* gp = p.g.p + n * p.g.sn + (int64_t)oy * p.g.sh + (int64_t)ox * p.g.sw + (int64_t)c * p.g.sc;
        if (GVEC && c + 3 < p.Cout) {
          const float4 v = ld4(gp);
          rg[0] = v.x; rg[1] = v.y; rg[2] = v.z; rg[3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (c + e < p.Cout) rg[e] = __ldg(gp + (int64_t)e * p.g.sc);
        }
      }
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      *reinterpret_cast<float4*>(&As[buf][a_p + 8 * j][a_kg * 4]) = make_float4(ra[j][0], ra[j][1], ra[j][2], ra[j][3]);
    *reinterpret_cast<float4*>(&Gs[buf][g_p][g_cg * 4]) = make_float4(rg[0], rg[1], rg[2], rg[3]);
  };

  const int tk = t >> 4, tc = t & 15;  // outputs: k rows {tk*4.., 64+tk*4..}, channels tc*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (mbeg < mend) {
    load(mbeg);
    store(0);
    __syncthreads();
    int cur = 0;
    for (int64_t mb = mbeg; mb < mend; mb += WPM) {
      const bool more = mb + WPM < mend;
      if (more) load(mb + WPM);
#pragma unroll
      for (int pm = 0; pm < WPM; ++pm) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][pm][tk * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][pm][64 + tk * 4]);
        const float4 g = *reinterpret_cast<const float4*>(&Gs[cur][pm][tc * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
      }
      if (more) store(cur ^ 1);
      __syncthreads();
      cur ^= 1;
    }
  }

  // ---- scatter-add into the parameter layout
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? tk * 4 + i : 64 + tk * 4 + (i - 4));
    if (k >= a.Ktot) continue;
    const int tap = k / p.Cin, ci = k - tap * p.Cin;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = c0 + tc * 4 + j;
      if (co >= p.Cout) continue;
      const int64_t idx = p.transposed ? (int64_t)ci * p.Cout + co
                                       : ((int64_t)co * p.Cin + ci) * (p.R * p.S) + tap;
      atomicAdd(p.dw + idx, acc[i][j]);
    }
  }
}

// per-channel sum over all pixels (bias gradient and generic column sums)
__global__ void colsum_kernel(FdgTensor x, int64_t M, int HW, int W, int C, float* out, int accumulate) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  // grid.x tiles channels by 32, grid.y splits pixels; block (32, 8)
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C) {
    for (int64_t m = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; m < M; m += (int64_t)gridDim.y * blockDim.y) {
      const int n = (int)(m / HW);
      const int rem = (int)(m - (int64_t)n * HW);
      const int h = rem / W, w = rem - h * W;
      s += __ldg(x.p + n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + (int64_t)c * x.sc);
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

// planar variant (NCHW planes: unit pixel stride, rows contiguous): one CTA sums a segment of one (image, channel) plane with
// coalesced loads; the bias gradient of the 3-channel head conv_refin3 took 0.2 ms in the generic kernel (3 of 32 lanes active)
__global__ void __launch_bounds__(256) colsum_planar_kernel(FdgTensor x, int HW, int C, int seg, float* out) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  __shared__ float red[8];
  const int plane = blockIdx.x, n = plane / C, c = plane - n * C;
  const float* src = x.p + n * x.sn + (int64_t)c * x.sc;
  const int beg = blockIdx.y * seg, end = beg + seg < HW ? beg + seg : HW;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = beg + threadIdx.x;
  for (; i + 768 < end; i += 1024) { s0 += __ldg(src + i); s1 += __ldg(src + i + 256); s2 += __ldg(src + i + 512); s3 += __ldg(src + i + 768); }
  for (; i < end; i += 256) s0 += __ldg(src + i);
  float s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w];
    atomicAdd(out + c, v);
  }
}

// 128-bit variant (unit channel stride, C % 4 == 0): thread = (4-channel group, pixel lane), four pixels in flight per
// iteration, partial sums reduced across the pixel lanes of the CTA in shared memory, one float atomic per channel
__global__ void __launch_bounds__(256) colsum_vec4_kernel(FdgTensor x, int64_t M, int HW, int W, int C, int cgroups, int pix_lanes,
                                                          float* out) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  __shared__ float4 red[256];
  const int cg = threadIdx.x % cgroups, pl = threadIdx.x / cgroups;
  const int c = (blockIdx.y * cgroups + cg) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C && pl < pix_lanes) {
    const bool small = M <= 0x7fffffffLL;
    const int64_t step = (int64_t)gridDim.x * pix_lanes;
    for (int64_t m0 = (int64_t)blockIdx.x * pix_lanes + pl; m0 < M; m0 += 4 * step) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t m = m0 + k * step;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M) {
          const int n = small ? (int)m / HW : (int)(m / HW);
          const int rem = small ? (int)m - n * HW : (int)(m - (int64_t)n * HW);
          const int h = rem / W, w = rem - h * W;
          v[k] = __ldg(reinterpret_cast<const float4*>(x.p + n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + c));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (pl == 0 && c < C) {
    for (int l = 1; l < pix_lanes; ++l) {
      const float4 q = red[l * cgroups + cg];
      s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
    }
    atomicAdd(out + c, s.x); atomicAdd(out + c + 1, s.y); atomicAdd(out + c + 2, s.z); atomicAdd(out + c + 3, s.w);
  }
}

}  // namespace fdg

namespace fdg {
int wgrad_umma_supported(const FdgWgrad* p);
int wgrad_umma(const FdgWgrad* p, cudaStream_t st);
int wgrad_halo_supported(const FdgWgrad* p);
int wgrad_thin(const FdgWgrad* p, cudaStream_t st);
int wgrad_halo(const FdgWgrad* p, cudaStream_t st);
int wgrad_k1_supported(const FdgWgrad* p);
int wgrad_k1(const FdgWgrad* p, cudaStream_t st);
}  // namespace fdg

using namespace fdg;

extern "C" int fdg_colsum(const FdgTensor* x, int N, int H, int W, int C, float* out, int accumulate,
                          fdg_stream_t stream) {
  FDG_REQUIRE(x && x->p && out && N > 0 && H > 0 && W > 0 && C > 0, "fdg_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    if (cudaMemsetAsync(out, 0, sizeof(float) * C, st) != cudaSuccess) { set_error("fdg_colsum: memset failed"); return FDG_ECUDA; }
  }
  const int64_t M = (int64_t)N * H * W;
  if (C % 4 == 0 && vec4_ok(*x)) {
    const int groups_total = C / 4;
    const int cgroups = groups_total < 64 ? groups_total : 64;
    const int pix_lanes = 256 / cgroups;
    const int gyc = cdiv(groups_total, cgroups);
    int64_t gx = cdiv64(M, (int64_t)pix_lanes * 16);
    const int64_t cap = (int64_t)device_sm_count() * 8 / gyc + 1;
    if (gx > cap) gx = cap;
    launch_k(colsum_vec4_kernel, dim3(dim3((unsigned)gx, gyc)), dim3(256), (size_t)(0), st, *x, M, H * W, W, C, cgroups, pix_lanes, out);
    return check_launch("fdg_colsum");
  }
  if (x->sw == 1 && x->sh == W && (int64_t)N * C <= 65535 && H * W >= 4096) {
    int splits = (int)cdiv64((int64_t)device_sm_count() * 4, (int64_t)N * C);
    const int max_splits = cdiv(H * W, 4096);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int seg = cdiv(cdiv(H * W, splits), 256) * 256;
    launch_k(colsum_planar_kernel, dim3((unsigned)(N * C), (unsigned)cdiv(H * W, seg)), dim3(256), (size_t)(0), st, *x, H * W, C, seg, out);
    return check_launch("fdg_colsum");
  }
  int64_t gy = cdiv64(M, 8 * 64);
  if (gy > 1024) gy = 1024;
  dim3 grid(cdiv(C, 32), (unsigned)gy);
  launch_k(colsum_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), st, *x, M, H * W, W, C, out, accumulate);
  return check_launch("fdg_colsum");
}

extern "C" int fdg_conv2d_wgrad(const FdgWgrad* p, fdg_stream_t stream) {
  FDG_REQUIRE(p && p->x.p && p->g.p && p->dw, "fdg_conv2d_wgrad: null pointer");
  FDG_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0 && p->R > 0 && p->S > 0 && p->stride > 0,
              "fdg_conv2d_wgrad: bad extents");
  FDG_REQUIRE(p->OH == (p->H + 2 * p->pad - p->R) / p->stride + 1 && p->OW == (p->W + 2 * p->pad - p->S) / p->stride + 1,
              "fdg_conv2d_wgrad: OH/OW inconsistent");
  FDG_REQUIRE(!p->transposed || (p->R == 1 && p->S == 1), "fdg_conv2d_wgrad: transposed layout is 1x1 only");
  FDG_REQUIRE(!p->has_affine || (p->scale && p->shift), "fdg_conv2d_wgrad: affine prologue without scale/shift");
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int rc = wgrad_thin(p, st);        // thin layers (stem, Fusion-D layer 1): direct kernel
    if (rc < 0) return rc;
    if (rc == 0) return p->dbias ? fdg_colsum(&p->g, p->N, p->OH, p->OW, p->Cout, p->dbias, 1, stream) : FDG_OK;
  }
  if (p->impl != 1) {
    const int k1 = wgrad_k1_supported(p);      // growth convolutions: filter columns concatenated along N (wgrad_k1.cu)
    const int ok = k1 || wgrad_umma_supported(p) || wgrad_halo_supported(p);
    if (p->impl == 2 && !ok) { set_error("fdg_conv2d_wgrad: impl=tcgen05 requested but shape/layout unsupported"); return FDG_ENOSUPPORT; }
    if (ok) {
      const int rc = k1 ? wgrad_k1(p, st) : (wgrad_halo_supported(p) ? wgrad_halo(p, st) : wgrad_umma(p, st));
      if (rc != FDG_OK) return rc;
      if (p->dbias) return fdg_colsum(&p->g, p->N, p->OH, p->OW, p->Cout, p->dbias, 1, stream);
      return FDG_OK;
    }
  }
  WgradArgs a;
  a.c = *p;
  a.ao = AOp{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.Ktot = p->R * p->S * p->Cin;
  a.avec = aop_vec_ok(a.ao, p->Cin);
  a.gvec = vec4_ok(p->g);
  const int tiles = cdiv(a.Ktot, WTK) * cdiv(p->Cout, WTC);
  int64_t splits = cdiv64(148 * 4, tiles);
  const int64_t max_splits = cdiv64(a.M, WPM * 16);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  a.m_per_split = cdiv64(cdiv64(a.M, splits), WPM) * WPM;
  splits = cdiv64(a.M, a.m_per_split);
  dim3 grid(cdiv(a.Ktot, WTK), cdiv(p->Cout, WTC), (unsigned)splits);
  int rc;
  {
  ProfScope prof(PF_WGRAD, 2.0 * (double)a.M * a.Ktot * p->Cout,
                 4.0 * ((double)a.M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  if (a.avec && a.gvec) wgrad_simt_kernel<true, true><<<grid, WNT, 0, st>>>(a);
  else if (a.avec) wgrad_simt_kernel<true, false><<<grid, WNT, 0, st>>>(a);
  else if (a.gvec) wgrad_simt_kernel<false, true><<<grid, WNT, 0, st>>>(a);
  else wgrad_simt_kernel<false, false><<<grid, WNT, 0, st>>>(a);
  rc = check_launch("fdg_conv2d_wgrad");
  }
  if (rc != FDG_OK) return rc;
  if (p->dbias) {
    return fdg_colsum(&p->g, p->N, p->OH, p->OW, p->Cout, p->dbias, 1, stream);
  }
  return FDG_OK;
}
