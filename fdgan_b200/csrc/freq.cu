// Frequency decomposition feeding the Fusion-discriminator: z = [x, Blur(x), Laplacian(x)].
//
// Reference semantics (loss.py survives only as bytecode):
//   Blur.forward       loss.pyc@L142-151: (x-mean)/std, ReflectionPad2d(7), one 15x15 Gaussian (sigma 3,
//                      isotropic_gaussian_kernel @L153-159) on every (batch, channel) plane
//   Laplacian.forward  loss.pyc@L286-301: depth-wise 3x3 [1 1 1; 1 -8 1; 1 1 1], zero padding 1
// The Gaussian is separable (outer(g,g)); the forward kernel stages a 46x46 halo tile of each plane in
// shared memory, runs the 15-tap row pass into shared memory and the 15-tap column pass out of it, and
// writes all nine channels of z in one pass (algorithmic traffic: 12 B read + 36 B written per pixel).
#include "common.cuh"

namespace fdg {

constexpr int FT = 32;        // output tile edge
constexpr int FR = 7;         // blur radius
constexpr int FH = FT + 2 * FR;

struct Gauss15 { float g[15]; };
__constant__ float c_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_std[3] = {0.229f, 0.224f, 0.225f};

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void __launch_bounds__(256) freq_fwd_kernel(FdgTensor x, FdgTensor z, int H, int W, const __grid_constant__ Gauss15 gk) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  __shared__ float raw[FH][FH + 1];
  __shared__ float nrm[FH][FH + 1];
  __shared__ float tmp[FH][FT + 1];
  const int n = blockIdx.z / 3, c = blockIdx.z % 3;
  const int y0 = blockIdx.y * FT, x0 = blockIdx.x * FT;
  const float mean = c_mean[c], sd = c_std[c];
  const float* xp = x.p + n * x.sn + (int64_t)c * x.sc;
  for (int i = threadIdx.x; i < FH * FH; i += blockDim.x) {
    const int ty = i / FH, tx = i - ty * FH;
    const int gy = reflect(y0 + ty - FR, H), gx = reflect(x0 + tx - FR, W);
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(xp + (int64_t)gy * x.sh + (int64_t)gx * x.sw);
    raw[ty][tx] = v;
    nrm[ty][tx] = (v - mean) / sd;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FH * FT; i += blockDim.x) {
    const int ty = i / FT, tx = i - ty * FT;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 15; ++t) s = fmaf(gk.g[t], nrm[ty][tx + t], s);
    tmp[ty][tx] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FT * FT; i += blockDim.x) {
    const int ty = i / FT, tx = i - ty * FT;
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy >= H || gx >= W) continue;
    float lf = 0.f;
#pragma unroll
    for (int t = 0; t < 15; ++t) lf = fmaf(gk.g[t], tmp[ty + t][tx], lf);
    const float ctr = raw[ty + FR][tx + FR];
    float hf = -8.f * ctr;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        if (dy == 0 && dx == 0) continue;
        const int yy = gy + dy, xx = gx + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) hf += raw[ty + FR + dy][tx + FR + dx];
      }
    float* zp = z.p + n * z.sn + (int64_t)gy * z.sh + (int64_t)gx * z.sw;
    zp[(int64_t)c * z.sc] = ctr;
    zp[(int64_t)(3 + c) * z.sc] = lf;
    zp[(int64_t)(6 + c) * z.sc] = hf;
  }
}

// adjoint of the 1-D "reflect-pad 7 then 15-tap correlate" along one axis: out[i] = sum over the padded
// positions j that fold onto i of sum_t g[t] * in[j - t + 7]
__device__ __forceinline__ float blur_adj_1d(const Gauss15& gk, const float* in, int64_t stride, int i, int n) {
  float s = 0.f;
  int js[3];
  int nj = 0;
  js[nj++] = i;
  if (i >= 1 && i <= FR) js[nj++] = -i;
  if (n - 1 - i >= 1 && n - 1 - i <= FR) js[nj++] = 2 * (n - 1) - i;
  for (int q = 0; q < nj; ++q) {
#pragma unroll
    for (int t = 0; t < 15; ++t) {
      const int o = js[q] - t + FR;
      if (o >= 0 && o < n) s = fmaf(gk.g[t], __ldg(in + (int64_t)o * stride), s);
    }
  }
  return s;
}

// pass 1: scratch[n][c][y][x] = adjoint along W of dz[:, 3+c]
__global__ void freq_bwd_rows_kernel(FdgTensor dz, float* scratch, int N, int H, int W, const __grid_constant__ Gauss15 gk) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const int64_t total = (int64_t)N * 3 * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % 3);
    const int n = (int)(r / 3);
    const float* row = dz.p + n * dz.sn + (int64_t)y * dz.sh + (int64_t)(3 + c) * dz.sc;
    scratch[i] = blur_adj_1d(gk, row, dz.sw, xw, W);
  }
}

// pass 2: dx = dz[:, c] + (1/std) * adjoint along H of scratch + Laplacian(dz[:, 6+c])  (symmetric kernel, zero pad)
__global__ void freq_bwd_cols_kernel(FdgTensor dz, const float* scratch, FdgTensor dx, int N, int H, int W,
                                     const __grid_constant__ Gauss15 gk) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const int64_t total = (int64_t)N * 3 * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % 3);
    const int n = (int)(r / 3);
    const float* col = scratch + ((int64_t)(n * 3 + c) * H) * W + xw;
    const float lf = blur_adj_1d(gk, col, W, y, H) / c_std[c];
    const float* d0 = dz.p + n * dz.sn + (int64_t)c * dz.sc;
    const float* d2 = dz.p + n * dz.sn + (int64_t)(6 + c) * dz.sc;
    float hf = -8.f * __ldg(d2 + (int64_t)y * dz.sh + (int64_t)xw * dz.sw);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dxx = -1; dxx <= 1; ++dxx) {
        if (dy == 0 && dxx == 0) continue;
        const int yy = y + dy, xx = xw + dxx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) hf += __ldg(d2 + (int64_t)yy * dz.sh + (int64_t)xx * dz.sw);
      }
    dx.p[n * dx.sn + (int64_t)y * dx.sh + (int64_t)xw * dx.sw + (int64_t)c * dx.sc] =
        __ldg(d0 + (int64_t)y * dz.sh + (int64_t)xw * dz.sw) + lf + hf;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Generic depth-wise l x l filter (one kernel shared by every (image, channel) plane): the channel-agnostic use of the
// reference's Laplacian (loss.pyc@L286-301: kernel.repeat(c,1,1,1), groups=c, zero padding) and Blur with a non-default
// (l, kernel, use_input_norm) (loss.pyc@L123-151: optional ImageNet normalisation, ReflectionPad2d(l//2), one kernel on
// every plane).  Pure HBM traffic (4 B read + 4 B written per element); a (TY+l-1) x (TX+l-1) halo tile and the filter
// taps live in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int DW_TX = 32, DW_TY = 16, DW_MAXL = 31;

__device__ __forceinline__ int reflect_multi(int i, int n) {      // ReflectionPad2d for pad < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) depthwise_fwd_kernel(FdgDepthwise d) {
  extern __shared__ float dsm[];
  const int l = d.l, r = l / 2;
  const int TW = DW_TX + l - 1, TH = DW_TY + l - 1;
  float* kw = dsm;                 // l*l taps
  float* tile = dsm + l * l;       // TH x (TW+1)
  const int plane = blockIdx.z, n = plane / d.c, c = plane - n * d.c;
  const int x0 = blockIdx.x * DW_TX, y0 = blockIdx.y * DW_TY;
  for (int i = threadIdx.x; i < l * l; i += blockDim.x) kw[i] = __ldg(d.kernel + i);
  const float mean = d.mean ? __ldg(d.mean + c) : 0.f;
  const float istd = d.inv_std ? __ldg(d.inv_std + c) : 1.f;
  const float* xp = d.x.p + (int64_t)n * d.x.sn + (int64_t)c * d.x.sc;
  for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
    const int ty = i / TW, tx = i - ty * TW;
    int gy = y0 + ty - r, gx = x0 + tx - r;
    float v = 0.f;
    if (d.pad_mode == 1) {
      gy = reflect_multi(gy, d.h); gx = reflect_multi(gx, d.w);
      if (gy >= 0 && gy < d.h && gx >= 0 && gx < d.w) v = (__ldg(xp + (int64_t)gy * d.x.sh + (int64_t)gx * d.x.sw) - mean) * istd;
    } else if (gy >= 0 && gy < d.h && gx >= 0 && gx < d.w) {
      v = (__ldg(xp + (int64_t)gy * d.x.sh + (int64_t)gx * d.x.sw) - mean) * istd;
    }
    tile[ty * (TW + 1) + tx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DW_TX * DW_TY; i += blockDim.x) {
    const int ty = i / DW_TX, tx = i - ty * DW_TX;
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy >= d.h || gx >= d.w) continue;
    float s = 0.f;
    for (int a = 0; a < l; ++a)
      for (int b = 0; b < l; ++b) s = fmaf(kw[a * l + b], tile[(ty + a) * (TW + 1) + tx + b], s);
    float* yp = d.y.p + (int64_t)n * d.y.sn + (int64_t)gy * d.y.sh + (int64_t)gx * d.y.sw + (int64_t)c * d.y.sc;
    *yp = d.accumulate ? *yp + s : s;
  }
}

// Adjoint: dx[n,c,fold(y+a-r),fold(x+b-r)] += inv_std[c] * k[a,b] * dy[n,c,y,x] (fold = reflection or drop-outside).
// dx must hold the values to accumulate onto (zeros for a plain gradient); atomics because folded taps collide.
__global__ void __launch_bounds__(256) depthwise_bwd_kernel(FdgDepthwise d) {
  extern __shared__ float dsm[];
  const int l = d.l, r = l / 2;
  float* kw = dsm;
  for (int i = threadIdx.x; i < l * l; i += blockDim.x) kw[i] = __ldg(d.kernel + i);
  __syncthreads();
  const int64_t total = (int64_t)d.n * d.c * d.h * d.w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xw = (int)(i % d.w);
    int64_t q = i / d.w;
    const int y = (int)(q % d.h); q /= d.h;
    const int c = (int)(q % d.c);
    const int n = (int)(q / d.c);
    // x = dy (the incoming gradient), y = dx (the gradient w.r.t. the filter input)
    const float g = __ldg(d.x.p + (int64_t)n * d.x.sn + (int64_t)y * d.x.sh + (int64_t)xw * d.x.sw + (int64_t)c * d.x.sc) *
                    (d.inv_std ? __ldg(d.inv_std + c) : 1.f);
    float* op = d.y.p + (int64_t)n * d.y.sn + (int64_t)c * d.y.sc;
    for (int a = 0; a < l; ++a) {
      int yy = y + a - r;
      if (d.pad_mode == 1) yy = reflect_multi(yy, d.h);
      if (yy < 0 || yy >= d.h) continue;
      for (int b = 0; b < l; ++b) {
        int xx = xw + b - r;
        if (d.pad_mode == 1) xx = reflect_multi(xx, d.w);
        if (xx < 0 || xx >= d.w) continue;
        atomicAdd(op + (int64_t)yy * d.y.sh + (int64_t)xx * d.y.sw, kw[a * l + b] * g);
      }
    }
  }
}

// isotropic_gaussian_kernel(l=15, sigma=3) is outer(g, g) with g = exp(-a^2/(2 sigma^2)) / sum (float64 math)
static Gauss15 make_gauss() {
  double g[15], s = 0.0;
  for (int t = 0; t < 15; ++t) { const double a = t - 7.0; g[t] = exp(-(a * a) / 18.0); s += g[t]; }
  Gauss15 k;
  for (int t = 0; t < 15; ++t) k.g[t] = (float)(g[t] / s);
  return k;
}

}  // namespace fdg

using namespace fdg;

extern "C" int fdg_freq_concat_fwd(const FdgTensor* x, const FdgTensor* z, int N, int H, int W, fdg_stream_t stream) {
  FDG_REQUIRE(x && z && x->p && z->p && N > 0, "fdg_freq_concat_fwd: bad arguments");
  FDG_REQUIRE(H > FR && W > FR, "fdg_freq_concat_fwd: reflection padding 7 needs H, W > 7 (got %d x %d)", H, W);
  static const Gauss15 gk = make_gauss();
  dim3 grid(cdiv(W, FT), cdiv(H, FT), N * 3);
  ProfScope prof(PF_FREQ, 2.0 * 234.0 * 3.0 * N * H * W, 48.0 * (double)N * H * W, (cudaStream_t)stream);
  launch_k(freq_fwd_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, *x, *z, H, W, gk);
  return check_launch("fdg_freq_concat_fwd");
}

extern "C" int fdg_freq_concat_bwd(const FdgTensor* dz, const FdgTensor* dx, float* scratch, int N, int H, int W,
                                   fdg_stream_t stream) {
  FDG_REQUIRE(dz && dx && dz->p && dx->p && scratch && N > 0, "fdg_freq_concat_bwd: bad arguments");
  FDG_REQUIRE(H > FR && W > FR, "fdg_freq_concat_bwd: reflection padding 7 needs H, W > 7");
  static const Gauss15 gk = make_gauss();
  const int64_t total = (int64_t)N * 3 * H * W;
  int64_t g = cdiv64(total, 256);
  if (g > 148 * 16) g = 148 * 16;
  launch_k(freq_bwd_rows_kernel, dim3((unsigned)g), dim3(256), (size_t)(0), (cudaStream_t)stream, *dz, scratch, N, H, W, gk);
  int rc = check_launch("fdg_freq_concat_bwd[rows]");
  if (rc != FDG_OK) return rc;
  launch_k(freq_bwd_cols_kernel, dim3((unsigned)g), dim3(256), (size_t)(0), (cudaStream_t)stream, *dz, scratch, *dx, N, H, W, gk);
  return check_launch("fdg_freq_concat_bwd[cols]");
}

static int depthwise_check(const FdgDepthwise* d, const char* who) {
  FDG_REQUIRE(d && d->x.p && d->y.p && d->kernel && d->n > 0 && d->c > 0 && d->h > 0 && d->w > 0, "%s: bad arguments", who);
  FDG_REQUIRE(d->l >= 1 && d->l <= DW_MAXL && (d->l & 1), "%s: filter size must be odd and <= %d (got %d)", who, DW_MAXL, d->l);
  FDG_REQUIRE(d->pad_mode == 0 || d->pad_mode == 1, "%s: pad_mode must be 0 (zero) or 1 (reflect)", who);
  FDG_REQUIRE(d->pad_mode == 0 || (d->h > d->l / 2 && d->w > d->l / 2), "%s: reflection padding %d needs H, W > %d (got %d x %d)", who,
              d->l / 2, d->l / 2, d->h, d->w);
  FDG_REQUIRE((int64_t)d->n * d->c <= 65535, "%s: more than 65535 (image, channel) planes", who);
  return FDG_OK;
}

extern "C" int fdg_depthwise2d_fwd(const FdgDepthwise* d, fdg_stream_t stream) {
  int rc = depthwise_check(d, "fdg_depthwise2d_fwd");
  if (rc != FDG_OK) return rc;
  const int TW = DW_TX + d->l - 1, TH = DW_TY + d->l - 1;
  const size_t smem = sizeof(float) * ((size_t)d->l * d->l + (size_t)TH * (TW + 1));
  dim3 grid(cdiv(d->w, DW_TX), cdiv(d->h, DW_TY), d->n * d->c);
  ProfScope prof(PF_FREQ, 2.0 * d->l * d->l * d->n * d->c * (double)d->h * d->w, 8.0 * d->n * d->c * (double)d->h * d->w, (cudaStream_t)stream);
  depthwise_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(*d);
  return check_launch("fdg_depthwise2d_fwd");
}

extern "C" int fdg_depthwise2d_bwd(const FdgDepthwise* d, fdg_stream_t stream) {
  int rc = depthwise_check(d, "fdg_depthwise2d_bwd");
  if (rc != FDG_OK) return rc;
  const int64_t total = (int64_t)d->n * d->c * d->h * d->w;
  int64_t g = cdiv64(total, 256);
  if (g > 148 * 16) g = 148 * 16;
  ProfScope prof(PF_FREQ, 2.0 * d->l * d->l * (double)total, 8.0 * (double)total, (cudaStream_t)stream);
  depthwise_bwd_kernel<<<(unsigned)g, 256, sizeof(float) * d->l * d->l, (cudaStream_t)stream>>>(*d);
  return check_launch("fdg_depthwise2d_bwd");
}
