// fdg_ssim_loss_grad: the SSIM term of the generator loss and its gradient.
//
// Reference arithmetic: pytorch_ssim._ssim (models/pytorch_ssim/__init__.py:17-37) -- five depth-wise 11x11 Gaussian
// (sigma 1.5) correlations with zero padding per call (mu1, mu2, E[x^2], E[y^2], E[xy]), the SSIM map and its mean; the
// backward runs five more through autograd.  Here: the window is separable, so one kernel filters the five quantities of
// a 32x32 tile in shared memory (rows, then columns), forms the SSIM map, adds it to the loss and writes the three
// partial derivatives w.r.t. mu1, E[x^2], E[xy]; a second kernel filters those three maps (the window is symmetric, so
// the adjoint of the zero-padded correlation is the same correlation) and combines
//     d sum(S) / dx = w * Dm + 2 x (w * D11) + y (w * D12).
#include "common.cuh"

namespace fdg {

constexpr int SS_T = 32;              // output tile
constexpr int SS_R = 5;               // window radius (11 taps)
constexpr int SS_I = SS_T + 2 * SS_R; // 42 input rows / columns
constexpr int SS_P = SS_I + 1;        // padded pitch


struct SsimArgs {
  FdgTensor x, y, g;
  int N, H, W, C;
  float lscale, gscale;
  int accumulate;
  double* loss;
  float* scratch;     // [3][N*C][H][W]
  float w[11];        // gaussian(11, 1.5): travels with the launch (no __constant__ upload -> legal under stream capture, nothing per device)
};

__device__ __forceinline__ float tget(const FdgTensor& t, int n, int c, int h, int w) {
  return __ldg(t.p + n * t.sn + (int64_t)h * t.sh + (int64_t)w * t.sw + (int64_t)c * t.sc);
}

__global__ void __launch_bounds__(256) ssim_fwd_kernel(const __grid_constant__ SsimArgs a) {
  __shared__ float xs[SS_I][SS_P], ys[SS_I][SS_P];
  __shared__ float hq[5][SS_I][SS_T];
  __shared__ float red[8];
  const int t = threadIdx.x;
  const int plane = blockIdx.z, n = plane / a.C, c = plane - n * a.C;
  const int oy0 = blockIdx.y * SS_T, ox0 = blockIdx.x * SS_T;
  for (int i = t; i < SS_I * SS_I; i += 256) {
    const int r = i / SS_I, q = i - r * SS_I;
    const int iy = oy0 - SS_R + r, ix = ox0 - SS_R + q;
    const bool in = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    xs[r][q] = in ? tget(a.x, n, c, iy, ix) : 0.f;
    ys[r][q] = in ? tget(a.y, n, c, iy, ix) : 0.f;
  }
  __syncthreads();
  for (int i = t; i < SS_I * SS_T; i += 256) {          // horizontal pass over all 42 rows
    const int r = i / SS_T, q = i - r * SS_T;
    float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = a.w[k], xv = xs[r][q + k], yv = ys[r][q + k];
      s1 = fmaf(w, xv, s1); s2 = fmaf(w, yv, s2);
      s11 = fmaf(w, xv * xv, s11); s22 = fmaf(w, yv * yv, s22); s12 = fmaf(w, xv * yv, s12);
    }
    hq[0][r][q] = s1; hq[1][r][q] = s2; hq[2][r][q] = s11; hq[3][r][q] = s22; hq[4][r][q] = s12;
  }
  __syncthreads();
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float ssum = 0.f;
  const int64_t plane_sz = (int64_t)a.H * a.W;
  float* dm = a.scratch + (int64_t)plane * plane_sz;
  float* d11 = dm + (int64_t)a.N * a.C * plane_sz;
  float* d12 = d11 + (int64_t)a.N * a.C * plane_sz;
  for (int i = t; i < SS_T * SS_T; i += 256) {          // vertical pass + SSIM map
    const int r = i / SS_T, q = i - r * SS_T;
    const int oy = oy0 + r, ox = ox0 + q;
    if (oy >= a.H || ox >= a.W) continue;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = a.w[k];
      mu1 = fmaf(w, hq[0][r + k][q], mu1); mu2 = fmaf(w, hq[1][r + k][q], mu2);
      e11 = fmaf(w, hq[2][r + k][q], e11); e22 = fmaf(w, hq[3][r + k][q], e22); e12 = fmaf(w, hq[4][r + k][q], e12);
    }
    const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
    const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2, B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s1 + s2 + C2;
    const float inv = 1.f / (B1 * B2);
    const float S = A1 * A2 * inv;
    ssum += S;
    const float dA1 = A2 * inv, dA2 = A1 * inv, dB1 = -S / B1, dB2 = -S / B2;
    const int64_t o = (int64_t)oy * a.W + ox;
    dm[o] = 2.f * mu2 * (dA1 - dA2) + 2.f * mu1 * (dB1 - dB2);   // dS/dmu1 (through A1, s12, B1, s1)
    d11[o] = dB2;                                                // dS/dE[x^2]
    d12[o] = 2.f * dA2;                                          // dS/dE[xy]
  }
  ssum = warp_sum(ssum);
  if ((t & 31) == 0) red[t >> 5] = ssum;
  __syncthreads();
  if (t == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(a.loss, (double)a.lscale * (double)s);
  }
}

__global__ void __launch_bounds__(256) ssim_bwd_kernel(const __grid_constant__ SsimArgs a) {
  __shared__ float ds[3][SS_I][SS_P];
  __shared__ float hq[3][SS_I][SS_T];
  const int t = threadIdx.x;
  const int plane = blockIdx.z, n = plane / a.C, c = plane - n * a.C;
  const int oy0 = blockIdx.y * SS_T, ox0 = blockIdx.x * SS_T;
  const int64_t plane_sz = (int64_t)a.H * a.W;
  const float* dm = a.scratch + (int64_t)plane * plane_sz;
  const float* d11 = dm + (int64_t)a.N * a.C * plane_sz;
  const float* d12 = d11 + (int64_t)a.N * a.C * plane_sz;
  for (int i = t; i < SS_I * SS_I; i += 256) {
    const int r = i / SS_I, q = i - r * SS_I;
    const int iy = oy0 - SS_R + r, ix = ox0 - SS_R + q;
    const bool in = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    const int64_t o = (int64_t)iy * a.W + ix;
    ds[0][r][q] = in ? dm[o] : 0.f;
    ds[1][r][q] = in ? d11[o] : 0.f;
    ds[2][r][q] = in ? d12[o] : 0.f;
  }
  __syncthreads();
  for (int i = t; i < SS_I * SS_T; i += 256) {
    const int r = i / SS_T, q = i - r * SS_T;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = a.w[k];
      s0 = fmaf(w, ds[0][r][q + k], s0); s1 = fmaf(w, ds[1][r][q + k], s1); s2 = fmaf(w, ds[2][r][q + k], s2);
    }
    hq[0][r][q] = s0; hq[1][r][q] = s1; hq[2][r][q] = s2;
  }
  __syncthreads();
  for (int i = t; i < SS_T * SS_T; i += 256) {
    const int r = i / SS_T, q = i - r * SS_T;
    const int oy = oy0 + r, ox = ox0 + q;
    if (oy >= a.H || ox >= a.W) continue;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = a.w[k];
      f0 = fmaf(w, hq[0][r + k][q], f0); f1 = fmaf(w, hq[1][r + k][q], f1); f2 = fmaf(w, hq[2][r + k][q], f2);
    }
    const float xv = tget(a.x, n, c, oy, ox), yv = tget(a.y, n, c, oy, ox);
    const float gr = a.gscale * (f0 + 2.f * xv * f1 + yv * f2);
    float* gp = a.g.p + n * a.g.sn + (int64_t)oy * a.g.sh + (int64_t)ox * a.g.sw + (int64_t)c * a.g.sc;
    *gp = a.accumulate ? *gp + gr : gr;
  }
}

}  // namespace fdg

using namespace fdg;

extern "C" int fdg_ssim_loss_grad(const FdgTensor* x, const FdgTensor* y, int N, int H, int W, int C, float lscale, float gscale,
                                  const FdgTensor* grad, int accumulate, double* loss, float* scratch, fdg_stream_t stream) {
  FDG_REQUIRE(x && y && x->p && y->p && loss && scratch && N > 0 && H > 0 && W > 0 && C > 0, "fdg_ssim_loss_grad: bad arguments");
  FDG_REQUIRE((int64_t)N * C <= 65535, "fdg_ssim_loss_grad: too many image planes");
  SsimArgs a;
  a.x = *x; a.y = *y;
  a.g = grad ? *grad : FdgTensor{nullptr, 0, 0, 0, 0};
  a.N = N; a.H = H; a.W = W; a.C = C;
  a.lscale = lscale; a.gscale = gscale; a.accumulate = accumulate;
  a.loss = loss; a.scratch = scratch;
  {   // gaussian(11, 1.5) in fp32 like the reference (models/pytorch_ssim/__init__.py:7-9)
    float sum = 0.f;
    for (int i = 0; i < 11; ++i) { a.w[i] = expf(-(float)((i - 5) * (i - 5)) / (2.f * 1.5f * 1.5f)); sum += a.w[i]; }
    for (int i = 0; i < 11; ++i) a.w[i] /= sum;
  }
  dim3 grid((unsigned)cdiv(W, SS_T), (unsigned)cdiv(H, SS_T), (unsigned)(N * C));
  ssim_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  int rc = check_launch("fdg_ssim_loss_grad[fwd]");
  if (rc != FDG_OK || !grad || !grad->p) return rc;
  ssim_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("fdg_ssim_loss_grad[bwd]");
}
