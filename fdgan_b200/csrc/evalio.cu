// Output path and quality metrics on the GPU: the image writer's normalisation and the metric
// script of the reference, so that PSNR / SSIM can be evaluated inside a benchmark or validation loop without a
// round trip through PNG files.
//
//  * fdg_image_minmax + fdg_image_pack_u8: torchvision.utils.save_image(normalize=True, scale_each=False) as called at
//    demo.py:151 -- min / max over the whole tensor, (x - min) / max(max - min, 1e-5), * 255 + 0.5, clamp, truncate to
//    uint8, HWC.  The fp32 operations are issued one by one in torch's order (no contraction) so the bytes are identical.
//  * fdg_psnr_ssim_u8: PSNRSSIM.py:201-240 -- 1-pixel border crop; PSNR on /255 values; SSIM per channel with Gaussian
//    weights (scipy.ndimage.gaussian_filter sigma 1.5: 13 taps, 'reflect' boundary), population covariance, data range
//    255, K1 .01, K2 .03, 5-pixel crop before the mean (PSNRSSIM.py:46-194).  The kernel accumulates the four sums
//    (squared error, SSIM map of each channel) in fp64; the host turns them into the two numbers.
#include <atomic>
#include "common.cuh"

namespace fdg {

__global__ void __launch_bounds__(1024) image_minmax_kernel(FdgTensor x, int64_t total, int H, int W, int C, float* out2) {
  __shared__ float smin[32], smax[32];
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const float v = __ldg(x.p + n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + (int64_t)c * x.sc);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    out2[0] = lo;
    out2[1] = hi;
  }
}

// out[n][h][w][c] (uint8, dense HWC per image)
__global__ void image_pack_u8_kernel(FdgTensor x, int64_t total, int H, int W, int C, const float* minmax, uint8_t* out) {
  const float lo = minmax[0], hi = minmax[1];
  const float d = fmaxf(__fsub_rn(hi, lo), 1e-5f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float v = __ldg(x.p + n * x.sn + (int64_t)h * x.sh + (int64_t)w * x.sw + (int64_t)c * x.sc);
    v = fminf(fmaxf(v, lo), hi);                       // clamp_(min, max)
    v = __fdiv_rn(__fsub_rn(v, lo), d);                // sub_(min).div_(max(max - min, 1e-5))
    v = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);          // mul(255).add_(0.5)
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = (uint8_t)v;                               // truncation, like .to(torch.uint8)
  }
}

constexpr int PS_T = 16, PS_R = 6, PS_I = PS_T + 2 * PS_R;   // 13-tap window: truncate 4.0 * sigma 1.5 -> radius 6
__constant__ double c_ps_w[13];

// scipy 'reflect' (half-sample symmetric) index into [0, n)
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// a, b: uint8 [H][W][3]; the metrics run on the 1-pixel-cropped images (h = H - 2, w = W - 2)
__global__ void __launch_bounds__(256) psnr_ssim_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W,
                                                        double* sums /* [0] sum (a-b)^2 / 255^2, [1..3] sum of the cropped SSIM map per channel */) {
  __shared__ float xs[PS_I][PS_I + 1], ys[PS_I][PS_I + 1];
  __shared__ double hq[5][PS_I][PS_T];
  __shared__ double red[2][8];
  const int h = H - 2, w = W - 2;
  const int c = blockIdx.z;
  const int oy0 = blockIdx.y * PS_T, ox0 = blockIdx.x * PS_T;
  const int t = threadIdx.x;
  for (int i = t; i < PS_I * PS_I; i += 256) {
    const int r = i / PS_I, q = i - r * PS_I;
    const int iy = reflect_idx(oy0 - PS_R + r, h), ix = reflect_idx(ox0 - PS_R + q, w);
    const int64_t o = ((int64_t)(iy + 1) * W + (ix + 1)) * 3 + c;
    xs[r][q] = (float)a[o];
    ys[r][q] = (float)b[o];
  }
  __syncthreads();
  for (int i = t; i < PS_I * PS_T; i += 256) {
    const int r = i / PS_T, q = i - r * PS_T;
    double s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      const double wv = c_ps_w[k], xv = xs[r][q + k], yv = ys[r][q + k];
      s1 += wv * xv; s2 += wv * yv; s11 += wv * xv * xv; s22 += wv * yv * yv; s12 += wv * xv * yv;
    }
    hq[0][r][q] = s1; hq[1][r][q] = s2; hq[2][r][q] = s11; hq[3][r][q] = s22; hq[4][r][q] = s12;
  }
  __syncthreads();
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  double ssum = 0, esum = 0;
  for (int i = t; i < PS_T * PS_T; i += 256) {
    const int r = i / PS_T, q = i - r * PS_T;
    const int oy = oy0 + r, ox = ox0 + q;
    if (oy >= h || ox >= w) continue;
    const double d = ((double)xs[r + PS_R][q + PS_R] - (double)ys[r + PS_R][q + PS_R]) / 255.0;
    esum += d * d;
    if (oy < 5 || oy >= h - 5 || ox < 5 || ox >= w - 5) continue;       // 5-pixel crop of the SSIM map
    double ux = 0, uy = 0, uxx = 0, uyy = 0, uxy = 0;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      const double wv = c_ps_w[k];
      ux += wv * hq[0][r + k][q]; uy += wv * hq[1][r + k][q];
      uxx += wv * hq[2][r + k][q]; uyy += wv * hq[3][r + k][q]; uxy += wv * hq[4][r + k][q];
    }
    const double vx = uxx - ux * ux, vy = uyy - uy * uy, vxy = uxy - ux * uy;
    ssum += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
  }
  for (int o = 16; o > 0; o >>= 1) {
    ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    esum += __shfl_xor_sync(0xffffffffu, esum, o);
  }
  if ((t & 31) == 0) { red[0][t >> 5] = ssum; red[1][t >> 5] = esum; }
  __syncthreads();
  if (t == 0) {
    double s = 0, e = 0;
    for (int i = 0; i < 8; ++i) { s += red[0][i]; e += red[1][i]; }
    atomicAdd(sums + 1 + c, s);
    atomicAdd(sums, e);
  }
}

}  // namespace fdg

using namespace fdg;

extern "C" int fdg_image_minmax(const FdgTensor* x, int N, int H, int W, int C, float* out2, fdg_stream_t stream) {
  FDG_REQUIRE(x && x->p && out2 && N > 0 && H > 0 && W > 0 && C > 0, "fdg_image_minmax: bad arguments");
  image_minmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*x, (int64_t)N * H * W * C, H, W, C, out2);
  return check_launch("fdg_image_minmax");
}

extern "C" int fdg_image_pack_u8(const FdgTensor* x, int N, int H, int W, int C, const float* minmax, uint8_t* out, fdg_stream_t stream) {
  FDG_REQUIRE(x && x->p && minmax && out && N > 0 && H > 0 && W > 0 && C > 0, "fdg_image_pack_u8: bad arguments");
  const int64_t total = (int64_t)N * H * W * C;
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  image_pack_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*x, total, H, W, C, minmax, out);
  return check_launch("fdg_image_pack_u8");
}

extern "C" int fdg_psnr_ssim_u8(const uint8_t* ref, const uint8_t* res, int H, int W, double* sums4, fdg_stream_t stream) {
  FDG_REQUIRE(ref && res && sums4 && H >= 13 && W >= 13, "fdg_psnr_ssim_u8: bad arguments (images of at least 13x13 after the crops)");
  static std::atomic<int> window_done[64];         // __constant__ memory is per device
  const int wdev = current_device();
  if (!window_done[wdev]) {   // scipy.ndimage.gaussian_filter(sigma=1.5): radius int(4.0 * 1.5 + 0.5) = 6, normalised exp(-x^2 / (2 sigma^2))
    double w[13], s = 0;
    for (int i = 0; i < 13; ++i) { w[i] = exp(-0.5 * (double)((i - 6) * (i - 6)) / (1.5 * 1.5)); s += w[i]; }
    for (int i = 0; i < 13; ++i) w[i] /= s;
    if (cudaMemcpyToSymbol(c_ps_w, w, sizeof(w)) != cudaSuccess) { set_error("fdg_psnr_ssim_u8: cannot upload the window"); return FDG_ECUDA; }
    window_done[wdev] = 1;
  }
  if (cudaMemsetAsync(sums4, 0, 4 * sizeof(double), (cudaStream_t)stream) != cudaSuccess) { set_error("fdg_psnr_ssim_u8: memset failed"); return FDG_ECUDA; }
  dim3 grid((unsigned)cdiv(W - 2, PS_T), (unsigned)cdiv(H - 2, PS_T), 3);
  psnr_ssim_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ref, res, H, W, sums4);
  return check_launch("fdg_psnr_ssim_u8");
}
