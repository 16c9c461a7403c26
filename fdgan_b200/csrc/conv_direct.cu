// fdg_conv2d, direct fp32 paths for the thin layers that are pure HBM traffic: there is no GEMM worth
// tiling when K = R*S*Cin is a few dozen or the input has one channel.
//
//  * conv_thin_kernel: K <= 160, Cout in {16, 36, 64} (stem conv_refin1 3->64, Vgg16 conv1_1 3->64, Fusion-D layer 1
//    9->36 4x4 stride 2).  One thread owns one output pixel and all its channels: the K input values are read once
//    (lanes = consecutive pixels), the weights come from shared memory as warp-wide broadcasts, the result goes through
//    a per-warp shared-memory tile so that the stores are coalesced rows and the BatchNorm statistics are column sums.
//  * conv_cin1_kernel: Cin == 1 (data gradient of Fusion-D layer 5, 1 -> 288 channels, with the LeakyReLU mask of the
//    layer input): one thread per (pixel, 4 channels), R*S taps, 128-bit stores.
#include <atomic>
#include <cstdlib>

#include "aop.cuh"

namespace fdg {

constexpr int TH_THREADS = 256;
constexpr int TH_MAXK = 160;

struct ThinArgs {
  FdgConv c;
  int64_t M;
  int K;
};

template <int NC4>   // Cout = 4 * NC4
__global__ void __launch_bounds__(TH_THREADS) conv_thin_kernel(const __grid_constant__ ThinArgs a) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  constexpr int CO = 4 * NC4;
  constexpr int PITCH = CO + 4;                       // floats per tile row (16-byte aligned rows, conflict-free 128-bit access)
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                                   // [K][CO]
  float* tiles = smem + TH_MAXK * CO;                 // [8 warps][32][PITCH]
  __shared__ float red[2][8][CO];
  const FdgConv& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < a.K * CO; i += TH_THREADS) ws[i] = __ldg(p.w + (int64_t)(i / CO) * p.w_ld + (i % CO));
  __syncthreads();
  float* tile = tiles + warp * 32 * PITCH;
  const int OHW = p.OH * p.OW;
  float st1[(CO + 31) / 32], st2[(CO + 31) / 32];     // running statistics of channels lane, lane + 32
#pragma unroll
  for (int g = 0; g < (CO + 31) / 32; ++g) { st1[g] = 0.f; st2[g] = 0.f; }
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t m0 = ((int64_t)blockIdx.x * 8 + warp) * 32; m0 < a.M; m0 += nwarps * 32) {
    const int64_t m = m0 + lane;
    const bool mv = m < a.M;
    float v[CO];
#pragma unroll
    for (int u = 0; u < CO; ++u) v[u] = 0.f;
    int64_t yoff = 0;
    if (mv) {
      const int n = (int)(m / OHW);
      const int rem = (int)(m - (int64_t)n * OHW);
      const int oy = rem / p.OW, ox = rem - oy * p.OW;
      yoff = n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox * p.y.sw;
      const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
      const float* xb = p.x.p + n * p.x.sn;
      int k = 0;
      for (int r = 0; r < p.R; ++r) {
        const int iy = iy0 + r;
        for (int s = 0; s < p.S; ++s) {
          const int ix = ix0 + s;
          const bool in = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
          const float* xp = xb + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw;
          for (int ci = 0; ci < p.Cin; ++ci, ++k) {
            float xv = in ? __ldg(xp + (int64_t)ci * p.x.sc) : 0.f;
            if (in) xv = prologue_act(xv, p.slope);
            const float4* wr = reinterpret_cast<const float4*>(ws + k * CO);
#pragma unroll
            for (int q = 0; q < NC4; ++q) {
              const float4 wv = wr[q];                 // same address in every lane: broadcast
              v[4 * q] = fmaf(xv, wv.x, v[4 * q]); v[4 * q + 1] = fmaf(xv, wv.y, v[4 * q + 1]);
              v[4 * q + 2] = fmaf(xv, wv.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(xv, wv.w, v[4 * q + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        float r = v[u] * p.alpha + (p.bias ? __ldg(p.bias + u) : 0.f);
        if (p.act == FDG_ACT_RELU) r = fmaxf(r, 0.f);
        else if (p.act == FDG_ACT_TANH) r = tanhf(r);
        else if (p.act == FDG_ACT_SIGMOID) r = 1.f / (1.f + expf(-r));
        v[u] = r;
      }
    }
    // ---- tile row of this pixel
#pragma unroll
    for (int q = 0; q < NC4; ++q)
      *reinterpret_cast<float4*>(tile + lane * PITCH + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
    // ---- coalesced rows: flat float4 index f = row * NC4 + c4
    for (int f = lane; f < 32 * NC4; f += 32) {
      const int row = f / NC4, c4 = f - row * NC4;
      const int64_t yo = __shfl_sync(0xffffffffu, yoff, row);
      if (m0 + row < a.M) *reinterpret_cast<float4*>(p.y.p + yo + 4 * c4) = *reinterpret_cast<const float4*>(tile + row * PITCH + 4 * c4);
    }
    if (p.stats) {
#pragma unroll
      for (int g = 0; g < (CO + 31) / 32; ++g) {
        const int c = g * 32 + lane;
        if (c < CO) {
          float s1 = 0.f, s2 = 0.f;
          for (int rr = 0; rr < 32; ++rr) {
            const float xv = tile[rr * PITCH + c];     // rows past the end hold zeros
            s1 += xv;
            s2 = fmaf(xv, xv, s2);
          }
          st1[g] += s1;
          st2[g] += s2;
        }
      }
    }
    __syncwarp();
  }
  if (p.stats) {
#pragma unroll
    for (int g = 0; g < (CO + 31) / 32; ++g) {
      const int c = g * 32 + lane;
      if (c < CO) { red[0][warp][c] = st1[g]; red[1][warp][c] = st2[g]; }
    }
    __syncthreads();
    for (int c = t; c < CO; c += TH_THREADS) {
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < 8; ++w) { s1 += (double)red[0][w][c]; s2 += (double)red[1][w][c]; }
      atomicAdd(p.stats + c, s1);
      atomicAdd(p.stats + p.stats_ld + c, s2);
    }
  }
}

int conv2d_thin_supported(const FdgConv* p) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  if (!on || p->impl != 0) return 0;                  // impl = 1 keeps the generic SIMT kernel (the tests' fp32 arbiter)
  const int K = p->R * p->S * p->Cin;
  if (K > TH_MAXK || p->Cin > 16 || !(p->Cout == 16 || p->Cout == 36 || p->Cout == 64)) return 0;
  if (p->gather != FDG_GATHER_DIRECT || p->has_affine || p->e.p || p->e_scale || p->store != FDG_STORE_NORMAL) return 0;
  if (!vec4_ok(p->y)) return 0;
  return 1;
}

template <int NC4>
static int launch_thin(const ThinArgs& a, cudaStream_t st) {
  constexpr int CO = 4 * NC4;
  constexpr int smem = (TH_MAXK * CO + 8 * 32 * (CO + 4)) * 4;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(conv_thin_kernel<NC4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d[thin]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  int64_t blocks = cdiv64(a.M, 8 * 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof(PF_CONV_SIMT, 2.0 * (double)a.M * a.K * CO, 4.0 * ((double)a.M * CO + (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  launch_k(conv_thin_kernel<NC4>, dim3((unsigned)blocks), dim3(TH_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d[thin]");
}

int conv2d_thin(const FdgConv* p, cudaStream_t st) {
  ThinArgs a;
  a.c = *p;
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.K = p->R * p->S * p->Cin;
  switch (p->Cout) {
    case 16: return launch_thin<4>(a, st);
    case 36: return launch_thin<9>(a, st);
    default: return launch_thin<16>(a, st);
  }
}

// ------------------------------------------------------------------------------------------------ Cin == 1
struct Cin1Args {
  FdgConv c;
  int64_t total;   // pixels * Cout / 4
  int c4n;         // Cout / 4
  int evec;
};

__global__ void __launch_bounds__(256) conv_cin1_kernel(const __grid_constant__ Cin1Args a) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const FdgConv& p = a.c;
  const int OHW = p.OH * p.OW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / a.c4n;
    const int c = (int)(i - m * a.c4n) * 4;
    const int n = (int)(m / OHW);
    const int rem = (int)(m - (int64_t)n * OHW);
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
    const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
    const float* xb = p.x.p + n * p.x.sn;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.R; ++r) {
      const int iy = iy0 + r;
      if (iy < 0 || iy >= p.H) continue;
      for (int s = 0; s < p.S; ++s) {
        const int ix = ix0 + s;
        if (ix < 0 || ix >= p.W) continue;
        const float xv = prologue_act(__ldg(xb + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw), p.slope);
        const float4 wv = ld4(p.w + (int64_t)(r * p.S + s) * p.w_ld + c);
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      }
    }
    acc.x *= p.alpha; acc.y *= p.alpha; acc.z *= p.alpha; acc.w *= p.alpha;
    if (p.e.p) {
      const float4 ev = ld4(p.e.p + n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw + c);
      acc.x *= ev.x > 0.f ? 1.f : p.eslope; acc.y *= ev.y > 0.f ? 1.f : p.eslope;
      acc.z *= ev.z > 0.f ? 1.f : p.eslope; acc.w *= ev.w > 0.f ? 1.f : p.eslope;
    }
    *reinterpret_cast<float4*>(p.y.p + n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox * p.y.sw + c) = acc;
  }
}

// Same arithmetic (taps accumulated in the same order), four horizontally adjacent pixels x four channels per thread
// (stride 1): a weight vector is loaded once per tap and used for four pixels, the (S + 3) input values of a filter row
// are shared by the four pixels, and four 128-bit mask loads are in flight per thread.  L1 wavefronts per output
// vector drop from ~22 to ~8.  (Tried and measured slower, 0.70 vs 0.55 ms: all R*S weights in registers with one pixel
// per thread -- 122 registers leave 16 warps per SM and one mask load in flight per thread.)
template <int R, int S>
__global__ void __launch_bounds__(256) conv_cin1_px4_kernel(const __grid_constant__ Cin1Args a, int owg, int64_t total) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  const FdgConv& p = a.c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c, og, oy, n;
    if (total <= 0x7fffffffLL) {      // 32-bit index arithmetic (a 64-bit division costs ~100 instructions)
      const int i32 = (int)i, r0 = i32 / a.c4n, r1 = r0 / owg;
      c = (i32 - r0 * a.c4n) * 4; og = r0 - r1 * owg; n = r1 / p.OH; oy = r1 - n * p.OH;
    } else {
      int64_t r_ = i / a.c4n;
      c = (int)(i - r_ * a.c4n) * 4;
      og = (int)(r_ % owg); r_ /= owg;
      oy = (int)(r_ % p.OH);
      n = (int)(r_ / p.OH);
    }
    const int ox0 = og * 4;
    const int npx = p.OW - ox0 < 4 ? p.OW - ox0 : 4;
    float4 ev[4];
    const float* ep = p.e.p ? p.e.p + n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox0 * p.e.sw + c : nullptr;
#pragma unroll
    for (int k = 0; k < 4; ++k) ev[k] = (ep && k < npx) ? ld4(ep + (int64_t)k * p.e.sw) : make_float4(1.f, 1.f, 1.f, 1.f);
    const int iy0 = oy - p.pad, ix0 = ox0 - p.pad;
    const float* xb = p.x.p + n * p.x.sn;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int iy = iy0 + r;
      const bool rin = iy >= 0 && iy < p.H;
      float xs[S + 3];
#pragma unroll
      for (int q = 0; q < S + 3; ++q) {
        const int ix = ix0 + q;
        xs[q] = (rin && ix >= 0 && ix < p.W) ? prologue_act(__ldg(xb + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw), p.slope) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < S; ++q) {
        const float4 wv = ld4(p.w + (int64_t)(r * S + q) * p.w_ld + c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k].x = fmaf(xs[q + k], wv.x, acc[k].x); acc[k].y = fmaf(xs[q + k], wv.y, acc[k].y);
          acc[k].z = fmaf(xs[q + k], wv.z, acc[k].z); acc[k].w = fmaf(xs[q + k], wv.w, acc[k].w);
        }
      }
    }
    float* yp = p.y.p + n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox0 * p.y.sw + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < npx) {
        float4 o = acc[k];
        o.x *= p.alpha; o.y *= p.alpha; o.z *= p.alpha; o.w *= p.alpha;
        if (ep) {
          o.x *= ev[k].x > 0.f ? 1.f : p.eslope; o.y *= ev[k].y > 0.f ? 1.f : p.eslope;
          o.z *= ev[k].z > 0.f ? 1.f : p.eslope; o.w *= ev[k].w > 0.f ? 1.f : p.eslope;
        }
        *reinterpret_cast<float4*>(yp + (int64_t)k * p.y.sw) = o;
      }
    }
  }
}

int conv2d_cin1_supported(const FdgConv* p) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  if (!on || p->impl != 0) return 0;
  if (p->Cin != 1 || p->Cout % 4 != 0 || p->gather != FDG_GATHER_DIRECT || p->has_affine) return 0;
  if (p->bias || p->act != FDG_ACT_NONE || p->stats || p->e_scale || p->store != FDG_STORE_NORMAL) return 0;
  if (!vec4_ok(p->y) || !aligned16(p->w) || p->w_ld % 4 != 0) return 0;
  if (p->e.p && !vec4_ok(p->e)) return 0;
  return 1;
}

int conv2d_cin1(const FdgConv* p, cudaStream_t st) {
  Cin1Args a;
  a.c = *p;
  a.c4n = p->Cout / 4;
  a.total = (int64_t)p->N * p->OH * p->OW * a.c4n;
  a.evec = 1;
  int64_t blocks = cdiv64(a.total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope prof(PF_CONV_SIMT, 2.0 * (double)a.total * 4 * p->R * p->S, 4.0 * (double)a.total * 4 * (p->e.p ? 2 : 1), st);
  static const int px4 = [] { const char* e = getenv("FDG_CIN1_PX4"); return e ? atoi(e) : 1; }();
  if (px4 && p->R == 4 && p->S == 4 && p->stride == 1) {   // Fusion-D layer 5 data gradient
    const int owg = cdiv(p->OW, 4);
    const int64_t total = (int64_t)p->N * p->OH * owg * a.c4n;
    int64_t nb = cdiv64(total, 256);
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if (nb > cap) nb = cap;
    launch_k(conv_cin1_px4_kernel<4, 4>, dim3((unsigned)nb), dim3(256), (size_t)(0), st, a, owg, total);
    return check_launch("fdg_conv2d[cin1]");
  }
  launch_k(conv_cin1_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), st, a);
  return check_launch("fdg_conv2d[cin1]");
}

}  // namespace fdg

// ------------------------------------------------------------------------------------------------ thin-layer gradients
// Data gradient of a strided convolution with few input channels (Fusion-D layer 1: 4x4 stride 2, 36 -> 9): one thread
// per input pixel and all its channels; the gradient rows are read as float4, the weights sit in shared memory as
// [tap][co][12] so that the 9 products of one (tap, co) are three 128-bit broadcasts.
namespace fdg {

constexpr int DS_CI = 12;

__global__ void __launch_bounds__(128) dgrad_strided_small_kernel(const __grid_constant__ FdgDgradStrided p, int64_t total) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ __align__(16) float ws[];      // [R*S][Cout][DS_CI]
  const int taps = p.R * p.S;
  for (int i = threadIdx.x; i < taps * p.Cout * DS_CI; i += blockDim.x) {
    const int ci = i % DS_CI;
    const int co = (i / DS_CI) % p.Cout;
    const int tap = i / (DS_CI * p.Cout);
    ws[i] = ci < p.Cin ? __ldg(p.w + ((int64_t)co * p.Cin + ci) * taps + tap) : 0.f;
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.W);
    int64_t r = i / p.W;
    const int y = (int)(r % p.H);
    const int n = (int)(r / p.H);
    float acc[DS_CI];
#pragma unroll
    for (int c = 0; c < DS_CI; ++c) acc[c] = 0.f;
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
        const float4* wt = reinterpret_cast<const float4*>(ws + (size_t)(kr * p.S + ks) * p.Cout * DS_CI);
        for (int co = 0; co < p.Cout; co += 4) {
          const float4 g4 = ld4(gp + co);
          const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 w0 = wt[(co + u) * 3], w1 = wt[(co + u) * 3 + 1], w2 = wt[(co + u) * 3 + 2];
            acc[0] = fmaf(gq[u], w0.x, acc[0]); acc[1] = fmaf(gq[u], w0.y, acc[1]); acc[2] = fmaf(gq[u], w0.z, acc[2]); acc[3] = fmaf(gq[u], w0.w, acc[3]);
            acc[4] = fmaf(gq[u], w1.x, acc[4]); acc[5] = fmaf(gq[u], w1.y, acc[5]); acc[6] = fmaf(gq[u], w1.z, acc[6]); acc[7] = fmaf(gq[u], w1.w, acc[7]);
            acc[8] = fmaf(gq[u], w2.x, acc[8]); acc[9] = fmaf(gq[u], w2.y, acc[9]); acc[10] = fmaf(gq[u], w2.z, acc[10]); acc[11] = fmaf(gq[u], w2.w, acc[11]);
          }
        }
      }
    }
    float* o = p.dx.p + n * p.dx.sn + (int64_t)y * p.dx.sh + (int64_t)x * p.dx.sw;
#pragma unroll
    for (int c = 0; c < DS_CI; ++c)
      if (c < p.Cin) {
        float* oc = o + (int64_t)c * p.dx.sc;
        *oc = p.accumulate ? *oc + acc[c] : acc[c];
      }
  }
}

int dgrad_strided_small(const FdgDgradStrided* p, cudaStream_t st) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  const int smem = p->R * p->S * p->Cout * DS_CI * 4;
  if (!on || p->Cin > DS_CI || p->Cout % 4 != 0 || smem > 48 * 1024 || !vec4_ok(p->g)) return 1;   // 1 = not taken
  const int64_t total = (int64_t)p->N * p->H * p->W;
  int64_t blocks = cdiv64(total, 128);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(dgrad_strided_small_kernel, dim3((unsigned)blocks), dim3(128), (size_t)(smem), st, *p, total);
  return check_launch("fdg_conv2d_dgrad_strided[small]");
}

// Weight gradient of a thin layer (K = R*S*Cin <= 160, Cout <= 64: stem conv_refin1, Fusion-D layer 1).  A CTA walks
// a range of output pixels in batches of 32: the im2col rows [32][K] and the gradient rows [32][Cout] are staged in
// shared memory with coalesced loads, every thread owns a KB x 4 block of dW in registers, and the CTA adds its partial
// sums atomically into the OIHW parameter layout at the end.
constexpr int WT_P = 32;
constexpr int WT_THREADS = 256;

struct WThinArgs {
  FdgWgrad c;
  int64_t M, m_per_cta;
  int K, kb, kgroups, cgroups;
  int gvec;      // gradient rows loadable as float4
};

template <int KB>
__global__ void __launch_bounds__(WT_THREADS) wgrad_thin_kernel(const __grid_constant__ WThinArgs a) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  const FdgWgrad& p = a.c;
  const int K = a.K, Cout = p.Cout;
  const int KP = a.kgroups * KB;                 // padded K (rows of the x tile)
  const int CP = a.cgroups * 4;                  // padded Cout
  float* xs = sm;                                // [WT_P][KP]
  float* gs = sm + WT_P * KP;                    // [WT_P][CP]
  const int t = threadIdx.x;
  const bool worker = t < a.kgroups * a.cgroups;
  const int kg = worker ? t / a.cgroups : 0, cg = worker ? t % a.cgroups : 0;
  float acc[KB][4];
#pragma unroll
  for (int i = 0; i < KB; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  const int OHW = p.OH * p.OW;
  const int64_t mbeg = (int64_t)blockIdx.x * a.m_per_cta;
  const int64_t mend = mbeg + a.m_per_cta < a.M ? mbeg + a.m_per_cta : a.M;
  const bool kfast = p.x.sc == 1;                // NHWC: consecutive k of one filter row are contiguous
  // index tables: the filter-tap decomposition of k once per CTA, the pixel coordinates once per batch -- the staging
  // loops below do no division by runtime values other than KP / CP
  __shared__ int k_r[TH_MAXK + 8], k_s[TH_MAXK + 8], k_c[TH_MAXK + 8];
  __shared__ int p_n[WT_P], p_y[WT_P], p_x[WT_P];
  for (int k = t; k < KP; k += WT_THREADS) {
    const int kk = k < K ? k : 0;
    const int tap = kk / p.Cin;
    k_c[k] = k < K ? kk - tap * p.Cin : -1;
    k_r[k] = tap / p.S;
    k_s[k] = tap - (tap / p.S) * p.S;
  }
  const bool gvec4 = a.gvec != 0;
  for (int64_t m0 = mbeg; m0 < mend; m0 += WT_P) {
    if (t < WT_P) {
      const int64_t m = m0 + t;
      if (m < mend) {
        const int n = (int)(m / OHW);
        const int rem = (int)(m - (int64_t)n * OHW);
        p_n[t] = n; p_y[t] = rem / p.OW; p_x[t] = rem - (rem / p.OW) * p.OW;
      } else {
        p_n[t] = -1;
      }
    }
    __syncthreads();
    // ---- stage the im2col rows and the gradient rows of 32 pixels
    for (int i = t; i < WT_P * KP; i += WT_THREADS) {
      const int pi = kfast ? i / KP : i & (WT_P - 1);
      const int k = kfast ? i - pi * KP : i >> 5;
      float v = 0.f;
      const int n = p_n[pi], ci = k_c[k];
      if (n >= 0 && ci >= 0) {
        const int iy = p_y[pi] * p.stride - p.pad + k_r[k], ix = p_x[pi] * p.stride - p.pad + k_s[k];
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
          v = __ldg(p.x.p + n * p.x.sn + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw + (int64_t)ci * p.x.sc);
          if (p.has_affine) v = fmaf(v, __ldg(p.scale + ci), __ldg(p.shift + ci));
          v = prologue_act(v, p.slope);
        }
      }
      xs[pi * KP + k] = v;
    }
    if (gvec4) {
      const int c4n = CP >> 2;
      for (int i = t; i < WT_P * c4n; i += WT_THREADS) {
        const int pi = i / c4n, c = (i - pi * c4n) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = p_n[pi];
        if (n >= 0 && c < Cout) v = ld4(p.g.p + n * p.g.sn + (int64_t)p_y[pi] * p.g.sh + (int64_t)p_x[pi] * p.g.sw + c);
        *reinterpret_cast<float4*>(gs + pi * CP + c) = v;
      }
    } else {
      for (int i = t; i < WT_P * CP; i += WT_THREADS) {
        const int pi = i / CP, c = i - pi * CP;
        float v = 0.f;
        const int n = p_n[pi];
        if (n >= 0 && c < Cout) v = __ldg(p.g.p + n * p.g.sn + (int64_t)p_y[pi] * p.g.sh + (int64_t)p_x[pi] * p.g.sw + (int64_t)c * p.g.sc);
        gs[pi * CP + c] = v;
      }
    }
    __syncthreads();
    if (worker) {
#pragma unroll 4
      for (int pi = 0; pi < WT_P; ++pi) {
        const float4 g4 = *reinterpret_cast<const float4*>(gs + pi * CP + cg * 4);
        const float* xr = xs + pi * KP + kg * KB;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
          const float xv = xr[i];
          acc[i][0] = fmaf(xv, g4.x, acc[i][0]); acc[i][1] = fmaf(xv, g4.y, acc[i][1]);
          acc[i][2] = fmaf(xv, g4.z, acc[i][2]); acc[i][3] = fmaf(xv, g4.w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }
  if (worker) {
    const int RS = p.R * p.S;
#pragma unroll
    for (int i = 0; i < KB; ++i) {
      const int k = kg * KB + i;
      if (k < K) {
        const int tap = k / p.Cin, ci = k - tap * p.Cin;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int co = cg * 4 + u;
          if (co < Cout) atomicAdd(p.dw + ((int64_t)co * p.Cin + ci) * RS + tap, acc[i][u]);
        }
      }
    }
  }
}

// returns 1 when the shape is not taken by this path
int wgrad_thin(const FdgWgrad* p, cudaStream_t st) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  const int K = p->R * p->S * p->Cin;
  if (!on || p->impl != 0 || K > TH_MAXK || p->Cin > 16 || p->Cout > 64 || p->gather != FDG_GATHER_DIRECT || p->transposed) return 1;
  WThinArgs a;
  a.c = *p;
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.K = K;
  a.cgroups = cdiv(p->Cout, 4);
  a.gvec = vec4_ok(p->g) && p->Cout % 4 == 0;
  const int kb = cdiv(K * a.cgroups, WT_THREADS) <= 3 ? 3 : 6;    // k rows per thread so that kgroups * cgroups <= 256
  a.kb = kb;
  a.kgroups = cdiv(K, kb);
  if (a.kgroups * a.cgroups > WT_THREADS) return 1;
  static const int cta_mul = [] { const char* e = getenv("FDG_WTHIN_CTAS"); return e ? atoi(e) : 4; }();
  int64_t ctas = 148 * cta_mul;
  a.m_per_cta = cdiv64(cdiv64(a.M, ctas), WT_P) * WT_P;
  ctas = cdiv64(a.M, a.m_per_cta);
  const int smem = WT_P * (a.kgroups * kb + a.cgroups * 4) * 4;
  ProfScope prof(PF_WGRAD, 2.0 * (double)a.M * K * p->Cout, 4.0 * ((double)a.M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  if (kb == 3) launch_k(wgrad_thin_kernel<3>, dim3((unsigned)ctas), dim3(WT_THREADS), (size_t)(smem), st, a);
  else launch_k(wgrad_thin_kernel<6>, dim3((unsigned)ctas), dim3(WT_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d_wgrad[thin]");
}

}  // namespace fdg
