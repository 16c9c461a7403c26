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

// FR, FS, FC > 0: filter extent and input channels are compile-time (stem / Vgg16 conv1_1 3x3x3, Fusion-D layer 1 4x4x9): the
// input values of a whole filter (K <= 32) or of one filter row are fetched into registers FIRST -- independent loads in flight --
// and multiplied afterwards.  The generic form below issues one dependent load per k, which left the kernel latency-bound at
// 16 warps per SM (stem: 0.29 ms against a 0.06 ms FMA floor).
template <int NC4, int FR = 0, int FS = 0, int FC = 0>   // Cout = 4 * NC4
__global__ void __launch_bounds__(TH_THREADS, 2) conv_thin_kernel(const __grid_constant__ ThinArgs a) {
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
      if constexpr (FR > 0) {
        constexpr int KROW = FS * FC;
        constexpr int RCH = (FR * KROW <= 32) ? FR : 1;      // filter rows per register chunk
        const float sl = p.slope;
#pragma unroll 1
        for (int r0 = 0; r0 < FR; r0 += RCH) {
          float xin[RCH * KROW];
#pragma unroll
          for (int rr = 0; rr < RCH; ++rr) {
            const int iy = iy0 + r0 + rr;
            const bool yin = iy >= 0 && iy < p.H;
            const float* xr = xb + (int64_t)iy * p.x.sh;
#pragma unroll
            for (int s = 0; s < FS; ++s) {
              const int ix = ix0 + s;
              const bool in = yin && ix >= 0 && ix < p.W;
#pragma unroll
              for (int ci = 0; ci < FC; ++ci)
                xin[(rr * FS + s) * FC + ci] = in ? __ldg(xr + (int64_t)ix * p.x.sw + (int64_t)ci * p.x.sc) : 0.f;
            }
          }
          if (sl != 1.f) {
#pragma unroll
            for (int i = 0; i < RCH * KROW; ++i) xin[i] = prologue_act(xin[i], sl);
          }
          const float4* wr = reinterpret_cast<const float4*>(ws + r0 * KROW * CO);
#pragma unroll
          for (int i = 0; i < RCH * KROW; ++i) {
            const float xv = xin[i];
#pragma unroll
            for (int q = 0; q < NC4; ++q) {
              const float4 wv = wr[i * NC4 + q];                 // same address in every lane: broadcast
              v[4 * q] = fmaf(xv, wv.x, v[4 * q]); v[4 * q + 1] = fmaf(xv, wv.y, v[4 * q + 1]);
              v[4 * q + 2] = fmaf(xv, wv.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(xv, wv.w, v[4 * q + 3]);
            }
          }
        }
      } else {
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

// Stem form (3x3 / stride 1 / pad 1, 3 -> 64: conv_refin1 and Vgg16 conv1_1).  With one pixel per thread every k costs 16 warp-wide
// 128-bit weight broadcasts for 64 FMAs per lane: the shared-memory pipe (4 cycles per 128-bit warp load) bounds the kernel at
// ~0.2 ms.  Here a thread owns FOUR horizontally adjacent pixels x 16 channels (channels 16q + 4*cg .. +3, q = 0..3, so that the
// four lanes of a pixel group store 64 contiguous bytes per instruction): 4 weight loads per 64 FMAs, the 18 x 3 input values of
// the four pixels are fetched into registers first.  BatchNorm statistics are per-thread sums folded by shuffles.
constexpr int ST_THREADS = 128;

struct StemArgs {
  FdgConv c;
  int owg;           // groups of four pixels per row
  int64_t groups;    // N * OH * owg
};

template <bool STATS>
__global__ void __launch_bounds__(ST_THREADS, 3) conv_stem4_kernel(const __grid_constant__ StemArgs a) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  __shared__ __align__(16) float ws[27 * 64];
  __shared__ __align__(16) float bs[64];
  __shared__ float red[ST_THREADS / 32][2][64];      // per-warp sums, added in a fixed order (fp32 partial sums stay reproducible)
  const FdgConv& p = a.c;
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < 27 * 64; i += ST_THREADS) ws[i] = __ldg(p.w + (int64_t)(i >> 6) * p.w_ld + (i & 63));
  if (t < 64) bs[t] = p.bias ? __ldg(p.bias + t) : 0.f;
  __syncthreads();
  const int cg = t & 3;
  const bool relu = p.act == FDG_ACT_RELU;
  float s1[STATS ? 16 : 1], s2[STATS ? 16 : 1];     // per-thread statistics of its 16 channels (kept in registers: 3 CTAs of 128 threads per SM)
#pragma unroll
  for (int u = 0; u < (STATS ? 16 : 1); ++u) { s1[u] = 0.f; s2[u] = 0.f; }
  const int64_t gstep = (int64_t)gridDim.x * (ST_THREADS / 4);
  for (int64_t gi = (int64_t)blockIdx.x * (ST_THREADS / 4) + (t >> 2); gi < a.groups; gi += gstep) {
    const int og = (int)(gi % a.owg);
    const int64_t r_ = gi / a.owg;
    const int oy = (int)(r_ % p.OH), n = (int)(r_ / p.OH);
    const int ox0 = og * 4;
    const float* xb = p.x.p + n * p.x.sn;
    // ---- the 3 x 6 x 3 input patch of the four pixels (zero outside the image)
    float xin[3][6][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy - 1 + r;
      const bool yin = iy >= 0 && iy < p.H;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int ix = ox0 - 1 + c;
        const bool in = yin && ix >= 0 && ix < p.W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
          xin[r][c][ci] = in ? __ldg(xb + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw + (int64_t)ci * p.x.sc) : 0.f;
      }
    }
    if (p.slope != 1.f) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) xin[r][c][ci] = prologue_act(xin[r][c][ci], p.slope);
    }
    float v[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int u = 0; u < 16; ++u) v[j][u] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* wk = ws + ((r * 3 + sx) * 3 + ci) * 64 + 4 * cg;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(wk + 16 * q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float xv = xin[r][sx + j][ci];
              v[j][4 * q] = fmaf(xv, wv.x, v[j][4 * q]); v[j][4 * q + 1] = fmaf(xv, wv.y, v[j][4 * q + 1]);
              v[j][4 * q + 2] = fmaf(xv, wv.z, v[j][4 * q + 2]); v[j][4 * q + 3] = fmaf(xv, wv.w, v[j][4 * q + 3]);
            }
          }
        }
    float* yb = p.y.p + n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox0 * p.y.sw + 4 * cg;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = *reinterpret_cast<const float4*>(bs + 16 * q + 4 * cg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float o[4];
        o[0] = fmaf(v[j][4 * q], p.alpha, b4.x); o[1] = fmaf(v[j][4 * q + 1], p.alpha, b4.y);
        o[2] = fmaf(v[j][4 * q + 2], p.alpha, b4.z); o[3] = fmaf(v[j][4 * q + 3], p.alpha, b4.w);
        if (relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
        if (ox0 + j < p.OW) {
          *reinterpret_cast<float4*>(yb + (int64_t)j * p.y.sw + 16 * q) = make_float4(o[0], o[1], o[2], o[3]);
          if (STATS) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { s1[4 * q + u] += o[u]; s2[4 * q + u] = fmaf(o[u], o[u], s2[4 * q + u]); }
          }
        }
      }
    }
  }
  if (STATS) {
    // all lanes are here (the loop has ended for everyone): the eight lanes with equal cg (lane & 3) fold their sums
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int off = 4; off <= 16; off <<= 1) {
        s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], off);
        s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], off);
      }
    }
    if (lane < 4) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int c = 16 * (u >> 2) + 4 * cg + (u & 3);
        red[t >> 5][0][c] = s1[u];
        red[t >> 5][1][c] = s2[u];
      }
    }
    __syncthreads();
    if (t < 64) {
      double d1 = 0.0, d2 = 0.0;
#pragma unroll
      for (int w = 0; w < ST_THREADS / 32; ++w) { d1 += (double)red[w][0][t]; d2 += (double)red[w][1][t]; }
      atomicAdd(p.stats + t, d1);
      atomicAdd(p.stats + p.stats_ld + t, d2);
    }
  }
}

static int conv2d_stem4(const FdgConv* p, cudaStream_t st) {
  StemArgs a;
  a.c = *p;
  a.owg = cdiv(p->OW, 4);
  a.groups = (int64_t)p->N * p->OH * a.owg;
  int64_t blocks = cdiv64(a.groups, ST_THREADS / 4);
  const int64_t cap = (int64_t)device_sm_count() * 12;
  if (blocks > cap) blocks = cap;
  const double M = (double)p->N * p->OH * p->OW;
  ProfScope prof(PF_CONV_SIMT, 2.0 * M * 27 * 64, 4.0 * (M * 64 + (double)p->N * p->H * p->W * 3), st);
  if (p->stats) launch_k(conv_stem4_kernel<true>, dim3((unsigned)blocks), dim3(ST_THREADS), (size_t)(0), st, a);
  else launch_k(conv_stem4_kernel<false>, dim3((unsigned)blocks), dim3(ST_THREADS), (size_t)(0), st, a);
  return check_launch("fdg_conv2d[thin]");
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

template <int NC4, int FR = 0, int FS = 0, int FC = 0>
static int launch_thin(const ThinArgs& a, cudaStream_t st) {
  constexpr int CO = 4 * NC4;
  constexpr int smem = (TH_MAXK * CO + 8 * 32 * (CO + 4)) * 4;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(conv_thin_kernel<NC4, FR, FS, FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d[thin]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  int64_t blocks = cdiv64(a.M, 8 * 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof(PF_CONV_SIMT, 2.0 * (double)a.M * a.K * CO, 4.0 * ((double)a.M * CO + (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  launch_k(conv_thin_kernel<NC4, FR, FS, FC>, dim3((unsigned)blocks), dim3(TH_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d[thin]");
}

int conv2d_thin(const FdgConv* p, cudaStream_t st) {
  static const int stem4 = [] { const char* e = getenv("FDG_STEM4"); return e ? atoi(e) : 1; }();
  if (stem4 && p->Cout == 64 && p->Cin == 3 && p->R == 3 && p->S == 3 && p->stride == 1 && p->pad == 1 &&
      (p->act == FDG_ACT_NONE || p->act == FDG_ACT_RELU))
    return conv2d_stem4(p, st);
  ThinArgs a;
  a.c = *p;
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.K = p->R * p->S * p->Cin;
  switch (p->Cout) {
    case 16: return launch_thin<4>(a, st);
    case 36: return (p->R == 4 && p->S == 4 && p->Cin == 9) ? launch_thin<9, 4, 4, 9>(a, st) : launch_thin<9>(a, st);
    default: return (p->R == 3 && p->S == 3 && p->Cin == 3) ? launch_thin<16, 3, 3, 3>(a, st) : launch_thin<16>(a, st);
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

// Stride-2 form (Fusion-D layer 1): the one-pixel kernel above fetches three 128-bit weight vectors per (tap, output channel) for 12
// FMAs -- the shared-memory pipe bounds it (0.35 ms against a 0.05 ms FMA floor).  Input pixels of equal column parity use the SAME
// filter taps, so a thread takes FOUR of them (x, x + 2, x + 4, x + 6: consecutive output columns) and every weight vector feeds 48 FMAs.
__global__ void __launch_bounds__(128) dgrad_stride2_px4_kernel(const __grid_constant__ FdgDgradStrided p, int xgroups, int64_t total) {
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
    // i -> (image, row, group of eight columns, column parity)
    const int par = (int)(i & 1);
    int64_t r = i >> 1;
    const int xg = (int)(r % xgroups); r /= xgroups;
    const int y = (int)(r % p.H);
    const int n = (int)(r / p.H);
    const int x0 = xg * 8 + par;                   // the thread's pixels: x0 + 2 j
    float acc[4][DS_CI];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < DS_CI; ++c) acc[j][c] = 0.f;
    for (int kr = 0; kr < p.R; ++kr) {
      const int ty = y + p.pad - kr;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= p.OH) continue;
      for (int ks = (x0 + p.pad) & 1; ks < p.S; ks += 2) {       // taps with (x + pad - ks) even
        const int ox0 = (x0 + p.pad - ks) >> 1;                   // may be -1 for the first group (arithmetic shift); pixel j reads ox0 + j
        const float* gp = p.g.p + n * p.g.sn + (int64_t)oy * p.g.sh;
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ok[j] = ox0 + j >= 0 && ox0 + j < p.OW && x0 + 2 * j < p.W;
        const float4* wt = reinterpret_cast<const float4*>(ws + (size_t)(kr * p.S + ks) * p.Cout * DS_CI);
        for (int co = 0; co < p.Cout; co += 4) {
          float gq[4][4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g4 = ok[j] ? ld4(gp + (int64_t)(ox0 + j) * p.g.sw + co) : make_float4(0.f, 0.f, 0.f, 0.f);
            gq[j][0] = g4.x; gq[j][1] = g4.y; gq[j][2] = g4.z; gq[j][3] = g4.w;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 w0 = wt[(co + u) * 3], w1 = wt[(co + u) * 3 + 1], w2 = wt[(co + u) * 3 + 2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float gv = gq[j][u];
              acc[j][0] = fmaf(gv, w0.x, acc[j][0]); acc[j][1] = fmaf(gv, w0.y, acc[j][1]); acc[j][2] = fmaf(gv, w0.z, acc[j][2]); acc[j][3] = fmaf(gv, w0.w, acc[j][3]);
              acc[j][4] = fmaf(gv, w1.x, acc[j][4]); acc[j][5] = fmaf(gv, w1.y, acc[j][5]); acc[j][6] = fmaf(gv, w1.z, acc[j][6]); acc[j][7] = fmaf(gv, w1.w, acc[j][7]);
              acc[j][8] = fmaf(gv, w2.x, acc[j][8]); acc[j][9] = fmaf(gv, w2.y, acc[j][9]); acc[j][10] = fmaf(gv, w2.z, acc[j][10]); acc[j][11] = fmaf(gv, w2.w, acc[j][11]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + 2 * j;
      if (x < p.W) {
        float* o = p.dx.p + n * p.dx.sn + (int64_t)y * p.dx.sh + (int64_t)x * p.dx.sw;
#pragma unroll
        for (int c = 0; c < DS_CI; ++c)
          if (c < p.Cin) {
            float* oc = o + (int64_t)c * p.dx.sc;
            *oc = p.accumulate ? *oc + acc[j][c] : acc[j][c];
          }
      }
    }
  }
}

int dgrad_strided_small(const FdgDgradStrided* p, cudaStream_t st) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  const int smem = p->R * p->S * p->Cout * DS_CI * 4;
  if (!on || p->Cin > DS_CI || p->Cout % 4 != 0 || smem > 48 * 1024 || !vec4_ok(p->g)) return 1;   // 1 = not taken
  static const int px4 = [] { const char* e = getenv("FDG_DGRAD_S2_PX4"); return e ? atoi(e) : 1; }();
  if (px4 && p->stride == 2) {
    const int xgroups = cdiv(p->W, 8);
    const int64_t total4 = (int64_t)p->N * p->H * xgroups * 2;
    int64_t blocks = cdiv64(total4, 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_k(dgrad_stride2_px4_kernel, dim3((unsigned)blocks), dim3(128), (size_t)(smem), st, *p, xgroups, total4);
    return check_launch("fdg_conv2d_dgrad_strided[small]");
  }
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

// Tiled form (default; FDG_WTHIN_TILE=0 selects the batch kernel above).  The batch kernel materialises im2col rows of 32 pixels
// with one scalar load and a tap decode per element and synchronises twice per 32 pixels: 0.74 ms for the head conv_refin3
// (16 -> 3), 0.33 ms for the stem, 0.28 ms for Fusion-D layer 1, 5-15x above their FMA / HBM floors.  Here a CTA owns tiles of
// WTL_TH x WTL_TW output pixels: the INPUT tile with its halo is staged once (prologue applied, zero padding after it), the
// gradient tile as 128-bit rows; a thread owns a KB x CB block of dW in registers for the whole kernel, its KB filter elements
// are fixed offsets into the input tile, and the tile's pixels are dealt round-robin to `nslices` thread groups.  Partial sums
// meet in shared memory (one atomic per element per CTA, two CTAs per SM).
constexpr int WTL_TW = 32, WTL_TH = 4, WTL_PX = WTL_TW * WTL_TH;
constexpr int WTL_KB = 6;

struct WTileArgs {
  FdgWgrad c;
  int K, kgroups, cgroups, nslices, CP;
  int HR, HC;                    // input tile rows / columns (with halo)
  int tiles_x, tiles_y, total_tiles;
  int gvec;
};

__device__ __forceinline__ void cp_async4_zfill(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

template <int CB>
__global__ void __launch_bounds__(256, 2) wgrad_tile_kernel(const __grid_constant__ WTileArgs a) {
  pdl_wait();      // PDL contract (common.cuh): nothing of an earlier kernel is read before this returns
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  const FdgWgrad& p = a.c;
  const int Cin = p.Cin, Cout = p.Cout, CP = a.CP;
  const int xs_n = a.HR * a.HC * Cin;
  const int xs_p = (xs_n + 3) & ~3;
  const int buf_words = xs_p + WTL_PX * CP;        // one stage: input tile [HR][HC][Cin] + gradient tile [WTL_PX][CP]
  const int t = threadIdx.x;
  const int per_slice = a.kgroups * a.cgroups;
  const int slice = t / per_slice;
  const bool worker = slice < a.nslices;
  const int wi = t - slice * per_slice;
  const int kg = worker ? wi / a.cgroups : 0, cg = worker ? wi - kg * a.cgroups : 0;
  int off[WTL_KB];
#pragma unroll
  for (int i = 0; i < WTL_KB; ++i) {
    const int k = kg * WTL_KB + i;
    off[i] = 0;
    if (k < a.K) {
      const int tap = k / Cin, ci = k - tap * Cin;
      const int r = tap / p.S, s2 = tap - r * p.S;
      off[i] = (r * a.HC + s2) * Cin + ci;
    }
  }
  float acc[WTL_KB][CB];
#pragma unroll
  for (int i = 0; i < WTL_KB; ++i)
#pragma unroll
    for (int u = 0; u < CB; ++u) acc[i][u] = 0.f;
  const int tiles_img = a.tiles_x * a.tiles_y;
  const bool nhwc = p.x.sc == 1;
  const int rowlen = a.HC * Cin, plane = a.HR * a.HC;
  const float sl = p.slope;
  const bool prologue = p.has_affine || sl != 1.f;

  // element e of the input tile -> (halo row, halo column, channel); the order follows the memory layout (NHWC: channel fastest,
  // NCHW planes: column fastest) so that the copies coalesce
  auto xdecode = [&](int e, int& hy, int& hx, int& ci) {
    if (nhwc) { hy = e / rowlen; const int q = e - hy * rowlen; hx = q / Cin; ci = q - hx * Cin; }
    else { ci = e / plane; const int q = e - ci * plane; hy = q / a.HC; hx = q - hy * a.HC; }
  };
  auto tile_origin = [&](int tile, int& n, int& oy0, int& ox0) {
    n = tile / tiles_img;
    const int rr = tile - n * tiles_img;
    const int tyi = rr / a.tiles_x;
    oy0 = tyi * WTL_TH;
    ox0 = (rr - tyi * a.tiles_x) * WTL_TW;
  };
  // asynchronous copies of one tile into stage `b` (zero-filled outside the image / past Cout): no registers held in flight
  auto stage = [&](int tile, int b) {
    float* xs = sm + b * buf_words;
    float* gs = xs + xs_p;
    int n, oy0, ox0;
    tile_origin(tile, n, oy0, ox0);
    const int iy0 = oy0 * p.stride - p.pad, ix0 = ox0 * p.stride - p.pad;
    const float* xb = p.x.p + n * p.x.sn;
    for (int e = t; e < xs_n; e += 256) {
      int hy, hx, ci;
      xdecode(e, hy, hx, ci);
      const int iy = iy0 + hy, ix = ix0 + hx;
      const bool in = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      cp_async4_zfill(xs + (hy * a.HC + hx) * Cin + ci, in ? xb + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw + (int64_t)ci * p.x.sc : xb, in);
    }
    const float* gb = p.g.p + n * p.g.sn;
    if (a.gvec) {
      const int c4n = CP >> 2;
      for (int e = t; e < WTL_PX * c4n; e += 256) {
        const int pi = e / c4n, c = (e - pi * c4n) * 4;
        const int oy = oy0 + (pi >> 5), ox = ox0 + (pi & 31);
        const bool in = oy < p.OH && ox < p.OW && c < Cout;
        cp_async16_zfill(gs + pi * CP + c, in ? gb + (int64_t)oy * p.g.sh + (int64_t)ox * p.g.sw + c : gb, in);
      }
    } else {
      // channel-planar gradients (NCHW head): pixel fastest so that the copies coalesce
      for (int e = t; e < WTL_PX * CP; e += 256) {
        const int c = e >> 7, pi = e & (WTL_PX - 1);
        const int oy = oy0 + (pi >> 5), ox = ox0 + (pi & 31);
        const bool in = oy < p.OH && ox < p.OW && c < Cout;
        cp_async4_zfill(gs + pi * CP + c, in ? gb + (int64_t)oy * p.g.sh + (int64_t)ox * p.g.sw + (int64_t)c * p.g.sc : gb, in);
      }
    }
  };

  int b = 0;
  if ((int)blockIdx.x < a.total_tiles) stage(blockIdx.x, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < a.total_tiles) stage(next, b ^ 1);            // stage b ^ 1 was released by the barrier that ended the previous tile
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");     // this thread's copies of the current tile have landed
    float* xs = sm + b * buf_words;
    const float* gs = xs + xs_p;
    if (prologue) {
      // BatchNorm scale/shift + activation in place, by the thread that copied the element; padding stays exactly zero
      int n, oy0, ox0;
      tile_origin(tile, n, oy0, ox0);
      const int iy0 = oy0 * p.stride - p.pad, ix0 = ox0 * p.stride - p.pad;
      for (int e = t; e < xs_n; e += 256) {
        int hy, hx, ci;
        xdecode(e, hy, hx, ci);
        const int iy = iy0 + hy, ix = ix0 + hx;
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
          float* q = xs + (hy * a.HC + hx) * Cin + ci;
          float v = *q;
          if (p.has_affine) v = fmaf(v, __ldg(p.scale + ci), __ldg(p.shift + ci));
          *q = prologue_act(v, sl);
        }
      }
    }
    __syncthreads();
    if (worker) {
      const float* gcol = gs + cg * CB;
#pragma unroll 2
      for (int pi = slice; pi < WTL_PX; pi += a.nslices) {
        const float* xr = xs + (((pi >> 5) * p.stride) * a.HC + (pi & 31) * p.stride) * Cin;
        float gv[CB];
#pragma unroll
        for (int u = 0; u < CB; u += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gcol + pi * CP + u);
          gv[u] = g4.x; gv[u + 1] = g4.y; gv[u + 2] = g4.z; gv[u + 3] = g4.w;
        }
#pragma unroll
        for (int i = 0; i < WTL_KB; ++i) {
          const float xv = xr[off[i]];
#pragma unroll
          for (int u = 0; u < CB; ++u) acc[i][u] = fmaf(xv, gv[u], acc[i][u]);
        }
      }
    }
    __syncthreads();
    b ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- partial sums of the slices meet in shared memory, then one atomic per element and CTA
  float* red = sm;                                 // [K][Cout] (fits: K * Cout <= 160 * 64 floats <= the stages)
  for (int e = t; e < a.K * Cout; e += 256) red[e] = 0.f;
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int i = 0; i < WTL_KB; ++i) {
      const int k = kg * WTL_KB + i;
      if (k < a.K) {
#pragma unroll
        for (int u = 0; u < CB; ++u) {
          const int co = cg * CB + u;
          if (co < Cout) atomicAdd(red + k * Cout + co, acc[i][u]);
        }
      }
    }
  }
  __syncthreads();
  const int RS = p.R * p.S;
  for (int e = t; e < a.K * Cout; e += 256) {
    const int k = e / Cout, co = e - k * Cout;
    const int tap = k / Cin, ci = k - tap * Cin;
    atomicAdd(p.dw + ((int64_t)co * Cin + ci) * RS + tap, red[e]);
  }
}

static int wgrad_tile(const FdgWgrad* p, cudaStream_t st) {
  WTileArgs a;
  a.c = *p;
  a.K = p->R * p->S * p->Cin;
  const int cb = p->Cout > 36 ? 8 : 4;
  a.kgroups = cdiv(a.K, WTL_KB);
  a.cgroups = cdiv(p->Cout, cb);
  if (a.kgroups * a.cgroups > 256) return 1;
  a.nslices = 256 / (a.kgroups * a.cgroups);
  if (a.nslices > 16) a.nslices = 16;
  a.CP = a.cgroups * cb;
  a.HR = (WTL_TH - 1) * p->stride + p->R;
  a.HC = (WTL_TW - 1) * p->stride + p->S;
  a.tiles_x = cdiv(p->OW, WTL_TW);
  a.tiles_y = cdiv(p->OH, WTL_TH);
  a.total_tiles = p->N * a.tiles_x * a.tiles_y;
  a.gvec = vec4_ok(p->g) && p->Cout % 4 == 0;
  const int xs_n = (a.HR * a.HC * p->Cin + 3) & ~3;
  int words = 2 * (xs_n + WTL_PX * a.CP);          // two stages (asynchronous copies of the next tile during the current one)
  if (words < a.K * p->Cout) words = a.K * p->Cout;
  const int smem = words * 4;
  if (smem > 110 * 1024) return 1;
  static std::atomic<int> attr_done[64][2];           // per device and instantiation: largest size configured so far
  const int adev = current_device();
  if (attr_done[adev][cb == 8] < smem) {
    const cudaError_t e = cb == 8 ? cudaFuncSetAttribute(wgrad_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                  : cudaFuncSetAttribute(wgrad_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("fdg_conv2d_wgrad[thin]: cannot raise dynamic shared memory to %d bytes", smem); return FDG_ECUDA; }
    attr_done[adev][cb == 8] = smem;
  }
  int ctas = device_sm_count() * 2;
  if (ctas > a.total_tiles) ctas = a.total_tiles;
  const double M = (double)p->N * p->OH * p->OW;
  ProfScope prof(PF_WGRAD, 2.0 * M * a.K * p->Cout, 4.0 * (M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  if (cb == 8) launch_k(wgrad_tile_kernel<8>, dim3((unsigned)ctas), dim3(256), (size_t)(smem), st, a);
  else launch_k(wgrad_tile_kernel<4>, dim3((unsigned)ctas), dim3(256), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d_wgrad[thin]");
}

// returns 1 when the shape is not taken by this path
int wgrad_thin(const FdgWgrad* p, cudaStream_t st) {
  static const int on = [] { const char* e = getenv("FDG_THIN"); return e ? atoi(e) : 1; }();
  const int K = p->R * p->S * p->Cin;
  if (!on || p->impl != 0 || K > TH_MAXK || p->Cin > 16 || p->Cout > 64 || p->gather != FDG_GATHER_DIRECT || p->transposed) return 1;
  static const int tile_on = [] { const char* e = getenv("FDG_WTHIN_TILE"); return e ? atoi(e) : 1; }();
  if (tile_on && p->stride <= 2) {
    const int rc = wgrad_tile(p, st);
    if (rc != 1) return rc;
  }
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
