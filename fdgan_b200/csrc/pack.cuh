// Weight-operand packing, one work item at a time: shared by the single-job entry points (fdg_pack_weight,
// fdg_pack_weight_umma, fdg_pack_weight_k1) and the table-driven batch kernel (fdg_pack_batch, pack_batch.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace fdg {

// fp32 GEMM operand of a convolution weight; item i = destination element (k, n) = (i / out_ld, i % out_ld)
//   mode 0: [(r,s,ci)][co] <- W[co][ci][r][s]           forward
//   mode 1: [(r,s,co)][ci] <- W[co][ci][R-1-r][S-1-s]   data gradient of a stride-1 convolution
//   mode 2: [(co)][ci]     <- W[ci][co]                 data gradient of the 1x1 ConvTranspose2d (TransitionBlockdy)
__device__ __forceinline__ void pack_w_item(const float* __restrict__ w, int Cout, int Cin, int R, int S, int mode,
                                            float* __restrict__ out, int out_ld, int64_t i) {
  const int n = (int)(i % out_ld);
  const int k = (int)(i / out_ld);
  float v = 0.f;
  if (mode == 0) {
    if (n < Cout) {
      const int tap = k / Cin, ci = k - tap * Cin;
      v = w[((int64_t)n * Cin + ci) * (R * S) + tap];
    }
  } else if (mode == 1) {
    if (n < Cin) {
      const int tap = k / Cout, co = k - tap * Cout;
      const int r = tap / S, s = tap - r * S;
      v = w[((int64_t)co * Cin + n) * (R * S) + (R - 1 - r) * S + (S - 1 - s)];
    }
  } else {
    if (n < Cin) v = w[(int64_t)n * Cout + k];
  }
  out[i] = v;
}

__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// tcgen05 image of the per-tap kernels: out[(ntile, kchunk)] = [hi: NT rows x 128 B, SWIZZLE_128B][lo: same]; source = fp32
// [K][ld] GEMM operand.  Item i = (ntile, kchunk, row n, 16-byte chunk j): 8 consecutive k of one output channel.
// pair (halo kernel, Cin <= 32 and >= 4 taps: the 32 -> 128 data gradients of the growth convolutions): a 64-deep chunk holds TWO
// filter taps, channels 0..31 of tap 2c in its first 64 bytes and of tap 2c + 1 in the second -- no padding streams through the ring.
__device__ __forceinline__ void pack_umma_item(const float* __restrict__ w, int ld, int taps, int Cin, int Cout, int NT, int cchunks,
                                               uint8_t* __restrict__ out, int64_t i, int pair = 0) {
  const int jj = (int)(i & 7);
  int64_t r = i >> 3;
  const int nrow = (int)(r % NT); r /= NT;
  const int nchunks = pair ? (taps + 1) / 2 : taps * cchunks;
  const int kc = (int)(r % nchunks);
  const int nt = (int)(r / nchunks);
  int tap, cbase;
  if (pair) { tap = 2 * kc + (jj >> 2); cbase = (jj & 3) * 8; }
  else { tap = kc / cchunks; cbase = (kc - tap * cchunks) * 64 + jj * 8; }
  const int co = nt * NT + nrow;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ci = cbase + e;
    f[e] = (co < Cout && ci < Cin && tap < taps) ? w[((int64_t)tap * Cin + ci) * ld + co] : 0.f;
  }
  uint4 hi, lo;
  split8(f, hi, lo);
  uint8_t* base = out + ((int64_t)nt * nchunks + kc) * (2 * NT * 128);
  const int off = nrow * 128 + ((jj ^ (nrow & 7)) << 4);
  *reinterpret_cast<uint4*>(base + off) = hi;
  *reinterpret_cast<uint4*>(base + NT * 128 + off) = lo;
}

// tcgen05 image of the growth-convolution kernel (conv_k1.cu): out[(chunk, kx)][n = 0..191][64 k, SWIZZLE_128B]: n < 96: hi of
// (ky = n / 32, co = n % 32); n >= 96: lo of the same.  Source: fp32 GEMM operand w[(ky*3 + kx)*Cin + ci][ld].
constexpr int PACK_K1_B_TILE = 192 * 128;
__device__ __forceinline__ void pack_k1_item(const float* __restrict__ w, int ld, int Cin, int Cout, uint8_t* __restrict__ out, int64_t i) {
  const int jj = (int)(i & 7);
  int64_t r = i >> 3;
  const int nrow = (int)(r % 96); r /= 96;
  const int kx = (int)(r % 3);
  const int cc = (int)(r / 3);
  const int ky = nrow / 32, co = nrow - ky * 32;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ci = cc * 64 + jj * 8 + e;
    f[e] = (co < Cout && ci < Cin) ? w[((int64_t)(ky * 3 + kx) * Cin + ci) * ld + co] : 0.f;
  }
  uint4 hi, lo;
  split8(f, hi, lo);
  uint8_t* base = out + (int64_t)(cc * 3 + kx) * PACK_K1_B_TILE;
  const int nlo = nrow + 96;
  *reinterpret_cast<uint4*>(base + nrow * 128 + ((jj ^ (nrow & 7)) << 4)) = hi;
  *reinterpret_cast<uint4*>(base + nlo * 128 + ((jj ^ (nlo & 7)) << 4)) = lo;
}

}  // namespace fdg
