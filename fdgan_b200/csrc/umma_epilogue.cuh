// Epilogue of the tcgen05 convolution kernels.  Four epilogue warps (one per TMEM lane quarter) walk the accumulator
// in 32-channel groups; after tcgen05.ld a lane holds one output pixel.  alpha / bias / activation happen in
// registers, then the group goes into a [128 pixels][32 channels] fp32 staging tile in shared memory laid out exactly
// like a SWIZZLE_128B TMA box (16-byte chunk index XOR row & 7), from where
//   * the common case (plain store into a unit-channel-stride view) leaves with ONE bulk tensor store per group
//     (cp.async.bulk.tensor, clipping of partial tiles / channel tails by the tensor map), or
//   * the other store modes (2x2 nearest-upsample replicas, accumulate, ReLU/LeakyReLU backward mask from a second
//     tensor, strided channels) leave as coalesced 128-bit rows: one instruction = 4 pixels x 128 bytes.
// The per-channel BatchNorm statistics are column sums over the same tile (lane = channel).
#pragma once
#include "aop.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace fdg {

constexpr int EP_TILE_BYTES = 128 * 128;   // staging tile: 128 pixels x 32 fp32 channels

struct EpiTma {
  const void* map;   // tensor map in kernel-parameter space, or nullptr: coalesced path
  int rank;          // 2: {channel, linear pixel}; 4: {channel, x, y, image}
  int c1, c2, c3;    // pixel coordinates of the tile (c1 only for rank 2)
  int wrows;         // 0: one store per group for the whole tile (thread 0, 128-thread barriers); > 0: every warp stores its own 32 staging rows
                     // (rank 2: 32 pixels; rank 4: wrows image rows of the tile) and synchronises with nobody else
};

// v:      the lane's pixel (tile row quarter*32 + lane), channels c0 .. c0+31 (accumulator values)
// mv:     the lane's pixel exists
// yoff:   element offset of the lane's pixel in y (for FDG_STORE_UP2: of its top-left replica), channel 0
// eoff:   element offset of the lane's pixel in e
// et:     thread index within the 128 epilogue threads
// st1, st2: (p.stats only) this warp's running per-channel sums for channels c0 .. c0+31 in shared memory; they receive
//         sum v, sum v^2 over the warp's 32 pixels -- or, in the BatchNorm-backward mode (p.e_scale), sum dz, sum dz*e
template <bool BN>   // BN: the BatchNorm-backward mode (p.e_scale) is compiled in
__device__ __forceinline__ void umma_epilogue_group(const FdgConv& p, bool yvec, bool evec, float (&v)[32], bool mv, int64_t yoff,
                                                    int64_t eoff, int c0, int lane, int quarter, int et, uint32_t tile, const EpiTma& tm,
                                                    float* st1, float* st2, int ncap = 32) {
  // ncap < 32: the group is the 16-channel tail of an 80-wide tile; channels beyond it belong to the NEXT tile and are not touched
  const int nrem = p.Cout - c0 < ncap ? p.Cout - c0 : ncap;
  const int nvalid = nrem < 32 ? nrem : 32;
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
  if (!mv || !full) {
#pragma unroll
    for (int u = 0; u < 32; ++u) if (!mv || u >= nvalid) v[u] = 0.f;
  }
  const bool use_tma = tm.map != nullptr && ncap >= 32;   // a tile-tail group leaves through the coalesced path (a 32-channel box would overrun)
  if (tm.map != nullptr) {
    // the previous group's bulk store must have finished READING the staging tile before it is overwritten
    if (tm.wrows) {
      if (lane == 0) bulk_wait_read0();      // the warp's own rows, its own bulk group
      __syncwarp();
    } else {
      if (et == 0) bulk_wait_read0();
      asm volatile("bar.sync 2, 128;" ::: "memory");
    }
  }
  // ---- staging tile row of this lane's pixel
  const uint32_t wrow0 = tile + (uint32_t)(quarter * 32) * 128u;      // first row of this warp
  {
    const uint32_t trow = wrow0 + (uint32_t)lane * 128u;
    const int sw = lane & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(trow + (uint32_t)((q ^ sw) << 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
  }
  const bool has_e = p.e.p != nullptr;
  const bool bnbwd = BN && p.e_scale != nullptr;
  if (use_tma) {
    fence_proxy_async();                                   // generic-proxy writes -> visible to the bulk-copy engine
    if (tm.wrows) {
      // per-warp stores: four 4 KB boxes per group instead of one 16 KB box, but no warp waits for another one's accumulator reads,
      // staging stores or statistics, and a warp only waits for the bulk engine to have read ITS rows
      __syncwarp();
      if (lane == 0) {
        if (tm.rank == 2) tma_store_2d(tm.map, wrow0, c0, tm.c1 + quarter * 32);
        else tma_store_4d(tm.map, wrow0, c0, tm.c1, tm.c2 + quarter * tm.wrows, tm.c3);
        bulk_commit();
      }
    } else {
      asm volatile("bar.sync 3, 128;" ::: "memory");
      if (et == 0) {
        if (tm.rank == 2) tma_store_2d(tm.map, tile, c0, tm.c1);
        else tma_store_4d(tm.map, tile, c0, tm.c1, tm.c2, tm.c3);
        bulk_commit();
      }
    }
  } else {
    __syncwarp();
    const uint32_t vmask = __ballot_sync(0xffffffffu, mv);
    const int up2 = p.store == FDG_STORE_UP2;
    if (yvec && (nvalid & 3) == 0 && (!has_e || evec)) {
      // ---- coalesced phase: instruction i covers pixels 4i .. 4i+3 of the warp, lane handles channels 4*(lane & 7) .. +3
      const int q = lane & 7, c4 = q * 4;
      const bool cv = c4 < nvalid;
      float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc, ps1 = bsc, ps2 = bsc;
      if (bnbwd && cv) { bsc = ld4(p.e_scale + c0 + c4); bsh = ld4(p.e_shift + c0 + c4); }
      // mask-tensor loads four rows at a time: four independent 128-bit loads in flight per lane, then their stores
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        float4 evs[4];
        if (has_e) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 4 * (4 * hb + i) + (lane >> 3);
            const int64_t eo = __shfl_sync(0xffffffffu, eoff, row);
            evs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (((vmask >> row) & 1u) && cv) evs[i] = ld4(p.e.p + eo + c0 + c4);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = 4 * (4 * hb + i) + (lane >> 3);
          const int64_t yo = __shfl_sync(0xffffffffu, yoff, row);
          if (((vmask >> row) & 1u) && cv) {
            float4 val;
            const uint32_t ta = wrow0 + (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(ta) : "memory");
            if (bnbwd) {
              const float4 ev = evs[i];
              val.x *= fmaf(ev.x, bsc.x, bsh.x) > 0.f ? 1.f : p.eslope; val.y *= fmaf(ev.y, bsc.y, bsh.y) > 0.f ? 1.f : p.eslope;
              val.z *= fmaf(ev.z, bsc.z, bsh.z) > 0.f ? 1.f : p.eslope; val.w *= fmaf(ev.w, bsc.w, bsh.w) > 0.f ? 1.f : p.eslope;
              ps1.x += val.x; ps1.y += val.y; ps1.z += val.z; ps1.w += val.w;
              ps2.x = fmaf(val.x, ev.x, ps2.x); ps2.y = fmaf(val.y, ev.y, ps2.y); ps2.z = fmaf(val.z, ev.z, ps2.z); ps2.w = fmaf(val.w, ev.w, ps2.w);
              val.x *= bsc.x; val.y *= bsc.y; val.z *= bsc.z; val.w *= bsc.w;
            } else if (has_e) {
              const float4 ev = evs[i];
              val.x *= ev.x > 0.f ? 1.f : p.eslope; val.y *= ev.y > 0.f ? 1.f : p.eslope;
              val.z *= ev.z > 0.f ? 1.f : p.eslope; val.w *= ev.w > 0.f ? 1.f : p.eslope;
              if (p.stats) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            }
            float* yp = p.y.p + yo + c0 + c4;
            if (up2) {
              *reinterpret_cast<float4*>(yp) = val;
              *reinterpret_cast<float4*>(yp + p.y.sw) = val;
              *reinterpret_cast<float4*>(yp + p.y.sh) = val;
              *reinterpret_cast<float4*>(yp + p.y.sh + p.y.sw) = val;
            } else if (p.store == FDG_STORE_ACCUM) {
              // fire-and-forget 128-bit reduction at L2: no round trip for the old value (every address is touched by
              // exactly one lane per launch, so the result is deterministic)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yp), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
            } else {
              *reinterpret_cast<float4*>(yp) = val;
            }
          }
        }
      }
      if (bnbwd && p.stats) {
        // lanes l, l+8, l+16, l+24 hold partial sums of the same four channels: fold them, lanes 0..7 publish
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
          ps1.x += __shfl_xor_sync(0xffffffffu, ps1.x, off); ps1.y += __shfl_xor_sync(0xffffffffu, ps1.y, off);
          ps1.z += __shfl_xor_sync(0xffffffffu, ps1.z, off); ps1.w += __shfl_xor_sync(0xffffffffu, ps1.w, off);
          ps2.x += __shfl_xor_sync(0xffffffffu, ps2.x, off); ps2.y += __shfl_xor_sync(0xffffffffu, ps2.y, off);
          ps2.z += __shfl_xor_sync(0xffffffffu, ps2.z, off); ps2.w += __shfl_xor_sync(0xffffffffu, ps2.w, off);
        }
        if (lane < 8 && cv) {
          st1[c4] += ps1.x; st1[c4 + 1] += ps1.y; st1[c4 + 2] += ps1.z; st1[c4 + 3] += ps1.w;
          st2[c4] += ps2.x; st2[c4 + 1] += ps2.y; st2[c4 + 2] += ps2.z; st2[c4 + 3] += ps2.w;
        }
      }
    } else {
      // ---- generic phase (odd channel counts, strided channels, unaligned views): the lane walks its own pixel
      if (mv) {
        const uint32_t trow = wrow0 + (uint32_t)lane * 128u;
#pragma unroll 1
        for (int u = 0; u < nvalid; ++u) {
          const uint32_t ta = trow + (uint32_t)((((u >> 2) ^ (lane & 7)) << 4) + ((u & 3) << 2));
          float val;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(ta) : "memory");
          if (has_e) {
            val *= __ldg(p.e.p + eoff + (int64_t)(c0 + u) * p.e.sc) > 0.f ? 1.f : p.eslope;
            if (p.stats) asm volatile("st.shared.f32 [%0], %1;" ::"r"(ta), "f"(val) : "memory");
          }
          float* yp = p.y.p + yoff + (int64_t)(c0 + u) * p.y.sc;
          if (up2) {
            yp[0] = val; yp[p.y.sw] = val; yp[p.y.sh] = val; yp[p.y.sh + p.y.sw] = val;
          } else if (p.store == FDG_STORE_ACCUM) {
            *yp += val;
          } else {
            *yp = val;
          }
        }
      }
    }
    __syncwarp();
  }
  if (p.stats && !bnbwd) {
    // lane = channel: column sums over the warp's 32 rows (conflict-free: one row per step, 32 distinct words)
    float a1 = 0.f, a2 = 0.f;
    const uint32_t cpos = (uint32_t)(lane >> 2), cw = (uint32_t)(lane & 3) << 2;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      float xv;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv) : "r"(wrow0 + (uint32_t)rr * 128u + (((cpos ^ (uint32_t)(rr & 7)) << 4) + cw)) : "memory");
      a1 += xv;
      a2 = fmaf(xv, xv, a2);
    }
    st1[lane] += a1;
    st2[lane] += a2;
  }
  if (!use_tma) __syncwarp();   // the warp's rows are rewritten by the next group
}

}  // namespace fdg
