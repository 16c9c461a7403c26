// fdg_conv2d_wgrad, tcgen05 path:  dW[k][co] += sum_pixels a[pixel][k] * g[pixel][co]
//
// GEMM with the PIXEL index as the contraction dimension.  One CTA owns a (filter tap, 128-channel block,
// NT-wide output-channel tile) of dW and a contiguous range of pixels (split-K over pixels across CTAs); it keeps
// the fp32 [128 x NT] partial result in TMEM for its whole lifetime and adds it to global memory once at the end.
//   * A operand = the conv input after the consumer prologue (BatchNorm scale/shift + ReLU/LeakyReLU, pooled /
//     upsampled gather, zero padding), B operand = the output gradient; both are fp32 NHWC in HBM, channel-
//     contiguous, i.e. MN-major for this GEMM, so the loaders store [pixel][64 channels] rows of 128 bytes into the
//     canonical MN-major SWIZZLE_128B layout and the MMA runs with a_major = b_major = MN.
//   * bf16 hi/lo split of both operands, three MMAs per 16-pixel slice (hi*hi + hi*lo + lo*hi), fp32 accumulate.
//   * 8 loader warps cp.async the fp32 rows straight into the operand ring and convert them in
//     place; one thread issues tcgen05.mma; mbarrier ring of STAGES chunks of 32 pixels, STAGES - 2 chunks in flight.
#include <atomic>
#include <cstdlib>

#include "aop.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace fdg {

constexpr int WU_K = 128;                 // channel rows per tile (MMA M)
constexpr int WU_P = 32;                  // pixels per chunk (two K = 16 slices)
constexpr int WU_LOAD_WARPS = 8;          // 256 loader threads: pixel row t >> 3, 8-channel chunk t & 7 of every 64-channel block
constexpr int WU_THREADS = (WU_LOAD_WARPS + 1) * 32;
constexpr int WU_BLK = WU_P * 128;        // one [32 pixels x 64 channels] bf16 block = 4 KB
constexpr int WU_A_BYTES = 2 * WU_BLK;    // 128 channels = two blocks

struct WUArgs {
  FdgWgrad c;
  AOp ao;
  int64_t M;
  int kblocks;       // ceil(R*S*Cin / 128): blocks of the flattened (tap, channel) index
  int co_tiles;      // ceil(Cout / NT)
  int tiles;         // kblocks * co_tiles
  int64_t m_per_split;
  int dbg;           // ablation: 1 no global loads, 2 no split/stores, 4 no MMA
  int gvec;          // gradient rows loadable as float4 (unit channel stride, aligned, Cout % 8 == 0)
  int g_split;       // the gradient operand arrives as split-bf16 planes through gmap_hi / gmap_lo (bulk tensor loads)
  int pf_ahead;      // FAST: chunks of L2 prefetch in front of the ring (0 = none)
  int upt;           // XS: 64-channel units per filter tap, ceil(Cin / 64); the 128 rows of a k block are units 2 kb, 2 kb + 1
  alignas(64) CUtensorMap gmap_hi;
  alignas(64) CUtensorMap gmap_lo;
  alignas(64) CUtensorMap xmap_hi;   // XS: the post-prologue input as split-bf16 planes, {Cin, W, H, N}
  alignas(64) CUtensorMap xmap_lo;
};

// In-place staging: the fp32 rows are cp.async'ed straight into the operand ring.  Each thread's 32 bytes of fp32 per
// 8-channel chunk land in the two 16-byte slots that will hold that chunk's bf16 hi / lo halves, so the conversion is a
// read-modify-write of the thread's own slots and every ring stage doubles as prefetch buffer: STAGES - 2 chunks of
// loads are in flight while one chunk is converted and one is consumed by the tensor core.
// FAST (default; FDG_WU_FAST=0 selects the general loader; round 2: parity suite green, 82.7 -> 81.2 ms/step): loader specialised for the
// dominant call -- 1x1 / stride 1 / no padding, direct gather from a pixel-linear input, split-bf16 gradient -- where a chunk's
// source address is a running pointer and its validity a comparison, so the per-chunk coordinate arithmetic, tap / border
// predicates and the shared-memory metadata word of the general loader disappear (~14 instructions per element
// against ~7.5 of useful work).
// XS (wide RxS weight gradients, stride 1, OW % 32 == 0): BOTH operands arrive as split-bf16 planes through the bulk-tensor engine.
// The generic loader converts every input element once per filter tap and every gradient element once per k block (10x redundant for
// a 160 -> 128 3x3 layer: the kernel ran at 12-40 % of the tensor pipe with the MMAs waiting on the eight loader warps); here the
// planes are written once by an element-wise pass, a filter tap is a coordinate offset of a [64 channels x 32 pixels] box (borders
// zero-filled by the tensor map) and one thread feeds the ring.
// CL (wide tiles with a split-plane gradient, even number of k blocks): CTAs 2i, 2i + 1 (k blocks 2i, 2i + 1 of the same output tile and pixel
// range) form a cluster and SHARE the gradient operand: rank 0 fetches the hi plane's boxes, rank 1 the lo plane's, each multicast
// into both CTAs.  The 320-wide tile moves 56 KB per 32-pixel chunk through the L2 -> SM path (60 B / clock / SM at the MMA rate, more
// than the L2 sustains): the gradient is 40 KB of it.  A stage is free when BOTH CTAs' MMAs have retired (multicast commits).
template <int NT, int STAGES, bool FAST = false, bool XS = false, bool CL = false>
__global__ void __launch_bounds__(WU_THREADS, 1) wgrad_umma_kernel(const __grid_constant__ WUArgs a) {
  static_assert(NT == 64 || NT == 128 || NT == 256 || NT == 320, "output-channel tile");
  constexpr bool CONCAT = NT <= 128;      // [G_hi | G_lo] as one operand of width 2*NT (see umma_chunk8); wide tiles run
                                          // the three passes as N = N1 + N2 MMAs into one [128 x NT] accumulator
  constexpr int GQ = (NT + 63) / 64;      // 64-channel gradient blocks
  constexpr int G_BYTES = GQ * WU_BLK;
  constexpr int STAGE_BYTES = 2 * WU_A_BYTES + 2 * G_BYTES;
  constexpr int TMEM_COLS = CONCAT ? 2 * NT : (NT <= 256 ? 256 : 512);
  constexpr int N1 = NT == 320 ? 192 : NT / 2, N2 = NT - N1;   // block-aligned column split of the wide tiles
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t meta_s[STAGES][WU_LOAD_WARPS * 32];

  const FdgWgrad& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tile = blockIdx.x % a.tiles, split = blockIdx.x / a.tiles;
  const int cot = tile % a.co_tiles;
  const int kb = tile / a.co_tiles;
  const int Ktot = p.R * p.S * p.Cin;
  const int64_t mbeg = (int64_t)split * a.m_per_split;
  const int64_t mend = mbeg + a.m_per_split < a.M ? mbeg + a.m_per_split : a.M;
  const int nchunks = mbeg < mend ? (int)((mend - mbeg + WU_P - 1) / WU_P) : 0;
  const int OHW = p.OH * p.OW;

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), XS ? 1 : WU_LOAD_WARPS + (a.g_split ? 1 : 0));
      mbar_init(smem_u32(&bar_empty[s]), CL ? 2 : 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == WU_LOAD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t crank = CL ? cluster_ctarank() : 0u;
  if (CL) cluster_sync_all();      // the peer's barriers are initialised before anything of this CTA can signal them

  if (XS) {
    if (t == 0 && nchunks > 0) {
      // =============================================================== bulk-tensor producer (both operands)
      const int total_units = p.R * p.S * a.upt;
      int xc[2], xdx[2], xdy[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int u = kb * 2 + h;
        if (u < total_units) {
          const int tap = u / a.upt, fr = tap / p.S;
          xc[h] = (u - tap * a.upt) * 64; xdy[h] = fr - p.pad; xdx[h] = tap - fr * p.S - p.pad;
        } else {
          xc[h] = a.upt * 64; xdy[h] = 0; xdx[h] = 0;       // past the last channel: the box is zero-filled
        }
      }
      int n = (int)(mbeg / OHW);
      const int rem = (int)(mbeg - (int64_t)n * OHW);
      int oy = rem / p.OW, ox = rem - oy * p.OW;              // first pixel of the chunk: OW % 32 == 0, a chunk never leaves its image row
      int64_t lm = mbeg;
      int s = 0;
      uint32_t ph = 0;
      for (int q = 0; q < nchunks; ++q) {
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t bar = smem_u32(&bar_full[s]);
        const uint32_t abase = smem_base + s * STAGE_BYTES, gbase = abase + 2 * WU_A_BYTES;
        mbar_arrive_expect_tx(bar, 2 * WU_A_BYTES + 2 * G_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tma_load_4d(abase + h * WU_BLK, &a.xmap_hi, xc[h], ox + xdx[h], oy + xdy[h], n, bar);
          tma_load_4d(abase + WU_A_BYTES + h * WU_BLK, &a.xmap_lo, xc[h], ox + xdx[h], oy + xdy[h], n, bar);
        }
#pragma unroll
        for (int g = 0; g < GQ; ++g) {
          tma_load_2d(gbase + g * WU_BLK, &a.gmap_hi, cot * NT + 64 * g, (int)lm, bar);
          tma_load_2d(gbase + G_BYTES + g * WU_BLK, &a.gmap_lo, cot * NT + 64 * g, (int)lm, bar);
        }
        lm += WU_P;
        ox += WU_P;
        if (ox >= p.OW) { ox = 0; if (++oy == p.OH) { oy = 0; ++n; } }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < WU_LOAD_WARPS && nchunks > 0) {
    // =============================================================== loaders
    const int pr = t >> 3, seg = t & 7;         // pixel row of the chunk, 8-channel chunk within a 64-channel block
    const bool direct = p.gather == FDG_GATHER_DIRECT;
    int pn, poy, pox;                           // running pixel coordinates of (chunk base + pr)
    {
      const int64_t m = mbeg + pr;
      const int64_t mm = m < a.M ? m : 0;
      pn = (int)(mm / OHW);
      const int rem = (int)(mm - (int64_t)pn * OHW);
      poy = rem / p.OW;
      pox = rem - poy * p.OW;
    }
    int64_t lm = mbeg + pr;                     // pixel index of the next chunk to load
    // A side: two 8-channel chunks in the flattened k = tap*Cin + ci index (Cin % 8 == 0: a chunk never straddles a tap);
    // each has its own filter tap, i.e. its own shifted source pixel
    const int k0 = kb * WU_K + seg * 8, k1 = k0 + 64;
    const int tap0 = k0 < Ktot ? k0 / p.Cin : 0, tap1 = k1 < Ktot ? k1 / p.Cin : 0;
    const int ca0 = k0 < Ktot ? k0 - tap0 * p.Cin : p.Cin, ca1 = k1 < Ktot ? k1 - tap1 * p.Cin : p.Cin;   // >= Cin: chunk is padding
    const int fr0 = tap0 / p.S, fs0 = tap0 - fr0 * p.S, fr1 = tap1 / p.S, fs1 = tap1 - fr1 * p.S;
    const int cg0 = cot * NT + seg * 8;         // G side: first gradient channel of this thread
    float4 scv[4], shv[4];                      // BatchNorm scale/shift of the A thread's 16 input channels
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h ? ca1 : ca0;
      const bool v = p.has_affine && c < p.Cin;
      scv[2 * h] = v ? ld4(p.scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      scv[2 * h + 1] = v ? ld4(p.scale + c + 4) : make_float4(1.f, 1.f, 1.f, 1.f);
      shv[2 * h] = v ? ld4(p.shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      shv[2 * h + 1] = v ? ld4(p.shift + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t roff_c = (uint32_t)pr * 128u + (uint32_t)((seg ^ (pr & 7)) << 4);
    const uint32_t meta0 = smem_u32(&meta_s[0][0]) + (uint32_t)t * 4u;
    const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
    constexpr int NLT = WU_LOAD_WARPS * 32;
    auto slot_a = [&](int st, int h, int lo) -> uint32_t {   // A chunk h (64-channel block h), hi (lo=0) or lo (lo=1) tile
      return smem_base + st * STAGE_BYTES + (uint32_t)lo * WU_A_BYTES + (uint32_t)h * WU_BLK + roff_c;
    };
    auto slot_g = [&](int st, int q, int lo) -> uint32_t {
      return smem_base + st * STAGE_BYTES + 2 * WU_A_BYTES + (uint32_t)lo * G_BYTES + (uint32_t)q * WU_BLK + roff_c;
    };
    auto sts4 = [](uint32_t addr, float4 f) {
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
    };
    auto lds4 = [](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
      return v;
    };
    auto issue_async = [&](int st, uint32_t eph) {
      mbar_wait(empty0 + st * 8, eph ^ 1u);          // the MMAs that read this stage have retired
      uint32_t ok = 0;
      if (lm < mend && !(a.dbg & 1)) {
        {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = h ? ca1 : ca0;
            const int iy = poy * p.stride - p.pad + (h ? fr1 : fr0), ix = pox * p.stride - p.pad + (h ? fs1 : fs0);
            if (c < p.Cin && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
              ok |= 1u << h;
              if (direct) {
                const float* xp = p.x.p + pn * p.x.sn + (int64_t)iy * p.x.sh + (int64_t)ix * p.x.sw + c;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, h, 0)), "l"(xp) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, h, 1)), "l"(xp + 4) : "memory");
              } else {
                sts4(slot_a(st, h, 0), fetch4(a.ao, pn, iy, ix, c));
                sts4(slot_a(st, h, 1), fetch4(a.ao, pn, iy, ix, c + 4));
              }
            }
          }
        }
        if (!a.g_split) {
          const float* gp = p.g.p + pn * p.g.sn + (int64_t)poy * p.g.sh + (int64_t)pox * p.g.sw;
#pragma unroll
          for (int q = 0; q < GQ; ++q) {
            const int c = cg0 + 64 * q;
            if (c < p.Cout) {
              ok |= 4u << q;
              if (a.gvec) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_g(st, q, 0)), "l"(gp + c) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_g(st, q, 1)), "l"(gp + c + 4) : "memory");
              } else {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = c + e < p.Cout ? __ldg(gp + (int64_t)(c + e) * p.g.sc) : 0.f;
                sts4(slot_g(st, q, 0), make_float4(f[0], f[1], f[2], f[3]));
                sts4(slot_g(st, q, 1), make_float4(f[4], f[5], f[6], f[7]));
              }
            }
          }
        }
      }
      if (a.g_split && t == 0) {
        // split-bf16 gradient: [32 pixels x 64 channels] boxes of the hi and lo planes straight into the operand slots
        const uint32_t bar = full0 + st * 8;
        const uint32_t gbase = smem_base + st * STAGE_BYTES + 2 * WU_A_BYTES;
        mbar_arrive_expect_tx(bar, 2 * G_BYTES);
        if (CL) {
          // this CTA's half of the shared gradient operand (rank 0: hi plane, rank 1: lo plane), delivered to both CTAs
#pragma unroll
          for (int q = 0; q < GQ; ++q)
            tma_load_2d_multicast(gbase + crank * G_BYTES + q * WU_BLK, crank ? &a.gmap_lo : &a.gmap_hi, cot * NT + 64 * q, (int)lm, bar, (uint16_t)3);
        } else {
#pragma unroll
          for (int q = 0; q < GQ; ++q) {
            tma_load_2d(gbase + q * WU_BLK, &a.gmap_hi, cot * NT + 64 * q, (int)lm, bar);
            tma_load_2d(gbase + G_BYTES + q * WU_BLK, &a.gmap_lo, cot * NT + 64 * q, (int)lm, bar);
          }
        }
      }
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(meta0 + st * (NLT * 4)), "r"(ok) : "memory");
      lm += WU_P;
      pox += WU_P;
      while (pox >= p.OW) { pox -= p.OW; if (++poy == p.OH) { poy = 0; ++pn; } }
    };
    // convert stage st in place and publish it
    auto convert = [&](int st) {
      uint32_t ok;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ok) : "r"(meta0 + st * (NLT * 4)) : "memory");
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!(a.dbg & 2)) {
        {
          const float sl = p.slope;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 a0 = z4, a1 = z4;
            if ((ok >> h) & 1u) {
              a0 = lds4(slot_a(st, h, 0));
              a1 = lds4(slot_a(st, h, 1));
              if (direct) {
                const float4 sc0 = scv[2 * h], sc1 = scv[2 * h + 1], sh0 = shv[2 * h], sh1 = shv[2 * h + 1];
                a0.x = prologue_act(fmaf(a0.x, sc0.x, sh0.x), sl); a0.y = prologue_act(fmaf(a0.y, sc0.y, sh0.y), sl);
                a0.z = prologue_act(fmaf(a0.z, sc0.z, sh0.z), sl); a0.w = prologue_act(fmaf(a0.w, sc0.w, sh0.w), sl);
                a1.x = prologue_act(fmaf(a1.x, sc1.x, sh1.x), sl); a1.y = prologue_act(fmaf(a1.y, sc1.y, sh1.y), sl);
                a1.z = prologue_act(fmaf(a1.z, sc1.z, sh1.z), sl); a1.w = prologue_act(fmaf(a1.w, sc1.w, sh1.w), sl);
              }
            }
            uint32_t hi[4], lo[4];
            split2(a0.x, a0.y, hi[0], lo[0]); split2(a0.z, a0.w, hi[1], lo[1]);
            split2(a1.x, a1.y, hi[2], lo[2]); split2(a1.z, a1.w, hi[3], lo[3]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_a(st, h, 0)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_a(st, h, 1)), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          }
        }
        if (!a.g_split) {
#pragma unroll
          for (int q = 0; q < GQ; ++q) {
            float4 g0 = z4, g1 = z4;
            if (ok & (4u << q)) { g0 = lds4(slot_g(st, q, 0)); g1 = lds4(slot_g(st, q, 1)); }
            uint32_t hi[4], lo[4];
            split2(g0.x, g0.y, hi[0], lo[0]); split2(g0.z, g0.w, hi[1], lo[1]);
            split2(g1.x, g1.y, hi[2], lo[2]); split2(g1.z, g1.w, hi[3], lo[3]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_g(st, q, 0)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_g(st, q, 1)), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + st * 8);
    };
    if constexpr (FAST) {
      const int wu_pf = a.pf_ahead;
      // ---- FAST: running pointer, validity by comparison (k == channel, tap 0; the thread's rows are pixels mbeg + pr + 32 q)
      const bool fv0 = k0 < p.Cin, fv1 = k1 < p.Cin;
      const float* fx = p.x.p + (mbeg + pr) * p.x.sw;       // pixel-linear view: pixel m lives at m * sw
      const int64_t fxstep = (int64_t)WU_P * p.x.sw;
      int64_t fl = mbeg + pr, fc = mbeg + pr;               // pixel of this thread's row in the next chunk to load / to convert
      auto issue_fast = [&](int st, uint32_t eph) {
        mbar_wait(empty0 + st * 8, eph ^ 1u);
        if (fl < mend) {
          if (fv0) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, 0, 0)), "l"(fx + k0) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, 0, 1)), "l"(fx + k0 + 4) : "memory");
          }
          if (fv1) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, 1, 0)), "l"(fx + k1) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slot_a(st, 1, 1)), "l"(fx + k1 + 4) : "memory");
          }
        }
        if (wu_pf > 0 && fl + (int64_t)wu_pf * WU_P < mend) {
          // L2 prefetch of this thread's rows wu_pf chunks ahead: the ring holds STAGES - 2 chunks of 16 KB in flight, too few bytes per
          // SM to cover the DRAM latency at full bandwidth; with the lines already in L2 the same ring covers the (shorter) L2 latency
          const float* pf = fx + (int64_t)wu_pf * fxstep;
          if (fv0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + k0) : "memory");
          if (fv1) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + k1) : "memory");
        }
        if (t == 0) {       // pr == 0: fl is the first pixel of the chunk
          const uint32_t bar = full0 + st * 8;
          const uint32_t gbase = smem_base + st * STAGE_BYTES + 2 * WU_A_BYTES;
          mbar_arrive_expect_tx(bar, 2 * G_BYTES);
  #pragma unroll
          for (int q = 0; q < GQ; ++q) {
            tma_load_2d(gbase + q * WU_BLK, &a.gmap_hi, cot * NT + 64 * q, (int)fl, bar);
            tma_load_2d(gbase + G_BYTES + q * WU_BLK, &a.gmap_lo, cot * NT + 64 * q, (int)fl, bar);
          }
        }
        fx += fxstep;
        fl += WU_P;
      };
      auto convert_fast = [&](int st) {
        const bool pv = fc < mend;
        const float sl = p.slope;
  #pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
          if (pv && (h ? fv1 : fv0)) {
            a0 = lds4(slot_a(st, h, 0));
            a1 = lds4(slot_a(st, h, 1));
            const float4 sc0 = scv[2 * h], sc1 = scv[2 * h + 1], sh0 = shv[2 * h], sh1 = shv[2 * h + 1];
            a0.x = prologue_act(fmaf(a0.x, sc0.x, sh0.x), sl); a0.y = prologue_act(fmaf(a0.y, sc0.y, sh0.y), sl);
            a0.z = prologue_act(fmaf(a0.z, sc0.z, sh0.z), sl); a0.w = prologue_act(fmaf(a0.w, sc0.w, sh0.w), sl);
            a1.x = prologue_act(fmaf(a1.x, sc1.x, sh1.x), sl); a1.y = prologue_act(fmaf(a1.y, sc1.y, sh1.y), sl);
            a1.z = prologue_act(fmaf(a1.z, sc1.z, sh1.z), sl); a1.w = prologue_act(fmaf(a1.w, sc1.w, sh1.w), sl);
          }
          uint32_t hi[4], lo[4];
          split2(a0.x, a0.y, hi[0], lo[0]); split2(a0.z, a0.w, hi[1], lo[1]);
          split2(a1.x, a1.y, hi[2], lo[2]); split2(a1.z, a1.w, hi[3], lo[3]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_a(st, h, 0)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot_a(st, h, 1)), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + st * 8);
        fc += WU_P;
      };
      int sl_ = 0, sf = 0;            // stages of the next chunk to load / to convert
      uint32_t lph = 0;               // parity of the load side's pass over the ring
  #pragma unroll 1
      for (int q = 0; q < STAGES - 2; ++q) {
        if (q < nchunks) issue_fast(sl_, lph);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++sl_ == STAGES) { sl_ = 0; lph ^= 1u; }
      }
  #pragma unroll 1
      for (int q = 0; q < nchunks; ++q) {
        // refill first (the stage released two chunks ago), then convert chunk q, whose copies have landed when at most
        // STAGES-2 newer groups are pending
        if (q + STAGES - 2 < nchunks) issue_fast(sl_, lph);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++sl_ == STAGES) { sl_ = 0; lph ^= 1u; }
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        convert_fast(sf);
        if (++sf == STAGES) sf = 0;
      }
    } else {
      int sl_ = 0, sf = 0;            // stages of the next chunk to load / to convert
      uint32_t lph = 0;               // parity of the load side's pass over the ring
  #pragma unroll 1
      for (int q = 0; q < STAGES - 2; ++q) {
        if (q < nchunks) issue_async(sl_, lph);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++sl_ == STAGES) { sl_ = 0; lph ^= 1u; }
      }
  #pragma unroll 1
      for (int q = 0; q < nchunks; ++q) {
        // refill first (the stage released two chunks ago), then convert chunk q, whose copies have landed when at most
        // STAGES-2 newer groups are pending
        if (q + STAGES - 2 < nchunks) issue_async(sl_, lph);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (++sl_ == STAGES) { sl_ = 0; lph ^= 1u; }
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        convert(sf);
        if (++sf == STAGES) sf = 0;
      }
    }
  }
  if (warp == WU_LOAD_WARPS && lane == 0 && nchunks > 0) {
    // =============================================================== MMA issue
    constexpr uint32_t idesc1 = umma_idesc_bf16_mn(WU_K, CONCAT ? NT : N1), idesc2 = umma_idesc_bf16_mn(WU_K, CONCAT ? 2 * NT : N2);
    const uint32_t mn_hw = umma_desc_hi(1024);
    int s = 0;
    uint32_t ph = 0;
    for (int kc = 0; kc < nchunks; ++kc) {
      mbar_wait(smem_u32(&bar_full[s]), ph);
      tc_fence_after();
      const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + WU_A_BYTES;
      const uint32_t g_hi = a_lo + WU_A_BYTES;      // [G_hi | G_lo] blocks are adjacent: one operand of width 2*NT
      // 16 pixels per slice = two 8-row swizzle atoms (SBO 1024 B); 64-channel blocks are WU_BLK bytes apart (LBO);
      // consecutive slices are 2048 B = 128 descriptor units apart
      if (!(a.dbg & 4)) {
        const uint32_t ah = umma_desc_lo(a_hi, WU_BLK), al = umma_desc_lo(a_lo, WU_BLK), gh = umma_desc_lo(g_hi, WU_BLK);
        if (CONCAT) {
          umma_concat_slice(tmem_base, ah, al, mn_hw, gh, mn_hw, idesc2, idesc1, kc > 0 ? 1u : 0u, (uint32_t)NT);
          umma_concat_slice(tmem_base, ah + 128u, al + 128u, mn_hw, gh + 128u, mn_hw, idesc2, idesc1, 1u, (uint32_t)NT);
        } else {
          const uint32_t gl = umma_desc_lo(g_hi + G_BYTES, WU_BLK);
          constexpr uint32_t b2 = (uint32_t)(N1 / 64) * (WU_BLK >> 4);      // descriptor units to the first block of the N2 columns
#pragma unroll
          for (uint32_t sl = 0; sl < 2; ++sl) {
            const uint32_t o = sl * 128u;
            const uint32_t acc = (kc > 0 || sl > 0) ? 1u : 0u;
            umma_single(tmem_base, ah + o, mn_hw, gh + o, mn_hw, idesc1, acc);
            umma_single(tmem_base + N1, ah + o, mn_hw, gh + o + b2, mn_hw, idesc2, acc);
            umma_single(tmem_base, ah + o, mn_hw, gl + o, mn_hw, idesc1, 1u);
            umma_single(tmem_base + N1, ah + o, mn_hw, gl + o + b2, mn_hw, idesc2, 1u);
            umma_single(tmem_base, al + o, mn_hw, gh + o, mn_hw, idesc1, 1u);
            umma_single(tmem_base + N1, al + o, mn_hw, gh + o + b2, mn_hw, idesc2, 1u);
          }
        }
      }
      if (CL) umma_commit_multicast(smem_u32(&bar_empty[s]), (uint16_t)3);
      else umma_commit(smem_u32(&bar_empty[s]));
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    umma_commit(smem_u32(&bar_acc));
  }

  // =============================================================== epilogue: TMEM -> atomicAdd into the parameter layout
  if (warp < 4 && nchunks > 0) {
    mbar_wait(smem_u32(&bar_acc), 0);
    __syncwarp();                                 // XS: lane 0 of warp 0 arrives from the producer loop; tcgen05.ld is warp-aligned
    tc_fence_after();
    const int krow = kb * WU_K + warp * 32 + lane;
    bool kvalid = krow < Ktot;
    int tap = kvalid ? krow / p.Cin : 0;
    int ci = kvalid ? krow - tap * p.Cin : 0;
    if (XS) {                                     // rows = two 64-channel units of (tap, channel block), channels padded per tap
      const int u = kb * 2 + (warp >> 1);
      tap = u / a.upt;
      ci = (u - tap * a.upt) * 64 + (warp & 1) * 32 + lane;
      kvalid = u < p.R * p.S * a.upt && ci < p.Cin;
      if (!kvalid) { tap = 0; ci = 0; }
    }
    const int RS = p.R * p.S;
#pragma unroll 1
    for (int g = 0; g < NT / 32; ++g) {
      float v[32];
      const uint32_t tcol = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 32);
      tmem_ld32_nowait(tcol, v);
      if (CONCAT) {
        float v2[32];
        tmem_ld32_nowait(tcol + NT, v2);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] += v2[u];
      } else {
        tmem_ld_wait();
      }
      if (kvalid && cot * NT + g * 32 < p.Cout) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int co = cot * NT + g * 32 + u;
          if (co < p.Cout) {
            const int64_t idx = p.transposed ? (int64_t)ci * p.Cout + co : ((int64_t)co * p.Cin + ci) * RS + tap;
            atomicAdd(p.dw + idx, v[u]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (CL) cluster_sync_all();      // shared memory and barriers stay alive until the peer's last multicast copy / commit has landed
  if (warp == WU_LOAD_WARPS) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// Output-channel tile: one CTA covers as many output channels as TMEM and shared memory allow, because every extra
// channel tile re-reads and re-converts the A operand (and every extra k block the gradient operand).
static inline int wu_ntile(int Cout) {
  static const int wide = [] { const char* e = getenv("FDG_WGRAD_WIDE"); return e ? atoi(e) : 1; }();
  if (Cout <= 64) return 64;
  if (Cout <= 128 || !wide) return 128;
  if (Cout <= 256) return 256;
  if (Cout <= 320) return 320;
  return Cout % 256 == 0 ? 256 : 128;
}

int wgrad_umma_supported(const FdgWgrad* p) {
  if (p->Cin % 8 != 0 || p->Cin < 16 || p->Cout < 1) return 0;
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

template <int NT, int STAGES, bool FAST = false, bool XS = false, bool CL = false>
static int launch_wu(WUArgs& a, cudaStream_t st) {
  constexpr int smem = STAGES * (2 * WU_A_BYTES + 2 * ((NT + 63) / 64) * WU_BLK) + 1024;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(wgrad_umma_kernel<NT, STAGES, FAST, XS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d_wgrad[tcgen05]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  const int num_sms = device_sm_count();
  a.co_tiles = cdiv(a.c.Cout, NT);
  a.tiles = a.kblocks * a.co_tiles;
  // split the pixels so that about one wave of CTAs covers the chip; every split is a whole number of chunks
  int64_t splits = a.tiles >= num_sms ? 1 : num_sms / a.tiles;
  const int64_t max_splits = cdiv64(a.M, 8 * WU_P);
  if (splits > max_splits) splits = max_splits;
  {
    // every split ends with [128 x NT] atomics into the same addresses; with few pixel chunks the epilogues cost more than the chunks
    // (see wgrad_k1.cu): cap the splits at ~sqrt(c * chunks)
    static const int kfac = [] { const char* e = getenv("FDG_WGRAD_SPLIT_FAC"); return e ? atoi(e) : 10; }();
    const int64_t chunks = cdiv64(a.M, WU_P);
    int64_t cap = 1;
    while (cap * cap * 4 < (int64_t)kfac * chunks) ++cap;      // a 32-pixel chunk is a quarter of wgrad_k1's 128-pixel tile
    if (kfac > 0 && splits > cap) splits = cap;
  }
  if (splits < 1) splits = 1;
  a.m_per_split = cdiv64(cdiv64(a.M, splits), WU_P) * WU_P;
  splits = cdiv64(a.M, a.m_per_split);
  ProfScope prof(PF_WGRAD, 2.0 * (double)a.M * a.c.R * a.c.S * a.c.Cin * a.c.Cout,
                 4.0 * ((double)a.M * a.c.Cout + (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  if (CL) launch_k_cluster(wgrad_umma_kernel<NT, STAGES, FAST, XS, CL>, dim3((unsigned)(a.tiles * splits)), dim3(WU_THREADS), (size_t)(smem), st, 2, a);
  else launch_k(wgrad_umma_kernel<NT, STAGES, FAST, XS, CL>, dim3((unsigned)(a.tiles * splits)), dim3(WU_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d_wgrad[tcgen05]");
}

int wgrad_umma(const FdgWgrad* p, cudaStream_t st) {
  WUArgs a;
  a.c = *p;
  a.ao = AOp{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.kblocks = cdiv(p->R * p->S * p->Cin, WU_K);
  a.gvec = vec4_ok(p->g) && (p->Cout % 8 == 0);
  a.dbg = dbg_flags();
  a.g_split = 0;
  if (p->g_split) {
    if (p->Cout % 8 != 0 || a.M >= (1ll << 31)) {      // rows of the planes must be 16-byte multiples; a partial last 64-channel box is zero-filled
      set_error("fdg_conv2d_wgrad[tcgen05]: g_split needs Cout %% 8 == 0");
      return FDG_ENOSUPPORT;
    }
    const uint64_t dims[2] = {(uint64_t)p->Cout, (uint64_t)a.M};
    const uint64_t strides[1] = {(uint64_t)p->Cout * 2};
    const uint32_t box[2] = {64, (uint32_t)WU_P};
    const uint8_t* hi = reinterpret_cast<const uint8_t*>(p->g_split);
    if (!make_tmap_bf16(&a.gmap_hi, hi, 2, dims, strides, box) || !make_tmap_bf16(&a.gmap_lo, hi + (size_t)a.M * p->Cout * 2, 2, dims, strides, box)) {
      set_error("fdg_conv2d_wgrad[tcgen05]: cannot build the tensor maps of the split-bf16 gradient");
      return FDG_ECUDA;
    }
    a.g_split = 1;
  }
  a.upt = 0;
  static const int pf_ahead = [] { const char* e = getenv("FDG_WU_PF"); return e ? atoi(e) : 0; }();   // measured: 0.341 ms without, 0.356 ms with 4 / 8 / 16 chunks (224 -> 128 @256^2): DRAM latency is not what bounds the loaders
  a.pf_ahead = pf_ahead;
  if (p->x_split) {
    // both operands as planes: stride-1 RxS filters whose output rows are whole 32-pixel chunks (see the kernel comment)
    const int nt = wu_ntile(p->Cout);
    if (!(a.g_split && p->stride == 1 && p->gather == FDG_GATHER_DIRECT && p->OW % WU_P == 0 && p->Cin % 8 == 0 && !p->transposed &&
          (nt == 128 || nt == 256) && !a.dbg)) {
      set_error("fdg_conv2d_wgrad[tcgen05]: x_split needs g_split, stride 1, a direct gather, OW %% 32 == 0, Cin %% 8 == 0 and 64 < Cout");
      return FDG_ENOSUPPORT;
    }
    const uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    const uint64_t strides[3] = {(uint64_t)p->Cin * 2, (uint64_t)p->W * p->Cin * 2, (uint64_t)p->H * p->W * p->Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)WU_P, 1, 1};
    const uint8_t* hi = reinterpret_cast<const uint8_t*>(p->x_split);
    const size_t plane = (size_t)p->N * p->H * p->W * p->Cin * 2;
    if (!make_tmap_bf16(&a.xmap_hi, hi, 4, dims, strides, box) || !make_tmap_bf16(&a.xmap_lo, hi + plane, 4, dims, strides, box)) {
      set_error("fdg_conv2d_wgrad[tcgen05]: cannot build the tensor maps of the split-bf16 input");
      return FDG_ECUDA;
    }
    a.upt = cdiv(p->Cin, 64);
    a.kblocks = cdiv(p->R * p->S * a.upt, 2);
    return nt == 128 ? launch_wu<128, 6, false, true>(a, st) : launch_wu<256, 4, false, true>(a, st);
  }
  static const int cl_on = [] { const char* e = getenv("FDG_WU_CLUSTER"); return e ? atoi(e) : 1; }();
  switch (wu_ntile(p->Cout)) {
    case 64: return launch_wu<64, 8>(a, st);      // 8 x 24 KB in-place staging ring
    case 128: {                                   // 6 x 32 KB
      static const int fast_on = [] { const char* e = getenv("FDG_WU_FAST"); return e ? atoi(e) : 1; }();
      const FdgTensor& x = p->x;
      const bool fast = fast_on && a.g_split && !a.dbg && p->R == 1 && p->S == 1 && p->stride == 1 && p->pad == 0 &&
                        p->gather == FDG_GATHER_DIRECT && x.sh == (int64_t)p->W * x.sw && x.sn == (int64_t)p->H * x.sh;
      return fast ? launch_wu<128, 6, true>(a, st) : launch_wu<128, 6, false>(a, st);
    }
    case 256: {                                   // 4 x 48 KB
      if (cl_on && a.g_split && !a.dbg && cdiv(p->Cout, 256) == 1 && a.kblocks % 2 == 0) return launch_wu<256, 4, false, false, true>(a, st);
      return launch_wu<256, 4>(a, st);
    }
    default: {                                    // 3 x 56 KB
      if (cl_on && a.g_split && !a.dbg && a.kblocks % 2 == 0) return launch_wu<320, 3, false, false, true>(a, st);
      return launch_wu<320, 3>(a, st);
    }
  }
}

}  // namespace fdg
