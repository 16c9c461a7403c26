// fdg_conv2d, tcgen05 path for the 3x3 / stride 1 / pad 1 convolutions with at most 32 output channels: the
// 128 -> 32 growth convolutions of the dense blocks (torchvision _DenseLayer.conv2, spec models/densenet.py:197-198;
// 43 per generator forward), BottleneckBlockdy.conv2 of dense_block6 and conv_refin3 (models/dehaze1113.py:264,753).
//
// With N = 32 a tcgen05.mma is bound by shared-memory reads of its A operand (4 KB per MMA), and
// the halo kernel issues one MMA group per filter tap.  Here the three filter ROWS of one filter column kx share a
// single MMA: the B operand is [W(ky=0,kx) | W(ky=1,kx) | W(ky=2,kx)] (3 x 32 columns, hi then lo = 192 rows), the A
// operand is the input halo WITHOUT a row shift, and the accumulator holds per A-pixel (hy, x) the three partial sums
// P_ky(hy, x) = sum_c in(hy, x + kx, c) W(ky, kx)(c, :).  The output pixel (oy, x) is P_0(oy, x) + P_1(oy + 1, x) +
// P_2(oy + 2, x): the row shift moved from the operand fetch (where it costs an MMA per tap) to the epilogue (where it
// is a shifted read of a shared-memory tile).  A tile of 16 x 8 A pixels (UMMA M = 128) yields 14 x 8 outputs.
// Per 64-channel chunk: 3 (kx) x 4 (K slices) x 2 MMAs (A_hi x [B_hi | B_lo], N = 192; A_lo x B_hi, N = 96) instead
// of 9 x 4 x 2, i.e. one third of the A reads and MMA issues.
#include <atomic>
#include <cstdlib>

#include "pack.cuh"
#include "aop.cuh"
#include "umma.cuh"
#include "umma_epilogue.cuh"

namespace fdg {

constexpr int K1_TW = 8, K1_TH = 16, K1_OH = 14;          // A tile 16 x 8 pixels, 14 x 8 outputs
constexpr int K1_HC = K1_TW + 2;                           // halo columns
constexpr int K1_ROWS = K1_TH * K1_HC;                     // 160 halo pixels
constexpr int K1_A_TILE = K1_ROWS * 128;                   // 20 KB per hi (or lo) halo tile of one 64-channel chunk
constexpr int K1_A_STAGE = 2 * K1_A_TILE;
constexpr int K1_B_TILE = 192 * 128;                       // 24 KB: [hi: ky0 ky1 ky2 | lo: ky0 ky1 ky2] x 32 channels
constexpr int K1_BSTAGES = 3;
constexpr int K1_LOAD_WARPS = 8;
constexpr int K1_MMA_WARP = 8, K1_W_WARP = 9, K1_EPI_WARP0 = 10;
constexpr int K1_THREADS = 16 * 32;                       // 4 warpgroups (setmaxnreg works per warpgroup); warps 14, 15 idle
constexpr int K1_ITEMS = K1_ROWS * 8 / (K1_LOAD_WARPS * 32);   // 5 sixteen-byte chunks per loader thread
constexpr int K1_MAX_AFF = 1024;
constexpr int K1_SMEM = 2 * K1_A_STAGE + K1_BSTAGES * K1_B_TILE + 3 * EP_TILE_BYTES + 1024;

struct K1Args {
  FdgConv c;
  int cchunks, tiles_x, tiles_y, total_tiles;
  int yvec;
  int tma_rank;
  alignas(64) CUtensorMap ymap;
};

__global__ void __launch_bounds__(K1_THREADS, 1) conv_k1_kernel(const __grid_constant__ K1Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full[2], bar_a_empty[2], bar_b_full[K1_BSTAGES], bar_b_empty[K1_BSTAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sred[2][4][32];
  __shared__ __align__(16) float aff_s[2][K1_MAX_AFF];

  const FdgConv& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + 2 * K1_A_STAGE;
  const uint32_t ep_base = b_base + K1_BSTAGES * K1_B_TILE;      // S1 | S2 | output staging, 16 KB each
  const int tiles_img = a.tiles_x * a.tiles_y;

  if (t == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_a_full[s]), K1_LOAD_WARPS);
      mbar_init(smem_u32(&bar_a_empty[s]), 1);
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), 4);
    }
    for (int s = 0; s < K1_BSTAGES; ++s) {
      mbar_init(smem_u32(&bar_b_full[s]), 1);
      mbar_init(smem_u32(&bar_b_empty[s]), 1);
    }
    fence_barrier_init();
  }
  if (warp == K1_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  const bool aff_smem = p.has_affine && p.Cin <= K1_MAX_AFF;
  if (aff_smem)
    for (int i = t; i < p.Cin; i += K1_THREADS) { aff_s[0][i] = __ldg(p.scale + i); aff_s[1][i] = __ldg(p.shift + i); }
  for (int i = t; i < 2 * 4 * 32; i += K1_THREADS) (&sred[0][0][0])[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  auto decode = [&](int tile, int& n, int& oy0, int& ox0) {
    n = tile / tiles_img;
    const int r = tile - n * tiles_img;
    const int tyi = r / a.tiles_x;
    oy0 = tyi * K1_OH;
    ox0 = (r - tyi * a.tiles_x) * K1_TW;
  };

  // register re-allocation between the roles: the loaders hold two halo chunks in flight (80 data registers)
  if (warp < K1_LOAD_WARPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;" ::: "memory");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 104;" ::: "memory");

  if (warp < K1_LOAD_WARPS) {
    // =============================================================== halo loaders
    const int j = t & 7;
    const int row0 = t >> 3;                    // halo pixel of item i: row0 + 32 i (coordinates recomputed, not kept in registers)
    const uint32_t full0 = smem_u32(&bar_a_full[0]), empty0 = smem_u32(&bar_a_empty[0]);
    const uint32_t aff0 = smem_u32(&aff_s[0][0]);
    auto lds4u = [](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      return v;
    };
    int buf = 0;
    uint32_t ph = 0;
    // Two tiles of this CTA form a group that shares every weight tile: the work items come in (group, chunk, tile)
    // order.  The loads of item k+1 are issued before item k is converted (register double buffer), so the HBM / L2
    // latency of one halo chunk hides behind the conversion and the stores of the previous one.
    int c_tile0 = blockIdx.x, c_cc = 0, c_tj = 0;               // cursor of the next item to load
    auto issue = [&](float4 (&v0)[K1_ITEMS], float4 (&v1)[K1_ITEMS]) -> uint32_t {
      if (c_tile0 >= a.total_tiles) return 0u;
      const int tile = c_tile0 + c_tj * gridDim.x;
      int n, oy0, ox0;
      decode(tile, n, oy0, ox0);
      const int iy0 = oy0 - 1, ix0 = ox0 - 1;
      const int c = c_cc * 64 + j * 8;
      const bool cvalid = c < p.Cin, cvalid2 = c + 4 < p.Cin;
      const float* tbase = p.x.p + n * p.x.sn + (int64_t)iy0 * p.x.sh + (int64_t)ix0 * p.x.sw + c;   // dereferenced only in range
      uint32_t okmask = 0;
#pragma unroll
      for (int i = 0; i < K1_ITEMS; ++i) {
        const int row = row0 + i * (K1_LOAD_WARPS * 4);
        const int hyi = row / K1_HC, hxi = row - hyi * K1_HC;
        const int iy = iy0 + hyi, ix = ix0 + hxi;
        v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[i] = v0[i];
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && cvalid) {
          okmask |= 1u << i;
          const float* src = tbase + (int64_t)hyi * p.x.sh + (int64_t)hxi * p.x.sw;
          v0[i] = ld4(src);
          if (cvalid2) v1[i] = ld4(src + 4);
        }
      }
      const uint32_t meta = 0x80000000u | okmask | ((uint32_t)c_cc << 8);
      // advance: tile within the group, then chunk, then group
      if (c_tj == 0 && c_tile0 + (int)gridDim.x < a.total_tiles) c_tj = 1;
      else { c_tj = 0; if (++c_cc == a.cchunks) { c_cc = 0; c_tile0 += 2 * gridDim.x; } }
      return meta;
    };
    auto finish = [&](float4 (&v0)[K1_ITEMS], float4 (&v1)[K1_ITEMS], uint32_t meta) {
      const int c = (int)((meta >> 8) & 0xffu) * 64 + j * 8;
      const bool cvalid = c < p.Cin, cvalid2 = c + 4 < p.Cin;
      const float sl = p.slope;
      if (p.has_affine) {
        float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
        if (cvalid) {
          if (aff_smem) {
            sc0 = lds4u(aff0 + c * 4); sh0 = lds4u(aff0 + (K1_MAX_AFF + c) * 4);
            if (cvalid2) { sc1 = lds4u(aff0 + c * 4 + 16); sh1 = lds4u(aff0 + (K1_MAX_AFF + c) * 4 + 16); }
          } else {
            sc0 = ld4(p.scale + c); sh0 = ld4(p.shift + c);
            if (cvalid2) { sc1 = ld4(p.scale + c + 4); sh1 = ld4(p.shift + c + 4); }
          }
        }
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) {
          v0[i].x = fmaf(v0[i].x, sc0.x, sh0.x); v0[i].y = fmaf(v0[i].y, sc0.y, sh0.y);
          v0[i].z = fmaf(v0[i].z, sc0.z, sh0.z); v0[i].w = fmaf(v0[i].w, sc0.w, sh0.w);
          v1[i].x = fmaf(v1[i].x, sc1.x, sh1.x); v1[i].y = fmaf(v1[i].y, sc1.y, sh1.y);
          v1[i].z = fmaf(v1[i].z, sc1.z, sh1.z); v1[i].w = fmaf(v1[i].w, sc1.w, sh1.w);
        }
      }
      if (sl != 1.f) {
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) {
          v0[i].x = prologue_act(v0[i].x, sl); v0[i].y = prologue_act(v0[i].y, sl); v0[i].z = prologue_act(v0[i].z, sl); v0[i].w = prologue_act(v0[i].w, sl);
          v1[i].x = prologue_act(v1[i].x, sl); v1[i].y = prologue_act(v1[i].y, sl); v1[i].z = prologue_act(v1[i].z, sl); v1[i].w = prologue_act(v1[i].w, sl);
        }
      }
      if (p.has_affine) {   // zero padding is applied AFTER the prologue
#pragma unroll
        for (int i = 0; i < K1_ITEMS; ++i) {
          if (!((meta >> i) & 1u)) v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(((meta >> i) & 1u) && cvalid2)) v1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(empty0 + buf * 8, ph ^ 1u);
      const uint32_t a_hi = smem_base + buf * K1_A_STAGE, a_lo = a_hi + K1_A_TILE;
#pragma unroll
      for (int i = 0; i < K1_ITEMS; ++i) {
        const int row = (t >> 3) + i * (K1_LOAD_WARPS * 4);
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        uint32_t h[4], l[4];
        split2(v0[i].x, v0[i].y, h[0], l[0]);
        split2(v0[i].z, v0[i].w, h[1], l[1]);
        split2(v1[i].x, v1[i].y, h[2], l[2]);
        split2(v1[i].z, v1[i].w, h[3], l[3]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + buf * 8);
      if (++buf == 2) { buf = 0; ph ^= 1u; }
    };
    {
      float4 A0[K1_ITEMS], A1[K1_ITEMS], B0[K1_ITEMS], B1[K1_ITEMS];
      uint32_t mA = issue(A0, A1), mB = 0;
      while (mA) {
        mB = issue(B0, B1);
        finish(A0, A1, mA);
        if (!mB) break;
        mA = issue(A0, A1);
        finish(B0, B1, mB);
      }
    }
  } else if (warp == K1_MMA_WARP) {
    // =============================================================== MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 192), idesc1 = umma_idesc_bf16(128, 96);
      const uint32_t a_hw = umma_desc_hi((uint32_t)K1_HC * 128u);     // stride between 8-pixel image rows of the halo
      const uint32_t b_hw = umma_desc_hi(1024);
      int buf = 0, g = 0;
      uint32_t aph = 0, bph = 0;
      // weight slot kx holds W(chunk, kx) for one (group, chunk) phase; both tiles of the group use it
      for (int tile0 = blockIdx.x; tile0 < a.total_tiles; tile0 += 2 * gridDim.x, ++g) {
        const int ntg = tile0 + (int)gridDim.x < a.total_tiles ? 2 : 1;
        for (int cc = 0; cc < a.cchunks; ++cc) {
          for (int tj = 0; tj < ntg; ++tj) {
            if (cc == 0) {                                            // the epilogue drained this accumulator (previous group)
              mbar_wait(smem_u32(&bar_acc_empty[tj]), ((uint32_t)g & 1u) ^ 1u);
              tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)(tj * 256);
            mbar_wait(smem_u32(&bar_a_full[buf]), aph);
            tc_fence_after();
            const uint32_t a_hi = smem_base + buf * K1_A_STAGE, a_lo = a_hi + K1_A_TILE;
            const uint32_t ah0 = umma_desc_lo(a_hi, 16), al0 = umma_desc_lo(a_lo, 16);
            for (int kx = 0; kx < 3; ++kx) {
              if (tj == 0) {
                mbar_wait(smem_u32(&bar_b_full[kx]), bph);
                tc_fence_after();
              }
              const uint32_t shift = (uint32_t)kx * 8u;               // kx halo rows of 128 B, in 16-byte descriptor units
              umma_chunk8(d_tmem, ah0 + shift, al0 + shift, a_hw, umma_desc_lo(b_base + kx * K1_B_TILE, 16), b_hw, idesc2, idesc1,
                          (cc > 0 || kx > 0) ? 1u : 0u, 2u, 2u, 96u);
              if (tj == ntg - 1) umma_commit(smem_u32(&bar_b_empty[kx]));   // last user of this weight tile in the phase
            }
            umma_commit(smem_u32(&bar_a_empty[buf]));
            if (++buf == 2) { buf = 0; aph ^= 1u; }
            if (cc == a.cchunks - 1) umma_commit(smem_u32(&bar_acc_full[tj]));
          }
          bph ^= 1u;
        }
      }
    }
  } else if (warp == K1_W_WARP) {
    // =============================================================== weight-tile producer (bulk-copy ring)
    if (lane == 0) {
      uint32_t bph = 0;
      for (int tile0 = blockIdx.x; tile0 < a.total_tiles; tile0 += 2 * gridDim.x) {
        for (int cc = 0; cc < a.cchunks; ++cc) {
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(smem_u32(&bar_b_empty[kx]), bph ^ 1u);
            const uint32_t bar = smem_u32(&bar_b_full[kx]);
            mbar_arrive_expect_tx(bar, K1_B_TILE);
            bulk_g2s(b_base + kx * K1_B_TILE, reinterpret_cast<const uint8_t*>(p.w_k1) + (size_t)(cc * 3 + kx) * K1_B_TILE, K1_B_TILE, bar);
          }
          bph ^= 1u;
        }
      }
    }
  } else if (warp < K1_EPI_WARP0 + 4) {
    // =============================================================== epilogue warps (TMEM lane quarter = warp & 3)
    const int quarter = warp & 3;
    const int et = t - K1_EPI_WARP0 * 32;
    const uint32_t s1 = ep_base, s2 = ep_base + EP_TILE_BYTES, so = ep_base + 2 * EP_TILE_BYTES;
    const bool evec = p.e.p && p.e.sc == 1 && aligned16_dev(p.e.p) && (p.e.sn % 4 == 0) && (p.e.sh % 4 == 0) && (p.e.sw % 4 == 0);
    const int m = quarter * 32 + lane;                 // A pixel of this lane: halo row hy, column x
    const int hy = m >> 3, x = m & 7;
    const uint32_t myrow = (uint32_t)m * 128u;
    const int sw = m & 7;
    auto sts_row = [&](uint32_t tile_base, const float (&v)[32]) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile_base + myrow + (uint32_t)((q ^ sw) << 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                     "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
    };
    auto add_row = [&](uint32_t tile_base, int row, float (&v)[32]) {
      const uint32_t rb = tile_base + (uint32_t)row * 128u;
      const int rs = row & 7;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 f;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(rb + (uint32_t)((q ^ rs) << 4)) : "memory");
        v[4 * q] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      int n, oy0, ox0;
      decode(tile, n, oy0, ox0);
      const int b = it & 1;
      while (!mbar_try_wait(smem_u32(&bar_acc_full[b]), ((uint32_t)it >> 1) & 1u)) __nanosleep(200);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 256);
      float v[32];
      {
        float v2[32];
        // P_1 (filter row 1) and P_2 (filter row 2) of this A pixel go to the shift tiles, P_0 stays in registers
        tmem_ld32_nowait(tcol + 32, v); tmem_ld32_nowait(tcol + 96 + 32, v2); tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] += v2[u];
        sts_row(s1, v);
        tmem_ld32_nowait(tcol + 64, v); tmem_ld32_nowait(tcol + 96 + 64, v2); tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] += v2[u];
        sts_row(s2, v);
        tmem_ld32_nowait(tcol, v); tmem_ld32_nowait(tcol + 96, v2); tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) v[u] += v2[u];
      }
      // the accumulator is drained: hand it back to the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
      asm volatile("bar.sync 4, 128;" ::: "memory");          // shift tiles complete (rows of the other warps included)
      const int oy = oy0 + hy, ox = ox0 + x;
      const bool mv = hy < K1_OH && oy < p.OH && ox < p.OW;
      if (hy < K1_OH) {
        add_row(s1, m + 8, v);                                 // P_1 of the pixel one image row below
        add_row(s2, m + 16, v);                                // P_2 two rows below
      }
      int64_t yoff = 0, eoff = 0;
      if (mv) {
        const int us = p.store == FDG_STORE_UP2 ? 2 : 1;
        yoff = n * p.y.sn + (int64_t)(us * oy) * p.y.sh + (int64_t)(us * ox) * p.y.sw;
        if (p.e.p) eoff = n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw;
      }
      const EpiTma tm{a.tma_rank ? (const void*)&a.ymap : nullptr, a.tma_rank, ox0, oy0, n, 0};   // 14 of the tile's 16 rows are outputs: one box for the tile
      umma_epilogue_group<true>(p, a.yvec, evec, v, mv, yoff, eoff, 0, lane, quarter, et, so, tm, &sred[0][quarter][0], &sred[1][quarter][0]);
      asm volatile("bar.sync 4, 128;" ::: "memory");          // every warp is done reading the shift tiles
    }
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < 32 && et < p.Cout) {
        atomicAdd(p.stats + et, (double)((sred[0][0][et] + sred[0][1][et]) + (sred[0][2][et] + sred[0][3][et])));
        atomicAdd(p.stats + p.stats_ld + et, (double)((sred[1][0][et] + sred[1][1][et]) + (sred[1][2][et] + sred[1][3][et])));
      }
    }
    if (a.tma_rank && et == 0) bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == K1_MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight image
// out[(chunk, kx)][n = 0..191][64 k, SWIZZLE_128B]: n < 96: hi of (ky = n / 32, co = n % 32); n >= 96: lo of the same.
// Source: fp32 GEMM operand w[(ky*3 + kx)*Cin + ci][ld].
__global__ void pack_k1_kernel(const float* __restrict__ w, int ld, int Cin, int Cout, int cchunks, uint8_t* __restrict__ out, int64_t total) {
  static_assert(PACK_K1_B_TILE == K1_B_TILE, "pack.cuh and conv_k1.cu disagree on the weight tile size");
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    pack_k1_item(w, ld, Cin, Cout, out, i);
}

static int g_k1_on = [] { const char* e = getenv("FDG_K1"); return e ? atoi(e) : 1; }();
void set_k1(int on) { g_k1_on = on; }

int conv2d_k1_supported(const FdgConv* p) {
  if (!g_k1_on || !p->w_k1) return 0;
  if (p->gather != FDG_GATHER_DIRECT || p->stride != 1 || p->pad != 1 || p->R != 3 || p->S != 3) return 0;
  if (p->Cin % 4 != 0 || p->Cin < 16 || p->Cout < 1 || p->Cout > 32) return 0;
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

int conv2d_k1(const FdgConv* p, cudaStream_t st) {
  K1Args a;
  a.c = *p;
  a.cchunks = cdiv(p->Cin, 64);
  a.tiles_x = cdiv(p->OW, K1_TW);
  a.tiles_y = cdiv(p->OH, K1_OH);
  a.total_tiles = p->N * a.tiles_x * a.tiles_y;
  a.yvec = vec4_ok(p->y);
  a.tma_rank = 0;
  static const int tma_on = [] { const char* e = getenv("FDG_TMA_STORE"); return e ? atoi(e) : 1; }();
  if (tma_on && a.yvec && p->store == FDG_STORE_NORMAL && !p->e.p) {
    const uint64_t dims[4] = {(uint64_t)p->Cout, (uint64_t)p->OW, (uint64_t)p->OH, (uint64_t)p->N};
    const uint64_t strides[3] = {(uint64_t)p->y.sw * 4, (uint64_t)p->y.sh * 4, (uint64_t)p->y.sn * 4};
    const uint32_t box[4] = {32, (uint32_t)K1_TW, (uint32_t)K1_OH, 1};
    if (make_tmap_f32(&a.ymap, p->y.p, 4, dims, strides, box)) a.tma_rank = 4;
  }
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(conv_k1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM) != cudaSuccess) {
      set_error("fdg_conv2d[tcgen05 k1]: cannot raise dynamic shared memory to %d bytes", K1_SMEM);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  const int num_sms = device_sm_count();
  dim3 grid((unsigned)(a.total_tiles < num_sms ? a.total_tiles : num_sms));
  const double M = (double)p->N * p->OH * p->OW;
  ProfScope prof(PF_CONV_UMMA, 2.0 * M * 9 * p->Cin * p->Cout, 4.0 * (M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  launch_k(conv_k1_kernel, dim3(grid), dim3(K1_THREADS), (size_t)(K1_SMEM), st, a);
  return check_launch("fdg_conv2d[tcgen05 k1]");
}

}  // namespace fdg

using namespace fdg;

extern "C" int64_t fdg_k1_weight_bytes(int Cin) { return (int64_t)cdiv(Cin, 64) * 3 * K1_B_TILE; }

extern "C" int fdg_pack_weight_k1(const float* w, int w_ld, int Cin, int Cout, void* out, fdg_stream_t stream) {
  FDG_REQUIRE(w && out && Cin > 0 && Cout > 0 && Cout <= 32 && w_ld >= Cout, "fdg_pack_weight_k1: bad arguments (Cout <= 32)");
  FDG_REQUIRE(aligned16(out), "fdg_pack_weight_k1: output must be 16-byte aligned");
  const int cch = cdiv(Cin, 64);
  const int64_t total = (int64_t)cch * 3 * 96 * 8;
  int64_t g = cdiv64(total, 256);
  if (g > 148 * 8) g = 148 * 8;
  pack_k1_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(w, w_ld, Cin, Cout, cch, (uint8_t*)out, total);
  return check_launch("fdg_pack_weight_k1");
}
