// fdg_conv2d, tcgen05 halo-tile path for stride-1 RxS (2..4) convolutions: 3x3 dense-block / VGG / decoder convs,
// their data gradients, and the 4x4 stride-1 layers of the Fusion-discriminator.
//
// conv_umma.cu converts every input pixel from fp32 to bf16 hi/lo once PER FILTER TAP.  Here a CTA owns a 16x8 block
// of output pixels (UMMA M = 128); for each 64-channel chunk the loaders fetch the (16+R-1) x (8+S-1) input halo
// ONCE, run the consumer prologue (BatchNorm scale/shift + ReLU/LeakyReLU, zero padding after it), split to bf16
// hi/lo and store one K-major SWIZZLE_128B tile per halo (row index = hy*HC + hx, swizzle taken from the absolute
// shared-memory address).  The R*S taps are then pure descriptor arithmetic: tap (ky,kx) reads the same tile through
// a descriptor whose start address is shifted by (ky*HC + kx) rows and whose stride between 8-row groups (SBO) is the
// halo pitch HC*128 B, so an 8-pixel output row maps onto 8 consecutive halo rows.  (tests/probes/probe_shift.cu
// verifies on hardware that row-shifted starts and SBO != 1024 are honoured with base_offset = 0.)
// Weights: the same pre-swizzled per-(N tile, tap, chunk) images as conv_umma.cu, streamed through a bulk-TMA ring.
// Persistent, warp-specialised: warps 0-7 halo loaders, 8-11 epilogue (double-buffered TMEM accumulators),
// warp 12 MMA issue, warp 13 weight-tile producer.
#include <atomic>
#include <cstdlib>

#include "aop.cuh"
#include "umma.cuh"
#include "umma_epilogue.cuh"

namespace fdg {

constexpr int UKC_H = 64;                          // K elements per chunk
constexpr int HT_W = 8, HT_H = 16;                 // output tile (8 wide so that one 8-row group = one image row)
constexpr int H_LOAD_WARPS = 8;
constexpr int H_MMA_WARP = H_LOAD_WARPS, H_W_WARP = H_LOAD_WARPS + 1, H_EPI_WARP0 = H_LOAD_WARPS + 2;
constexpr int H_THREADS = (H_LOAD_WARPS + 2 + 4) * 32;
constexpr int H_MAXROWS = (HT_H + 3) * (HT_W + 3); // 4x4 filter: 19 x 11 halo pixels
constexpr int H_A_TILE = ((H_MAXROWS * 128 + 1023) / 1024) * 1024;   // 27 KB per hi (or lo) halo tile
constexpr int H_MAX_AFF = 1024;
constexpr int H_ITEMS = (H_MAXROWS * 8 + H_LOAD_WARPS * 32 - 1) / (H_LOAD_WARPS * 32);   // 16-byte chunks per loader thread

struct HaloArgs {
  FdgConv c;
  int cchunks, tiles_x, tiles_y, n_tiles, total_tiles;
  int HR, HC;      // halo rows / columns
  int a_tile;      // bytes of one bf16 halo tile (hi or lo), multiple of 1024
  int yvec;
  int dbg;   // ablation: 1 no global loads, 2 no split/stores, 4 no MMA, 8 no epilogue
  int pair;  // weight image holds two filter taps per 64-deep chunk (Cin <= 32, pack.cuh)
  int wstages;   // weight stages per tile: taps * cchunks, or ceil(taps / 2) in pair mode
  int tma_rank;                 // 4: the epilogue stores through ymap {channel, x, y, image}; 0: coalesced stores
  int epi_wrows;                // EpiTma.wrows: 4 = every epilogue warp stores its own 4 tile rows (box {32, 8, 4, 1}); 0 = one box per group
  alignas(64) CUtensorMap ymap;
};

// BN2 (own instantiation, NT = 128, Cin <= 32, no prologue): the BatchNorm-backward epilogue (FdgConv.e_scale) with the mask-tensor rows of
// the NEXT 32-channel group (or of the next tile's first group) in flight while the current group is processed -- the dense-layer norm2
// backward inside the conv2 data gradient: dz = acc * [e_scale * e + e_shift > 0], y = e_scale * dz, stats += (sum dz, sum dz * e).
// A 32 -> 128 data gradient is bound by its epilogue (K = 288 against 128 output channels), and a 32-channel halo is half the loader
// work: warps 0-3 load (4 eight-channel chunks per halo pixel instead of 8), warps 4-7 become a SECOND set of epilogue warps with its
// own staging tile; set s takes the 32-channel groups g = s (mod 2).
// CL (one output-channel tile, even tile count): CTAs 2i, 2i + 1 form a cluster and SHARE the weight stream.  Every 128-pixel tile needs the
// whole filter (160 KB for a 3x3 32 -> 128 data gradient), streamed from L2 because it does not fit beside the halo ring: 1.3 GB of L2 -> SM
// traffic for 0.5 GB of output, the kernel runs at the L2 fabric's rate (0.169 ms with MMAs and epilogue switched off).  Rank 0 fetches the
// hi half of each stage, rank 1 the lo half, each multicast into both CTAs (same shared-memory offset, each CTA's own full barrier counts
// all bytes); a stage is refilled when BOTH CTAs' MMAs have retired (multicast commits on an empty barrier that counts two arrivals).
template <int NT, int BSTAGES, bool BN2 = false, bool CL = false>
__global__ void __launch_bounds__(H_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloArgs a) {
  constexpr int B_TILE_BYTES = NT * 128;
  const int A_STAGE = 2 * a.a_tile;
  static_assert(NT <= 128 && NT % 16 == 0, "two double-width accumulators must fit the 512 TMEM columns");
  constexpr int TMEM_COLS = 4 * NT <= 128 ? 128 : (4 * NT <= 256 ? 256 : 512);   // two accumulator buffers of [hi*hi | hi*lo + lo*hi] halves (power of two)
  constexpr int NTP = (NT + 31) / 32 * 32;  // 32-channel epilogue groups; the last one of an 80-wide tile holds 16 channels
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a_full[2], bar_a_empty[2], bar_b_full[BSTAGES], bar_b_empty[BSTAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sred[2][4][NTP];
  __shared__ __align__(1024) uint8_t ep_stage[EP_TILE_BYTES];   // epilogue staging tile (SWIZZLE_128B box layout)
  __shared__ __align__(1024) uint8_t ep_stage2[BN2 ? EP_TILE_BYTES : 16];   // BN2: staging tile of the second epilogue set
  __shared__ __align__(16) float aff_s[2][BN2 ? 4 : H_MAX_AFF];             // (BN2 runs without a prologue: the space goes to ep_stage2)
  constexpr int LW = BN2 ? 4 : H_LOAD_WARPS;          // loader warps
  constexpr int JB = BN2 ? 2 : 3;                     // log2 of the 8-channel chunks per halo pixel a pass of the loaders covers

  const FdgConv& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + 2 * A_STAGE;
  const int taps = p.R * p.S;
  const int HP = a.HR * a.HC;
  const int tiles_img = a.tiles_x * a.tiles_y;

  if (t == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_a_full[s]), LW);
      mbar_init(smem_u32(&bar_a_empty[s]), 1);
      mbar_init(smem_u32(&bar_acc_full[s]), 1);
      mbar_init(smem_u32(&bar_acc_empty[s]), BN2 ? 8 : 4);
    }
    for (int s = 0; s < BSTAGES; ++s) {
      mbar_init(smem_u32(&bar_b_full[s]), 1);
      mbar_init(smem_u32(&bar_b_empty[s]), CL ? 2 : 1);
    }
    fence_barrier_init();
  }
  if (warp == H_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  const bool aff_smem = !BN2 && p.has_affine && p.Cin <= H_MAX_AFF;
  if (aff_smem)
    for (int i = t; i < p.Cin; i += H_THREADS) { aff_s[0][i] = __ldg(p.scale + i); aff_s[1][i] = __ldg(p.shift + i); }
  for (int i = t; i < 2 * 4 * NTP; i += H_THREADS) (&sred[0][0][0])[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t crank = CL ? cluster_ctarank() : 0u;
  if (CL) cluster_sync_all();      // the peer's barriers are initialised before anything of this CTA can signal them

  // tile id -> (output-channel tile, image, tile row, tile column)
  auto decode = [&](int tile, int& ntile, int& n, int& oy0, int& ox0) {
    ntile = tile / (p.N * tiles_img);
    int r = tile - ntile * (p.N * tiles_img);
    n = r / tiles_img;
    r -= n * tiles_img;
    const int tyi = r / a.tiles_x;
    oy0 = tyi * HT_H;
    ox0 = (r - tyi * a.tiles_x) * HT_W;
  };

  const bool epi2 = BN2 && warp >= LW && warp < H_LOAD_WARPS;      // BN2: second epilogue set
  if (warp < LW) {
    // =============================================================== halo loaders
    // item i of this thread: halo row (pixel) hrow[i], 16-byte bf16 chunk j (8 channels); fixed for the whole kernel
    const int j = t & ((1 << JB) - 1);
    int hy[H_ITEMS], hx[H_ITEMS];
    bool iv[H_ITEMS];
#pragma unroll
    for (int i = 0; i < H_ITEMS; ++i) {
      const int row = (t >> JB) + i * 32;
      iv[i] = row < HP;
      hy[i] = iv[i] ? row / a.HC : 0;
      hx[i] = iv[i] ? row - hy[i] * a.HC : 0;
    }
    const uint32_t full0 = smem_u32(&bar_a_full[0]), empty0 = smem_u32(&bar_a_empty[0]);
    const uint32_t aff0 = smem_u32(&aff_s[0][0]);
    auto lds4u = [](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      return v;
    };
    int buf = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      int ntile, n, oy0, ox0;
      decode(tile, ntile, n, oy0, ox0);
      const int iy0 = oy0 - p.pad, ix0 = ox0 - p.pad;
      const float* tbase = p.x.p + n * p.x.sn + (int64_t)iy0 * p.x.sh + (int64_t)ix0 * p.x.sw + j * 8;   // dereferenced only in range
      uint32_t okmask = 0;
#pragma unroll
      for (int i = 0; i < H_ITEMS; ++i) {
        const int iy = iy0 + hy[i], ix = ix0 + hx[i];
        if (iv[i] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) okmask |= 1u << i;
      }
      for (int cc = 0; cc < a.cchunks; ++cc) {
        const int c = cc * UKC_H + j * 8;
        const bool cvalid = c < p.Cin, cvalid2 = c + 4 < p.Cin;     // Cin % 4 == 0: the second half of a chunk may be padding
        // ---- all loads of the chunk first (independent, up to 2*H_ITEMS 128-bit loads in flight per thread)
        float4 v0[H_ITEMS], v1[H_ITEMS];
#pragma unroll
        for (int i = 0; i < H_ITEMS; ++i) {
          v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          v1[i] = v0[i];
          if (((okmask >> i) & 1u) && cvalid && !(a.dbg & 1)) {
            const float* src = tbase + (int64_t)hy[i] * p.x.sh + (int64_t)hx[i] * p.x.sw + cc * UKC_H;
            v0[i] = ld4(src);
            if (cvalid2) v1[i] = ld4(src + 4);
          }
        }
        // ---- consumer prologue
        const float sl = p.slope;
        if (p.has_affine) {
          float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
          if (cvalid) {
            if (aff_smem) {
              sc0 = lds4u(aff0 + c * 4); sh0 = lds4u(aff0 + (H_MAX_AFF + c) * 4);
              if (cvalid2) { sc1 = lds4u(aff0 + c * 4 + 16); sh1 = lds4u(aff0 + (H_MAX_AFF + c) * 4 + 16); }
            } else {
              sc0 = ld4(p.scale + c); sh0 = ld4(p.shift + c);
              if (cvalid2) { sc1 = ld4(p.scale + c + 4); sh1 = ld4(p.shift + c + 4); }
            }
          }
#pragma unroll
          for (int i = 0; i < H_ITEMS; ++i) {
            v0[i].x = fmaf(v0[i].x, sc0.x, sh0.x); v0[i].y = fmaf(v0[i].y, sc0.y, sh0.y);
            v0[i].z = fmaf(v0[i].z, sc0.z, sh0.z); v0[i].w = fmaf(v0[i].w, sc0.w, sh0.w);
            v1[i].x = fmaf(v1[i].x, sc1.x, sh1.x); v1[i].y = fmaf(v1[i].y, sc1.y, sh1.y);
            v1[i].z = fmaf(v1[i].z, sc1.z, sh1.z); v1[i].w = fmaf(v1[i].w, sc1.w, sh1.w);
          }
        }
        if (sl == 0.f) {
#pragma unroll
          for (int i = 0; i < H_ITEMS; ++i) {
            v0[i].x = fmaxf(v0[i].x, 0.f); v0[i].y = fmaxf(v0[i].y, 0.f); v0[i].z = fmaxf(v0[i].z, 0.f); v0[i].w = fmaxf(v0[i].w, 0.f);
            v1[i].x = fmaxf(v1[i].x, 0.f); v1[i].y = fmaxf(v1[i].y, 0.f); v1[i].z = fmaxf(v1[i].z, 0.f); v1[i].w = fmaxf(v1[i].w, 0.f);
          }
        } else if (sl != 1.f) {
#pragma unroll
          for (int i = 0; i < H_ITEMS; ++i) {
            v0[i].x = prologue_act(v0[i].x, sl); v0[i].y = prologue_act(v0[i].y, sl); v0[i].z = prologue_act(v0[i].z, sl); v0[i].w = prologue_act(v0[i].w, sl);
            v1[i].x = prologue_act(v1[i].x, sl); v1[i].y = prologue_act(v1[i].y, sl); v1[i].z = prologue_act(v1[i].z, sl); v1[i].w = prologue_act(v1[i].w, sl);
          }
        }
        if (p.has_affine) {   // zero padding is applied AFTER the prologue
#pragma unroll
          for (int i = 0; i < H_ITEMS; ++i)
            if (!(((okmask >> i) & 1u) && cvalid)) { v0[i] = make_float4(0.f, 0.f, 0.f, 0.f); v1[i] = v0[i]; }
        }
        // ---- split and store into the halo tile of buffer `buf`
        mbar_wait(empty0 + buf * 8, ph ^ 1u);
        const uint32_t a_hi = smem_base + buf * A_STAGE, a_lo = a_hi + a.a_tile;
#pragma unroll
        for (int i = 0; i < H_ITEMS; ++i) {
          if (iv[i] && !(a.dbg & 2)) {
            const int row = (t >> JB) + i * 32;
            const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
            uint32_t h[4], l[4];
            split2(v0[i].x, v0[i].y, h[0], l[0]);
            split2(v0[i].z, v0[i].w, h[1], l[1]);
            split2(v1[i].x, v1[i].y, h[2], l[2]);
            split2(v1[i].z, v1[i].w, h[3], l[3]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + buf * 8);
        if (++buf == 2) { buf = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= H_EPI_WARP0 || epi2) {
    // =============================================================== epilogue warps (TMEM lane quarter = warp & 3, two per quarter)
    const int quarter = warp & 3;
    const int wset = epi2 ? 1 : 0;                      // BN2: epilogue set, 32-channel groups g = wset (mod 2)
    constexpr int NSETS = BN2 ? 2 : 1;
    const int et = epi2 ? t : t - H_EPI_WARP0 * 32;     // set 1 (threads 128..255) keeps its thread index
    const uint32_t stage = epi2 ? smem_u32(ep_stage2) : smem_u32(ep_stage);
    const bool evec = p.e.p && p.e.sc == 1 && aligned16_dev(p.e.p) && (p.e.sn % 4 == 0) && (p.e.sh % 4 == 0) && (p.e.sw % 4 == 0);
    // BN2: lane -> (row of four, 4-channel chunk) of the coalesced phase; evn[i] = mask-tensor values of row 4 i + brs of this warp's 32 pixels
    float4 evn[BN2 ? 8 : 1];
    const int bq = lane & 7, bc4 = bq * 4, brs = lane >> 3;
    auto bn_prefetch = [&](int tile_, int g_) {
      if constexpr (BN2) {
        int nt_, n_, oy_, ox_;
        decode(tile_, nt_, n_, oy_, ox_);
        const int c_ = nt_ * NT + g_ * 32 + bc4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int oy = oy_ + quarter * 4 + (i >> 1), ox = ox_ + 4 * (i & 1) + brs;
          evn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (oy < p.OH && ox < p.OW && c_ < p.Cout) evn[i] = ld4(p.e.p + n_ * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw + c_);
        }
      }
    };
    int it = 0;
    if (BN2 && (int)blockIdx.x < a.total_tiles) bn_prefetch(blockIdx.x, wset);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      int ntile, n, oy0, ox0;
      decode(tile, ntile, n, oy0, ox0);
      const int b = it & 1;
      while (!mbar_try_wait(smem_u32(&bar_acc_full[b]), ((uint32_t)it >> 1) & 1u)) __nanosleep(200);
      tc_fence_after();
      const int m = quarter * 32 + lane;
      const int oy = oy0 + (m >> 3), ox = ox0 + (m & 7);
      const bool mv = oy < p.OH && ox < p.OW;
      int64_t yoff = 0, eoff = 0;
      if (mv) {
        const int us = p.store == FDG_STORE_UP2 ? 2 : 1;
        yoff = n * p.y.sn + (int64_t)(us * oy) * p.y.sh + (int64_t)(us * ox) * p.y.sw;
        if (p.e.p) eoff = n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw;
      }
      const EpiTma tm{a.tma_rank ? (const void*)&a.ymap : nullptr, a.tma_rank, ox0, oy0, n, a.epi_wrows};
      const int cbase = ntile * NT;
      if constexpr (BN2) {
        const int ngroups = (p.Cout - cbase + 31) / 32 < NT / 32 ? (p.Cout - cbase + 31) / 32 : NT / 32;
        const uint32_t wrow0 = stage + (uint32_t)(quarter * 32) * 128u;
#pragma unroll 1
        for (int g = wset; g < ngroups; g += NSETS) {      // Cout % 64 == 0: every tile has groups for both sets
          const int c0 = cbase + g * 32;
          {
            float v[32], v2[32];
            const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 2 * NT + g * 32);
            tmem_ld32_nowait(tcol, v);
            tmem_ld32_nowait(tcol + NT, v2);
            tmem_ld_wait();
            const uint32_t trow = wrow0 + (uint32_t)lane * 128u;
            const int sw = lane & 7;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(trow + (uint32_t)((qq ^ sw) << 4)), "f"(v[4 * qq] + v2[4 * qq]),
                           "f"(v[4 * qq + 1] + v2[4 * qq + 1]), "f"(v[4 * qq + 2] + v2[4 * qq + 2]), "f"(v[4 * qq + 3] + v2[4 * qq + 3]) : "memory");
          }
          __syncwarp();
          const bool cv = c0 + bc4 < p.Cout;
          float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc, ps1 = bsc, ps2 = bsc;
          if (cv) { bsc = ld4(p.e_scale + c0 + bc4); bsh = ld4(p.e_shift + c0 + bc4); }
          const float al = p.alpha, als = p.alpha * p.eslope;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + brs;
            const int oy = oy0 + quarter * 4 + (i >> 1), ox = ox0 + 4 * (i & 1) + brs;
            if (oy < p.OH && ox < p.OW && cv) {
              float4 val;
              const uint32_t ta = wrow0 + (uint32_t)row * 128u + (uint32_t)((bq ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(ta) : "memory");
              const float4 ev = evn[i];
              val.x *= fmaf(ev.x, bsc.x, bsh.x) > 0.f ? al : als; val.y *= fmaf(ev.y, bsc.y, bsh.y) > 0.f ? al : als;
              val.z *= fmaf(ev.z, bsc.z, bsh.z) > 0.f ? al : als; val.w *= fmaf(ev.w, bsc.w, bsh.w) > 0.f ? al : als;
              ps1.x += val.x; ps1.y += val.y; ps1.z += val.z; ps1.w += val.w;
              ps2.x = fmaf(val.x, ev.x, ps2.x); ps2.y = fmaf(val.y, ev.y, ps2.y); ps2.z = fmaf(val.z, ev.z, ps2.z); ps2.w = fmaf(val.w, ev.w, ps2.w);
              val.x *= bsc.x; val.y *= bsc.y; val.z *= bsc.z; val.w *= bsc.w;
              float* yp = p.y.p + n * p.y.sn + (int64_t)oy * p.y.sh + (int64_t)ox * p.y.sw + c0 + bc4;
              if (p.store == FDG_STORE_ACCUM)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yp), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
              else
                *reinterpret_cast<float4*>(yp) = val;
            }
          }
          // mask rows of the next group / of the next tile's first group: in flight during the statistics fold, the next tcgen05.ld and
          // staging stores and, across tiles, the wait for the accumulator
          if (g + NSETS < ngroups) bn_prefetch(tile, g + NSETS);
          else if (tile + (int)gridDim.x < a.total_tiles) bn_prefetch(tile + gridDim.x, wset);
          if (p.stats) {
#pragma unroll
            for (int off = 8; off <= 16; off <<= 1) {
              ps1.x += __shfl_xor_sync(0xffffffffu, ps1.x, off); ps1.y += __shfl_xor_sync(0xffffffffu, ps1.y, off);
              ps1.z += __shfl_xor_sync(0xffffffffu, ps1.z, off); ps1.w += __shfl_xor_sync(0xffffffffu, ps1.w, off);
              ps2.x += __shfl_xor_sync(0xffffffffu, ps2.x, off); ps2.y += __shfl_xor_sync(0xffffffffu, ps2.y, off);
              ps2.z += __shfl_xor_sync(0xffffffffu, ps2.z, off); ps2.w += __shfl_xor_sync(0xffffffffu, ps2.w, off);
            }
            if (lane < 8 && cv) {
              float* st1 = &sred[0][quarter][g * 32 + bc4];
              float* st2 = &sred[1][quarter][g * 32 + bc4];
              st1[0] += ps1.x; st1[1] += ps1.y; st1[2] += ps1.z; st1[3] += ps1.w;
              st2[0] += ps2.x; st2[1] += ps2.y; st2[2] += ps2.z; st2[3] += ps2.w;
            }
          }
          __syncwarp();
        }
      } else {
#pragma unroll 1
      for (int g = 0; g < NTP / 32; ++g) {
        const int c0 = cbase + g * 32;
        if (c0 < p.Cout && !(a.dbg & 8)) {
          float v[32];
          {
            float v2[32];
            const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * 2 * NT + g * 32);
            tmem_ld32_nowait(tcol, v);
            tmem_ld32_nowait(tcol + NT, v2);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] += v2[u];
          }
          umma_epilogue_group<true>(p, a.yvec, evec, v, mv, yoff, eoff, c0, lane, quarter, et, stage, tm, &sred[0][quarter][g * 32],
                              &sred[1][quarter][g * 32], NT - g * 32);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
      if (p.stats) {
        const int next = tile + gridDim.x;
        if (next >= a.total_tiles || next / (p.N * tiles_img) != ntile) {
          if (BN2) asm volatile("bar.sync 1, 256;" ::: "memory"); else
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (!epi2)
          for (int cidx = et; cidx < NT; cidx += 128) {
            const int c = ntile * NT + cidx;
            if (c < p.Cout) {
              atomicAdd(p.stats + c, (double)((sred[0][0][cidx] + sred[0][1][cidx]) + (sred[0][2][cidx] + sred[0][3][cidx])));
              atomicAdd(p.stats + p.stats_ld + c, (double)((sred[1][0][cidx] + sred[1][1][cidx]) + (sred[1][2][cidx] + sred[1][3][cidx])));
            }
            sred[0][0][cidx] = 0.f; sred[0][1][cidx] = 0.f; sred[0][2][cidx] = 0.f; sred[0][3][cidx] = 0.f;
            sred[1][0][cidx] = 0.f; sred[1][1][cidx] = 0.f; sred[1][2][cidx] = 0.f; sred[1][3][cidx] = 0.f;
          }
          if (BN2) asm volatile("bar.sync 1, 256;" ::: "memory"); else
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
    if (a.tma_rank && !epi2 && (a.epi_wrows ? lane == 0 : et == 0)) bulk_wait_read0();
  } else if (warp == H_MMA_WARP) {
    // =============================================================== MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 2 * NT), idesc1 = umma_idesc_bf16(128, NT);
      const uint32_t sbo = (uint32_t)a.HC * 128u;
      const uint32_t b_hw = umma_desc_hi(1024);
      const uint32_t bfull0 = smem_u32(&bar_b_full[0]), bempty0 = smem_u32(&bar_b_empty[0]);
      int buf = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t tshift[16];                               // pair mode: A-descriptor row shift of every filter tap (rows * 128 B in 16-byte units)
#pragma unroll
      for (int tp = 0; tp < 16; ++tp) {
        const int ky = tp / p.S, kx = tp - ky * p.S;
        tshift[tp] = (uint32_t)(ky * a.HC + kx) * 8u;
      }
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(smem_u32(&bar_acc_empty[b]), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * 2 * NT);
        for (int cc = 0; cc < a.cchunks; ++cc) {
          mbar_wait(smem_u32(&bar_a_full[buf]), aph);
          tc_fence_after();
          const uint32_t a_hi = smem_base + buf * A_STAGE, a_lo = a_hi + a.a_tile;
          const uint32_t ah_lo0 = umma_desc_lo(a_hi, 16), al_lo0 = umma_desc_lo(a_lo, 16), a_hw = umma_desc_hi(sbo);
          const int cleft = p.Cin - cc * UKC_H;
          const int kslices = cleft >= UKC_H ? 4 : (cleft + 15) / 16;     // K = 16 slices of this chunk that hold channels
          if (a.pair) {
            // two filter taps per weight stage: tap 2s in bytes 0..63 of the rows, tap 2s + 1 in bytes 64..127 (K slices 2, 3 of the
            // B descriptor); the A operand is the same halo tile at each tap's own row shift, channels 0..31.  The issuing thread's
            // instruction latency bounds the MMA rate: shifts come from a table, a stage is ONE block of eight MMAs.
#pragma unroll
            for (int st = 0; st < 8; ++st) {
              if (st < a.wstages) {
                mbar_wait(bfull0 + bs * 8, bph);
                tc_fence_after();
                const uint32_t bl = umma_desc_lo(b_base + bs * (2 * B_TILE_BYTES), 16);
                if (!(a.dbg & 4)) {
                  if (kslices == 2 && 2 * st + 1 < taps) {
                    umma_pair8(d_tmem, ah_lo0 + tshift[2 * st], al_lo0 + tshift[2 * st], ah_lo0 + tshift[2 * st + 1], al_lo0 + tshift[2 * st + 1], a_hw,
                               bl, b_hw, idesc2, idesc1, st > 0 ? 1u : 0u, (uint32_t)NT);
                  } else {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                      if (2 * st + half < taps) {
                        for (int sl = 0; sl < kslices; ++sl)
                          umma_concat_slice(d_tmem, ah_lo0 + tshift[2 * st + half] + 2u * sl, al_lo0 + tshift[2 * st + half] + 2u * sl, a_hw,
                                            bl + 4u * half + 2u * sl, b_hw, idesc2, idesc1, (st > 0 || half > 0 || sl > 0) ? 1u : 0u, (uint32_t)NT);
                      }
                    }
                  }
                }
                if (CL) umma_commit_multicast(bempty0 + bs * 8, (uint16_t)3); else umma_commit(bempty0 + bs * 8);
                if (++bs == BSTAGES) { bs = 0; bph ^= 1u; }
              }
            }
            umma_commit(smem_u32(&bar_a_empty[buf]));
            if (++buf == 2) { buf = 0; aph ^= 1u; }
            continue;
          }
          int ky = 0, kx = 0;
          for (int tap = 0; tap < taps; ++tap) {
            const uint32_t shift = (uint32_t)(ky * a.HC + kx) * 8u;            // rows * 128 B, in 16-byte descriptor units
            mbar_wait(bfull0 + bs * 8, bph);
            tc_fence_after();
            const uint32_t b_hi = b_base + bs * (2 * B_TILE_BYTES);
            if (!(a.dbg & 4)) {
              if (kslices == 4) {
                umma_chunk8(d_tmem, ah_lo0 + shift, al_lo0 + shift, a_hw, umma_desc_lo(b_hi, 16), b_hw, idesc2, idesc1,
                            (cc > 0 || tap > 0) ? 1u : 0u, 2u, 2u, (uint32_t)NT);
              } else {   // channel tail of the last chunk (e.g. the 32-channel data gradients): only the K slices that hold data
                const uint32_t bl = umma_desc_lo(b_hi, 16);
                for (int sl = 0; sl < kslices; ++sl)
                  umma_concat_slice(d_tmem, ah_lo0 + shift + 2u * sl, al_lo0 + shift + 2u * sl, a_hw, bl + 2u * sl, b_hw, idesc2, idesc1,
                                    (cc > 0 || tap > 0 || sl > 0) ? 1u : 0u, (uint32_t)NT);
              }
            }
            if (CL) umma_commit_multicast(bempty0 + bs * 8, (uint16_t)3); else umma_commit(bempty0 + bs * 8);
            if (++bs == BSTAGES) { bs = 0; bph ^= 1u; }
            if (++kx == p.S) { kx = 0; ++ky; }
          }
          umma_commit(smem_u32(&bar_a_empty[buf]));
          if (++buf == 2) { buf = 0; aph ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[b]));
      }
    }
  } else if (warp == H_W_WARP) {
    // =============================================================== weight-tile producer (bulk TMA ring)
    if (lane == 0) {
      int bs = 0;
      uint32_t bph = 0;
      const int nchunks = a.wstages;
      const int inner = a.pair ? a.wstages : taps;       // stages per channel chunk (pair mode: a single chunk)
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        int ntile, n, oy0, ox0;
        decode(tile, ntile, n, oy0, ox0);
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.w_umma) + (size_t)ntile * nchunks * (2 * B_TILE_BYTES);
        for (int cc = 0; cc < a.cchunks; ++cc)
          for (int tap = 0; tap < inner; ++tap) {
            mbar_wait(smem_u32(&bar_b_empty[bs]), bph ^ 1u);
            const uint32_t bar = smem_u32(&bar_b_full[bs]);
            mbar_arrive_expect_tx(bar, 2 * B_TILE_BYTES);
            const int idx = a.pair ? tap : tap * a.cchunks + cc;
            if (CL) {
              // this CTA's half of the stage (rank 0: hi tile, rank 1: lo tile), delivered to both CTAs of the pair
              bulk_g2s_multicast(b_base + bs * (2 * B_TILE_BYTES) + crank * B_TILE_BYTES, wimg + (size_t)idx * (2 * B_TILE_BYTES) + crank * B_TILE_BYTES,
                                 B_TILE_BYTES, bar, (uint16_t)3);
            } else {
              bulk_g2s(b_base + bs * (2 * B_TILE_BYTES), wimg + (size_t)idx * (2 * B_TILE_BYTES), 2 * B_TILE_BYTES, bar);
            }
            if (++bs == BSTAGES) { bs = 0; bph ^= 1u; }
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == H_MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
  if (CL) cluster_sync_all();      // neither CTA leaves while the other may still multicast into it or signal its barriers
}

int umma_tap_pair(int taps, int Cin);
static int g_halo_on = [] { const char* e = getenv("FDG_HALO"); return e ? atoi(e) : 1; }();
int halo_enabled() { return g_halo_on; }

int conv2d_halo_supported(const FdgConv* p) {
  if (!g_halo_on || !p->w_umma) return 0;
  if (p->gather != FDG_GATHER_DIRECT || p->stride != 1) return 0;
  if (p->R < 2 || p->R > 4 || p->S < 2 || p->S > 4) return 0;
  if (p->Cin % 4 != 0 || p->Cin < 16 || p->Cout < 1) return 0;      // 8-channel loader chunks, second half optional
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

template <int NT, int BSTAGES, bool BN2 = false, bool CL = false>
static int launch_halo(const HaloArgs& a, cudaStream_t st) {
  const int smem = 2 * (2 * a.a_tile) + BSTAGES * (2 * NT * 128) + 1024;
  static std::atomic<int> attr_done[64];           // per device: largest size configured so far
  const int adev = current_device();
  if (attr_done[adev] < smem) {
    if (cudaFuncSetAttribute(conv_halo_kernel<NT, BSTAGES, BN2, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d[tcgen05 halo]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = smem;
  }
  const int num_sms = device_sm_count();
  dim3 grid((unsigned)(a.total_tiles < num_sms ? a.total_tiles : num_sms));
  const double M = (double)a.c.N * a.c.OH * a.c.OW;
  ProfScope prof(PF_CONV_UMMA, 2.0 * M * a.c.R * a.c.S * a.c.Cin * a.c.Cout,
                 4.0 * (M * a.c.Cout * (1.0 + (a.c.e.p ? 1.0 : 0.0) + (a.c.store == FDG_STORE_ACCUM ? 1.0 : 0.0)) + (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  if (CL) {
    // persistent pairs: as many CTAs as clusters can be resident at once (a cluster that waits for a free pair of SMs would run its
    // whole share of the tiles after everybody else)
    static std::atomic<int> max_clusters[64];
    if (!max_clusters[adev]) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(num_sms & ~1));
      cfg.blockDim = dim3(H_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, conv_halo_kernel<NT, BSTAGES, BN2, CL>, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); nc = num_sms / 2 - 2; }
      max_clusters[adev] = nc;
    }
    int g = 2 * max_clusters[adev];
    if (g > a.total_tiles) g = a.total_tiles;      // even (checked by the caller)
    launch_k_cluster(conv_halo_kernel<NT, BSTAGES, BN2, CL>, dim3((unsigned)g), dim3(H_THREADS), (size_t)(smem), st, 2, a);
  } else
  launch_k(conv_halo_kernel<NT, BSTAGES, BN2, CL>, dim3(grid), dim3(H_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d[tcgen05 halo]");
}

int conv2d_halo(const FdgConv* p, int nt, cudaStream_t st) {
  HaloArgs a;
  a.c = *p;
  a.cchunks = cdiv(p->Cin, UKC_H);
  a.tiles_x = cdiv(p->OW, HT_W);
  a.tiles_y = cdiv(p->OH, HT_H);
  a.n_tiles = cdiv(p->Cout, nt);
  a.total_tiles = a.n_tiles * p->N * a.tiles_x * a.tiles_y;
  a.HR = HT_H + p->R - 1;
  a.HC = HT_W + p->S - 1;
  a.a_tile = ((a.HR * a.HC * 128 + 1023) / 1024) * 1024;
  a.yvec = vec4_ok(p->y);
  a.dbg = dbg_flags();
  a.pair = umma_tap_pair(p->R * p->S, p->Cin);
  a.wstages = a.pair ? (p->R * p->S + 1) / 2 : p->R * p->S * a.cchunks;
  a.tma_rank = 0;
  a.epi_wrows = 0;
  static const int tma_on = [] { const char* e = getenv("FDG_TMA_STORE"); return e ? atoi(e) : 1; }();
  if (tma_on && a.yvec && p->store == FDG_STORE_NORMAL && !p->e.p) {
    const uint64_t dims[4] = {(uint64_t)p->Cout, (uint64_t)p->OW, (uint64_t)p->OH, (uint64_t)p->N};
    const uint64_t strides[3] = {(uint64_t)p->y.sw * 4, (uint64_t)p->y.sh * 4, (uint64_t)p->y.sn * 4};
    a.epi_wrows = epi_warp_stores() ? 32 / HT_W : 0;
    const uint32_t box[4] = {32, (uint32_t)HT_W, a.epi_wrows ? (uint32_t)a.epi_wrows : (uint32_t)HT_H, 1};
    if (make_tmap_f32(&a.ymap, p->y.p, 4, dims, strides, box)) a.tma_rank = 4;
  }
  // weight-tile ring depth: as deep as shared memory allows (the ring hides the L2 latency of the bulk copies)
  static const int cl_on = [] { const char* e = getenv("FDG_HALO_CLUSTER"); return e ? atoi(e) : 1; }();
  // pairs of CTAs share the weight stream (CL) when tiles 2i, 2i + 1 always lie in the same output-channel tile.  Measured on the plain
  // instantiations (3x3 32 -> 128, 64 -> 64, 72 -> 144, 128 -> 128 ...): no difference, so only the epilogue-bound BN2 kernel takes it
  // (0.450 -> 0.388 ms at 256^2: its mask-tensor reads compete with the weight stream for the L2 -> SM path)
  const bool cl = cl_on && (p->N * a.tiles_x * a.tiles_y) % 2 == 0 && a.total_tiles >= 4 && !a.dbg;
  switch (nt) {
    case 32: return launch_halo<32, 10>(a, st);
    case 64: return launch_halo<64, 5>(a, st);
    case 80: return launch_halo<80, 4>(a, st);     // 144 = 2 x 80, 72, 160 (Fusion-D layer 3, data gradients of layers 3 / 4, conv_refine4)
    case 96: return launch_halo<96, 3>(a, st);     // 288 = 3 x 96 (Fusion-D layer 4)
    default: {
      static const int bn2_on = [] { const char* e = getenv("FDG_HALO_BN2"); return e ? atoi(e) : 1; }();
      // BatchNorm-backward epilogue on 128-bit views (validated by fdg_conv2d): the instantiation that prefetches its mask rows
      if (bn2_on && p->e_scale && a.a_tile <= 23 * 1024 && a.pair && a.cchunks == 1 && !p->has_affine && p->Cout % 64 == 0)
        return cl ? launch_halo<128, 3, true, true>(a, st) : launch_halo<128, 3, true>(a, st);
      return a.a_tile <= 23 * 1024 ? launch_halo<128, 3>(a, st) : launch_halo<128, 2>(a, st);
    }
  }
}

}  // namespace fdg

// Runtime options (tests / benchmarks): "halo" = 0 routes stride-1 RxS convolutions through the generic per-tap
// tcgen05 kernel instead of the halo-tile kernel.
namespace fdg { void set_wgrad_halo(int on); }

extern "C" int fdg_set_option(const char* name, int value) {
  if (name && name[0] == 'h' && name[1] == 'a' && name[2] == 'l' && name[3] == 'o' && name[4] == 0) {
    fdg::g_halo_on = value;
    fdg::set_wgrad_halo(value);
    return FDG_OK;
  }
  if (name && name[0] == 'd' && name[1] == 'b' && name[2] == 'g' && name[3] == 0) {
    fdg::set_dbg_flags(value);
    return FDG_OK;
  }
  fdg::set_error("fdg_set_option: unknown option");
  return FDG_EINVAL;
}
