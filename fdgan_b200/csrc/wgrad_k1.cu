// fdg_conv2d_wgrad, tcgen05 path for the growth convolutions (3x3 / stride 1 / pad 1, Cout <= 32): the 42 dense-layer conv2
// weight gradients of the generator (128 -> 32), 8.4 ms of an 83 ms step in round 1.
//
//   dW[ky][kx][c][co] = sum over pixels (y, x) of  A(y, x)[c] * G(y - ky + 1, x - kx + 1)[co]
//
// wgrad_halo.cu keeps G fixed and shifts the 128-channel activation operand per filter tap: 27 MMAs of N = 32 per 16-pixel
// slice, each re-reading a [128 x 16] A slice from shared memory -- with N = 32 the tensor core waits on those reads (ncu r01h:
// tensor-core operand reads 57 % of the L1 data pipe, tensor pipe 23 %).  Here A stays UNSHIFTED and the three filter
// columns become three shifted views of the 32-channel gradient tile side by side along N:
//   * the gradient halo tile is stored [halo pixel][32 channels] = 64-byte rows, MN-major SWIZZLE_64B, swizzle from absolute
//     shared-memory address bits; an N block is 32 elements, and LBO = 64 B makes N block v the SAME tile shifted by v pixels
//     (overlapping N blocks: tests/probes/probe_mn64.cu verifies the hardware reads exactly that, for any 64-byte start shift);
//   * a K = 16 slice is one 16-pixel tile row; the filter row ky is a start-address shift of one halo row;
//   * per slice: 3 filter rows x 3 split passes (hi*hi, hi*lo, lo*hi) = 9 MMAs of N = 96 into three [128 x 96] fp32
//     accumulators in TMEM (288 columns) -- a third of the A reads, no tensor work wasted.
// The activation tile needs no halo any more (wgrad_halo converts a 10x10 halo per 8x8 block: 1.56x the loader work).
#include <atomic>
#include <cstdlib>

#include "aop.cuh"
#include "umma.cuh"

namespace fdg {

constexpr int WK_TW = 16, WK_TH = 8;                       // pixel tile: one K = 16 slice per tile row
constexpr int WK_LOAD_WARPS = 8;
constexpr int WK_THREADS = (WK_LOAD_WARPS + 1) * 32;
constexpr int WK_ABLK = WK_TW * WK_TH * 128;               // one 64-channel block of the activation tile (hi or lo): 16 KB
constexpr int WK_HC = WK_TW + 2, WK_HR = WK_TH + 2;        // gradient halo: 18 x 10 pixels
constexpr int WK_GT = ((WK_HC * WK_HR * 64 + 1023) / 1024) * 1024;   // gradient halo tile (hi or lo): 12 KB
constexpr int WK_STAGE = 4 * WK_ABLK + 2 * WK_GT;          // A hi/lo x 2 channel blocks + G hi/lo = 88 KB
constexpr int WK_GITEMS = (WK_HC * WK_HR * 4 + WK_LOAD_WARPS * 32 - 1) / (WK_LOAD_WARPS * 32);   // 16-byte gradient chunks per thread

struct WKArgs {
  FdgWgrad c;
  int cblocks, co_tiles, tiles_x, tiles_y, total_ptiles, ptiles_per_split, splits;
  int gvec;
};

// descriptor high word: SBO | version 1 | layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint32_t wk_desc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }

__global__ void __launch_bounds__(WK_THREADS, 1) wgrad_k1_kernel(const __grid_constant__ WKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const FdgWgrad& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  int bid = blockIdx.x;
  const int split = bid % a.splits; bid /= a.splits;
  const int cot = bid % a.co_tiles;                        // 32-channel tile of the output channels (Cout <= 96: Fusion-D layer 2)
  const int cb = bid / a.co_tiles;
  const int pt0 = split * a.ptiles_per_split;
  const int pt1 = pt0 + a.ptiles_per_split < a.total_ptiles ? pt0 + a.ptiles_per_split : a.total_ptiles;
  const int ntiles = pt1 > pt0 ? pt1 - pt0 : 0;
  const int tiles_img = a.tiles_x * a.tiles_y;

  if (t == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_full[s]), WK_LOAD_WARPS); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == WK_LOAD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < WK_LOAD_WARPS && ntiles > 0) {
    // =============================================================== loaders
    // A: thread = 8-channel chunk cj of tile column xx, all 8 tile rows (row r of the tile = K slice r)
    const int cj = t & 15, xx = t >> 4;
    const int ca = cb * 128 + cj * 8;
    const bool cav = ca < p.Cin, cav2 = ca + 4 < p.Cin;       // Cin % 4 == 0: the second half of a chunk may be padding
    const uint32_t a_off = (uint32_t)(cj >> 3) * WK_ABLK + (uint32_t)xx * 128u + (uint32_t)(((cj & 7) ^ (xx & 7)) << 4);   // + r * 2048
    float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
    if (p.has_affine && cav) { sc0 = ld4(p.scale + ca); sh0 = ld4(p.shift + ca); }
    if (p.has_affine && cav2) { sc1 = ld4(p.scale + ca + 4); sh1 = ld4(p.shift + ca + 4); }
    // G: items idx = t + 256 i -> halo pixel idx >> 2, 8-channel chunk idx & 3
    int ghy[WK_GITEMS], ghx[WK_GITEMS];
    uint32_t g_off[WK_GITEMS];
    bool giv[WK_GITEMS];
#pragma unroll
    for (int i = 0; i < WK_GITEMS; ++i) {
      const int idx = t + i * (WK_LOAD_WARPS * 32);
      const int hp = idx >> 2;
      giv[i] = hp < WK_HC * WK_HR;
      ghy[i] = giv[i] ? hp / WK_HC : 0;
      ghx[i] = giv[i] ? hp - ghy[i] * WK_HC : 0;
      g_off[i] = (uint32_t)hp * 64u + (uint32_t)(idx & 3) * 16u;     // linear; the swizzle needs the absolute address (per stage)
    }
    const int cg = cot * 32 + (t & 3) * 8;
    const float sl = p.slope;
    int buf = 0;
    uint32_t ph = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      const int n = pt / tiles_img;
      const int rr = pt - n * tiles_img;
      const int tyi = rr / a.tiles_x;
      const int oy0 = tyi * WK_TH, ox0 = (rr - tyi * a.tiles_x) * WK_TW;
      // ---- loads first (all independent)
      float4 v0[WK_TH], v1[WK_TH];
      uint32_t ok = 0;
      {
        const int ix = ox0 + xx;
        const float* src = p.x.p + n * p.x.sn + (int64_t)oy0 * p.x.sh + (int64_t)ix * p.x.sw + ca;
#pragma unroll
        for (int r = 0; r < WK_TH; ++r) {
          v0[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          v1[r] = v0[r];
          if (cav && ix < p.W && oy0 + r < p.H) {
            ok |= 1u << r;
            v0[r] = ld4(src + (int64_t)r * p.x.sh);
            if (cav2) v1[r] = ld4(src + (int64_t)r * p.x.sh + 4);
          }
        }
      }
      float4 g0[WK_GITEMS], g1[WK_GITEMS];
#pragma unroll
      for (int i = 0; i < WK_GITEMS; ++i) {
        g0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        g1[i] = g0[i];
        const int gy = oy0 - 1 + ghy[i], gx = ox0 - 1 + ghx[i];
        if (giv[i] && gy >= 0 && gy < p.OH && gx >= 0 && gx < p.OW && cg < p.Cout) {
          const float* gp = p.g.p + n * p.g.sn + (int64_t)gy * p.g.sh + (int64_t)gx * p.g.sw;
          if (a.gvec) { g0[i] = ld4(gp + cg); g1[i] = ld4(gp + cg + 4); }
          else {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = cg + e < p.Cout ? __ldg(gp + (int64_t)(cg + e) * p.g.sc) : 0.f;
            g0[i] = make_float4(f[0], f[1], f[2], f[3]);
            g1[i] = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
      // ---- consumer prologue (zero padding after it)
#pragma unroll
      for (int r = 0; r < WK_TH; ++r) {
        if ((ok >> r) & 1u) {
          v0[r].x = prologue_act(fmaf(v0[r].x, sc0.x, sh0.x), sl); v0[r].y = prologue_act(fmaf(v0[r].y, sc0.y, sh0.y), sl);
          v0[r].z = prologue_act(fmaf(v0[r].z, sc0.z, sh0.z), sl); v0[r].w = prologue_act(fmaf(v0[r].w, sc0.w, sh0.w), sl);
          if (cav2) {
            v1[r].x = prologue_act(fmaf(v1[r].x, sc1.x, sh1.x), sl); v1[r].y = prologue_act(fmaf(v1[r].y, sc1.y, sh1.y), sl);
            v1[r].z = prologue_act(fmaf(v1[r].z, sc1.z, sh1.z), sl); v1[r].w = prologue_act(fmaf(v1[r].w, sc1.w, sh1.w), sl);
          }
        }
      }
      // ---- split + store
      mbar_wait(smem_u32(&bar_empty[buf]), ph ^ 1u);
      const uint32_t st = smem_base + buf * WK_STAGE;
#pragma unroll
      for (int r = 0; r < WK_TH; ++r) {
        uint32_t h[4], l[4];
        split2(v0[r].x, v0[r].y, h[0], l[0]); split2(v0[r].z, v0[r].w, h[1], l[1]);
        split2(v1[r].x, v1[r].y, h[2], l[2]); split2(v1[r].z, v1[r].w, h[3], l[3]);
        const uint32_t dst = st + a_off + (uint32_t)r * (WK_TW * 128u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * WK_ABLK), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
      }
#pragma unroll
      for (int i = 0; i < WK_GITEMS; ++i) {
        if (giv[i]) {
          uint32_t h[4], l[4];
          split2(g0[i].x, g0[i].y, h[0], l[0]); split2(g0[i].z, g0[i].w, h[1], l[1]);
          split2(g1[i].x, g1[i].y, h[2], l[2]); split2(g1[i].z, g1[i].w, h[3], l[3]);
          const uint32_t lin = st + 4 * WK_ABLK + g_off[i];
          const uint32_t dst = lin ^ (((lin >> 7) & 3u) << 4);          // Swizzle<2,4,3> on the absolute address (WK_GT keeps bits 7, 8)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + WK_GT), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[buf]));
      if (++buf == 2) { buf = 0; ph ^= 1u; }
    }
  } else if (warp == WK_LOAD_WARPS && lane == 0 && ntiles > 0) {
    // =============================================================== MMA issue: per tile 8 slices x 3 filter rows x 3 passes
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 96);
    const uint32_t a_hw = wk_desc_hi(1024, 2);                // A: SWIZZLE_128B, 8 pixel rows of 128 B per group
    const uint32_t g_hw = wk_desc_hi(512, 4);                 // G: SWIZZLE_64B, 8 pixel rows of 64 B per group
    int buf = 0;
    uint32_t ph = 0;
    for (int it = 0; it < ntiles; ++it) {
      mbar_wait(smem_u32(&bar_full[buf]), ph);
      tc_fence_after();
      const uint32_t st = smem_base + buf * WK_STAGE;
      const uint32_t ah0 = umma_desc_lo(st, WK_ABLK), al0 = umma_desc_lo(st + 2 * WK_ABLK, WK_ABLK);
      const uint32_t gh0 = umma_desc_lo(st + 4 * WK_ABLK, 64), gl0 = umma_desc_lo(st + 4 * WK_ABLK + WK_GT, 64);   // LBO 64 B: N block v = v pixels further
#pragma unroll 1
      for (int r = 0; r < WK_TH; ++r) {
        const uint32_t ah = ah0 + (uint32_t)r * (WK_TW * 128u / 16u), al = al0 + (uint32_t)r * (WK_TW * 128u / 16u);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          // halo row of the views for tile row r and filter row ky: (r + 2 - ky); view v <-> filter column kx = 2 - v
          const uint32_t gsh = (uint32_t)((r + 2 - ky) * WK_HC) * (64u / 16u);
          const uint32_t d = tmem_base + (uint32_t)(ky * 96);
          const uint32_t acc = (it > 0 || r > 0) ? 1u : 0u;
          umma_single(d, ah, a_hw, gh0 + gsh, g_hw, idesc, acc);
          umma_single(d, ah, a_hw, gl0 + gsh, g_hw, idesc, 1u);
          umma_single(d, al, a_hw, gh0 + gsh, g_hw, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar_empty[buf]));
      if (++buf == 2) { buf = 0; ph ^= 1u; }
    }
    umma_commit(smem_u32(&bar_acc));
  }

  // =============================================================== epilogue: 3 accumulators [128 x 96] -> dW (OIHW), reduced at L2
  // For one output channel the CTA's 128 input channels x 9 taps are 1152 CONSECUTIVE floats of dW.  A lane holds one input channel,
  // so reducing straight from registers is 288 scalar reductions per lane at a 36-byte lane stride (one 32-byte sector per lane and
  // instruction: ~37 k sector operations per CTA, ~19 us -- more than the main loop at small batches).  The accumulators are
  // therefore transposed through the (now idle) operand ring into dW's own order and leave as 128-bit reductions, 512 contiguous
  // bytes per warp instruction (8x fewer sector operations).
  if (warp < WK_LOAD_WARPS && ntiles > 0) {
    mbar_wait(smem_u32(&bar_acc), 0);                        // every MMA has retired: the ring is free
    tc_fence_after();
    // The ring is about to be overwritten by OTHER threads than those that filled it.  The mbarrier chain (loader stores -> arrive ->
    // tcgen05.mma -> commit -> this wait) already orders that; the CTA barrier among the eight loader warps makes the hand-over explicit
    // (and visible to compute-sanitizer's racecheck, which does not model mbarrier / tensor-core completion).
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int quarter = warp & 3;                            // TMEM lanes this warp may read
    const int cil = quarter * 32 + lane;                     // input channel within the block
    const uint32_t stg = smem_base;                          // [32 co][128 ci][9 taps] floats = 147456 B <= 2 * WK_STAGE
    const bool vec = (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
#pragma unroll 1
    for (int q = (warp >> 2) * 5; q < ((warp >> 2) ? 9 : 5); ++q) {      // warps 0-3: q = 0..4, warps 4-7: q = 5..8 (q = ky * 3 + view)
      const int ky = q / 3, kx = 2 - (q - ky * 3);
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * 32), v);
      if (vec) {
        const uint32_t dst = stg + (uint32_t)(cil * 9 + ky * 3 + kx) * 4u;      // lane stride 9 words: conflict-free
#pragma unroll
        for (int u = 0; u < 32; ++u) asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + (uint32_t)u * (1152u * 4u)), "f"(v[u]) : "memory");
      } else {                                               // unaligned dW (a view at an odd offset): scalar reductions
        const int ci = cb * 128 + cil;
        if (ci < p.Cin) {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (cot * 32 + u < p.Cout) atomicAdd(p.dw + ((int64_t)(cot * 32 + u) * p.Cin + ci) * 9 + ky * 3 + kx, v[u]);
        }
      }
    }
    tc_fence_before();
    if (vec) {
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight epilogue warps
      const int valid4 = (p.Cin - cb * 128 < 128 ? p.Cin - cb * 128 : 128) * 9 / 4;      // float4s per output channel that hold data (Cin % 4 == 0)
      float* base = p.dw + ((int64_t)cot * 32 * p.Cin + (int64_t)cb * 128) * 9;
      const int nco = p.Cout - cot * 32 < 32 ? p.Cout - cot * 32 : 32;
      for (int i = t; i < nco * 288; i += WK_LOAD_WARPS * 32) {
        const int co = i / 288, r = i - co * 288;
        if (r < valid4) {
          float4 val;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(stg + (uint32_t)(co * 1152 + r * 4) * 4u) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + (int64_t)co * p.Cin * 9 + r * 4), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
        }
      }
    }
  }
  __syncthreads();
  if (warp == WK_LOAD_WARPS) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int g_wgrad_k1_on = [] { const char* e = getenv("FDG_WGRAD_K1"); return e ? atoi(e) : 1; }();

int wgrad_k1_supported(const FdgWgrad* p) {
  if (!g_wgrad_k1_on) return 0;
  if (p->gather != FDG_GATHER_DIRECT || p->stride != 1 || p->pad != 1 || p->R != 3 || p->S != 3 || p->transposed) return 0;
  // a 32-channel output tile re-reads the activation tile: beyond three tiles the wide kernels of wgrad_umma.cu win (when they apply)
  if (p->Cin % 4 != 0 || p->Cin < 16 || p->Cout < 1 || p->Cout > (p->Cin % 8 != 0 ? 96 : 32)) return 0;
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

int wgrad_k1(const FdgWgrad* p, cudaStream_t st) {
  WKArgs a;
  a.c = *p;
  a.cblocks = cdiv(p->Cin, 128);
  a.co_tiles = cdiv(p->Cout, 32);
  a.tiles_x = cdiv(p->OW, WK_TW);
  a.tiles_y = cdiv(p->OH, WK_TH);
  a.total_ptiles = p->N * a.tiles_x * a.tiles_y;
  a.gvec = vec4_ok(p->g) && (p->Cout % 8 == 0);
  const int num_sms = device_sm_count();
  const int groups = a.cblocks * a.co_tiles;
  int splits = groups >= num_sms ? 1 : num_sms / groups;
  {
    // every split ends with 128 x 288 atomics into the same addresses: with few pixel tiles (small batches / deep layers) the epilogues
    // cost more than the tiles.  T(s) = tiles/s * t_tile + s * t_epi is minimal at s = sqrt(tiles * t_tile / t_epi) ~ sqrt(10 tiles)
    // (measured: ~1.7 us per tile, ~0.17 us of atomics contention per split); batch 16 keeps one split per SM.
    static const int kfac = [] { const char* e = getenv("FDG_WGRAD_SPLIT_FAC"); return e ? atoi(e) : 10; }();
    int cap = 1;
    while ((int64_t)cap * cap < (int64_t)kfac * a.total_ptiles) ++cap;
    if (kfac > 0 && splits > cap) splits = cap;
  }
  if (splits > a.total_ptiles) splits = a.total_ptiles;
  if (splits < 1) splits = 1;
  a.ptiles_per_split = cdiv(a.total_ptiles, splits);
  a.splits = cdiv(a.total_ptiles, a.ptiles_per_split);
  constexpr int smem = 2 * WK_STAGE + 1024;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(wgrad_k1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d_wgrad[tcgen05 k1]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  const double M = (double)p->N * p->OH * p->OW;
  ProfScope prof(PF_WGRAD, 2.0 * M * 9.0 * p->Cin * p->Cout, 4.0 * (M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  launch_k(wgrad_k1_kernel, dim3((unsigned)(groups * a.splits)), dim3(WK_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d_wgrad[tcgen05 k1]");
}

}  // namespace fdg
