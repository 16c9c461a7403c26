// fdg_conv2d_wgrad, tcgen05 halo path for stride-1 RxS (2..4) convolutions with few output channels (Cout <= 64):
// the 3x3 128->32 growth convolutions of the dense blocks.
//
// wgrad_umma.cu gives every filter tap its own CTA, so the input is re-read from L2 and re-converted R*S times
// (the 128->32 3x3 layers are L2-bandwidth bound there).  Here a CTA owns a 128-channel block, 32 output channels and
// ALL R*S taps: it keeps R*S accumulators [128 x 32] fp32 in TMEM (R*S*32 <= 512 columns).  Per 8x8 block of
// output pixels the loaders fetch the (8+R-1) x (8+S-1) input halo once, apply the consumer prologue, split to bf16
// hi/lo and store it MN-major (row = halo pixel, 64 channels = 128 B, SWIZZLE_128B from absolute address bits); the
// gradient block is stored the same way.  Tap (ky,kx) is a descriptor shift of (ky*HC + kx) rows; a K=16 slice is two
// image rows of 8 pixels (stride between 8-row groups = halo pitch).  Double-buffered over pixel blocks.
#include <atomic>
#include <cstdlib>

#include "aop.cuh"
#include "umma.cuh"

namespace fdg {

constexpr int WH_T = 8;                        // 8x8 output pixels per block (K' = 64)
constexpr int WH_LOAD_WARPS = 8;
constexpr int WH_THREADS = (WH_LOAD_WARPS + 1) * 32;
constexpr int WH_MAXROWS = (WH_T + 3) * (WH_T + 3);          // 11 x 11 halo pixels for a 4x4 filter
constexpr int WH_ABLK = ((WH_MAXROWS * 128 + 1023) / 1024) * 1024;   // one 64-channel block of the halo tile (hi or lo)
constexpr int WH_GT = WH_T * WH_T * 128;                      // gradient tile: 64 pixels x 128 B (32 channels used)
constexpr int WH_STAGE = 4 * WH_ABLK + 2 * WH_GT;             // A hi/lo x 2 blocks + G hi/lo
constexpr int WH_AITEMS = (WH_MAXROWS * 16 + WH_LOAD_WARPS * 32 - 1) / (WH_LOAD_WARPS * 32);

struct WHArgs {
  FdgWgrad c;
  int cblocks, co_tiles, tiles_x, tiles_y, total_ptiles, ptiles_per_split, splits;
  int HR, HC;
  int gvec;
  int dbg;
};

__global__ void __launch_bounds__(WH_THREADS, 1) wgrad_halo_kernel(const __grid_constant__ WHArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc;
  __shared__ uint32_t tmem_base_s;
  const FdgWgrad& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int taps = p.R * p.S;
  const int HP = a.HR * a.HC;
  // block -> (channel block, output-channel tile, pixel-block range)
  int bid = blockIdx.x;
  const int split = bid % a.splits; bid /= a.splits;
  const int cot = bid % a.co_tiles;
  const int cb = bid / a.co_tiles;
  const int pt0 = split * a.ptiles_per_split;
  const int pt1 = pt0 + a.ptiles_per_split < a.total_ptiles ? pt0 + a.ptiles_per_split : a.total_ptiles;
  const int ntiles = pt1 > pt0 ? pt1 - pt0 : 0;
  const int tiles_img = a.tiles_x * a.tiles_y;

  if (t == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_full[s]), WH_LOAD_WARPS); mbar_init(smem_u32(&bar_empty[s]), 1); }
    mbar_init(smem_u32(&bar_acc), 1);
    fence_barrier_init();
  }
  if (warp == WH_LOAD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < WH_LOAD_WARPS && ntiles > 0) {
    // =============================================================== loaders
    // A item i: halo pixel row (t >> 4) + 16 i, 8-channel chunk cj = t & 15 of the 128-channel block
    const int cj = t & 15;
    const int ca = cb * 128 + cj * 8;                 // first channel of this thread's chunk
    const bool cav = ca < p.Cin, cav2 = ca + 4 < p.Cin;     // Cin % 4 == 0: the second half of a chunk may be padding
    const int ablk = cj >> 3, acj = cj & 7;
    float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
    if (p.has_affine && cav) { sc0 = ld4(p.scale + ca); sh0 = ld4(p.shift + ca); }
    if (p.has_affine && cav2) { sc1 = ld4(p.scale + ca + 4); sh1 = ld4(p.shift + ca + 4); }
    int hy[WH_AITEMS], hx[WH_AITEMS];
    bool iv[WH_AITEMS];
#pragma unroll
    for (int i = 0; i < WH_AITEMS; ++i) {
      const int row = (t >> 4) + i * (WH_LOAD_WARPS * 2);
      iv[i] = row < HP;
      hy[i] = iv[i] ? row / a.HC : 0;
      hx[i] = iv[i] ? row - hy[i] * a.HC : 0;
    }
    // G item: pixel (t >> 2) of the 8x8 block, 8-channel chunk gj = t & 3 of the 32-channel tile
    const int gpix = t >> 2, gj = t & 3;
    const int cg = cot * 32 + gj * 8;
    const float sl = p.slope;
    int buf = 0;
    uint32_t ph = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      const int n = pt / tiles_img;
      const int r = pt - n * tiles_img;
      const int tyi = r / a.tiles_x;
      const int oy0 = tyi * WH_T, ox0 = (r - tyi * a.tiles_x) * WH_T;
      const int iy0 = oy0 - p.pad, ix0 = ox0 - p.pad;
      // ---- loads first (all independent)
      float4 v0[WH_AITEMS], v1[WH_AITEMS];
      uint32_t ok = 0;
      const float* tbase = p.x.p + n * p.x.sn + (int64_t)iy0 * p.x.sh + (int64_t)ix0 * p.x.sw + ca;
#pragma unroll
      for (int i = 0; i < WH_AITEMS; ++i) {
        v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[i] = v0[i];
        const int iy = iy0 + hy[i], ix = ix0 + hx[i];
        if (iv[i] && cav && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && !(a.dbg & 1)) {
          ok |= 1u << i;
          const float* src = tbase + (int64_t)hy[i] * p.x.sh + (int64_t)hx[i] * p.x.sw;
          v0[i] = ld4(src);
          if (cav2) v1[i] = ld4(src + 4);
        }
      }
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      {
        const int oy = oy0 + (gpix >> 3), ox = ox0 + (gpix & 7);
        if (oy < p.OH && ox < p.OW && cg < p.Cout && !(a.dbg & 1)) {
          const float* gp = p.g.p + n * p.g.sn + (int64_t)oy * p.g.sh + (int64_t)ox * p.g.sw;
          if (a.gvec) { g0 = ld4(gp + cg); g1 = ld4(gp + cg + 4); }
          else {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = cg + e < p.Cout ? __ldg(gp + (int64_t)(cg + e) * p.g.sc) : 0.f;
            g0 = make_float4(f[0], f[1], f[2], f[3]);
            g1 = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
      // ---- consumer prologue (zero padding after it)
#pragma unroll
      for (int i = 0; i < WH_AITEMS; ++i) {
        if ((ok >> i) & 1u) {
          v0[i].x = prologue_act(fmaf(v0[i].x, sc0.x, sh0.x), sl); v0[i].y = prologue_act(fmaf(v0[i].y, sc0.y, sh0.y), sl);
          v0[i].z = prologue_act(fmaf(v0[i].z, sc0.z, sh0.z), sl); v0[i].w = prologue_act(fmaf(v0[i].w, sc0.w, sh0.w), sl);
          v1[i].x = prologue_act(fmaf(v1[i].x, sc1.x, sh1.x), sl); v1[i].y = prologue_act(fmaf(v1[i].y, sc1.y, sh1.y), sl);
          v1[i].z = prologue_act(fmaf(v1[i].z, sc1.z, sh1.z), sl); v1[i].w = prologue_act(fmaf(v1[i].w, sc1.w, sh1.w), sl);
        }
      }
      // ---- split + store
      mbar_wait(smem_u32(&bar_empty[buf]), ph ^ 1u);
      const uint32_t st = smem_base + buf * WH_STAGE;
#pragma unroll
      for (int i = 0; i < WH_AITEMS; ++i) {
        if (iv[i] && !(a.dbg & 2)) {
          const int row = (t >> 4) + i * (WH_LOAD_WARPS * 2);
          const uint32_t off = (uint32_t)ablk * WH_ABLK + (uint32_t)row * 128u + (uint32_t)((acj ^ (row & 7)) << 4);
          uint32_t h[4], l[4];
          split2(v0[i].x, v0[i].y, h[0], l[0]); split2(v0[i].z, v0[i].w, h[1], l[1]);
          split2(v1[i].x, v1[i].y, h[2], l[2]); split2(v1[i].z, v1[i].w, h[3], l[3]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + 2 * WH_ABLK + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
      }
      {
        const uint32_t off = (uint32_t)gpix * 128u + (uint32_t)((gj ^ (gpix & 7)) << 4);
        uint32_t h[4], l[4];
        split2(g0.x, g0.y, h[0], l[0]); split2(g0.z, g0.w, h[1], l[1]);
        split2(g1.x, g1.y, h[2], l[2]); split2(g1.z, g1.w, h[3], l[3]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + 4 * WH_ABLK + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + 4 * WH_ABLK + WH_GT + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[buf]));
      if (++buf == 2) { buf = 0; ph ^= 1u; }
    }
  } else if (warp == WH_LOAD_WARPS && lane == 0 && ntiles > 0) {
    // =============================================================== MMA issue: per pixel block, 12 MMAs per tap
    constexpr uint32_t idesc = umma_idesc_bf16_mn(128, 32);
    const uint32_t a_hw = umma_desc_hi((uint32_t)a.HC * 128u);   // stride between 8-pixel image rows
    const uint32_t g_hw = umma_desc_hi(1024);
    const uint32_t kstep_a = (uint32_t)(2 * a.HC) * 8u;          // a K=16 slice = two image rows of the 8x8 block
    int buf = 0;
    uint32_t ph = 0;
    for (int it = 0; it < ntiles; ++it) {
      mbar_wait(smem_u32(&bar_full[buf]), ph);
      tc_fence_after();
      const uint32_t st = smem_base + buf * WH_STAGE;
      const uint32_t ah0 = umma_desc_lo(st, WH_ABLK), al0 = umma_desc_lo(st + 2 * WH_ABLK, WH_ABLK);
      const uint32_t gh0 = umma_desc_lo(st + 4 * WH_ABLK, 1024), gl0 = umma_desc_lo(st + 4 * WH_ABLK + WH_GT, 1024);
      int ky = 0, kx = 0;
      for (int tap = 0; tap < taps; ++tap) {
        const uint32_t shift = (uint32_t)(ky * a.HC + kx) * 8u;
        if (!(a.dbg & 4)) umma_chunk12_ab(tmem_base + (uint32_t)(tap * 32), ah0 + shift, al0 + shift, a_hw, gh0, gl0, g_hw, idesc, it > 0 ? 1u : 0u,
                        kstep_a, 128u);
        if (++kx == p.S) { kx = 0; ++ky; }
      }
      umma_commit(smem_u32(&bar_empty[buf]));
      if (++buf == 2) { buf = 0; ph ^= 1u; }
    }
    umma_commit(smem_u32(&bar_acc));
  }

  // =============================================================== epilogue: R*S accumulators -> atomicAdd into dW (OIHW)
  if (warp < 4 && ntiles > 0) {
    mbar_wait(smem_u32(&bar_acc), 0);
    tc_fence_after();
    const int ci = cb * 128 + warp * 32 + lane;
#pragma unroll 1
    for (int tap = 0; tap < taps; ++tap) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * 32), v);
      if (ci < p.Cin) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int co = cot * 32 + u;
          if (co < p.Cout) atomicAdd(p.dw + ((int64_t)co * p.Cin + ci) * taps + tap, v[u]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == WH_LOAD_WARPS) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int g_wgrad_halo_on = [] { const char* e = getenv("FDG_WGRAD_HALO"); return e ? atoi(e) : 1; }();
void set_wgrad_halo(int on) { g_wgrad_halo_on = on; }

int wgrad_halo_supported(const FdgWgrad* p) {
  if (!g_wgrad_halo_on) return 0;
  if (p->gather != FDG_GATHER_DIRECT || p->stride != 1 || p->transposed) return 0;
  if (p->R < 2 || p->R > 4 || p->S < 2 || p->S > 4) return 0;
  if (p->Cin % 4 != 0 || p->Cin < 16 || p->Cout < 1) return 0;
  // every 32-channel output tile re-reads the input halo: wide layers go to wgrad_umma when it can take them
  if (p->Cout > 64 && !(p->Cin % 8 != 0 && p->Cout <= 96)) return 0;
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

int wgrad_halo(const FdgWgrad* p, cudaStream_t st) {
  WHArgs a;
  a.c = *p;
  a.cblocks = cdiv(p->Cin, 128);
  a.co_tiles = cdiv(p->Cout, 32);
  a.tiles_x = cdiv(p->OW, WH_T);
  a.tiles_y = cdiv(p->OH, WH_T);
  a.total_ptiles = p->N * a.tiles_x * a.tiles_y;
  a.HR = WH_T + p->R - 1;
  a.HC = WH_T + p->S - 1;
  a.gvec = vec4_ok(p->g) && (p->Cout % 8 == 0);
  a.dbg = dbg_flags();
  const int num_sms = device_sm_count();
  const int groups = a.cblocks * a.co_tiles;
  int splits = groups >= num_sms ? 1 : num_sms / groups;
  if (splits > a.total_ptiles) splits = a.total_ptiles;
  if (splits < 1) splits = 1;
  a.ptiles_per_split = cdiv(a.total_ptiles, splits);
  a.splits = cdiv(a.total_ptiles, a.ptiles_per_split);
  constexpr int smem = 2 * WH_STAGE + 1024;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d_wgrad[tcgen05 halo]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  const double M = (double)p->N * p->OH * p->OW;
  ProfScope prof(PF_WGRAD, 2.0 * M * p->R * p->S * p->Cin * p->Cout, 4.0 * (M * p->Cout + (double)p->N * p->H * p->W * p->Cin), st);
  launch_k(wgrad_halo_kernel, dim3((unsigned)(groups * a.splits)), dim3(WH_THREADS), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d_wgrad[tcgen05 halo]");
}

}  // namespace fdg
