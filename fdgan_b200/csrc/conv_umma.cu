// fdg_conv2d, tcgen05 (5th-gen tensor core) implicit-GEMM path for sm_100a.
//
//   D[128 pixels x NT channels] (fp32, TMEM) += A[128 x 64] (bf16, smem) * B[NT x 64]^T (bf16, smem)   per K chunk
//
// * Split precision: every fp32 operand value is split hi = bf16(x), lo = bf16(x - hi) and each K chunk
//   issues three MMAs per 16-deep slice (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM accumulator, i.e.
//   ~16 operand mantissa bits (plain bf16/tf32 operands miss the 1e-3 parity bar).
// * A operand: the activations are fp32 NHWC in HBM and need the consumer-side prologue (BatchNorm
//   scale/shift + ReLU/LeakyReLU, avg-pool / nearest-upsample gather, zero padding AFTER the prologue),
//   which TMA cannot apply -- so 8 loader warps read 128-bit, transform in registers, split, and store the
//   bf16 hi/lo tiles into shared memory in the canonical K-major SWIZZLE_128B layout, then
//   fence.proxy.async + mbarrier.arrive.  K order = (filter tap, 64-channel chunk).
// * B operand: weights are pre-split and pre-swizzled by fdg_pack_weight_umma into the exact shared-memory
//   image of every (N tile, K chunk), so one 1-D bulk TMA (cp.async.bulk, mbarrier complete_tx) per stage
//   brings hi and lo.
// * One elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=NT, K=16); tcgen05.commit
//   releases the smem stage / publishes the accumulator.  Epilogue warps read TMEM with tcgen05.ld
//   (32x32b.x32), apply alpha/bias/activation/mask, reduce the per-channel BatchNorm statistics with a
//   butterfly of warp shuffles and store 128-bit rows.
#include <atomic>
#include <cstdlib>

#include "aop.cuh"
#include "pack.cuh"
#include "umma.cuh"
#include "umma_epilogue.cuh"

namespace fdg {

constexpr int UM = 128;          // pixels per tile (UMMA M)
constexpr int UKC = 64;          // K elements per chunk (128 B of bf16 = one swizzle row)
constexpr int ULOAD_WARPS = 8;
constexpr int UTHREADS = (ULOAD_WARPS + 1) * 32;   // + 1 control warp (MMA issue, B bulk copies, TMEM alloc)
constexpr int A_TILE_BYTES = UM * 128;             // one bf16 [128 x 64] tile
constexpr int UMAX_AFF = 1024;                     // input channels whose scale/shift are cached in shared memory

struct UmmaArgs {
  FdgConv c;
  AOp ao;
  int64_t M;
  int cchunks;   // ceil(Cin / 64)
  int nchunks;   // taps * cchunks
  int yvec;
  int dbg;   // ablation: 1 no global loads, 2 no split/stores, 4 no MMA, 8 no epilogue
  int tma_rank;                 // 2: the epilogue stores through ymap {channel, linear pixel}; 0: coalesced stores
  int epi_wrows;                // EpiTma.wrows: 32 = every epilogue warp stores its own 32 pixels (box {32, 32}); 0 = one {32, 128} box per group
  int bn_linear;                // BatchNorm-backward epilogue: y and e are pixel-linear views (pipelined variant)
  int a_split;                  // the A operand arrives as split-bf16 planes through xmap_hi / xmap_lo (no loader work)
  int pf_ahead;                 // FAST: chunks of L2 prefetch in front of the staging ring (0 = none)
  alignas(64) CUtensorMap xmap_hi;
  alignas(64) CUtensorMap xmap_lo;
  alignas(64) CUtensorMap ymap;
};

__device__ __forceinline__ float epi_act_u(float v, int act) {
  switch (act) {
    case FDG_ACT_RELU: return fmaxf(v, 0.f);
    case FDG_ACT_TANH: return tanhf(v);
    case FDG_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// column sums of a 32-lane x 32-column register tile: afterwards lane l holds the sum of column l in v[0]
__device__ __forceinline__ float butterfly_colsum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: grid = min(#tiles, #SMs) CTAs of 13 warps
//   warps 0..7   A loaders (thread 0 also issues the bulk-TMA copy of each stage's weight tile)
//   warp  8      control: one thread issues tcgen05.mma / tcgen05.commit; the warp owns the TMEM allocation
//   warps 9..12  epilogue: TMEM -> registers -> global, overlapped with the next tile's main loop through
//                two TMEM accumulator buffers
// The shared-memory ring and the loaders' register double buffer run straight across tile boundaries.
constexpr int UEPI_WARP0 = ULOAD_WARPS + 1;
constexpr int UTHREADS_P = (ULOAD_WARPS + 1 + 4) * 32;

// DEPTH > 0: the loaders fetch through cp.async into a thread-private shared-memory staging ring (DEPTH chunks in
// flight, no registers tied up by loads in flight); DEPTH == 0: two-chunk register double buffer.
//
// FAST (1x1 / stride 1 / direct gather over a pixel-linear view, DEPTH >= 2): the fp32 rows reach the staging ring through the
// bulk-tensor engine instead of per-thread cp.async -- a 14th warp streams [128 pixels x 64 channels] chunks as two SWIZZLE_128B
// boxes of 32 floats (zero fill past Cin / M comes from the tensor map) and also owns the weight-tile copies; the loader warps
// only wait on an mbarrier, read their rows conflict-free, run the prologue + bf16 hi/lo split and store the operand tiles: no
// address arithmetic, bounds predicates or cursor bookkeeping on the LSU-bound warps (ncu r01g: 18 instructions per fp32
// element in the generic loader against ~6 here).
template <int NT, int STAGES, int DEPTH, bool BNBWD, bool FAST = false>
__global__ void __launch_bounds__(FAST ? UTHREADS_P + 32 : UTHREADS_P, 1) conv_umma_kernel(const __grid_constant__ UmmaArgs a) {
  constexpr int B_TILE_BYTES = NT * 128;
  constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  constexpr int STAGING_BYTES = UM * UKC * 4;          // one raw fp32 chunk
  static_assert(NT <= 128, "two double-width accumulators must fit the 512 TMEM columns");
  constexpr int TMEM_COLS = 4 * NT;        // two accumulator buffers of [hi*hi | hi*lo + lo*hi] halves
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_stg_full[DEPTH > 0 ? DEPTH : 1];     // FAST: staging slot filled by the bulk-tensor engine
  __shared__ __align__(8) uint64_t bar_stg_empty[DEPTH > 0 ? DEPTH : 1];    // FAST: every loader warp has read the slot
  __shared__ uint32_t tmem_base_s;
  __shared__ float sred[2][4][NT];
  __shared__ __align__(1024) uint8_t ep_stage[EP_TILE_BYTES];   // epilogue staging tile: 128 pixels x 32 channels, SWIZZLE_128B box layout
  __shared__ __align__(16) float aff_s[2][UMAX_AFF];   // BatchNorm scale / shift of the input channels (when they fit)

  const FdgConv& p = a.c;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-byte alignment
  const int OHW = p.OH * p.OW;
  const int m_tiles = (int)((a.M + UM - 1) / UM);
  const int n_tiles = (p.Cout + NT - 1) / NT;
  const int total_tiles = m_tiles * n_tiles;

  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), (BNBWD && a.a_split) ? 1 : ULOAD_WARPS + 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    if (FAST) {
      for (int d = 0; d < DEPTH; ++d) { mbar_init(smem_u32(&bar_stg_full[d]), 1); mbar_init(smem_u32(&bar_stg_empty[d]), ULOAD_WARPS); }
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_acc_full[b]), 1);
      mbar_init(smem_u32(&bar_acc_empty[b]), (BNBWD && a.a_split) ? 8 : 4);   // split input: a second epilogue set (warps 4-7)
    }
    fence_barrier_init();
  }
  if (warp == ULOAD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();      // PDL contract (common.cuh): barriers and TMEM are set up while the previous kernel drains; no global memory before this
  pdl_trigger();
  const bool aff_smem = p.has_affine && p.Cin <= UMAX_AFF;
  if (aff_smem) {
    for (int i = t; i < p.Cin; i += (int)blockDim.x) { aff_s[0][i] = __ldg(p.scale + i); aff_s[1][i] = __ldg(p.shift + i); }
  }
  for (int i = t; i < 2 * 4 * NT; i += (int)blockDim.x) (&sred[0][0][0])[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // With a split-bf16 input the loader warps have nothing to load: warp 0 drives the bulk copies and warps 4-7 become a
  // second set of epilogue warps (the kernel is then bound by its epilogue: mask reads, L2 reductions, statistics).
  const bool epi2 = BNBWD && a.a_split && warp >= 4 && warp < ULOAD_WARPS;
  if (BNBWD && a.a_split && warp < 4) {
    // =============================================================== split-bf16 input: both operands by the bulk-copy engine
    // The producer of the input (fdg_ew_bwd, out_split) already wrote it as bf16 hi / lo planes, i.e. in the operand
    // format; one thread streams [128 pixels x 64 channels] boxes of both planes (SWIZZLE_128B tensor maps place them
    // exactly like the loaders would) and the weight tile of the stage when it changed.  No loading, converting and
    // re-storing through the LSU pipe.
    if (t == 0) {
      uint32_t held[STAGES];
#pragma unroll
      for (int j = 0; j < STAGES; ++j) held[j] = 0xffffffffu;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        for (int kc = 0; kc < a.nchunks; ++kc) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES, b_hi = a_lo + A_TILE_BYTES;
          const uint32_t bar = smem_u32(&bar_full[s]);
          const uint32_t want = (uint32_t)(nt * a.nchunks + kc);
          uint32_t have = 0xffffffffu;
#pragma unroll
          for (int j = 0; j < STAGES; ++j) if (j == s) have = held[j];
          mbar_arrive_expect_tx(bar, 2 * A_TILE_BYTES + (have != want ? 2 * B_TILE_BYTES : 0));
          tma_load_2d(a_hi, &a.xmap_hi, kc * UKC, mt * UM, bar);
          tma_load_2d(a_lo, &a.xmap_lo, kc * UKC, mt * UM, bar);
          if (have != want) {
            bulk_g2s(b_hi, reinterpret_cast<const uint8_t*>(p.w_umma) + (size_t)want * (2 * B_TILE_BYTES), 2 * B_TILE_BYTES, bar);
#pragma unroll
            for (int j = 0; j < STAGES; ++j) if (j == s) held[j] = want;
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (FAST && warp == UEPI_WARP0 + 4) {
    // =============================================================== FAST: bulk-tensor producer (activations + weight tiles)
    if (lane == 0) {
      uint32_t held[STAGES];
#pragma unroll
      for (int j = 0; j < STAGES; ++j) held[j] = 0xffffffffu;
      const uint32_t stg0 = smem_base + STAGES * STAGE_BYTES;
      int my_tiles = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++my_tiles;
      const int total_chunks = my_tiles * a.nchunks;
      // cursor of the next chunk to stage (runs DEPTH chunks ahead of the weight-tile cursor)
      int s_tile = blockIdx.x, s_kc = 0, s_q = 0;
      // L2 prefetch cursor, a.pf_ahead chunks in front of the staging cursor: two 32 KB slots in flight per SM are ~9.5 MB on the chip,
      // about what HBM's loaded latency needs, but a slot is refilled only after the loaders drained it -- with the rows already in L2
      // the refill is an L2 hit
      int f_tile = blockIdx.x, f_kc = 0, f_q = 0;
      auto prefetch_to = [&](int upto) {
        while (f_q < upto && f_q < total_chunks) {
          const int mt = f_tile % m_tiles;
          if (f_kc * UKC < p.Cin) tma_prefetch_2d(&a.xmap_hi, f_kc * UKC, mt * UM);
          if (f_kc * UKC + 32 < p.Cin) tma_prefetch_2d(&a.xmap_hi, f_kc * UKC + 32, mt * UM);
          ++f_q;
          if (++f_kc == a.nchunks) { f_kc = 0; f_tile += gridDim.x; }
        }
      };
      auto stage_next = [&]() {
        if (a.pf_ahead > 0) prefetch_to(s_q + 1 + a.pf_ahead);
        const int slot = s_q % DEPTH;
        if (s_q >= DEPTH) mbar_wait(smem_u32(&bar_stg_empty[slot]), (uint32_t)((s_q / DEPTH) - 1) & 1u);
        const uint32_t bar = smem_u32(&bar_stg_full[slot]);
        const uint32_t dst = stg0 + (uint32_t)slot * STAGING_BYTES;
        const int mt = s_tile % m_tiles;
        mbar_arrive_expect_tx(bar, STAGING_BYTES);
        tma_load_2d(dst, &a.xmap_hi, s_kc * UKC, mt * UM, bar);                          // channels [0, 32) of the chunk
        tma_load_2d(dst + STAGING_BYTES / 2, &a.xmap_hi, s_kc * UKC + 32, mt * UM, bar);   // channels [32, 64)
        ++s_q;
        if (++s_kc == a.nchunks) { s_kc = 0; s_tile += gridDim.x; }
      };
      for (int q = 0; q < DEPTH && q < total_chunks; ++q) stage_next();
      int s = 0;
      uint32_t ph = 0;
      int q = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / m_tiles;
        for (int kc = 0; kc < a.nchunks; ++kc, ++q) {
          // weight tile of chunk q into ring stage s (kept while the same tile is needed again)
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          const uint32_t bar = smem_u32(&bar_full[s]);
          const uint32_t want = (uint32_t)(nt * a.nchunks + kc);
          uint32_t have = 0xffffffffu;
#pragma unroll
          for (int j = 0; j < STAGES; ++j) if (j == s) have = held[j];
          if (have != want) {
            mbar_arrive_expect_tx(bar, 2 * B_TILE_BYTES);
            bulk_g2s(smem_base + s * STAGE_BYTES + 2 * A_TILE_BYTES, reinterpret_cast<const uint8_t*>(p.w_umma) + (size_t)want * (2 * B_TILE_BYTES),
                     2 * B_TILE_BYTES, bar);
#pragma unroll
            for (int j = 0; j < STAGES; ++j) if (j == s) held[j] = want;
          } else {
            mbar_arrive(bar);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (s_q < total_chunks) stage_next();      // refill the slot the loaders free while they work on chunk q
        }
      }
    }
  } else if (FAST && warp < ULOAD_WARPS) {
    // =============================================================== FAST: A loaders fed by the staging ring
    // thread = 16-byte bf16 chunk j (8 channels) of rows rbase + 32 i; in the fp32 boxes those channels are the two 16-byte
    // chunks 2(j&3), 2(j&3)+1 of box j>>2, swizzled by row & 7.  Threads j >= 4 read their two chunks in the opposite order,
    // so that the 8 threads of a row always touch 8 different bank groups (box 0 and box 1 rows alias bank-wise).
    constexpr int RPT = UM * 8 / (ULOAD_WARPS * 32);   // 4 rows per thread
    const int j = t & 7, rbase = t >> 3;
    const uint32_t swz = (uint32_t)(rbase & 7);
    const uint32_t c_even = ((uint32_t)(2 * (j & 3)) ^ swz) << 4, c_odd = ((uint32_t)(2 * (j & 3) + 1) ^ swz) << 4;
    const uint32_t src0 = smem_base + STAGES * STAGE_BYTES + (uint32_t)(j >> 2) * (STAGING_BYTES / 2) + (uint32_t)rbase * 128u;
    const uint32_t dst0 = (uint32_t)rbase * 128u + (((uint32_t)j ^ swz) << 4);
    const bool flip = (j & 4) != 0;
    const uint32_t aff0 = smem_u32(&aff_s[0][0]);
    const float sl = p.slope;
    auto lds4u = [](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
      return v;
    };
    int s = 0, q = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile % m_tiles;
      const int64_t mleft = a.M - (int64_t)mt * UM - rbase;      // rows rbase + 32 i exist while 32 i < mleft
      for (int kc = 0; kc < a.nchunks; ++kc, ++q) {
        const int slot = q % DEPTH;
        const int c = kc * UKC + j * 8;
        const bool cvalid = c < p.Cin;
        float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
        if (p.has_affine && cvalid) {
          sc0 = lds4u(aff0 + c * 4); sc1 = lds4u(aff0 + c * 4 + 16);
          sh0 = lds4u(aff0 + (UMAX_AFF + c) * 4); sh1 = lds4u(aff0 + (UMAX_AFF + c) * 4 + 16);
        }
        mbar_wait(smem_u32(&bar_stg_full[slot]), (uint32_t)(q / DEPTH) & 1u);
        if (a.dbg & 2) {      // ablation: barrier protocol only (no shared-memory reads, arithmetic or operand stores)
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_stg_empty[slot]));
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          continue;
        }
        const uint32_t src = src0 + (uint32_t)slot * STAGING_BYTES;
        float4 v0[RPT], v1[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint32_t rs = src + (uint32_t)i * (32u * 128u);
          if (flip) { v1[i] = lds4u(rs + c_odd); v0[i] = lds4u(rs + c_even); }
          else { v0[i] = lds4u(rs + c_even); v1[i] = lds4u(rs + c_odd); }
        }
        uint32_t h[RPT][4], l[RPT][4];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          float4 x0 = v0[i], x1 = v1[i];
          if (p.has_affine) {
            x0.x = fmaf(x0.x, sc0.x, sh0.x); x0.y = fmaf(x0.y, sc0.y, sh0.y); x0.z = fmaf(x0.z, sc0.z, sh0.z); x0.w = fmaf(x0.w, sc0.w, sh0.w);
            x1.x = fmaf(x1.x, sc1.x, sh1.x); x1.y = fmaf(x1.y, sc1.y, sh1.y); x1.z = fmaf(x1.z, sc1.z, sh1.z); x1.w = fmaf(x1.w, sc1.w, sh1.w);
          }
          if (sl == 0.f) {
            x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
            x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
          } else if (sl != 1.f) {
            x0.x = prologue_act(x0.x, sl); x0.y = prologue_act(x0.y, sl); x0.z = prologue_act(x0.z, sl); x0.w = prologue_act(x0.w, sl);
            x1.x = prologue_act(x1.x, sl); x1.y = prologue_act(x1.y, sl); x1.z = prologue_act(x1.z, sl); x1.w = prologue_act(x1.w, sl);
          }
          if (!cvalid || (int64_t)(32 * i) >= mleft) {      // zero padding AFTER the prologue (a shifted zero is not zero)
            x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0;
          }
          split2(x0.x, x0.y, h[i][0], l[i][0]);
          split2(x0.z, x0.w, h[i][1], l[i][1]);
          split2(x1.x, x1.y, h[i][2], l[i][2]);
          split2(x1.z, x1.w, h[i][3], l[i][3]);
        }
        // the slot's values are in registers (consumed by the arithmetic above): hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_stg_empty[slot]));
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t a_hi = smem_base + s * STAGE_BYTES + dst0, a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const uint32_t off = (uint32_t)i * (32u * 128u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[i][0]), "r"(h[i][1]), "r"(h[i][2]), "r"(h[i][3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[i][0]), "r"(l[i][1]), "r"(l[i][2]), "r"(l[i][3]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < ULOAD_WARPS && !epi2) {
    // =============================================================== A loaders
    // thread handles the 16-byte bf16 chunk j (8 channels) of rows rbase + 64*i, i = 0..1
    constexpr int RPT = UM * 8 / (ULOAD_WARPS * 32);   // rows per thread
    constexpr int RSTEP = ULOAD_WARPS * 4;             // row stride between them
    const int j = t & 7, rbase = t >> 3;
    const bool direct = p.gather == FDG_GATHER_DIRECT;
    const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);   // 8 bytes per stage
    const uint32_t aff0 = smem_u32(&aff_s[0][0]);
    auto lds4u = [](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      return v;
    };
    // ---- load cursor: (tile, chunk) of the next chunk to fetch, with the tile's pixel coordinates
    int l_tile = blockIdx.x, l_kc = 0, l_r = 0, l_sx = 0, l_cc = 0, l_nt = 0;
    int pn[RPT], piy[RPT], pix[RPT];
    const float* rowp[RPT];
    uint32_t pvmask = 0;
    auto set_tile = [&](int tile) {
      const int mt = tile % m_tiles;
      l_nt = tile / m_tiles;
      pvmask = 0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int64_t m = (int64_t)mt * UM + rbase + RSTEP * i;
        const bool v = m < a.M;
        pvmask |= (v ? 1u : 0u) << i;
        const int64_t mm = v ? m : 0;
        pn[i] = (int)(mm / OHW);
        const int rem = (int)(mm - (int64_t)pn[i] * OHW);
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        piy[i] = oy * p.stride - p.pad;
        pix[i] = ox * p.stride - p.pad;
        rowp[i] = p.x.p + pn[i] * p.x.sn + (int64_t)piy[i] * p.x.sh + (int64_t)pix[i] * p.x.sw + j * 8;  // dereferenced only in range
      }
    };
    // issue the global loads of the cursor's chunk into registers (raw values for the direct gather; the pooled /
    // upsampled gathers apply the prologue inside fetch4 because it has to precede the averaging); returns metadata
    // (bits 0..3 in-range mask, bits 8.. channel-chunk index, bits 16.. weight-tile index) and advances the cursor
    auto issue = [&](float4 (&v0)[RPT], float4 (&v1)[RPT]) -> uint32_t {
      const int c = l_cc * UKC + j * 8;
      const bool cvalid = c < p.Cin;
      const int64_t toff = (int64_t)l_r * p.x.sh + (int64_t)l_sx * p.x.sw + l_cc * UKC;
      uint32_t ok = 0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int iy = piy[i] + l_r, ix = pix[i] + l_sx;
        v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[i] = v0[i];
        if (((pvmask >> i) & 1u) && cvalid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
          ok |= 1u << i;
          if (direct) {
            v0[i] = ld4(rowp[i] + toff);
            v1[i] = ld4(rowp[i] + toff + 4);
          } else {
            v0[i] = fetch4(a.ao, pn[i], iy, ix, c);
            v1[i] = fetch4(a.ao, pn[i], iy, ix, c + 4);
          }
        }
      }
      const uint32_t meta = ok | ((uint32_t)l_cc << 8) | ((uint32_t)(l_nt * a.nchunks + l_kc) << 16);
      if (++l_cc == a.cchunks) { l_cc = 0; if (++l_sx == p.S) { l_sx = 0; ++l_r; } }
      if (++l_kc == a.nchunks) {
        l_kc = 0; l_r = 0; l_sx = 0; l_cc = 0;
        l_tile += gridDim.x;
        if (l_tile < total_tiles) set_tile(l_tile);
      }
      return meta;
    };
    // prologue (direct gather), bf16 hi/lo split and swizzled store of one K chunk into stage s
    uint32_t held[STAGES];            // thread 0: which weight tile each stage holds
#pragma unroll
    for (int j = 0; j < STAGES; ++j) held[j] = 0xffffffffu;
    auto finish = [&](float4 (&v0)[RPT], float4 (&v1)[RPT], uint32_t meta, int s, uint32_t ph) {
      if (direct) {
        const int c = (int)((meta >> 8) & 0xffu) * UKC + j * 8;
        const float sl = p.slope;
        if (p.has_affine) {
          float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sh1 = sh0;
          if (c < p.Cin) {
            if (aff_smem) {
              sc0 = lds4u(aff0 + c * 4); sc1 = lds4u(aff0 + c * 4 + 16);
              sh0 = lds4u(aff0 + (UMAX_AFF + c) * 4); sh1 = lds4u(aff0 + (UMAX_AFF + c) * 4 + 16);
            } else {
              sc0 = ld4(p.scale + c); sc1 = ld4(p.scale + c + 4);
              sh0 = ld4(p.shift + c); sh1 = ld4(p.shift + c + 4);
            }
          }
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            v0[i].x = fmaf(v0[i].x, sc0.x, sh0.x); v0[i].y = fmaf(v0[i].y, sc0.y, sh0.y);
            v0[i].z = fmaf(v0[i].z, sc0.z, sh0.z); v0[i].w = fmaf(v0[i].w, sc0.w, sh0.w);
            v1[i].x = fmaf(v1[i].x, sc1.x, sh1.x); v1[i].y = fmaf(v1[i].y, sc1.y, sh1.y);
            v1[i].z = fmaf(v1[i].z, sc1.z, sh1.z); v1[i].w = fmaf(v1[i].w, sc1.w, sh1.w);
          }
        }
        if (sl == 0.f) {            // ReLU
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            v0[i].x = fmaxf(v0[i].x, 0.f); v0[i].y = fmaxf(v0[i].y, 0.f); v0[i].z = fmaxf(v0[i].z, 0.f); v0[i].w = fmaxf(v0[i].w, 0.f);
            v1[i].x = fmaxf(v1[i].x, 0.f); v1[i].y = fmaxf(v1[i].y, 0.f); v1[i].z = fmaxf(v1[i].z, 0.f); v1[i].w = fmaxf(v1[i].w, 0.f);
          }
        } else if (sl != 1.f) {     // LeakyReLU
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            v0[i].x = prologue_act(v0[i].x, sl); v0[i].y = prologue_act(v0[i].y, sl); v0[i].z = prologue_act(v0[i].z, sl); v0[i].w = prologue_act(v0[i].w, sl);
            v1[i].x = prologue_act(v1[i].x, sl); v1[i].y = prologue_act(v1[i].y, sl); v1[i].z = prologue_act(v1[i].z, sl); v1[i].w = prologue_act(v1[i].w, sl);
          }
        }
        if (p.has_affine) {         // zero padding is applied AFTER the prologue (a shifted zero is not zero)
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            if (!((meta >> i) & 1u)) { v0[i] = make_float4(0.f, 0.f, 0.f, 0.f); v1[i] = v0[i]; }
        }
      }
      mbar_wait(empty0 + s * 8, ph ^ 1u);
      const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
      if (t == 0) {   // this stage's weight tile (hi + lo) through the bulk-copy engine
        // A stage keeps its weight tile until a different one is needed: with one or two K chunks per tile (1x1 convs of
        // up to 128 input channels, every dense-layer conv1 data gradient) the tiles stay resident across M tiles and
        // their L2 latency disappears from the ring.
        const uint32_t bar = full0 + s * 8;
        const uint32_t want = meta >> 16;
        uint32_t have = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < STAGES; ++j) if (j == s) have = held[j];
        if (have != want) {
          mbar_arrive_expect_tx(bar, 2 * B_TILE_BYTES);
          bulk_g2s(a_lo + A_TILE_BYTES, reinterpret_cast<const uint8_t*>(p.w_umma) + (size_t)want * (2 * B_TILE_BYTES), 2 * B_TILE_BYTES, bar);
#pragma unroll
          for (int j = 0; j < STAGES; ++j) if (j == s) held[j] = want;
        } else {
          mbar_arrive(bar);
        }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        if (a.dbg & 2) break;
        const int row = rbase + RSTEP * i;
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
        uint32_t h[4], l[4];
        split2(v0[i].x, v0[i].y, h[0], l[0]);
        split2(v0[i].z, v0[i].w, h[1], l[1]);
        split2(v1[i].x, v1[i].y, h[2], l[2]);
        split2(v1[i].z, v1[i].w, h[3], l[3]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
      }
      fence_proxy_async();          // make this thread's generic-proxy stores visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + s * 8);   // one arrival per loader warp
    };
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++my_tiles;
    const int total_chunks = my_tiles * a.nchunks;
    if (total_chunks > 0) {
      set_tile(l_tile);
      int s = 0;
      uint32_t ph = 0;
      if (DEPTH > 0) {
        // ---- cp.async staging: slot d of this thread = RPT rows x 32 bytes, laid out [d][row i][thread]
        constexpr int NLT = ULOAD_WARPS * 32;
        const uint32_t stg = smem_base + STAGES * STAGE_BYTES + (uint32_t)t * 16u;   // [slot][row][half][thread] x 16 B
        __shared__ uint32_t meta_s[DEPTH > 0 ? DEPTH : 1][ULOAD_WARPS * 32];
        const uint32_t meta0 = smem_u32(&meta_s[0][0]) + (uint32_t)t * 4u;
        auto issue_async = [&](int d) {
          const int c = l_cc * UKC + j * 8;
          const bool cvalid = c < p.Cin;
          const int64_t toff = (int64_t)l_r * p.x.sh + (int64_t)l_sx * p.x.sw + l_cc * UKC;
          uint32_t ok = 0;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int iy = piy[i] + l_r, ix = pix[i] + l_sx;
            const bool v = ((pvmask >> i) & 1u) && cvalid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && !(a.dbg & 1);
            const uint32_t dst = stg + (uint32_t)((d * RPT + i) * 2 * NLT) * 16u;
            if (direct) {
              if (v) {
                ok |= 1u << i;
                const float* src = rowp[i] + toff;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + NLT * 16), "l"(src + 4) : "memory");
              }
            } else {
              float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
              if (v) { ok |= 1u << i; f0 = fetch4(a.ao, pn[i], iy, ix, c); f1 = fetch4(a.ao, pn[i], iy, ix, c + 4); }
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f0.x), "f"(f0.y), "f"(f0.z), "f"(f0.w) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + NLT * 16), "f"(f1.x), "f"(f1.y), "f"(f1.z), "f"(f1.w) : "memory");
            }
          }
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(meta0 + d * (NLT * 4)), "r"(ok | ((uint32_t)l_cc << 8) | ((uint32_t)(l_nt * a.nchunks + l_kc) << 16)) : "memory");
          if (++l_cc == a.cchunks) { l_cc = 0; if (++l_sx == p.S) { l_sx = 0; ++l_r; } }
          if (++l_kc == a.nchunks) {
            l_kc = 0; l_r = 0; l_sx = 0; l_cc = 0;
            l_tile += gridDim.x;
            if (l_tile < total_tiles) set_tile(l_tile);
          }
        };
        int dl = 0, df = 0;     // staging slots of the next chunk to load / to finish
#pragma unroll 1
        for (int q = 0; q < DEPTH - 1; ++q) {
          if (q < total_chunks) issue_async(dl);
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (++dl == DEPTH) dl = 0;
        }
#pragma unroll 1
        for (int q = 0; q < total_chunks; ++q) {
          if (q + DEPTH - 1 < total_chunks) issue_async(dl);
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (++dl == DEPTH) dl = 0;
          asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH > 0 ? DEPTH - 1 : 0) : "memory");
          uint32_t meta;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(meta) : "r"(meta0 + df * (NLT * 4)) : "memory");
          float4 v0[RPT], v1[RPT];
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const uint32_t src = stg + (uint32_t)((df * RPT + i) * 2 * NLT) * 16u;
            if ((meta >> i) & 1u) {
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0[i].x), "=f"(v0[i].y), "=f"(v0[i].z), "=f"(v0[i].w) : "r"(src) : "memory");
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v1[i].x), "=f"(v1[i].y), "=f"(v1[i].z), "=f"(v1[i].w) : "r"(src + NLT * 16) : "memory");
            } else {
              v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              v1[i] = v0[i];
            }
          }
          finish(v0, v1, meta, s, ph);
          if (++df == DEPTH) df = 0;
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      } else {
        float4 A0[RPT], A1[RPT], B0[RPT], B1[RPT];
        uint32_t metaA, metaB = 0;
        metaA = issue(A0, A1);
        for (int q = 0; q < total_chunks; q += 2) {       // two chunks of loads in flight per thread
          if (q + 1 < total_chunks) metaB = issue(B0, B1);
          finish(A0, A1, metaA, s, ph);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (q + 1 < total_chunks) {
            if (q + 2 < total_chunks) metaA = issue(A0, A1);
            finish(B0, B1, metaB, s, ph);
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == ULOAD_WARPS) {
    // =============================================================== control thread: MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(UM, 2 * NT), idesc1 = umma_idesc_bf16(UM, NT);
      const uint32_t k_hw = umma_desc_hi(1024);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(smem_u32(&bar_acc_empty[b]), (((uint32_t)it >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * 2 * NT);
        for (int kc = 0; kc < a.nchunks; ++kc) {
          mbar_wait(smem_u32(&bar_full[s]), ph);
          tc_fence_after();
          const uint32_t a_hi = smem_base + s * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_lo + A_TILE_BYTES;     // [B_hi | B_lo] are adjacent: one operand of 2*NT rows
          const int cleft = p.Cin - (kc % a.cchunks) * UKC;
          const int kslices = cleft >= UKC ? 4 : (cleft + 15) / 16;      // K = 16 slices of this chunk that hold channels
          if (!(a.dbg & 4)) {
            if (kslices == 4) {
              umma_chunk8(d_tmem, umma_desc_lo(a_hi, 16), umma_desc_lo(a_lo, 16), k_hw, umma_desc_lo(b_hi, 16), k_hw, idesc2, idesc1,
                          kc > 0 ? 1u : 0u, 2u, 2u, (uint32_t)NT);
            } else {
              const uint32_t ah = umma_desc_lo(a_hi, 16), al = umma_desc_lo(a_lo, 16), bl = umma_desc_lo(b_hi, 16);
              for (int sl = 0; sl < kslices; ++sl)
                umma_concat_slice(d_tmem, ah + 2u * sl, al + 2u * sl, k_hw, bl + 2u * sl, k_hw, idesc2, idesc1, (kc > 0 || sl > 0) ? 1u : 0u, (uint32_t)NT);
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));                // frees this stage when the MMAs above retire
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[b]));               // accumulator of this tile complete
      }
    }
  } else {
    // =============================================================== epilogue warps (TMEM lane quarter = warp & 3)
    const int quarter = warp & 3;                      // TMEM lanes this warp may read
    const int wset = epi2 ? 1 : 0, nsets = (BNBWD && a.a_split) ? 2 : 1;   // epilogue set: 32-channel groups g = wset (mod nsets)
    const int et = epi2 ? t : t - UEPI_WARP0 * 32;     // set 1 (threads 128..255) keeps its thread index
    const uint32_t stage = epi2 ? smem_base + STAGES * STAGE_BYTES : smem_u32(ep_stage);   // set 1 stages in the idle cp.async ring
    const bool evec = p.e.p && p.e.sc == 1 && aligned16_dev(p.e.p) && (p.e.sn % 4 == 0) && (p.e.sh % 4 == 0) && (p.e.sw % 4 == 0);
    // BatchNorm-backward epilogue on pixel-linear views (dense-block conv1 data gradient): the mask-tensor rows of the
    // NEXT 32-channel group (or of the next tile's first group) are fetched while the current group is processed, so
    // their HBM latency never sits between tcgen05.ld and the stores.
    const bool bn_lin = BNBWD;      // separate instantiation: its register needs must not spill the common kernel
    float4 evn[8];
    const int bq = lane & 7, bc4 = bq * 4, brs = lane >> 3;
    auto bn_prefetch = [&](int tile_, int g_) {
      const int mt_ = tile_ % m_tiles, c0_ = (tile_ / m_tiles) * NT + g_ * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m_ = (int64_t)mt_ * UM + quarter * 32 + 4 * i + brs;
        evn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m_ < a.M && c0_ + bc4 < p.Cout) evn[i] = ld4(p.e.p + m_ * p.e.sw + c0_ + bc4);
      }
    };
    auto bn_ngroups = [&](int tile_) {
      const int left = (p.Cout - (tile_ / m_tiles) * NT + 31) / 32;
      return left < NT / 32 ? left : NT / 32;
    };
    // first (tile, group) of this set at or after tile_: tiles whose channel tail has no group for the set are skipped
    auto bn_first = [&](int tile_) {
      while (tile_ < total_tiles && wset >= bn_ngroups(tile_)) tile_ += gridDim.x;
      return tile_;
    };
    if (BNBWD) { const int t0_ = bn_first(blockIdx.x); if (t0_ < total_tiles) bn_prefetch(t0_, wset); }
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int mt = tile % m_tiles, ntile = tile / m_tiles;
      const int b = it & 1;
      while (!mbar_try_wait(smem_u32(&bar_acc_full[b]), ((uint32_t)it >> 1) & 1u)) __nanosleep(200);   // leave the issue slots to the loaders
      tc_fence_after();
      const int64_t m = (int64_t)mt * UM + quarter * 32 + lane;
      const bool mv = m < a.M;
      int64_t yoff = 0, eoff = 0;
      if (mv) {
        const int n = (int)(m / OHW);
        const int rem = (int)(m - (int64_t)n * OHW);
        const int oy = rem / p.OW, ox = rem - oy * p.OW;
        const int us = p.store == FDG_STORE_UP2 ? 2 : 1;
        yoff = n * p.y.sn + (int64_t)(us * oy) * p.y.sh + (int64_t)(us * ox) * p.y.sw;
        if (p.e.p) eoff = n * p.e.sn + (int64_t)oy * p.e.sh + (int64_t)ox * p.e.sw;
      }
      const EpiTma tm{a.tma_rank ? (const void*)&a.ymap : nullptr, a.tma_rank, mt * UM, 0, 0, a.epi_wrows};
      const int cbase = ntile * NT;
      if constexpr (BNBWD) {
        const int ngroups = (p.Cout - cbase + 31) / 32 < NT / 32 ? (p.Cout - cbase + 31) / 32 : NT / 32;
        const uint32_t wrow0 = stage + (uint32_t)(quarter * 32) * 128u;
#pragma unroll 1
        for (int g = wset; g < ngroups; g += nsets) {
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
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + brs;
            const int64_t mr = (int64_t)mt * UM + quarter * 32 + row;
            if (mr < a.M && cv) {
              float4 val;
              const uint32_t ta = wrow0 + (uint32_t)row * 128u + (uint32_t)((bq ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w) : "r"(ta) : "memory");
              const float4 ev = evn[i];
              const float al = p.alpha;
              val.x *= fmaf(ev.x, bsc.x, bsh.x) > 0.f ? al : al * p.eslope; val.y *= fmaf(ev.y, bsc.y, bsh.y) > 0.f ? al : al * p.eslope;
              val.z *= fmaf(ev.z, bsc.z, bsh.z) > 0.f ? al : al * p.eslope; val.w *= fmaf(ev.w, bsc.w, bsh.w) > 0.f ? al : al * p.eslope;
              ps1.x += val.x; ps1.y += val.y; ps1.z += val.z; ps1.w += val.w;
              ps2.x = fmaf(val.x, ev.x, ps2.x); ps2.y = fmaf(val.y, ev.y, ps2.y); ps2.z = fmaf(val.z, ev.z, ps2.z); ps2.w = fmaf(val.w, ev.w, ps2.w);
              val.x *= bsc.x; val.y *= bsc.y; val.z *= bsc.z; val.w *= bsc.w;
              float* yp = p.y.p + mr * p.y.sw + c0 + bc4;
              if (p.store == FDG_STORE_ACCUM)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(yp), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
              else
                *reinterpret_cast<float4*>(yp) = val;
            }
          }
          // mask rows of the next group (or of the next tile's first group): in flight during the statistics fold, the
          // next tcgen05.ld / staging stores and, across tiles, the wait for the accumulator
          if (g + nsets < ngroups) bn_prefetch(tile, g + nsets);
          else { const int tn_ = bn_first(tile + gridDim.x); if (tn_ < total_tiles) bn_prefetch(tn_, wset); }
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
      for (int g = 0; g < NT / 32; ++g) {
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
          umma_epilogue_group<false>(p, a.yvec, evec, v, mv, yoff, eoff, c0, lane, quarter, et, stage, tm, &sred[0][quarter][g * 32],
                              &sred[1][quarter][g * 32]);
        }
      }
      }
      // release the accumulator buffer to the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[b]));
      // statistics are per output-channel tile: flush when the next tile of this CTA belongs to another one
      if (p.stats) {
        const int next = tile + gridDim.x;
        if (next >= total_tiles || next / m_tiles != ntile) {
          if (nsets == 2) asm volatile("bar.sync 1, 256;" ::: "memory"); else
          asm volatile("bar.sync 1, 128;" ::: "memory");      // the epilogue warps
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
          if (nsets == 2) asm volatile("bar.sync 1, 256;" ::: "memory"); else
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
    if (a.tma_rank && (a.epi_wrows ? lane == 0 : et == 0) && !epi2) bulk_wait_read0();   // the staging tile must outlive the last bulk store's read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ULOAD_WARPS) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight image
// out[(ntile, kchunk)] = [hi: NT rows x 128 B, SWIZZLE_128B][lo: same]; source = fp32 [K][ld] GEMM operand
__global__ void pack_umma_kernel(const float* __restrict__ w, int ld, int taps, int Cin, int Cout, int NT, int cchunks,
                                 uint8_t* __restrict__ out, int64_t total_pairs, int pair) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_pairs; i += (int64_t)gridDim.x * blockDim.x)
    pack_umma_item(w, ld, taps, Cin, Cout, NT, cchunks, out, i, pair);
}

int halo_enabled();

// Output-channel tile of the packed weight images.  Per-tap kernel (1x1 and everything the halo kernel does not take): 32 / 64 / 128.
// Filters with >= 4 taps run on the halo-tile kernel, which also has 80- and 96-wide instantiations: the MMAs of a partly filled
// tile are issued at full width, so 144 = 2 x 80 (not 2 x 128), 288 = 3 x 96, 72 -> 80, 160 = 2 x 80 (Fusion-D layers 3 / 4 and
// their data gradients: 25-44 % fewer tensor-core cycles).  Cost model: cycles per K = 16 slice of the concatenated-B pair
// (N = 2NT and N = NT; below N = 128 the shared-memory read of the A operand, 32 cycles, is the floor of an MMA).
int umma_ntile(int taps, int Cout) {
  static const int cap = [] { const char* e = getenv("FDG_UMMA_NT_CAP"); return e ? atoi(e) : 128; }();
  static const int fine = [] { const char* e = getenv("FDG_UMMA_NT_FINE"); return e ? atoi(e) : 1; }();
  int nt = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);
  if (fine && taps >= 4 && Cout > 64 && halo_enabled()) {
    const int cand[3] = {128, 96, 80}, cost[3] = {192, 152, 132};
    int best = 1 << 30;
    for (int i = 0; i < 3; ++i) {
      const int c = cdiv(Cout, cand[i]) * cost[i];
      if (c < best) { best = c; nt = cand[i]; }
    }
  }
  return nt > cap ? cap : nt;
}

// Two filter taps per 64-deep weight chunk (pack.cuh): filters of >= 4 taps over at most 32 input channels, halo kernel only.
int umma_tap_pair(int taps, int Cin) {
  static const int on = [] { const char* e = getenv("FDG_UMMA_TAP_PAIR"); return e ? atoi(e) : 1; }();
  return on && taps >= 4 && Cin <= 32 && halo_enabled();
}
int umma_nchunks(int taps, int Cin) { return umma_tap_pair(taps, Cin) ? (taps + 1) / 2 : taps * cdiv(Cin, UKC); }

int conv2d_halo_supported(const FdgConv* p);

int conv2d_umma_supported(const FdgConv* p) {
  if (!p->w_umma) return 0;
  if (p->Cin % 8 != 0) return p->Cin % 4 == 0 && conv2d_halo_supported(p);   // only the halo kernel takes half chunks
  {
    const int nt = umma_ntile(p->R * p->S, p->Cout);
    if (nt != 32 && nt != 64 && nt != 128 && !conv2d_halo_supported(p)) return 0;   // 80 / 96-wide images exist for the halo kernel only
    if (umma_tap_pair(p->R * p->S, p->Cin) && !conv2d_halo_supported(p)) return 0;  // so do tap-pair images
  }
  if (p->Cin < 16 || p->Cout < 1) return 0;
  AOp ao{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  if (!aop_vec_ok(ao, p->Cin)) return 0;
  return 1;
}

template <int NT, int STAGES, int DEPTH, bool BNBWD = false, bool FAST = false>
static int launch_umma(const UmmaArgs& a, cudaStream_t st) {
  constexpr int smem = STAGES * (2 * A_TILE_BYTES + 2 * NT * 128) + DEPTH * (UM * UKC * 4) + 1024;
  static std::atomic<int> attr_done[64];           // per device
  const int adev = current_device();
  if (!attr_done[adev]) {
    if (cudaFuncSetAttribute(conv_umma_kernel<NT, STAGES, DEPTH, BNBWD, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("fdg_conv2d[tcgen05]: cannot raise dynamic shared memory to %d bytes", smem);
      return FDG_ECUDA;
    }
    attr_done[adev] = 1;
  }
  const int64_t tiles = cdiv64(a.M, UM) * cdiv(a.c.Cout, NT);
  const int num_sms = device_sm_count();
  dim3 grid((unsigned)(tiles < num_sms ? tiles : num_sms));
  const double gmul = a.c.gather == FDG_GATHER_AVGPOOL2 ? 4.0 : 1.0;
  ProfScope prof(PF_CONV_UMMA, 2.0 * (double)a.M * a.c.R * a.c.S * a.c.Cin * a.c.Cout,
                 // output + input once, plus what the epilogue reads: the mask tensor (e) and, for accumulating stores, the old values
                 4.0 * ((double)a.M * a.c.Cout * (1.0 + (a.c.e.p ? 1.0 : 0.0) + (a.c.store == FDG_STORE_ACCUM ? 1.0 : 0.0)) +
                        gmul * (double)a.c.N * a.c.H * a.c.W * a.c.Cin), st);
  launch_k(conv_umma_kernel<NT, STAGES, DEPTH, BNBWD, FAST>, dim3(grid), dim3(FAST ? UTHREADS_P + 32 : UTHREADS_P), (size_t)(smem), st, a);
  return check_launch("fdg_conv2d[tcgen05]");
}

int conv2d_halo_supported(const FdgConv* p);
int conv2d_halo(const FdgConv* p, int nt, cudaStream_t st);

int conv2d_umma(const FdgConv* p, cudaStream_t st) {
  const int ntile = umma_ntile(p->R * p->S, p->Cout);
  if (ntile <= 128 && conv2d_halo_supported(p)) return conv2d_halo(p, ntile, st);
  UmmaArgs a;
  a.c = *p;
  a.ao = AOp{p->x, p->H, p->W, p->gather, p->has_affine, p->scale, p->shift, p->slope};
  a.M = (int64_t)p->N * p->OH * p->OW;
  a.cchunks = cdiv(p->Cin, UKC);
  a.nchunks = p->R * p->S * a.cchunks;
  a.yvec = vec4_ok(p->y);
  a.dbg = dbg_flags();
  a.tma_rank = 0;
  a.epi_wrows = 0;
  {
    auto lin = [&](const FdgTensor& t) { return t.sh == (int64_t)p->OW * t.sw && t.sn == (int64_t)p->OH * t.sh; };
    a.bn_linear = p->e_scale && p->e.p && lin(p->y) && lin(p->e) && vec4_ok(p->y) && vec4_ok(p->e) && p->store != FDG_STORE_UP2;
  }
  a.a_split = 0;
  static const int pf_ahead = [] { const char* e = getenv("FDG_CONV_PF"); return e ? atoi(e) : 2; }();   // measured: 0.341 -> 0.335 ms (224->128 @256^2), 0.163 -> 0.153 ms (480->128 @128^2); 8 is slower
  a.pf_ahead = pf_ahead;
  if (p->x_split) {
    if (!(a.bn_linear && p->R == 1 && p->S == 1 && p->stride == 1 && p->pad == 0 && p->gather == FDG_GATHER_DIRECT && !p->has_affine &&
          p->slope == 1.f && p->Cin % 64 == 0 && a.M < (1ll << 31))) {
      set_error("fdg_conv2d[tcgen05]: x_split needs a 1x1 / stride 1 / direct / prologue-free conv with Cin %% 64 == 0 and the BatchNorm-backward epilogue");
      return FDG_ENOSUPPORT;
    }
    const uint64_t dims[2] = {(uint64_t)p->Cin, (uint64_t)a.M};
    const uint64_t strides[1] = {(uint64_t)p->Cin * 2};
    const uint32_t box[2] = {64, 128};
    const uint8_t* hi = reinterpret_cast<const uint8_t*>(p->x_split);
    if (!make_tmap_bf16(&a.xmap_hi, hi, 2, dims, strides, box) || !make_tmap_bf16(&a.xmap_lo, hi + (size_t)a.M * p->Cin * 2, 2, dims, strides, box)) {
      set_error("fdg_conv2d[tcgen05]: cannot build the tensor maps of the split-bf16 input");
      return FDG_ECUDA;
    }
    a.a_split = 1;
  }
  static const int tma_on = [] { const char* e = getenv("FDG_TMA_STORE"); return e ? atoi(e) : 1; }();
  // bulk tensor stores: plain store into a unit-channel-stride, pixel-linear view (dense NHWC or a channel slice of one)
  if (tma_on && a.yvec && p->store == FDG_STORE_NORMAL && !p->e.p && p->y.sh == (int64_t)p->OW * p->y.sw &&
      p->y.sn == (int64_t)p->OH * p->y.sh && a.M < (1ll << 31)) {
    const uint64_t dims[2] = {(uint64_t)p->Cout, (uint64_t)a.M};
    const uint64_t strides[1] = {(uint64_t)p->y.sw * 4};
    a.epi_wrows = epi_warp_stores() ? 32 : 0;
    const uint32_t box[2] = {32, a.epi_wrows ? 32u : 128u};
    if (make_tmap_f32(&a.ymap, p->y.p, 2, dims, strides, box)) a.tma_rank = 2;
  }
  static const int reg_path = [] { const char* e = getenv("FDG_CONV_REG"); return e ? atoi(e) : 0; }();
  if (reg_path && ntile == 128) return launch_umma<128, 3, 0>(a, st);   // register double buffer, no staging ring
  if (p->e_scale) {   // BatchNorm-backward epilogue: own instantiation (pixel-linear 128-bit views, wide outputs)
    if (!a.bn_linear || ntile < 64) {
      set_error("fdg_conv2d[tcgen05]: the BatchNorm-backward epilogue of the per-tap kernel needs pixel-linear y / e views and Cout > 32");
      return FDG_ENOSUPPORT;
    }
    return ntile == 64 ? launch_umma<64, 2, 3, true>(a, st) : launch_umma<128, 2, 2, true>(a, st);
  }
  // 1x1 / stride 1 / direct gather over a pixel-linear view: the staging ring is fed by the bulk-tensor engine (FAST)
  static const int fast_on = [] { const char* e = getenv("FDG_CONV_FAST"); return e ? atoi(e) : 1; }();
  if (fast_on && ntile == 128 && p->R == 1 && p->S == 1 && p->stride == 1 && p->pad == 0 && p->gather == FDG_GATHER_DIRECT &&
      vec4_ok(p->x) && p->x.sh == (int64_t)p->W * p->x.sw && p->x.sn == (int64_t)p->H * p->x.sh && p->Cin % 8 == 0 &&
      (!p->has_affine || p->Cin <= UMAX_AFF) && a.M < (1ll << 31)) {
    const uint64_t dims[2] = {(uint64_t)p->Cin, (uint64_t)a.M};
    const uint64_t strides[1] = {(uint64_t)p->x.sw * 4};
    const uint32_t box[2] = {32, 128};
    if (make_tmap_f32(&a.xmap_hi, p->x.p, 2, dims, strides, box)) return launch_umma<128, 2, 2, false, true>(a, st);   // xmap_hi = the fp32 input view
  }
  if (ntile != 32 && ntile != 64 && ntile != 128) {
    set_error("fdg_conv2d[tcgen05]: the %d-wide weight image of this filter belongs to the halo-tile kernel, which does not take the shape", ntile);
    return FDG_ENOSUPPORT;
  }
  switch (ntile) {
    case 32: return launch_umma<32, 2, 3>(a, st);     // ring 2 x 40 KB + staging 3 x 32 KB
    case 64: return launch_umma<64, 2, 3>(a, st);     // ring 2 x 48 KB + staging 3 x 32 KB
    default: return launch_umma<128, 2, 2>(a, st);    // ring 2 x 64 KB + staging 2 x 32 KB
  }
}

}  // namespace fdg

using namespace fdg;

extern "C" int fdg_umma_ntile(int taps, int Cout) { return umma_ntile(taps, Cout); }
extern "C" int fdg_umma_tile_code(int taps, int Cin, int Cout) { return umma_ntile(taps, Cout) | (umma_tap_pair(taps, Cin) << 16); }

extern "C" int64_t fdg_umma_weight_bytes(int taps, int Cin, int Cout) {
  const int NT = umma_ntile(taps, Cout);
  return (int64_t)cdiv(Cout, NT) * umma_nchunks(taps, Cin) * 2 * NT * 128;
}

extern "C" int fdg_pack_weight_umma(const float* w, int w_ld, int taps, int Cin, int Cout, void* out, fdg_stream_t stream) {
  FDG_REQUIRE(w && out && taps > 0 && Cin > 0 && Cout > 0 && w_ld >= Cout, "fdg_pack_weight_umma: bad arguments");
  FDG_REQUIRE(aligned16(out), "fdg_pack_weight_umma: output must be 16-byte aligned");
  const int NT = umma_ntile(taps, Cout);
  const int cch = cdiv(Cin, UKC);
  const int pair = umma_tap_pair(taps, Cin);
  const int64_t total = (int64_t)cdiv(Cout, NT) * umma_nchunks(taps, Cin) * NT * 8;
  int64_t g = cdiv64(total, 256);
  if (g > 148 * 8) g = 148 * 8;
  pack_umma_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(w, w_ld, taps, Cin, Cout, NT, cch, (uint8_t*)out, total, pair);
  return check_launch("fdg_pack_weight_umma");
}
