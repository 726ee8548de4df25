// Fisher-vector product of layers >= 2 on the 5th-generation tensor cores (tcgen05 + TMEM), split-precision 3xTF32:
// R-forward (Pearlmutter), Fisher metric at the head, reverse sweep, hidden-layer weight / bias gradient partials and
// the delta_1 operand of the layer-1 gradient GEMM (trpo.py:45-58).  Replaces the warp-level mma.sync chain
// (mlp_chain.cu, chain_bwd_kernel<FVP>) for the Fvp, which was issue-bound on the legacy tensor path.
//
// One CTA (16 warps) walks 128-timestep tiles (M = 128 = the TMEM lanes) of its slab.  Every GEMM of the chain runs in
// TS mode: the A operand (activations of the 128 timesteps) is written into tensor memory by the epilogue warps that
// produced it, as (hi, lo) TF32 pairs, the B operand (weights, pre-split on the host side of the launch into the UMMA
// K-major core-matrix order) is streamed from L2 through a shared-memory ring by bulk async copies:
//
//   stage 0         Rz_2  = [Rh_1 | h_1] . [W_2 ; V_2]         Rh_1 = act'(h_1) (x.V_1 + vb_1),  x.V_1 from l1_forward_tc
//   stage l-2       Rz_l  = [Rh_{l-1} | h_{l-1}] . [W_l ; V_l]  ... l = 3..L
//   stage L-1       d_{L-1}' = delta_L . W_L^T                  delta_L = M (Rz_L + vb_L)   (Fisher metric of the head)
//   ...             d_{l-1}' = delta_l . W_l^T                  delta_l = d_l' * act'(h_l)
//   final           delta_1 -> DG (HBM, operand of l1_grad_tc_kernel) + layer-1 bias partials
//
// An epilogue thread owns ONE timestep (its TMEM lane) and all features of it, so head metrics need no shuffles.
// The A operand of a stage is handed over in slots of 16 features through a two-slot TMEM ring (one slot per epilogue
// warp group), so the MMAs of a stage start while its A operand is still being produced.
//
// Hidden-layer weight gradients G_l = h_{l-1}^T delta_l (K = timesteps) need the transposed orientation.  delta_l is
// written by the epilogue threads into shared memory as the (hi, lo) K-major B operand [n][timestep]; h_{l-1}^T goes
// to tensor memory through four converter warps whose lanes are FEATURE rows (they read the tile-major activation
// cache, where a feature's timesteps are contiguous); a second MMA-issuing warp accumulates the (k) tiles over the whole
// slab in TMEM.  A row of ones in each (k) tile yields the bias gradients.  Several layers share one 128-row tile when
// their rows fit (Humanoid: h_2 and h_3 against [delta_3 | delta_4]).
//
// Warp roles (512 threads): 0 producer of the weight ring (+ L2 prefetch of the next tile's activations), 1 chain MMA
// issuer (+ TMEM allocation), 2 (k) MMA issuer, 3 idle, 4-7 / 8-11 epilogue groups 0 / 1 (TMEM lane quarter = warp % 4),
// 12-15 converters.  All hand-overs are mbarriers with a spin watchdog: a protocol bug traps, it never hangs the GPU.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <string.h>

#define FT_THREADS 512
#define FT_WSTAGES 4          // weight ring stages
#define FT_KSLOTS 4           // (k) A-operand slots (8 timesteps x one tile = 16 TMEM columns each)
#define FT_MAX_STAGES 6
#define FT_MAX_PASS 3
#define FT_TMEM_COLS 512

struct FtStage {
  int rfwd;      // 1: Rz_l = [Rh_{l-1} | h_{l-1}] . [W_l ; V_l]     0: d_{l-1}' = delta_l . W_l^T
  int l;         // layer of W
  int ul;        // layer whose units form the A operand (l - 1 for R-forward, l for delta stages)
  int kgs;       // k-groups (8 features) of the A operand
  int N;         // MMA N (multiple of 16)
  int acc_col;   // TMEM column of the accumulator
  int w_off;     // float offset of the stage's B image in WC
  int v_off;     // float offset of the tangent image in VC (R-forward stages)
};
struct FtPlan {
  int L, S;                                  // layers, chain stages = 2 (L - 1)
  int Kg[MRL_MAX_LAYERS + 1], Np[MRL_MAX_LAYERS + 1];   // k-groups (round8 / 8) and MMA N (round16) of layers 1..L
  int vboff[MRL_MAX_LAYERS + 1];             // offset of layer l's tangent bias in the shared-memory copy
  FtStage st[FT_MAX_STAGES];
  int ring_col, kslot_col;                   // TMEM columns: chain A ring (2 x 64), (k) A slots (FT_KSLOTS x 16)
  int wstage_floats;                         // floats per weight ring stage
  int n_pass;                                // (k) tiles
  int pass_N[FT_MAX_PASS], pass_acc[FT_MAX_PASS], pass_buf[FT_MAX_PASS];   // N, TMEM column, float offset of the B buffer
  int pass_first_l[FT_MAX_PASS], pass_last_l[FT_MAX_PASS];                 // first / last produced delta layer (max / min l)
  int lay_pass[MRL_MAX_LAYERS + 1], lay_col[MRL_MAX_LAYERS + 1];           // layer l >= 2: its pass and column in the B tile
  short row_cache[FT_MAX_PASS][128];         // cache feature row of (k) tile row m; -1 none; -2 ones
  short row_lay[FT_MAX_PASS][128];           // layer l whose W_l gradient the row feeds
  short row_f[FT_MAX_PASS][128];             // in-feature index
  int wc_floats, vc_floats, kbuf_floats, smem_bytes;
};

// hi = rna_tf32(x) (integer rounding), lo = x - hi pre-rounded for the tensor core's truncation (mma_tf32.cuh)
__device__ __forceinline__ void ft_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void ft_rot8(uint32_t (&v)[8], int r) {   // v[i] <- v[(i + r) & 7]
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (r & 1) ? v[(i + 1) & 7] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (r & 2) ? a[(i + 2) & 7] : a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (r & 4) ? v[(i + 4) & 7] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = a[i];
}
__device__ __forceinline__ void ft_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// Register redistribution between the warp roles (all four warps of a warpgroup execute the same instruction): the
// launch gives every thread 128 registers (512 threads = the whole register file); the producer / MMA-issuer warps and
// the converters hand most of theirs back, the epilogue warps - which keep a slot of prefetched activations in
// registers to hide the L2 latency - take them.
#define FT_REGS_CTRL 40
#define FT_REGS_CONV 96
#define FT_REGS_EPI 176
template <int N> __device__ __forceinline__ void ft_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void ft_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
static_assert(4 * 32 * FT_REGS_CTRL + 4 * 32 * FT_REGS_CONV + 8 * 32 * FT_REGS_EPI <= 65536, "register file");

struct FtArgs {
  const float* WC;      // chain weight images of theta
  const float* VC;      // tangent weight images
  const float* vflat;   // tangent, flat (biases are read from it)
  const float* logstd;  // theta image logstd block (DiagGauss) or nullptr
  const float* Zt;      // x . V_1, tile-major [tile][d1][LDT]
  const float* cache;   // activations of theta, tile-major [tile][act_rows][LDT]
  float* DG;            // delta_1 operand of the layer-1 gradient GEMM
  float* partm;         // [n_slabs][pmid]
  float* dbg;           // debug dump of the first tile (nullptr: off)
  long long N;
  int n_tiles, n_mtiles, slab_mt, n_slabs, nu;
};

template <int ACT>
__global__ void __launch_bounds__(FT_THREADS, 1) fvp_tc_kernel(NetGeom g, FtPlan P, FtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* w_full = bars;                        // [FT_WSTAGES]
  uint64_t* w_empty = w_full + FT_WSTAGES;        // [FT_WSTAGES]
  uint64_t* a_full = w_empty + FT_WSTAGES;        // [2]
  uint64_t* a_empty = a_full + 2;                 // [2]
  uint64_t* acc_full = a_empty + 2;               // [FT_MAX_STAGES]
  uint64_t* kconv = acc_full + FT_MAX_STAGES;     // [FT_KSLOTS]
  uint64_t* kempty = kconv + FT_KSLOTS;           // [FT_KSLOTS]
  uint64_t* d_full = kempty + FT_KSLOTS;          // [FT_MAX_PASS]
  uint64_t* d_free = d_full + FT_MAX_PASS;        // [FT_MAX_PASS]
  uint64_t* gacc_full = d_free + FT_MAX_PASS;     // [1]
  uint64_t* gacc_empty = gacc_full + 1;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gacc_empty + 1);
  float* vb_s = reinterpret_cast<float*>(smem_raw + 512);       // tangent biases [vboff[l] + j], 512 floats
  float* ivar_s = vb_s + 512;                                   // 64 floats
  float* gb1s = ivar_s + 64;                                    // [8 epilogue warps][128] layer-1 bias partials
  float* wring = reinterpret_cast<float*>(smem_raw + 8192);     // [FT_WSTAGES][wstage_floats]
  float* kbuf = wring + (size_t)FT_WSTAGES * P.wstage_floats;   // (k) B operands: per pass [hi | lo][ngroup][32 k-chunks][8][4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = P.L, S = P.S;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FT_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FT_MAX_STAGES; ++i) mbar_init(&acc_full[i], 1);
    for (int i = 0; i < FT_KSLOTS; ++i) { mbar_init(&kconv[i], 4); mbar_init(&kempty[i], 1); }
    for (int i = 0; i < FT_MAX_PASS; ++i) { mbar_init(&d_full[i], 8); mbar_init(&d_free[i], 1); }
    mbar_init(gacc_full, 1);
    mbar_init(gacc_empty, 4);
    fence_barrier_init();
  }
  // tangent biases, 1/sigma^2, zeroed bias partials and (k) operand buffers (their padding columns stay zero)
  for (int i = threadIdx.x; i < 512; i += FT_THREADS) vb_s[i] = 0.f;
  for (int i = threadIdx.x; i < 8 * 128; i += FT_THREADS) gb1s[i] = 0.f;
  for (int i = threadIdx.x; i < P.kbuf_floats; i += FT_THREADS) kbuf[i] = 0.f;
  __syncthreads();
  for (int l = 1; l <= L; ++l)
    for (int j = threadIdx.x; j < g.d[l]; j += FT_THREADS) vb_s[P.vboff[l] + j] = a.vflat[g.off_flat_b[l] + j];
  for (int j = threadIdx.x; j < 64; j += FT_THREADS)
    ivar_s[j] = (g.head == MRL_HEAD_GAUSS && j < g.d[L]) ? expf(-2.f * a.logstd[j]) : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)FT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();     // the zeroed (k) buffers are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const size_t tile_c = (size_t)g.act_rows * MRL_LDT;   // floats per 64-timestep cache tile
  const size_t tile_z = (size_t)g.d[1] * MRL_LDT;

  if (warp < 4) {
  ft_reg_dec<FT_REGS_CTRL>();
  if (warp == 0) {
    // ================================================================ producer: weight ring + L2 prefetch
    if (lane == 0) {
      uint32_t wc = 0;
      for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
        const int mt0 = slab * a.slab_mt, mt1 = min(mt0 + a.slab_mt, a.n_mtiles);
        for (int mt = mt0; mt < mt1; ++mt) {
          if (mt + 1 < mt1) {   // next tile's activations and x.V_1 into L2 while this one computes
            const int t0 = 2 * (mt + 1), nt = min(2, a.n_tiles - t0);
            if (nt > 0) {
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.cache + (size_t)t0 * tile_c),
                           "r"((uint32_t)(nt * tile_c * 4)) : "memory");
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.Zt + (size_t)t0 * tile_z),
                           "r"((uint32_t)(nt * tile_z * 4)) : "memory");
            }
          }
          for (int s = 0; s < S; ++s) {
            const FtStage st = P.st[s];
            const int kgf = 2 * st.N * 8;                       // floats per k-group of one image
            for (int kg0 = 0; kg0 < st.kgs; kg0 += 2, ++wc) {
              const int nk = min(2, st.kgs - kg0);
              const int ws = wc % FT_WSTAGES;
              mbar_wait_sleep(&w_empty[ws], ((wc / FT_WSTAGES) & 1) ^ 1);
              float* dst = wring + (size_t)ws * P.wstage_floats;
              const uint32_t bytes = (uint32_t)(nk * kgf * 4);
              mbar_expect_tx(&w_full[ws], st.rfwd ? 2 * bytes : bytes);
              bulk_g2s(dst, a.WC + st.w_off + (size_t)kg0 * kgf, bytes, &w_full[ws]);
              if (st.rfwd) bulk_g2s(dst + 2 * kgf, a.VC + st.v_off + (size_t)kg0 * kgf, bytes, &w_full[ws]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ chain MMA issuer
    const uint32_t desc_hi = (uint32_t)(umma_desc(0, 0, 128) >> 32);
    const uint32_t wring_u32 = smem_u32(wring);
    uint32_t wc = 0, u = 0, tcount = 0;
    for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
      const int mt0 = slab * a.slab_mt, mt1 = min(mt0 + a.slab_mt, a.n_mtiles);
      for (int mt = mt0; mt < mt1; ++mt, ++tcount) {
        for (int s = 0; s < S; ++s) {
          const FtStage st = P.st[s];
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(st.N >> 3) << 17) | ((128u >> 4) << 24);
          const uint32_t lbo = (uint32_t)st.N * 16u;          // bytes between the two K halves of a B k-group
          const uint32_t kgb = (uint32_t)st.N * 64u;          // bytes per k-group of one image (hi | lo)
          const uint32_t d_tmem = tmem_base + st.acc_col;
          for (int kg0 = 0; kg0 < st.kgs; kg0 += 2, ++wc, ++u) {
            const int nk = min(2, st.kgs - kg0);
            const int ws = wc % FT_WSTAGES, e = u & 1;
            mbar_wait_sleep(&w_full[ws], (wc / FT_WSTAGES) & 1);
            mbar_wait_sleep(&a_full[e], (u >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t bw = umma_desc_lo(wring_u32 + (uint32_t)ws * (uint32_t)P.wstage_floats * 4u, lbo);
              const uint32_t ta = tmem_base + P.ring_col + 64 * e;
              for (int c = 0; c < nk; ++c) {
                const uint32_t first = (kg0 + c) ? 1u : 0u;
                if (st.rfwd) {
                  const uint32_t tr = ta + 32 * c, th = tr + 16;                 // [Rh hi 8 | Rh lo 8 | h hi 8 | h lo 8]
                  const uint32_t bW = bw + ((c * kgb) >> 4), bV = bw + ((2 * kgb + c * kgb) >> 4);
                  umma_tf32_ts(d_tmem, tr + 8, bW, desc_hi, idesc, first);               // Rh_lo . W_hi
                  umma_tf32_ts(d_tmem, tr, bW + (kgb >> 5), desc_hi, idesc, 1u);         // Rh_hi . W_lo
                  umma_tf32_ts(d_tmem, tr, bW, desc_hi, idesc, 1u);                      // Rh_hi . W_hi
                  umma_tf32_ts(d_tmem, th + 8, bV, desc_hi, idesc, 1u);                  // h_lo . V_hi
                  umma_tf32_ts(d_tmem, th, bV + (kgb >> 5), desc_hi, idesc, 1u);         // h_hi . V_lo
                  umma_tf32_ts(d_tmem, th, bV, desc_hi, idesc, 1u);                      // h_hi . V_hi
                } else {
                  const uint32_t td = ta + 16 * c;                                       // [delta hi 8 | delta lo 8]
                  const uint32_t bW = bw + ((c * kgb) >> 4);
                  umma_tf32_ts(d_tmem, td + 8, bW, desc_hi, idesc, first);
                  umma_tf32_ts(d_tmem, td, bW + (kgb >> 5), desc_hi, idesc, 1u);
                  umma_tf32_ts(d_tmem, td, bW, desc_hi, idesc, 1u);
                }
              }
              tc_commit(&a_empty[e]);
              tc_commit(&w_empty[ws]);
              if (kg0 + 2 >= st.kgs) tc_commit(&acc_full[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================================================ (k) MMA issuer: G tiles += h^T delta over the slab
    const uint32_t desc_hi = (uint32_t)(umma_desc(0, 0, 4096) >> 32);   // SBO = 32 k-chunks x 128 B between n-groups
    const uint32_t kbuf_u32 = smem_u32(kbuf);
    uint32_t kc = 0, tcount = 0, scount = 0;
    for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x, ++scount) {
      const int mt0 = slab * a.slab_mt, mt1 = min(mt0 + a.slab_mt, a.n_mtiles);
      mbar_wait_sleep(gacc_empty, (scount & 1) ^ 1);      // the previous slab's accumulators have been flushed
      tc_fence_after();
#ifdef FT_NO_K
      if (elect_one()) tc_commit(gacc_full);
      __syncwarp();
      continue;
#endif
      for (int mt = mt0; mt < mt1; ++mt, ++tcount) {
        for (int p = 0; p < P.n_pass; ++p) {
          const int Np_ = P.pass_N[p];
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np_ >> 3) << 17) | ((128u >> 4) << 24);
          const uint32_t d_tmem = tmem_base + P.pass_acc[p];
          const uint32_t bbase = umma_desc_lo(kbuf_u32 + (uint32_t)P.pass_buf[p] * 4u, 128);   // LBO = 128 B: adjacent k-chunks
          const uint32_t lo_off = ((uint32_t)Np_ * 512u) >> 4;                                  // lo half of the buffer
          mbar_wait_sleep(&d_full[p], tcount & 1);          // this tile's delta blocks of the pass are in shared memory
          for (int ks = 0; ks < 16; ++ks, ++kc) {
            const int sl = kc % FT_KSLOTS;
            mbar_wait_sleep(&kconv[sl], (kc / FT_KSLOTS) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t ta = tmem_base + P.kslot_col + 16 * sl;     // [h^T hi 8 | lo 8]
              const uint32_t db = bbase + ((uint32_t)ks * 256u >> 4);
              const uint32_t first = (mt > mt0 || ks > 0) ? 1u : 0u;
              umma_tf32_ts(d_tmem, ta + 8, db, desc_hi, idesc, first);            // h_lo . d_hi
              umma_tf32_ts(d_tmem, ta, db + lo_off, desc_hi, idesc, 1u);          // h_hi . d_lo
              umma_tf32_ts(d_tmem, ta, db, desc_hi, idesc, 1u);                   // h_hi . d_hi
              tc_commit(&kempty[sl]);
              if (ks == 15) tc_commit(&d_free[p]);
              if (ks == 15 && p == P.n_pass - 1 && mt == mt1 - 1) tc_commit(gacc_full);
            }
            __syncwarp();
          }
        }
      }
    }
  }
  } else if (warp >= 12) {
    // ================================================================ converters: h^T -> tensor memory, slab flush
    ft_reg_dec<FT_REGS_CONV>();
    const int q = warp & 3, m = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t kc = 0, scount = 0;
    for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x, ++scount) {
      const int mt0 = slab * a.slab_mt, mt1 = min(mt0 + a.slab_mt, a.n_mtiles);
#ifndef FT_NO_K
      for (int mt = mt0; mt < mt1; ++mt) {
        for (int p = 0; p < P.n_pass; ++p) {
          const int crow = P.row_cache[p][m];
          // the 16 k-steps (8 timesteps each) of the tile in groups of four; group g + 1 is requested before group g is
          // converted, so the L2 latency is covered by the conversion of four k-steps
          float4 xa[4][2], xb[4][2];
          auto request = [&](int g4, float4 (&x)[4][2]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int ks = 4 * g4 + i, t64 = 2 * mt + (ks >> 3);
              if (crow >= 0 && t64 < a.n_tiles) {
                const float4* src = reinterpret_cast<const float4*>(a.cache + (size_t)t64 * tile_c + (size_t)crow * MRL_LDT + (ks & 7) * 8);
                x[i][0] = __ldg(src);
                x[i][1] = __ldg(src + 1);
              } else {
                const float c = crow == -2 ? 1.f : 0.f;
                x[i][0] = make_float4(c, c, c, c);
                x[i][1] = x[i][0];
              }
            }
          };
          auto convert = [&](const float4 (&xx)[4][2]) {
#pragma unroll
            for (int i = 0; i < 4; ++i, ++kc) {
              const int sl = kc % FT_KSLOTS;
              const float x[8] = {xx[i][0].x, xx[i][0].y, xx[i][0].z, xx[i][0].w, xx[i][1].x, xx[i][1].y, xx[i][1].z, xx[i][1].w};
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) ft_split(x[k], hi[k], lo[k]);
              mbar_wait_sleep(&kempty[sl], ((kc / FT_KSLOTS) & 1) ^ 1);
              tc_fence_after();
              const uint32_t ta = tlane + P.kslot_col + 16 * sl;
              tmem_st8(ta, hi);
              tmem_st8(ta + 8, lo);
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&kconv[sl]);
            }
          };
          request(0, xa);
          request(1, xb);
          convert(xa);
          request(2, xa);
          convert(xb);
          request(3, xb);
          convert(xa);
          convert(xb);
        }
      }
#endif
      // ---- slab flush: accumulator rows -> fp32 slab partial (weights of layers >= 2, their biases from the ones row)
      mbar_wait_sleep(gacc_full, scount & 1);
      tc_fence_after();
      float* part = a.partm + (size_t)slab * g.pmid;
      for (int p = 0; p < P.n_pass; ++p) {
        const int crow = P.row_cache[p][m], rl = P.row_lay[p][m], rf = P.row_f[p][m];
        for (int c0 = 0; c0 < P.pass_N[p]; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tlane + P.pass_acc[p] + c0, v);
          if (crow == -1) continue;
          for (int l = P.pass_last_l[p]; l <= P.pass_first_l[p]; ++l) {
            if (crow >= 0 && l != rl) continue;
            const int col0 = P.lay_col[l];
            float* dst = crow >= 0 ? part + g.off_W[l] + (size_t)rf * g.ldw[l] : part + g.off_b[l];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = c0 + j - col0;
              if (n >= 0 && n < g.d[l]) dst[n] = __uint_as_float(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(gacc_empty);
    }
  } else {
    // ================================================================ epilogue groups: one thread = one timestep
    ft_reg_inc<FT_REGS_EPI>();
    const int e = (warp - 4) >> 2, q = warp & 3, m = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t ring = tlane + P.ring_col + 64 * e;
    float* gb1w = gb1s + ((warp - 4) * 128);
    const int rot = (m >> 2) & 7;
    uint32_t u = 0, ue = 0, tcount = 0;
    const bool cat = g.head == MRL_HEAD_CAT;
    for (int slab = blockIdx.x; slab < a.n_slabs; slab += gridDim.x) {
      const int mt0 = slab * a.slab_mt, mt1 = min(mt0 + a.slab_mt, a.n_mtiles);
      for (int mt = mt0; mt < mt1; ++mt, ++tcount) {
        const int t64 = 2 * mt + (m >> 6), r = m & 63;
        const bool ok = t64 < a.n_tiles;
        const long long T = (long long)mt * 128 + m;
        const bool valid = T < a.N;
        // Loads are unconditional with immediate offsets: a tile beyond the batch reads the last one (its rows are masked
        // at the head), features beyond a layer's width read the following rows (finite activations; the buffers have
        // one zeroed tile of slack) and meet zero weight rows / zero accumulator columns.
        const int t64c = min(t64, a.n_tiles - 1);
        const float* cb = a.cache + (size_t)t64c * tile_c + r;      // + feature row * LDT
        const float* zb = a.Zt + (size_t)t64c * tile_z + r;
        // ---- stages 0 .. S-1: produce the A operand of stage s in slots of 16 features.  The cached activations (and
        // x.V_1 in stage 0) of a group's NEXT slot are requested before its current slot is processed, and those of the
        // first slot before the wait for the previous stage's accumulator: the L2 latency hides behind the work.
        for (int s = 0; s < S; ++s) {
          const FtStage st = P.st[s];
          const int ul = st.ul;                      // layer whose units are produced
          const int du = g.d[ul];
          const float* hrow = cb + (size_t)g.off_act[ul] * MRL_LDT;
          const float* vbl = vb_s + P.vboff[ul];
          const bool is_head = !st.rfwd && ul == L;
          const bool need_h = !(is_head && !cat);    // the DiagGauss metric needs no cached row
          const int nslots = (st.kgs + 1) >> 1;
          int j = (e - (int)(u & 1)) & 1;            // this group's first slot of the stage
          u += nslots;
          float hc[16], zc[16], hn[16], zn[16];
          auto request = [&](int kg0, float (&hb)[16], float (&zb_)[16]) {
#ifdef FT_NO_LOADS
#pragma unroll
            for (int i = 0; i < 16; ++i) { hb[i] = 0.25f + kg0; zb_[i] = 0.5f; }
#else
            if (need_h) {
              const float* ph = hrow + (size_t)(8 * kg0) * MRL_LDT;
#pragma unroll
              for (int i = 0; i < 16; ++i) hb[i] = __ldg(ph + i * MRL_LDT);
            }
            if (s == 0) {
              const float* pz = zb + (size_t)(8 * kg0) * MRL_LDT;
#pragma unroll
              for (int i = 0; i < 16; ++i) zb_[i] = __ldg(pz + i * MRL_LDT);
            }
#endif
          };
          if (j < nslots) request(2 * j, hc, zc);
          uint32_t src_acc = 0;
          if (s > 0) {
            mbar_wait_sleep(&acc_full[s - 1], tcount & 1);
            tc_fence_after();
            src_acc = tlane + P.st[s - 1].acc_col;
          }
          float sdot = 0.f;     // Categorical: p . Rz over the whole row
          if (is_head && cat) {
            for (int c0 = 0; c0 < P.Np[L]; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(src_acc + c0, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int f = c0 + i;
                const float p = (ok && f < du) ? __ldg(hrow + (size_t)f * MRL_LDT) : 0.f;
                sdot += p * (__uint_as_float(v[i]) + vbl[f]);
              }
            }
          }
          int pass = -1;
          float* kb = nullptr;
          if (!st.rfwd) {
            pass = P.lay_pass[ul];
            kb = kbuf + P.pass_buf[pass] + (size_t)(P.lay_col[ul] >> 3) * 1024 + (m >> 2) * 32 + (m & 3);
#ifndef FT_NO_K
            if (ul == P.pass_first_l[pass]) mbar_wait_sleep(&d_free[pass], (tcount & 1) ^ 1);   // previous tile's (k) MMAs are done
#endif
          }
          for (; j < nslots; j += 2) {
            const int kg0 = 2 * j;
            const int nk = min(2, st.kgs - kg0);
            if (j + 2 < nslots) request(kg0 + 4, hn, zn);
            // unit values of features 8 kg0 .. 8 kg0 + 15
            float val[16];
            if (s == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) val[i] = dact_from_h<ACT>(hc[i]) * (zc[i] + vbl[8 * kg0 + i]);
            } else {
              uint32_t v[16];
              tmem_ld16(src_acc + 8 * kg0, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int f = 8 * kg0 + i;
                const float acc = __uint_as_float(v[i]);
                float o;
                if (st.rfwd) o = dact_from_h<ACT>(hc[i]) * (acc + vbl[f]);                     // Rh_ul
                else if (is_head) {
                  const float rz = acc + vbl[f];
                  o = cat ? hc[i] * (rz - sdot) : rz * ivar_s[f < 64 ? f : 63];               // Fisher metric
                  if (!valid) o = 0.f;
                  if (f >= du) o = 0.f;
                } else o = acc * dact_from_h<ACT>(hc[i]);                                      // delta_ul
                val[i] = o;                        // padding features: zero accumulator columns and tangent biases
              }
            }
            if (a.dbg && mt == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < 8 * nk) a.dbg[((size_t)s * 128 + m) * 128 + 8 * kg0 + i] = val[i];
            }
            mbar_wait_sleep(&a_empty[e], (ue & 1) ^ 1);
            tc_fence_after();
            ++ue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c < nk) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) ft_split(val[8 * c + k], hi[k], lo[k]);
                if (st.rfwd) {
                  tmem_st8(ring + 32 * c, hi);
                  tmem_st8(ring + 32 * c + 8, lo);
                  uint32_t h2[8], l2[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) ft_split(hc[8 * c + k], h2[k], l2[k]);
                  tmem_st8(ring + 32 * c + 16, h2);
                  tmem_st8(ring + 32 * c + 24, l2);
                } else {
                  tmem_st8(ring + 16 * c, hi);
                  tmem_st8(ring + 16 * c + 8, lo);
                  // the same delta block as the K-major B operand of the (k) GEMM: [n][timestep], rotated so that the
                  // 32 timesteps of a warp hit 32 distinct banks
#ifndef FT_NO_KBUF
                  ft_rot8(hi, rot);
                  ft_rot8(lo, rot);
                  float* kp = kb + (size_t)(kg0 + c) * 1024;
                  const int lo_off = P.pass_N[pass] * 128;
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int nn = ((i + rot) & 7) * 4;
                    kp[nn] = __uint_as_float(hi[i]);
                    kp[nn + lo_off] = __uint_as_float(lo[i]);
                  }
#endif
                }
              }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[e]);
#pragma unroll
            for (int i = 0; i < 16; ++i) { hc[i] = hn[i]; zc[i] = zn[i]; }
          }
          if (!st.rfwd && ul == P.pass_last_l[pass]) {     // all delta blocks of the pass are written
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_full[pass]);
          }
        }
        // ---- final: delta_1 = d_1' * act'(h_1) -> DG (tcgen05 B operand of the layer-1 gradient) + bias partial sums
        {
          const float* hrow = cb + (size_t)g.off_act[1] * MRL_LDT;
          const int d1 = g.d[1], nu = a.nu;
          const int nch = nu >> 4;
          float hc[16], hn[16];
          auto request = [&](int c0, float (&hb)[16]) {
            const float* ph = hrow + (size_t)c0 * MRL_LDT;
#pragma unroll
            for (int i = 0; i < 16; ++i) hb[i] = __ldg(ph + i * MRL_LDT);
          };
          int jj = e;
          if (jj < nch) request(16 * jj, hc);
          mbar_wait_sleep(&acc_full[S - 1], tcount & 1);
          tc_fence_after();
          const uint32_t src_acc = tlane + P.st[S - 1].acc_col;
          float* dgp = a.DG + (size_t)(T >> 3) * (2 * nu * 8) + ((T >> 2) & 1) * (nu * 4) + (T & 3);
          for (; jj < nch; jj += 2) {
            const int c0 = 16 * jj;
            if (jj + 2 < nch) request(c0 + 32, hn);
            uint32_t v[16];
            tmem_ld16(src_acc + c0, v);
            float val[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              val[i] = __uint_as_float(v[i]) * dact_from_h<ACT>(hc[i]);     // padding columns of the accumulator are zero
#ifndef FT_NO_DG
              if (ok) {
                uint32_t hi, lo;
                ft_split(val[i], hi, lo);
                float* p = dgp + ((c0 + i) >> 3) * 32 + ((c0 + i) & 7) * 4;
                p[0] = __uint_as_float(hi);
                p[nu * 8] = __uint_as_float(lo);
              }
#endif
            }
            if (a.dbg && mt == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) a.dbg[((size_t)S * 128 + m) * 128 + c0 + i] = val[i];
            }
            // column sums over the warp's 32 timesteps: 16 -> 8 -> 4 -> 2 -> 1 values per lane, then the pair
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float send = (lane & 16) ? val[i] : val[i + 8];
              const float keep = (lane & 16) ? val[i + 8] : val[i];
              val[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float send = (lane & 8) ? val[i] : val[i + 4];
              const float keep = (lane & 8) ? val[i + 4] : val[i];
              val[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float send = (lane & 4) ? val[i] : val[i + 2];
              const float keep = (lane & 4) ? val[i + 2] : val[i];
              val[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            {
              const float send = (lane & 2) ? val[0] : val[1];
              const float keep = (lane & 2) ? val[1] : val[0];
              val[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            val[0] += __shfl_xor_sync(0xffffffffu, val[0], 1);
            // lane holds column c0 + 8 b4 + 4 b3 + 2 b2 + b1 (b4 = bit 4 of the lane, ...)
            if ((lane & 1) == 0) {
              const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
              gb1w[col] += val[0];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) hc[i] = hn[i];
          }
          tc_fence_before();
        }
      }
      // ---- slab end: layer-1 bias partial (fixed order over the 8 epilogue warps), logstd block = 0 (set by the reduce)
      ft_epi_bar();
      float* part = a.partm + (size_t)slab * g.pmid;
      const int et = threadIdx.x - 128;              // 0..255 over the epilogue warps
      for (int f = et; f < g.d[1]; f += 256) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { sum += gb1s[w * 128 + f]; }
        part[g.off_b[1] + f] = sum;
      }
      for (int j = et; j < g.d[L]; j += 256) part[g.off_pm_logstd + j] = 0.f;
      ft_epi_bar();
      for (int f = et; f < 8 * 128; f += 256) gb1s[f] = 0.f;
      ft_epi_bar();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)FT_TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------ operand images
// src (flat parameter or tangent vector) -> the B images of the chain stages: per stage [k-group][hi | lo][khalf][n-group
// N/8][8 n][4 k]; R-forward stages hold W_l as B[n = out][k = in], delta stages W_l^T as B[n = in][k = out].
struct FtPackJob { int off, N, kgs, l, transposed, end; };
struct FtPackJobs { FtPackJob j[FT_MAX_STAGES]; int n; };
__global__ void ft_pack_kernel(NetGeom g, FtPackJobs jobs, const float* __restrict__ src, float* __restrict__ dst, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int q = 0, base = 0;
  while (q < jobs.n - 1 && i >= jobs.j[q].end) { base = jobs.j[q].end; ++q; }
  const FtPackJob jb = jobs.j[q];
  const int e = i - base, K = jb.kgs * 8;
  const int n = e / K, k = e % K;
  const int l = jb.l;
  float x = 0.f;
  if (!jb.transposed) {          // n = out unit, k = in unit
    if (n < g.d[l] && k < g.d[l - 1]) x = src[g.off_flat_W[l] + k * g.d[l] + n];
  } else {                       // n = in unit, k = out unit
    if (n < g.d[l - 1] && k < g.d[l]) x = src[g.off_flat_W[l] + n * g.d[l] + k];
  }
  const float h = tf32_rna(x);
  float* p = dst + jb.off + (size_t)(k >> 3) * (2 * jb.N * 8) + ((k & 7) >> 2) * (jb.N * 4) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);
  p[0] = h;
  p[jb.N * 8] = tf32_rna(x - h);
}

// ------------------------------------------------------------------------------------ host side
static bool ft_build_plan(const NetGeom& g, FtPlan* P) {
  memset(P, 0, sizeof(*P));
  const int L = g.L;
  if (L < 3 || L > 4) return false;
  if (g.head != MRL_HEAD_GAUSS && g.head != MRL_HEAD_CAT) return false;
  if (g.d[L] > 64) return false;
  P->L = L;
  P->S = 2 * (L - 1);
  int vb = 0;
  for (int l = 1; l <= L; ++l) {
    if (g.d[l] > 128) return false;
    P->Kg[l] = (g.d[l] + 7) / 8;
    P->Np[l] = round_up(g.d[l], 16);
    P->vboff[l] = vb;
    vb += round_up(g.d[l], 16) + 16;      // the epilogue reads up to 16 k-group-padded entries past d[l]
  }
  if (vb > 512) return false;
  if (P->Np[1] != l1tc_nu(g)) return false;
  // chain stages and their B images
  int s = 0, wc = 0, vc = 0, maxstage = 0;
  for (int l = 2; l <= L; ++l, ++s) {
    FtStage& st = P->st[s];
    st.rfwd = 1; st.l = l; st.ul = l - 1; st.kgs = P->Kg[l - 1]; st.N = P->Np[l];
    st.w_off = wc; st.v_off = vc;
    wc += st.kgs * 2 * st.N * 8;
    vc += st.kgs * 2 * st.N * 8;
    maxstage = maxstage > 2 * 2 * 2 * st.N * 8 ? maxstage : 2 * 2 * 2 * st.N * 8;   // 2 k-groups x (W, V) x (hi, lo)
  }
  for (int l = L; l >= 2; --l, ++s) {
    FtStage& st = P->st[s];
    st.rfwd = 0; st.l = l; st.ul = l; st.kgs = P->Kg[l]; st.N = P->Np[l - 1];
    st.w_off = wc; st.v_off = -1;
    wc += st.kgs * 2 * st.N * 8;
    maxstage = maxstage > 2 * 2 * st.N * 8 ? maxstage : 2 * 2 * st.N * 8;
  }
  P->wc_floats = wc;
  P->vc_floats = vc;
  P->wstage_floats = maxstage;
  // (k) tiles: layers are produced in the order delta_L, ..., delta_2; a tile takes consecutive layers while their input
  // rows plus one row of ones fit into 128 lanes
  int np = 0;
  for (int l = L; l >= 2;) {
    if (np == FT_MAX_PASS) return false;
    int rows = 0, cols = 0, lo_l = l;
    for (int k = l; k >= 2; --k) {
      if (rows + g.d[k - 1] + 1 > 128 || cols + P->Np[k] > 128) break;
      rows += g.d[k - 1];
      cols += P->Np[k];
      lo_l = k;
    }
    if (rows == 0) return false;
    P->pass_first_l[np] = l;
    P->pass_last_l[np] = lo_l;
    P->pass_N[np] = cols;
    for (int m = 0; m < 128; ++m) { P->row_cache[np][m] = -1; P->row_lay[np][m] = 0; P->row_f[np][m] = 0; }
    int m = 0, col = 0;
    for (int k = l; k >= lo_l; --k) {
      P->lay_pass[k] = np;
      P->lay_col[k] = col;
      col += P->Np[k];
      for (int f = 0; f < g.d[k - 1]; ++f, ++m) {
        P->row_cache[np][m] = (short)(g.off_act[k - 1] + f);
        P->row_lay[np][m] = (short)k;
        P->row_f[np][m] = (short)f;
      }
    }
    P->row_cache[np][m] = -2;   // ones
    ++np;
    l = lo_l - 1;
  }
  P->n_pass = np;
  // tensor memory: (k) accumulators | chain accumulators (two alternating regions) | chain A ring | (k) A slots
  int col = 0;
  for (int p = 0; p < np; ++p) { P->pass_acc[p] = col; col += P->pass_N[p]; }
  int reg[2] = {0, 0};
  for (int i = 0; i < P->S; ++i) reg[i & 1] = reg[i & 1] > P->st[i].N ? reg[i & 1] : P->st[i].N;
  for (int i = 0; i < P->S; ++i) P->st[i].acc_col = col + ((i & 1) ? reg[0] : 0);
  col += reg[0] + reg[1];
  P->ring_col = col;
  col += 128;
  P->kslot_col = col;
  col += 16 * FT_KSLOTS;
  if (col > FT_TMEM_COLS) return false;
  // shared memory: barriers + small (8 KB) | weight ring | (k) B buffers
  int kb = 0;
  for (int p = 0; p < np; ++p) { P->pass_buf[p] = kb; kb += 2 * P->pass_N[p] * 128; }
  P->kbuf_floats = kb;
  P->smem_bytes = 8192 + (FT_WSTAGES * P->wstage_floats + kb) * 4;
  if (P->smem_bytes > 227 * 1024) return false;
  return true;
}

bool fvp_tc_supported(const NetGeom& g) {
  FtPlan P;
  return g.act == MRL_ACT_TANH && ft_build_plan(g, &P);
}
size_t fvp_tc_image_floats(const NetGeom& g, int tangent) {
  FtPlan P;
  if (!ft_build_plan(g, &P)) return 0;
  return tangent ? P.vc_floats : P.wc_floats;
}
cudaError_t launch_fvp_tc_pack(const NetGeom& g, const float* src_flat, float* dst, int tangent, cudaStream_t st) {
  FtPlan P;
  if (!ft_build_plan(g, &P)) return cudaErrorInvalidConfiguration;
  FtPackJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  int total = 0;
  for (int s = 0; s < P.S; ++s) {
    const FtStage& sg = P.st[s];
    if (tangent && !sg.rfwd) continue;
    FtPackJob& jb = jobs.j[jobs.n++];
    jb.off = tangent ? sg.v_off : sg.w_off;
    jb.N = sg.N;
    jb.kgs = sg.kgs;
    jb.l = sg.l;
    jb.transposed = sg.rfwd ? 0 : 1;
    total += sg.N * sg.kgs * 8;
    jb.end = total;
  }
  ft_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(g, jobs, src_flat, dst, total);
  return cudaGetLastError();
}

cudaError_t launch_fvp_tc(const NetGeom& g, const FvpTcArgs& x, cudaStream_t st) {
  FtPlan P;
  if (g.act != MRL_ACT_TANH || !ft_build_plan(g, &P)) return cudaErrorInvalidConfiguration;
  if (x.slab_tiles % 2) return cudaErrorInvalidValue;
  FtArgs a;
  a.WC = x.WC; a.VC = x.VC; a.vflat = x.vflat;
  a.logstd = g.head == MRL_HEAD_GAUSS ? x.img + g.off_pm_logstd : nullptr;
  a.Zt = x.Zt; a.cache = x.cache; a.DG = x.DG; a.partm = x.partm; a.dbg = x.dbg;
  a.N = x.N;
  a.n_tiles = x.n_tiles;
  a.n_mtiles = (x.n_tiles + 1) / 2;
  a.slab_mt = x.slab_tiles / 2;
  a.n_slabs = x.n_slabs;
  a.nu = l1tc_nu(g);
  cudaError_t e = mrl_func_smem((const void*)fvp_tc_kernel<MRL_ACT_TANH>, P.smem_bytes);
  if (e != cudaSuccess) return e;
  const int sms = mrl_sm_count();
  fvp_tc_kernel<MRL_ACT_TANH><<<x.n_slabs < sms ? x.n_slabs : sms, FT_THREADS, P.smem_bytes, st>>>(g, P, a);
  return cudaGetLastError();
}
