// G2 — weight-gradient GEMM: the contraction over PIXELS.
//
//   P[split][tap][m][n] = sum_{pixels p in split} A[p, m] * B[p (+) tap, n]        (fp32)
//
//   conv3x3 wgrad : A = dY (m = out channel), B = layer input shifted by the tap (n = in channel)
//   convT   wgrad : A = layer input (m = in channel), B = 2x2-gathered dOut (n = out channel)
//   plain         : one tap, no shift (the im2col'ed first layer)
//
// Both operands are read straight from NHWC bf16 tensors by TMA.  Because the reduction
// dimension (pixels) is the slow axis of NHWC, both are "MN-major" tcgen05 operands: one
// 128-byte swizzled row per pixel holding 64 channels.  Each CTA owns a 128(m) x BNW(n) block
// of up to four taps (T*BNW <= 512 TMEM columns) and a slice of the pixel range (split-K);
// partials are written as plain fp32 and summed deterministically by wgrad_reduce (no atomics).
//
// Replaces: the weight-gradient half of conv2d / conv_transpose2d backward as dispatched by
// loss.backward() in /root/reference/train.py:208 (cuDNN wgrad in the reference).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

// Pixels per pipeline stage (kp) is chosen per launch (32, 64 or 128) so that one stage carries
// several hundred tensor-core cycles of work: a stage hand-off costs the single issuing thread an
// mbarrier round trip, which a 32-pixel stage of a 64-channel layer (192 cycles of MMA) cannot hide.
constexpr int WG_THREADS = 192;

struct WgradParams {
  int mode;        // SUNET_A_CONV3X3 / SUNET_A_PLAIN / SUNET_A_GATHER2X2 (how B is addressed)
  int T;           // taps handled per CTA (3, 1 or 4)
  int tap_groups;  // 3, 1, 1
  int nb64;        // BNW / 64
  int m_tiles, n_tiles, splits;
  int kb_total, kb_per_split;
  int tw, th, nb, tiles_x, tiles_y;
  int H;
  int c0_blocks;   // n tiles that come from B source 0 (rest from source 1)
  int Ca, Nb;      // real A channels (rows written), total B channels
  int taps_total;
  int tmem_cols;   // power of two >= T * BNW
  int dbg_shift, dbg_boff;  // SUNET_DBG_SHIFT / SUNET_DBG_BOFF: descriptor-swizzle experiment
  int shifted;     // conv3x3 with W % 32 == 0: ONE (KP+2)-pixel B box per 64 channels serves the 3 column taps
                   // as shifted descriptor windows (tcgen05 swizzles on absolute smem address bits)
  int bslot;       // bytes reserved per B box in a stage (1024-aligned)
  int b_tx;        // bytes one B box delivers
  int bboxes;      // B boxes per stage
  int stages;
  int kp;          // pixels per stage
  int a_reuse;     // keep A in the tensor core's collector across the taps of a K step (SUNET_WGRAD_A_REUSE, default 1)
  int abox;        // bytes of one A box = kp * 128
  float* out;      // [splits][taps_total][Ca][Nb]
  // training prologue on B (CTA-pair kernel, shifted-window box only): B holds the producer block's raw conv output y;
  // the four epilogue warps, idle during the main loop, turn every landed box into relu(y * scale + shift) in place
  const float* b_pro_scale;
  const float* b_pro_shift;
  int W;           // image width (p.H is the height)
};
constexpr int MAX_STAGES = 12;
constexpr int WG_BAR_BYTES = (4 * MAX_STAGES + 2) * 8;   // full, empty, done, tmem slot, (pair + prologue) blocal, bready

// MMA issue for one pipeline stage, taps unrolled at compile time.  The generic (runtime T) loop spent ~25 SASS
// instructions of the single issuing thread per UTCHMMA — more than the 64 tensor-core cycles an N=128 MMA takes,
// which capped these kernels at ~67 % tensor-pipe utilisation (profiles/r01/ncu_full_r01g.md).
// The T taps of one K step multiply the SAME A tile (dY, or the ConvT input) with shifted B windows: the first MMA
// fills the tensor core's A collector, the others reuse it (UTCHMMA .A_KEEP / .A_REUSE), so A crosses the shared-memory
// port once per K step instead of T times — these kernels sit on that port (profiles/r01/ncu_full_r01g.md:
// smem->tensor wavefronts 64-69 % with the tensor pipe at 33-58 %).  REUSE = false restores the plain form (A/B timing).
template <int T, bool PAIR, bool REUSE>
__device__ __forceinline__ void wg_issue_taps(uint32_t tmem_base, uint32_t bnw, uint64_t adesc, uint64_t bdesc,
                                              uint32_t tstep16, uint32_t idesc, uint32_t acc) {
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const uint32_t d = tmem_base + t * bnw;
    const uint64_t b = bdesc + t * tstep16;
    if (!REUSE || T == 1) umma_bf16_col<0, PAIR>(d, adesc, b, idesc, acc);
    else if (t == 0) umma_bf16_col<1, PAIR>(d, adesc, b, idesc, acc);
    else if (t == T - 1) umma_bf16_col<3, PAIR>(d, adesc, b, idesc, acc);
    else umma_bf16_col<2, PAIR>(d, adesc, b, idesc, acc);
  }
}
template <int T, bool PAIR, bool REUSE>
__device__ __forceinline__ void wg_issue_stage(uint32_t tmem_base, uint32_t bnw, uint64_t adesc, uint64_t bdesc,
                                               uint32_t tstep16, uint32_t idesc, int nkk, uint32_t acc_first) {
  wg_issue_taps<T, PAIR, REUSE>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, acc_first);
#pragma unroll 1
  for (int kk = 1; kk < nkk; ++kk) {
    adesc += 2048 >> 4;               // next 16 pixels (16 rows of 128 B)
    bdesc += 2048 >> 4;
    wg_issue_taps<T, PAIR, REUSE>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, 1u);
  }
}
template <bool PAIR>
__device__ __forceinline__ void wg_issue_stage_T(int T, bool reuse, uint32_t tmem_base, uint32_t bnw, uint64_t adesc,
                                                 uint64_t bdesc, uint32_t tstep16, uint32_t idesc, int nkk,
                                                 uint32_t acc_first) {
  if (reuse) {
    if (T == 3) wg_issue_stage<3, PAIR, true>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
    else if (T == 4) wg_issue_stage<4, PAIR, true>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
    else wg_issue_stage<1, PAIR, false>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
  } else {
    if (T == 3) wg_issue_stage<3, PAIR, false>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
    else if (T == 4) wg_issue_stage<4, PAIR, false>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
    else wg_issue_stage<1, PAIR, false>(tmem_base, bnw, adesc, bdesc, tstep16, idesc, nkk, acc_first);
  }
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0,
                  const __grid_constant__ CUtensorMap mapB1, const WgradParams p, const int stage_bytes) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* done_bar = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // work item
  int item = blockIdx.x;
  const int n_tile = item % p.n_tiles; item /= p.n_tiles;
  const int m_tile = item % p.m_tiles; item /= p.m_tiles;
  const int tg = item % p.tap_groups;  item /= p.tap_groups;
  const int split = item;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int BNW = p.nb64 * 64;

  if (warp == 0) {
    {
      const CUtensorMap* mapB = (n_tile < p.c0_blocks) ? &mapB0 : &mapB1;
      const int cB = ((n_tile < p.c0_blocks) ? n_tile : (n_tile - p.c0_blocks)) * BNW;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int xt = kb % p.tiles_x;
        const int yt = (kb / p.tiles_x) % p.tiles_y;
        const int nt = kb / (p.tiles_x * p.tiles_y);
        const int x0 = xt * p.tw, y0 = yt * p.th, n0 = nt * p.nb;
        uint8_t* st = smem + stage * stage_bytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if ((p.dbg_boff & 16) && kb >= kb0 + STAGES) {   // timing experiment: no TMA traffic after the first fill
          if (elect_one()) mbar_arrive(&full_bar[stage]);
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        uint8_t* sb = st + 2 * p.abox;
        if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], 2 * p.abox + p.bboxes * p.b_tx);
        // A: two 64-channel boxes (the second may be fully out of range -> zeros)
        tma_load_5d(st, &mapA, &full_bar[stage], m_tile * 128, x0, y0, n0, 0);
        tma_load_5d(st + p.abox, &mapA, &full_bar[stage], m_tile * 128 + 64, x0, y0, n0, 0);
        if (p.shifted) {
          // one box of KP+2 pixels starting one pixel to the left, on input row y + (filter row - 1)
          for (int j = 0; j < p.nb64; ++j)
            tma_load_5d(sb + j * p.bslot, mapB, &full_bar[stage], cB + j * 64, x0 - 1, y0 + tg - 1, n0, 0);
        } else
        for (int t = 0; t < p.T; ++t) {
          for (int j = 0; j < p.nb64; ++j) {
            uint8_t* dst = sb + (t * p.nb64 + j) * p.bslot;
            const int c = cB + j * 64;
            if (p.mode == SUNET_A_GATHER2X2) {
              tma_load_5d(dst, mapB, &full_bar[stage], c, t & 1, x0, t >> 1, n0 * p.H + y0);
            } else if (p.mode == SUNET_A_CONV3X3) {
              tma_load_5d(dst, mapB, &full_bar[stage], c, x0 + t - 1, y0 + tg - 1, n0, 0);
            } else {
              tma_load_5d(dst, mapB, &full_bar[stage], c, x0, y0, n0, 0);
            }
          }
        }
        }  // elect_one
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      // dbg_boff bit 3 (SUNET_DBG_BOFF=8): timing-only experiment, pretend the operands are K-major
      const bool mn = !(p.dbg_boff & 8);
      const uint32_t idesc = make_idesc_bf16(128, BNW, mn, mn);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + 2 * p.abox;
        const uint32_t tstep = p.shifted ? 128u : (uint32_t)(p.nb64 * p.bslot);   // smem distance between taps
        const int nkk = (p.dbg_boff & 32) ? 0 : p.kp / 16;                        // bit 5: timing experiment, no MMAs
        if (elect_one()) {
          uint64_t adesc = make_smem_desc_sw128(sa, p.abox, 1024);
          uint64_t bdesc0 = make_smem_desc_sw128(sb + p.dbg_shift * 128, p.bslot, 1024);
          bdesc0 |= static_cast<uint64_t>(p.dbg_boff & 7) << 49;                  // experiment hook, 0 in production
          // taps: shifted windows of one box (or consecutive box groups), tstep bytes apart
          if (nkk > 0)
            wg_issue_stage_T<false>(p.T, p.a_reuse != 0, tmem_base, (uint32_t)BNW, adesc, bdesc0, tstep >> 4, idesc,
                                    nkk, kb > kb0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(done_bar);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int mrow = m_tile * 128 + quad * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    for (int t = 0; t < p.T; ++t) {
      const int tap = tg * p.T + t;
      float* orow = p.out + ((static_cast<size_t>(split) * p.taps_total + tap) * p.Ca + mrow) * p.Nb + n_tile * BNW;
      for (int c = 0; c < BNW; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * BNW + c, v);
        tmem_ld_wait();
        if (mrow < p.Ca) {
          float4* o4 = reinterpret_cast<float4*>(orow + c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// ------------------------------------------------------------------ CTA-pair variant (A channels >= 256)
// tcgen05 cta_group::2: one M=256 MMA spans two CTAs; each loads its own 128 dY channels but only HALF
// of the B channels (64 of the 128-wide tile), so B TMA traffic and per-SM operand reads drop by a third.
// Launched as a 2-CTA cluster; barriers the leader's MMA warp waits on live in the leader CTA.
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB0,
                       const __grid_constant__ CUtensorMap mapB1, const WgradParams p, const int stage_bytes) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* done_bar = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);
  uint64_t* blocal = bars + 2 * MAX_STAGES + 2;       // prologue: this CTA's own B box has landed
  uint64_t* bready = blocal + MAX_STAGES;             // prologue (leader's copy): both CTAs' boxes are transformed
  const bool pro = p.b_pro_scale != nullptr;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);      // leader's arrive.expect_tx covers both CTAs' bytes
      mbar_init(&empty_bar[i], 1);
      mbar_init(&blocal[i], 1);
      mbar_init(&bready[i], 2);        // one arrival per CTA of the pair
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB0);
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_alloc_pair(tmem_slot, p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  int item = blockIdx.x >> 1;
  const int n_tile = item % p.n_tiles; item /= p.n_tiles;
  const int m_pair = item % p.m_tiles; item /= p.m_tiles;      // m_tiles counts 256-channel pairs here
  const int tg = item % p.tap_groups;  item /= p.tap_groups;
  const int split = item;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int BNW = p.nb64 * 64;                                   // 128: N of the pair MMA
  const int a_base = m_pair * 256 + (int)rank * 128;

  if (warp == 0) {
    const CUtensorMap* mapB = (n_tile < p.c0_blocks) ? &mapB0 : &mapB1;
    const int cB = ((n_tile < p.c0_blocks) ? n_tile : (n_tile - p.c0_blocks)) * BNW + (int)rank * 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      const int xt = kb % p.tiles_x;
      const int yt = (kb / p.tiles_x) % p.tiles_y;
      const int nt = kb / (p.tiles_x * p.tiles_y);
      const int x0 = xt * p.tw, y0 = yt * p.th, n0 = nt * p.nb;
      uint8_t* st = smem + stage * stage_bytes;
      uint8_t* sb = st + 2 * p.abox;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        if (rank == 0)
          mbar_arrive_expect_tx(&full_bar[stage], pro ? 2 * (2 * p.abox) : 2 * (2 * p.abox + p.bboxes * p.b_tx));
        tma_load_5d_pair(st, &mapA, &full_bar[stage], a_base, x0, y0, n0, 0);
        tma_load_5d_pair(st + p.abox, &mapA, &full_bar[stage], a_base + 64, x0, y0, n0, 0);
        if (pro) {
          // the B box completes on this CTA's own barrier: its transform warps take it from there (shifted box only)
          mbar_arrive_expect_tx(&blocal[stage], p.b_tx);
          tma_load_5d(sb, mapB, &blocal[stage], cB, x0 - 1, y0 + tg - 1, n0, 0);
        } else if (p.shifted) {
          tma_load_5d_pair(sb, mapB, &full_bar[stage], cB, x0 - 1, y0 + tg - 1, n0, 0);
        } else {
          for (int t = 0; t < p.T; ++t) {
            uint8_t* dst = sb + t * p.bslot;
            if (p.mode == SUNET_A_GATHER2X2) {
              tma_load_5d_pair(dst, mapB, &full_bar[stage], cB, t & 1, x0, t >> 1, n0 * p.H + y0);
            } else if (p.mode == SUNET_A_CONV3X3) {
              tma_load_5d_pair(dst, mapB, &full_bar[stage], cB, x0 + t - 1, y0 + tg - 1, n0, 0);
            } else {
              tma_load_5d_pair(dst, mapB, &full_bar[stage], cB, x0, y0, n0, 0);
            }
          }
        }
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, BNW, true, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (pro) mbar_wait(&bready[stage], phase);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + 2 * p.abox;
        const uint32_t tstep = p.shifted ? 128u : (uint32_t)p.bslot;
        const int nkk = p.kp / 16;
        if (elect_one()) {
          uint64_t adesc = make_smem_desc_sw128(sa, p.abox, 1024);
          uint64_t bdesc0 = make_smem_desc_sw128(sb, p.bslot, 1024);
          wg_issue_stage_T<true>(p.T, p.a_reuse != 0, tmem_base, (uint32_t)BNW, adesc, bdesc0, tstep >> 4, idesc, nkk,
                                 kb > kb0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit_pair(done_bar);
      __syncwarp();
    }
  } else {
    if (pro) {
      // -------- prologue transform of the B boxes (both CTAs; these four warps have nothing else to do until the
      // accumulators are complete): same fmaf / max / round as bn_relu_flat_kernel; pixels outside the image are TMA
      // zero fill (the conv's padding) and stay zero.  Each WARP owns every fourth stage, so four boxes are in
      // flight per CTA and the wait -> LDS -> FMA -> STS -> fence -> arrive chain of one box hides behind the
      // others (one 128-thread group per box measured 0.503 ms against 0.376 ms unfused at 128 x 64^2 x 256).
      const int pw = warp - 2;                     // 0..3
      const int j = lane & 7, r_first = lane >> 3;
      const int cB = n_tile * BNW + (int)rank * 64;          // one B source only (checked by the launcher)
      float sc[8], sh[8];
      {
        const float4* ps = reinterpret_cast<const float4*>(p.b_pro_scale + cB + j * 8);
        const float4* ph = reinterpret_cast<const float4*>(p.b_pro_shift + cB + j * 8);
        const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), h0 = __ldg(ph), h1 = __ldg(ph + 1);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (((kb - kb0) & 3) == pw) {
          const int xt = kb % p.tiles_x;
          const int yt = (kb / p.tiles_x) % p.tiles_y;
          const int x0 = xt * p.tw, yy = yt * p.th + tg - 1;
          uint8_t* sbp = smem + stage * stage_bytes + 2 * p.abox;
          mbar_wait(&blocal[stage], phase);
          if (yy >= 0 && yy < p.H && !(p.dbg_boff & 64)) {     // bit 6: timing experiment, no transform
            // nine rows per pass, loads first: the LDS -> FMA -> STS chains of one thread overlap
            for (int rb = r_first; rb < p.kp + 2; rb += 36) {
              uint4 v[9];
              bool on[9];
#pragma unroll
              for (int i = 0; i < 9; ++i) {
                const int r = rb + 4 * i;
                const int xx = x0 - 1 + r;
                on[i] = r < p.kp + 2 && xx >= 0 && xx < p.W;
                if (on[i]) v[i] = *reinterpret_cast<const uint4*>(sbp + r * 128 + ((j ^ (r & 7)) << 4));
              }
#pragma unroll
              for (int i = 0; i < 9; ++i) {
                if (!on[i]) continue;
                const int r = rb + 4 * i;
                uint32_t* ww = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float lo = fmaxf(fmaf(bf16lo(ww[k]), sc[2 * k], sh[2 * k]), 0.f);
                  const float hi = fmaxf(fmaf(bf16hi(ww[k]), sc[2 * k + 1], sh[2 * k + 1]), 0.f);
                  ww[k] = pack_bf16x2(lo, hi);
                }
                *reinterpret_cast<uint4*>(sbp + r * 128 + ((j ^ (r & 7)) << 4)) = v[i];
              }
            }
            if (!(p.dbg_boff & 128)) fence_proxy_async_smem();     // bit 7: timing experiment, no proxy fence
          }
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&bready[stage]);
            else mbar_arrive_remote(&bready[stage], 0);
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    const int quad = warp & 3;
    const int mrow = a_base + quad * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    for (int t = 0; t < p.T; ++t) {
      const int tap = tg * p.T + t;
      float* orow = p.out + ((static_cast<size_t>(split) * p.taps_total + tap) * p.Ca + mrow) * p.Nb + n_tile * BNW;
      for (int c = 0; c < BNW; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * BNW + c, v);
        tmem_ld_wait();
        if (mrow < p.Ca) {
          float4* o4 = reinterpret_cast<float4*>(orow + c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------ 64-output-channel variant
// conv3x3 layers with Cout = 64 (the three full-resolution layers = the largest pixel counts) would
// waste half of an M=128 tile on dY.  Here the roles are swapped: N = the 64 dY channels, and M = 128 =
// TWO filter taps x 64 input channels, stacked through the descriptor's leading-dimension stride: the
// second 64-row block of the A operand is simply the same halo tile shifted by (tap_b - tap_a) pixels.
// All nine taps (five tap pairs, 5 x 64 TMEM columns) and all three halo rows live in one CTA, so each
// dY pixel block is read once for the whole 3x3 filter.
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad64_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ CUtensorMap mapX0,
               const __grid_constant__ CUtensorMap mapX1, const WgradParams p, const int stage_bytes) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* done_bar = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapDy);
    tma_prefetch_desc(&mapX0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tile = blockIdx.x % p.n_tiles;       // which 64-channel block of the (concatenated) input
  const int split = blockIdx.x / p.n_tiles;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);

  if (warp == 0) {
    {
      const CUtensorMap* mapX = (n_tile < p.c0_blocks) ? &mapX0 : &mapX1;
      const int cX = ((n_tile < p.c0_blocks) ? n_tile : (n_tile - p.c0_blocks)) * 64;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int xt = kb % p.tiles_x;
        const int y = (kb / p.tiles_x) % p.tiles_y;
        const int n = kb / (p.tiles_x * p.tiles_y);
        const int x0 = xt * p.kp;
        uint8_t* st = smem + stage * stage_bytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], p.abox + 3 * p.b_tx);
          tma_load_5d(st, &mapDy, &full_bar[stage], 0, x0, y, n, 0);
          for (int r = 0; r < 3; ++r)
            tma_load_5d(st + p.abox + r * p.bslot, mapX, &full_bar[stage], cX, x0 - 1, y + r - 1, n, 0);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool mn = !(p.dbg_boff & 8);
      const uint32_t idesc = make_idesc_bf16(128, 64, mn, mn);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t sdy = smem_u32(smem + stage * stage_bytes);
        const uint32_t sx = sdy + p.abox;
        if (elect_one()) {
          // five tap-pair descriptors for this stage; every K step only advances them by 16 rows
          uint64_t ad[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const int ta = 2 * j, tb = 2 * j + 1;
            const uint32_t offa = (ta / 3) * p.bslot + (ta % 3) * 128;
            const uint32_t offb = (j < 4) ? (uint32_t)((tb / 3) * p.bslot + (tb % 3) * 128) : offa + 128;
            ad[j] = make_smem_desc_sw128(sx + offa, offb - offa, 1024);
          }
          uint64_t bdesc = make_smem_desc_sw128(sdy, 1024, 1024);
          const uint32_t acc_first = kb > kb0 ? 1u : 0u;
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            umma_bf16(tmem_base + j * 64, ad[j], bdesc, idesc, acc_first);
            ad[j] += 2048 >> 4;
          }
          bdesc += 2048 >> 4;
#pragma unroll 1
          for (int kk = 1; kk < p.kp / 16; ++kk) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              umma_bf16(tmem_base + j * 64, ad[j], bdesc, idesc, 1u);
              ad[j] += 2048 >> 4;
            }
            bdesc += 2048 >> 4;
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(done_bar);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int half = row >> 6, ci = row & 63;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    for (int j = 0; j < 5; ++j) {
      const int tap = 2 * j + half;
      uint32_t v[64];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + j * 64;
      tmem_ld_32x32b_x32(taddr, v);
      tmem_ld_32x32b_x32(taddr + 32, v + 32);
      tmem_ld_wait();
      if (tap < 9) {
        // out[split][tap][co][ci]: for a fixed co the 32 lanes write 32 consecutive floats
        float* o = p.out + ((static_cast<size_t>(split) * 9 + tap) * 64) * p.Nb + n_tile * 64 + ci;
#pragma unroll
        for (int c = 0; c < 64; ++c) o[static_cast<size_t>(c) * p.Nb] = __uint_as_float(v[c]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ 64-output-channel variant, N = 192
// Same problem as wgrad64_kernel with the roles arranged so that the MMAs are wide instead of tall.  An M=128 x N=64
// MMA reads 4 KB of A and 2 KB of B from shared memory for 32 cycles of math: the 128 B/clk port caps it at 62 % of
// the tensor rate (profiles/r01/mma_rate.log: 51.5 cycles), and five of them per K step leave wgrad64_kernel at
// ~45 % of peak.  Here a stage is ONE row segment of the layer input (KP+2 pixels, 64 channels) and the THREE dY row
// segments above / at / below it:
//   B (N = 192) = the three column taps as shifted windows of the input row   (leading-dimension stride 128 B)
//   A (M = 128) = two dY rows stacked                                          (leading-dimension stride one row slot)
// MMA 1 = rows (y, y+1) -> taps dr = 1 (lanes 0-63) and dr = 0 (lanes 64-127); MMA 2 = rows (y-1, y) -> tap dr = 2 in
// lanes 0-63 (lanes 64-127 repeat dr = 1 and are dropped: M = 64 would run at the same 96 cycles).  Two 96-cycle MMAs
// per 16 pixels against 10 KB of operand reads each: math-bound at 75 % of the tensor rate by construction.
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad64n_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ CUtensorMap mapX0,
                const __grid_constant__ CUtensorMap mapX1, const WgradParams p, const int stage_bytes) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* done_bar = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapDy);
    tma_prefetch_desc(&mapX0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tile = blockIdx.x % p.n_tiles;       // which 64-channel block of the (concatenated) input
  const int m_tile = (blockIdx.x / p.n_tiles) % p.m_tiles;     // which 64-channel block of dY
  const int split = blockIdx.x / (p.n_tiles * p.m_tiles);
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int TH = p.th;                             // input rows per stage; dY rows y0-1 .. y0+TH come as one TMA box
  const int dy_bytes = (TH + 2) * p.abox;

  if (warp == 0) {
    const CUtensorMap* mapX = (n_tile < p.c0_blocks) ? &mapX0 : &mapX1;
    const int cX = ((n_tile < p.c0_blocks) ? n_tile : (n_tile - p.c0_blocks)) * 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      const int xt = kb % p.tiles_x;
      const int y = ((kb / p.tiles_x) % p.tiles_y) * TH;   // first row of the layer INPUT (zero fill outside the image)
      const int n = kb / (p.tiles_x * p.tiles_y);
      const int x0 = xt * p.kp;
      uint8_t* st = smem + stage * stage_bytes;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], dy_bytes + TH * p.b_tx);
        tma_load_5d(st, &mapDy, &full_bar[stage], m_tile * 64, x0, y - 1, n, 0);
        tma_load_5d(st + dy_bytes, mapX, &full_bar[stage], cX, x0 - 1, y, n, 0);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 192, true, true);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after_sync();
      const uint32_t sdy = smem_u32(smem + stage * stage_bytes);
      const uint32_t sx = sdy + dy_bytes;
      if (elect_one()) {
        uint32_t acc = kb > kb0 ? 1u : 0u;
#pragma unroll 1
        for (int r = 0; r < TH; ++r) {          // input row y0 + r against dY rows y0 + r - 1 .. y0 + r + 1
          uint64_t a1 = make_smem_desc_sw128(sdy + (r + 1) * p.abox, p.abox, 1024);     // dY rows (y, y+1)
          uint64_t a2 = make_smem_desc_sw128(sdy + r * p.abox, p.abox, 1024);           // dY rows (y-1, y)
          uint64_t bd = make_smem_desc_sw128(sx + r * p.b_tx, 128, 1024);    // column taps 0, 1, 2: one pixel apart
#pragma unroll 1
          for (int kk = 0; kk < p.kp / 16; ++kk) {
            umma_bf16(tmem_base, a1, bd, idesc, acc);
            umma_bf16(tmem_base + 192, a2, bd, idesc, acc);
            acc = 1u;
            a1 += 2048 >> 4;
            a2 += 2048 >> 4;
            bd += 2048 >> 4;
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int co = (quad & 1) * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    // accumulator 0: lanes 0-63 tap row 1, lanes 64-127 tap row 0; accumulator 1: lanes 0-63 tap row 2
    for (int acc = 0; acc < (quad < 2 ? 2 : 1); ++acc) {
      const int dr = acc ? 2 : (quad < 2 ? 1 : 0);
      for (int dc = 0; dc < 3; ++dc) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 192 + dc * 64;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_32x32b_x32(taddr + 32, v + 32);
        tmem_ld_wait();
        // out[split][tap][co][ci]: this lane owns 64 consecutive input channels of its output channel
        float4* o = reinterpret_cast<float4*>(p.out + ((static_cast<size_t>(split) * 9 + dr * 3 + dc) * p.Ca + m_tile * 64 + co) *
                                                  p.Nb + n_tile * 64);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          o[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                             __uint_as_float(v[4 * c + 3]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

static int pow2_div(int v, int cap) {
  int t = 1;
  while (t < cap && (v % (t * 2)) == 0) t *= 2;
  return t;
}

struct WgPlan {
  int T, tap_groups, taps_total, BNW, m_tiles, n_tiles, splits, kb_total, kb_per_split;
  int tw, th, nb, tiles_x, tiles_y, tiles_n;
  int Nb;
  int shifted;
  int kp;
  int stacked;     // wgrad64_kernel
  int pair;        // wgrad_gemm_pair_kernel (cta_group::2)
};

// Waves of CTAs per launch (one CTA per SM at a time) that the split-K plan aims for.  ONE: a second wave repeats the
// ~10 us pipeline fill + partial-slab store of every CTA and doubles what wgrad_reduce has to read, and buys nothing —
// there is no second CTA on the SM to overlap with.  Measured on the whole step (profiles/r02/wgrad_waves_ab.log):
// 31.88 -> 31.59 ms at batch 128, 4.34 -> 4.26 ms at batch 16.  SUNET_WGRAD_WAVES=2 restores the old plan.
static int wgrad_waves() {
  static const int waves = [] {
    const char* e = getenv("SUNET_WGRAD_WAVES");
    const int v = e ? atoi(e) : 1;
    return v == 2 ? 2 : 1;
  }();
  return waves;
}

static bool wgrad64_wide() {
  static const bool wide = [] {
    const char* e = getenv("SUNET_WGRAD64_WIDE");     // 0: the tall (M = two taps, N = 64) form, wgrad64_kernel
    return e ? atoi(e) != 0 : true;
  }();
  return wide;
}

static int plan_wgrad(const sunet_wgrad_gemm_args* a, WgPlan* w) {
  switch (a->b_mode) {
    case SUNET_A_CONV3X3: w->T = 3; w->tap_groups = 3; break;
    case SUNET_A_PLAIN: w->T = 1; w->tap_groups = 1; break;
    case SUNET_A_GATHER2X2: w->T = 4; w->tap_groups = 1; break;
    default: return set_error(SUNET_ERR_INVALID, "wgrad_gemm: bad b_mode %d", a->b_mode);
  }
  w->taps_total = w->T * w->tap_groups;
  w->Nb = a->b0_channels + (a->b1 ? a->b1_channels : 0);
  if (a->a_channels <= 0 || a->a_channels % 64 || a->b0_channels <= 0 || a->b0_channels % 64 ||
      (a->b1 && a->b1_channels % 64))
    return set_error(SUNET_ERR_INVALID, "wgrad_gemm: channel counts must be multiples of 64");
  w->BNW = (a->b0_channels % 128 == 0 && (!a->b1 || a->b1_channels % 128 == 0)) ? 128 : 64;
  w->m_tiles = (a->a_channels + 127) / 128;
  w->n_tiles = w->Nb / w->BNW;
  w->pair = (a->a_channels % 256 == 0 && w->BNW == 128 && getenv("SUNET_WGRAD_NO_PAIR") == nullptr) ? 1 : 0;
  if (w->pair) w->m_tiles = a->a_channels / 256;
  const int B = a->batch, H = a->height, W = a->width;
  // the 64-channel dY-block kernels serve Cout = 64.  The N = 192 form can also take wider dY in 64-channel blocks
  // (SUNET_WGRAD64_MAXC=128), but its 25 % idle M rows lose to the M = 128 x N = 128 kernel there: 0.506 vs 0.456 ms
  // for 128 x 128^2 x 128->128 (profiles/r02/wgrad64_wide_ab.log)
  static const int stack_max = [] {
    const char* e = getenv("SUNET_WGRAD64_MAXC");
    return e ? atoi(e) : 64;
  }();
  w->stacked = (a->b_mode == SUNET_A_CONV3X3 && W % 64 == 0 && getenv("SUNET_WGRAD_NO_STACK") == nullptr &&
                (a->a_channels == 64 || (wgrad64_wide() && a->a_channels % 64 == 0 && a->a_channels <= stack_max))) ? 1 : 0;
  if (w->stacked) {
    w->kp = 64;
    w->tw = 64; w->th = 1; w->nb = 1;
    if (wgrad64_wide()) {
      // input rows per stage: the TH + 2 dY rows of a stage serve TH input rows, so dY is fetched (TH+2)/TH times
      // instead of three (shared-memory fill traffic is what the N = 192 form has left to give)
      int th = 4;
      if (const char* e = getenv("SUNET_WGRAD64_TH")) th = atoi(e);
      if (th != 1 && th != 2 && th != 4) th = 4;
      while (H % th) th >>= 1;
      w->th = th;
    }
    w->tiles_x = W / 64; w->tiles_y = H / w->th; w->tiles_n = B;
    w->kb_total = w->tiles_x * w->tiles_y * B;
    w->n_tiles = w->Nb / 64;
    w->m_tiles = a->a_channels / 64;
    w->pair = 0;
    w->shifted = 1;
    const int items = w->n_tiles * w->m_tiles;
    int want = (wgrad_waves() * num_sms()) / items;
    // short problems: one wave of longer CTAs (prologue + partial-store epilogue cost ~10 us per CTA)
    if (want > 1 && w->kb_total * w->th / want < 48) want = num_sms() / items;
    if (want < 1) want = 1;
    int max_splits = (w->kb_total * w->th + 7) / 8;
    if (max_splits > w->kb_total) max_splits = w->kb_total;
    if (want > max_splits) want = max_splits;
    w->kb_per_split = (w->kb_total + want - 1) / want;
    w->splits = (w->kb_total + w->kb_per_split - 1) / w->kb_per_split;
    return SUNET_OK;
  }
  // stage size: aim at >= ~768 MMA cycles per stage = (kp/16) * T * BNW/2
  int kp = (w->BNW == 128) ? 64 : 128;
  if (w->T == 1) kp = 128;
  {
    const char* e = getenv("SUNET_WGRAD_KP");
    if (e && (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128)) kp = atoi(e);
  }
  // never make a stage larger than the whole problem (tiny test shapes)
  while (kp > 32 && (long long)B * H * W < 2LL * kp) kp /= 2;
  // a stage must be whole image rows (or a power-of-two piece of one): shrink until it tiles
  for (;;) {
    w->tw = pow2_div(W, kp);
    w->th = pow2_div(H, kp / w->tw);
    w->nb = kp / (w->tw * w->th);
    if (!(w->nb > 1 && w->th != H)) break;
    if (kp == 32)
      return set_error(SUNET_ERR_INVALID, "wgrad_gemm: cannot tile %d x %d x %d into pixel blocks", B, H, W);
    kp /= 2;
  }
  w->kp = kp;
  w->tiles_x = W / w->tw;
  w->tiles_y = H / w->th;
  w->tiles_n = (B + w->nb - 1) / w->nb;
  w->kb_total = w->tiles_x * w->tiles_y * w->tiles_n;
  w->shifted = (a->b_mode == SUNET_A_CONV3X3 && w->tw == kp && getenv("SUNET_WGRAD_NO_SHIFT") == nullptr) ? 1 : 0;
  const int base_items = w->m_tiles * w->n_tiles * w->tap_groups * (w->pair ? 2 : 1);   // CTAs per split
  // full waves of CTAs (one CTA per SM at a time): round DOWN so no extra, nearly empty wave appears
  int want = (wgrad_waves() * num_sms()) / base_items;
  // short problems: one wave of longer CTAs (prologue + partial-store epilogue cost ~10 us per CTA)
  if (want > 1 && w->kb_total / want < 48) want = num_sms() / base_items;
  if (want < 1) want = 1;
  int max_splits = (w->kb_total + 7) / 8;  // at least 8 k-blocks per CTA when possible
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  w->kb_per_split = (w->kb_total + want - 1) / want;
  w->splits = (w->kb_total + w->kb_per_split - 1) / w->kb_per_split;
  return SUNET_OK;
}

static int make_map(CUtensorMap* m, const void* base, int gs, int C, int S, int B, int H, int W, const WgPlan& w,
                    int halo = 0) {
  uint64_t dims[5], str[4];
  uint32_t box[5];
  const uint64_t e = 2;
  if (!gs) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = B; dims[4] = 1;
    str[0] = (uint64_t)S * e; str[1] = (uint64_t)W * S * e; str[2] = (uint64_t)H * W * S * e;
    str[3] = (uint64_t)B * H * W * S * e;
    box[0] = 64; box[1] = w.tw + halo; box[2] = w.th; box[3] = w.nb; box[4] = 1;
  } else {
    dims[0] = C; dims[1] = 2; dims[2] = W; dims[3] = 2; dims[4] = (uint64_t)B * H;
    str[0] = (uint64_t)S * e; str[1] = (uint64_t)2 * S * e; str[2] = (uint64_t)2 * W * S * e;
    str[3] = (uint64_t)4 * W * S * e;
    box[0] = 64; box[1] = 1; box[2] = w.tw; box[3] = 1; box[4] = w.th * w.nb;
  }
  return make_tmap_bf16_5d(m, base, dims, str, box);
}

}  // namespace sunet

using namespace sunet;

extern "C" int sunet_wgrad_gemm_pro_supported(const sunet_wgrad_gemm_args* a) {
  WgPlan w{};
  if (!a || a->batch <= 0 || a->height <= 0 || a->width <= 0 || plan_wgrad(a, &w)) return 0;
  return (w.pair && w.shifted && !w.stacked && a->b1 == nullptr && a->b_mode == SUNET_A_CONV3X3) ? 1 : 0;
}

extern "C" int sunet_wgrad_gemm_splits(const sunet_wgrad_gemm_args* a) {
  WgPlan w{};
  if (!a || plan_wgrad(a, &w)) return -1;
  return w.splits;
}

extern "C" int sunet_wgrad_gemm(const sunet_wgrad_gemm_args* a, sunet_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: null args");
  if (!a->a || !a->b0 || !a->partials) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: null tensor");
  if (a->batch <= 0 || a->height <= 0 || a->width <= 0) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: bad grid");
  if (a->b_mode == SUNET_A_GATHER2X2 && a->b1) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: gather takes one B");
  WgPlan w{};
  int e = plan_wgrad(a, &w);
  if (e) return e;
  const size_t need = (size_t)w.splits * w.taps_total * a->a_channels * w.Nb * sizeof(float);
  if (a->partials_bytes < need)
    return set_error(SUNET_ERR_WORKSPACE, "wgrad_gemm: partials buffer %zu < %zu bytes", (size_t)a->partials_bytes, need);
  if (a->b1 && (a->b0_channels % w.BNW)) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: b0 channels vs tile");
  const bool pro = a->b_pro_scale != nullptr;
  if (pro && (!a->b_pro_shift || !sunet_wgrad_gemm_pro_supported(a)))
    return set_error(SUNET_ERR_INVALID, "wgrad_gemm: the B prologue (b_pro_*) needs both vectors and a shape served by the "
                                        "CTA-pair kernel; check sunet_wgrad_gemm_pro_supported()");

  CUtensorMap mA, mB0, mB1;
  const int B = a->batch, H = a->height, W = a->width;
  if ((e = make_map(&mA, a->a, 0, a->a_channels, a->a_pix_stride, B, H, W, w))) return e;
  const int halo = w.shifted ? 2 : 0;
  if ((e = make_map(&mB0, a->b0, a->b_mode == SUNET_A_GATHER2X2, a->b0_channels, a->b0_pix_stride, B, H, W, w, halo)))
    return e;
  if (a->b1) {
    if ((e = make_map(&mB1, a->b1, 0, a->b1_channels, a->b1_pix_stride, B, H, W, w, halo))) return e;
  } else {
    mB1 = mB0;
  }
  if (w.stacked) {
    if (a->b1 && (a->b0_channels % 64)) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: b0 channels vs tile");
    WgradParams q{};
    q.mode = a->b_mode;
    q.n_tiles = w.n_tiles; q.m_tiles = w.m_tiles; q.splits = w.splits; q.kb_total = w.kb_total;
    q.kb_per_split = w.kb_per_split;
    q.tiles_x = w.tiles_x; q.tiles_y = w.tiles_y;
    q.c0_blocks = a->b0_channels / 64;
    q.Ca = a->a_channels; q.Nb = w.Nb; q.taps_total = 9;
    q.out = a->partials;
    {
      const char* b = getenv("SUNET_DBG_BOFF");
      q.dbg_boff = b ? atoi(b) : 0;
    }
    q.kp = w.kp; q.abox = w.kp * 128;
    q.b_tx = (w.kp + 2) * 128;
    q.bslot = (q.b_tx + 1023) / 1024 * 1024;
    if (wgrad64_wide()) {
      // TH input rows (B, N = 192 as three shifted windows each) against TH + 2 dY rows (A, stacked in pairs)
      CUtensorMap mDy3;
      WgPlan w3 = w;
      w3.th = w.th + 2;
      if ((e = make_map(&mDy3, a->a, 0, a->a_channels, a->a_pix_stride, B, H, W, w3))) return e;
      q.th = w.th;
      const int sbytes = ((w.th + 2) * q.abox + w.th * q.b_tx + 1023) / 1024 * 1024;
      const int bbytes = (2 * MAX_STAGES + 2) * 8;
      int st = (227 * 1024 - 1024 - bbytes) / sbytes;
      if (st > MAX_STAGES) st = MAX_STAGES;
      q.stages = st;
      static bool attr64n = false;
      if (!attr64n) {
        if ((e = check_cuda(cudaFuncSetAttribute(wgrad64n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024), "cudaFuncSetAttribute(wgrad64n)")))
          return e;
        attr64n = true;
      }
      launch_k(wgrad64n_kernel, dim3(w.n_tiles * w.m_tiles * w.splits), dim3(WG_THREADS), st * sbytes + 1024 + bbytes, stream, mDy3,
               mB0, mB1, q, sbytes);
      return check_launch("wgrad64n_kernel");
    }
    const int sbytes = q.abox + 3 * q.bslot;
    const int bbytes = (2 * MAX_STAGES + 2) * 8;
    int st = (227 * 1024 - 1024 - bbytes) / sbytes;
    if (st > MAX_STAGES) st = MAX_STAGES;
    q.stages = st;
    static bool attr64 = false;
    if (!attr64) {
      if ((e = check_cuda(cudaFuncSetAttribute(wgrad64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024), "cudaFuncSetAttribute(wgrad64)")))
        return e;
      attr64 = true;
    }
    launch_k(wgrad64_kernel, dim3(w.n_tiles * w.splits), dim3(WG_THREADS), st * sbytes + 1024 + bbytes, stream, mA, mB0, mB1, q, sbytes);
    return check_launch("wgrad64_kernel");
  }
  WgradParams p;
  p.mode = a->b_mode;
  p.T = w.T; p.tap_groups = w.tap_groups; p.nb64 = w.BNW / 64;
  p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles; p.splits = w.splits;
  p.kb_total = w.kb_total; p.kb_per_split = w.kb_per_split;
  p.tw = w.tw; p.th = w.th; p.nb = w.nb; p.tiles_x = w.tiles_x; p.tiles_y = w.tiles_y;
  p.H = H;
  p.c0_blocks = a->b0_channels / w.BNW;
  p.Ca = a->a_channels; p.Nb = w.Nb;
  p.taps_total = w.taps_total;
  p.out = a->partials;
  {
    const char* s = getenv("SUNET_DBG_SHIFT");
    const char* b = getenv("SUNET_DBG_BOFF");
    p.dbg_shift = s ? atoi(s) : 0;
    p.dbg_boff = b ? atoi(b) : 0;
  }
  {
    const char* r = getenv("SUNET_WGRAD_A_REUSE");
    p.a_reuse = r ? atoi(r) : 1;
  }
  p.b_pro_scale = pro ? a->b_pro_scale : nullptr;
  p.b_pro_shift = pro ? a->b_pro_shift : nullptr;
  p.W = W;
  p.tmem_cols = 32;
  while (p.tmem_cols < w.T * w.BNW) p.tmem_cols *= 2;
  p.shifted = w.shifted;
  p.kp = w.kp;
  p.abox = w.kp * 128;
  const int nb64_cta = w.pair ? p.nb64 / 2 : p.nb64;   // 64-channel B boxes each CTA loads per tap
  if (w.shifted) {
    p.b_tx = (w.kp + 2) * 128;                       // kp + 2 halo pixels
    p.bslot = (p.b_tx + 1023) / 1024 * 1024;         // padded to the 1024-byte swizzle atom
    p.bboxes = nb64_cta;
  } else {
    p.bslot = p.abox;
    p.b_tx = p.abox;
    p.bboxes = w.T * nb64_cta;
  }
  const int stage_bytes = 2 * p.abox + p.bboxes * p.bslot;
  const int bar_bytes = WG_BAR_BYTES;
  int stages = (227 * 1024 - 1024 - bar_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return set_error(SUNET_ERR_INVALID, "wgrad_gemm: stage of %d bytes does not fit", stage_bytes);
  p.stages = stages;
  const int smem = stages * stage_bytes + 1024 + bar_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = check_cuda(cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024),
                        "cudaFuncSetAttribute(wgrad_gemm)")))
      return e;
    attr_set = true;
  }
  const int grid = w.m_tiles * w.n_tiles * w.tap_groups * w.splits;
  if (w.pair) {
    static bool attr_pair = false;
    if (!attr_pair) {
      if ((e = check_cuda(cudaFuncSetAttribute(wgrad_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               227 * 1024), "cudaFuncSetAttribute(wgrad_gemm_pair)")))
        return e;
      attr_pair = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid * 2);
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if ((e = check_cuda(cudaLaunchKernelEx(&cfg, wgrad_gemm_pair_kernel, mA, mB0, mB1, p, stage_bytes),
                        "cudaLaunchKernelEx(wgrad_gemm_pair)")))
      return e;
    return check_launch("wgrad_gemm_pair_kernel");
  }
  launch_k(wgrad_gemm_kernel, dim3(grid), dim3(WG_THREADS), smem, stream, mA, mB0, mB1, p, stage_bytes);
  return check_launch("wgrad_gemm_kernel");
}

// ------------------------------------------------------------------ split-K reduction
namespace sunet {
// grad[a][b][tap] = sum_k P[k][tap][a][b]  (layout 0: a = co, b = ci, 9 taps; layout 1: a = ci, b = co, 4 taps).
// One thread per (tap, a, b) in the PARTIALS' order, so every split is read as contiguous rows; the TAPS-strided
// 4-byte writes touch each output line TAPS times but the output is 1/splits of the traffic.
template <int TAPS>
__global__ void __launch_bounds__(256)
wgrad_reduce_taps_kernel(const float* __restrict__ P, int splits, int ab, float* __restrict__ grad) {
  pdl_wait();
  pdl_trigger();
  const int slab = TAPS * ab;                     // < 2^31 (checked by the launcher)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < slab; i += gridDim.x * blockDim.x) {
    const int tap = i / ab;
    const int j = i - tap * ab;
    const float* src = P + i;
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = 0.f;
    int k = 0;
    for (; k + 8 <= splits; k += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += src[(size_t)(k + u) * slab];
    }
    for (; k < splits; ++k) s[0] += src[(size_t)k * slab];
    grad[(size_t)j * TAPS + tap] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
}
// (A variant that stages TAPS x 64 sums in shared memory so that the output is written as one contiguous run was
// measured and is slower — 0.41 vs 0.25 ms per step at 16 patches, 0.47 vs 0.30 at 128: the partials are read with 9x
// fewer threads in flight, and the reads, not the strided writes, are what this kernel waits for.)

// first layer: grad[co][cin][tap] from P[k][co][tap*cin + ci] (Nb = padded K of the im2col'ed input)
__global__ void __launch_bounds__(256)
wgrad_reduce_first_kernel(const float* __restrict__ P, int splits, int Ca, int Nb, int real_cin,
                          float* __restrict__ grad, long long total) {
  pdl_wait();
  pdl_trigger();
  const long long slab = (long long)Ca * Nb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % real_cin);
    const int co = (int)(i / (9LL * real_cin));
    const long long src = (long long)co * Nb + tap * real_cin + ci;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += P[k * slab + src];
    grad[i] = s;
  }
}
// paired-pixel first layer: P[k][128][64] with A = (pixel parity, co), B = (pixel parity, 32-wide k);
// grad[co][cin][tap] = sum_k P[k][co][tap*cin+ci] + P[k][64+co][32 + tap*cin+ci]   (the two diagonal blocks)
__global__ void __launch_bounds__(256)
wgrad_reduce_first_pair_kernel(const float* __restrict__ P, int splits, int real_cin, float* __restrict__ grad,
                               int total) {
  pdl_wait();
  pdl_trigger();
  const int slab = 128 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int ci = (i / 9) % real_cin;
    const int co = i / (9 * real_cin);
    const int k = tap * real_cin + ci;
    const int s0 = co * 64 + k, s1 = (64 + co) * 64 + 32 + k;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += P[(size_t)sp * slab + s0] + P[(size_t)sp * slab + s1];
    grad[i] = s;
  }
}
}  // namespace sunet

extern "C" int sunet_wgrad_reduce(const float* partials, int splits, int taps, int a_channels, int b_channels,
                                  int layout, int real_cin, float* grad, sunet_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!partials || !grad || splits <= 0 || a_channels <= 0 || b_channels <= 0)
    return set_error(SUNET_ERR_INVALID, "wgrad_reduce: bad arguments");
  long long total;
  if (layout == 0 && taps == 9) total = 9LL * a_channels * b_channels;
  else if (layout == 1 && taps == 4) total = 4LL * a_channels * b_channels;
  else if (layout == 2 && taps == 1 && real_cin > 0 && real_cin * 9 <= b_channels) total = 9LL * a_channels * real_cin;
  else if (layout == 3 && taps == 1 && a_channels == 128 && b_channels == 64 && real_cin > 0 && real_cin * 9 <= 32)
    total = 9LL * 64 * real_cin;
  else return set_error(SUNET_ERR_INVALID, "wgrad_reduce: layout %d / taps %d mismatch", layout, taps);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "wgrad_reduce: tensor too large");
  const int ab = a_channels * b_channels;
  if (layout == 0)
    launch_k(wgrad_reduce_taps_kernel<9>, dim3((int)blocks), dim3(256), 0, stream, partials, splits, ab, grad);
  else if (layout == 1)
    launch_k(wgrad_reduce_taps_kernel<4>, dim3((int)blocks), dim3(256), 0, stream, partials, splits, ab, grad);
  else if (layout == 3)
    launch_k(wgrad_reduce_first_pair_kernel, dim3((int)blocks), dim3(256), 0, stream, partials, splits, real_cin, grad,
             (int)total);
  else
    launch_k(wgrad_reduce_first_kernel, dim3((int)blocks), dim3(256), 0, stream, partials, splits, a_channels, b_channels, real_cin,
                                                               grad, total);
  return check_launch("wgrad_reduce");
}
