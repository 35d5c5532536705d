// G1h2 — the halo-tile conv3x3 kernel on CTA PAIRS (tcgen05 cta_group::2).
//
// conv3_halo_kernel<64/128> is bound by shared-memory bandwidth, not by the tensor pipe: an M=128,
// N=128 MMA reads 8 KB of operands per 64 tensor cycles (128 B/clk, the whole SMEM port), N=64 needs
// 192 B/clk, and TMA fills + epilogue staging compete for the same port (scripts/mma_rate.py: N=64 tops
// out at 62 % of tensor peak even with nothing else touching SMEM).  With cta_group::2 the two CTAs of a
// cluster issue ONE M=256 MMA: each supplies its own 128 pixel rows (its own halo tile) but only HALF
// of the weight rows, so per-SM operand reads and weight TMA traffic drop by a quarter to a third.
//
// Differences from conv3_halo.cu: cluster of 2 CTAs; the pair works on blocks (2j, 2j+1); "full" barriers
// live in the leader CTA and are signalled by both CTAs' TMA loads (peer-bit-masked barrier address) plus
// one plain arrival per CTA; the leader alone issues tcgen05.mma.cta_group::2 and multicasts its commits
// to both CTAs' "empty" / "tmem full" barriers; both epilogues release the accumulators to the leader.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

struct Halo2Params {
  int cpt0, cpt1;
  int blocks_x, blocks_y;
  int m_blocks, n_tiles;
  float* stats;              // [gridDim.x / n_tiles][n_total][2] or nullptr (one row per CTA)
  int n_total;
  // BNB (fused BatchNorm-backward reduction): per-channel constants of the BN whose input gradient this launch
  // produces; stats then holds (sum g, sum g * xhat) with g = dst masked by relu(bn(y)) > 0
  const float* bnb_scale;
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  const float* ep_scale;     // inference epilogue: dst = relu(acc * ep_scale[n] + ep_shift[n]) (nullptr = off)
  const float* ep_shift;
  int bnb_col0;              // BNB covers output columns [bnb_col0, n_total); lower columns get (sum, sum sq)
  // PRO (training prologue): the source tensor holds the RAW conv output y of the producer block; every landed halo
  // tile is transformed in place to relu(y * pro_scale[c] + pro_shift[c]) (bf16) before the MMAs read it.  Indexed by
  // concatenated input channel (source 0 first); pro_mask bit s = transform the chunks of source s.
  const float* pro_scale;
  const float* pro_shift;
  int pro_mask;
};

// TILES = M=128 tiles per CTA and block: 2 (16x16 block) for BN <= 128, 1 (8 wide x 16 tall) for BN = 256
// BNB = fused BN-backward reduction in the epilogue: two extra 16 KB slots hold the matching tiles of y (the
// conv output the BN normalised), paid for with one or two weight stages.
// Two halo (A) stages.  A third stage (paid for with weight stages and y slots) was measured on B200 and changes
// nothing: 0.550 vs 0.548 ms for 128 x 256^2 x 64->64, 33.06 vs 33.18 ms per step (profiles/r02/halo_stages_ab.log) —
// the N = 64 kernel is bound by the shared-memory port (operand reads 65 % + epilogue 27 % of its wavefronts), not
// by the latency of the halo load.
// Also measured and rejected (profiles/r02/epilogue_groups_ab.log): EIGHT epilogue warps in two groups that take the
// 64-channel output chunks alternately (own staging tile, y pipeline, store thread and named barrier each).  Parity
// green, but 0.713 vs 0.736 ms for the level-1 BNB dgrad, 0.479 vs 0.463 at level 2, 0.403 vs 0.381 at level 3 and
// 32.65 vs 32.52 ms per step: the cost of the BNB epilogue is its extra traffic through the shared-memory port
// (y tile in, statistics loop out), not the latency of four warps — more warps only contend harder.
//
// PRO = training-mode prologue fusion (north_star: "BatchNorm ... scale/shift and ReLU fused into the ... prologue"):
// four extra warps transform every landed halo tile from the producer's raw conv output y to relu(bn(y)) in place.
// Each CTA's halo load then completes on a barrier of its OWN (the transform warps of that CTA wait on it) and the
// leader's MMA warp waits on a second barrier that both CTAs' transform warps arrive on.
template <int BN, bool BNB = false, bool PRO = false>
struct H2Cfg {
  static constexpr int TILES = (BN == 256) ? 1 : 2;
  static constexpr int PITCH = 8 * TILES + 2;
  static constexpr int A_TX = 18 * PITCH * 128;
  static constexpr int A_SLOT = (A_TX + 1023) / 1024 * 1024;
  static constexpr int A_STAGES = 2;
  static constexpr int B_HALF = (BN / 2) * 128;      // this CTA's half of one weight tile
  static constexpr int B_STAGES = (BN == 256) ? (BNB ? 7 : 8) : ((BNB && BN == 128) ? 9 : 10);
  static constexpr int STG_BYTES = 128 * 128;
  static constexpr int Y_SLOTS = BNB ? ((BN == 64) ? 4 : 2) : 0;      // power of two
  static constexpr int THREADS = PRO ? 320 : 192;    // warp 0 TMA, 1 MMA, 2-5 epilogue, (PRO) 6-9 prologue transform
  static_assert(!(PRO && BNB), "the prologue transform is a forward-pass feature");
  static constexpr int EP_BYTES = BNB ? 0 : 2 * BN * 4;   // inference epilogue: this CTA's scale / shift columns
  static constexpr int SMEM =
      A_STAGES * A_SLOT + B_STAGES * B_HALF + (2 + Y_SLOTS) * STG_BYTES + 1024 + 512 + EP_BYTES;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr int TMEM_COLS = 2 * TILES * BN;   // TILES x BN columns x 2 accumulator stages
};

template <int BN, bool BNB, bool PRO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((H2Cfg<BN, BNB, PRO>::THREADS), 1)
conv3_halo2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapD,
                   const __grid_constant__ CUtensorMap mapY, const Halo2Params p) {
  pdl_wait();
  pdl_trigger();
  using C = H2Cfg<BN, BNB, PRO>;
  constexpr int kH2Threads = C::THREADS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::A_STAGES * C::A_SLOT;
  uint8_t* sStg = sB + C::B_STAGES * C::B_HALF;
  uint8_t* sY = sStg + 2 * C::STG_BYTES;                  // BNB: 2 slots of y tiles (same box / swizzle as mapD)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + C::Y_SLOTS * C::STG_BYTES);
  uint64_t* afull = bars;
  uint64_t* aempty = afull + C::A_STAGES;
  uint64_t* bfull = aempty + C::A_STAGES;
  uint64_t* bempty = bfull + C::B_STAGES;
  uint64_t* tfull = bempty + C::B_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* yfull = tempty + 2;                           // [4], BNB only
  uint64_t* alocal = yfull + 4;                           // [A_STAGES], PRO only: this CTA's own halo tile has landed
  uint64_t* aready = alocal + C::A_STAGES;                // [A_STAGES], PRO only (leader's copy): both tiles transformed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + C::A_STAGES);
  float* sEp = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);      // [2][BN], !BNB only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::A_STAGES; ++i) {
      mbar_init(&afull[i], 1);      // the leader's arrive.expect_tx covers BOTH CTAs' TMA bytes
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < C::B_STAGES; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);     // 4 epilogue warps x 2 CTAs (only the leader's copy is waited on)
    }
    for (int i = 0; i < 4; ++i) mbar_init(&yfull[i], 1);
    for (int i = 0; i < C::A_STAGES; ++i) {
      mbar_init(&alocal[i], 1);
      mbar_init(&aready[i], 2);     // one arrival per CTA of the pair
    }
    fence_barrier_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapD);
    if (BNB) tma_prefetch_desc(&mapY);
  }
  if (!BNB && p.ep_scale != nullptr) {
    const int c0 = ((blockIdx.x >> 1) % p.n_tiles) * BN;
    for (int i = threadIdx.x; i < BN; i += kH2Threads) {
      sEp[i] = __ldg(p.ep_scale + c0 + i);
      sEp[BN + i] = __ldg(p.ep_shift + c0 + i);
    }
  }
  __syncthreads();
  cluster_sync_all();               // both CTAs' barriers exist before anything signals across the pair
  if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = blockIdx.x >> 1;
  const int n_tile = pair % p.n_tiles;
  const int m_first = pair / p.n_tiles;
  const int m_step = (gridDim.x >> 1) / p.n_tiles;
  const int m_pairs = (p.m_blocks + 1) >> 1;
  const int cpt = p.cpt0 + p.cpt1;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (both CTAs)
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int mp = m_first; mp < m_pairs; mp += m_step) {
      const int mb = 2 * mp + (int)rank;           // may be == m_blocks for the odd tail: n lands out of range -> zeros
      const int bx = mb % p.blocks_x;
      const int by = (mb / p.blocks_x) % p.blocks_y;
      const int n = mb / (p.blocks_x * p.blocks_y);
      for (int cc = 0; cc < cpt; ++cc) {
        const CUtensorMap* mapA = (cc < p.cpt0) ? &mapA0 : &mapA1;
        const int c0 = ((cc < p.cpt0) ? cc : cc - p.cpt0) * 64;
        mbar_wait(&aempty[as], aph ^ 1);
        if (elect_one()) {
          if (PRO) {
            // each CTA's tile completes on its own barrier: its transform warps take it from there
            mbar_arrive_expect_tx(&alocal[as], C::A_TX);
            tma_load_5d(sA + as * C::A_SLOT, mapA, &alocal[as], c0, bx * (8 * C::TILES) - 1, by * 16 - 1, n, 0);
          } else {
            // Only the leader arrives.  The peer cannot run a phase ahead: it refills slot s only after the
            // leader's MMAs that consumed the previous contents have committed to its aempty[s].
            if (rank == 0) mbar_arrive_expect_tx(&afull[as], 2 * C::A_TX);
            tma_load_5d_pair(sA + as * C::A_SLOT, mapA, &afull[as], c0, bx * (8 * C::TILES) - 1, by * 16 - 1, n, 0);
          }
        }
        __syncwarp();
        if (++as == C::A_STAGES) {
          as = 0;
          aph ^= 1;
        }
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&bempty[bs], bph ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&bfull[bs], 2 * C::B_HALF);
            tma_load_2d_pair(sB + bs * C::B_HALF, &mapB, &bfull[bs], (tap * cpt + cc) * 64,
                             n_tile * BN + (int)rank * (BN / 2));
          }
          __syncwarp();
          if (++bs == C::B_STAGES) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, false, false);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int mp = m_first; mp < m_pairs; mp += m_step, ++it) {
        const int acs = it & 1;
        const uint32_t acph = (it >> 1) & 1;
        mbar_wait(&tempty[acs], acph ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acs * (C::TILES * BN);
        for (int cc = 0; cc < cpt; ++cc) {
          mbar_wait(PRO ? &aready[as] : &afull[as], aph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + as * C::A_SLOT);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&bfull[bs], bph);
            tc_fence_after_sync();
            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + bs * C::B_HALF), 16, 1024);
            const int r = tap / 3, s = tap - r * 3;
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < C::TILES; ++t) {
                const uint64_t adesc =
                    make_smem_desc_sw128(a_base + (r * C::PITCH + 8 * t + s) * 128, 16, C::PITCH * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_pair(tmem_d + t * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (cc | tap | k) != 0 ? 1u : 0u);
              }
              umma_commit_pair(&bempty[bs]);
            }
            __syncwarp();
            if (++bs == C::B_STAGES) {
              bs = 0;
              bph ^= 1;
            }
          }
          if (elect_one()) umma_commit_pair(&aempty[as]);
          __syncwarp();
          if (++as == C::A_STAGES) {
            as = 0;
            aph ^= 1;
          }
        }
        if (elect_one()) umma_commit_pair(&tfull[acs]);
        __syncwarp();
      }
    }
  } else if (PRO && warp >= 6) {
    // ------------------------------------------------------------- prologue transform (both CTAs, 128 threads each)
    // y -> relu(y * scale + shift), bf16, in place on the landed halo tile: the same fmaf / max / round-to-nearest as
    // bn_relu_flat_kernel, so the MMA operands are bit-identical to the materialised activation.  Pixels outside the
    // image are TMA zero fill = the conv's zero padding and stay zero (relu(shift) != 0 in general).
    const int tt = threadIdx.x - 192;            // 0..127
    const int j = tt & 7;                        // logical 16-byte chunk of a pixel row: channels 8j .. 8j+7
    const int r_first = tt >> 3;                 // rows r_first, r_first + 16, ...
    const int Himg = p.blocks_y * 16, Wimg = p.blocks_x * (8 * C::TILES);
    int as = 0;
    uint32_t aph = 0;
    for (int mp = m_first; mp < m_pairs; mp += m_step) {
      const int mb = 2 * mp + (int)rank;
      const int bx = mb % p.blocks_x;
      const int by = (mb / p.blocks_x) % p.blocks_y;
      const bool block_valid = mb < p.m_blocks;
      for (int cc = 0; cc < cpt; ++cc) {
        const int src = (cc < p.cpt0) ? 0 : 1;
        const bool active = ((p.pro_mask >> src) & 1) && block_valid;
        float sc[8], sh[8];
        if (active) {
          const float4* ps = reinterpret_cast<const float4*>(p.pro_scale + cc * 64 + j * 8);
          const float4* ph = reinterpret_cast<const float4*>(p.pro_shift + cc * 64 + j * 8);
          const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), h0 = __ldg(ph), h1 = __ldg(ph + 1);
          sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
          sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
        }
        mbar_wait(&alocal[as], aph);
        if (active) {
          uint8_t* tile = sA + as * C::A_SLOT;
          // four rows per pass, loads first: the LDS -> FMA -> STS chains of one thread overlap
          for (int rb = r_first; rb < 18 * C::PITCH; rb += 64) {
            uint4 v[4];
            bool on[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rb + 16 * i;
              const int ry = r / C::PITCH, rx = r - ry * C::PITCH;
              const int yy = by * 16 - 1 + ry, xx = bx * (8 * C::TILES) - 1 + rx;
              on[i] = r < 18 * C::PITCH && yy >= 0 && yy < Himg && xx >= 0 && xx < Wimg;
              if (on[i]) v[i] = *reinterpret_cast<const uint4*>(tile + r * 128 + ((j ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!on[i]) continue;
              const int r = rb + 16 * i;
              uint32_t* ww = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float lo = fmaxf(fmaf(bf16lo(ww[k]), sc[2 * k], sh[2 * k]), 0.f);
                const float hi = fmaxf(fmaf(bf16hi(ww[k]), sc[2 * k + 1], sh[2 * k + 1]), 0.f);
                ww[k] = pack_bf16x2(lo, hi);
              }
              *reinterpret_cast<uint4*>(tile + r * 128 + ((j ^ (r & 7)) << 4)) = v[i];
            }
          }
          fence_proxy_async_smem();               // generic-proxy writes -> visible to the tensor core's async proxy
        }
        named_bar_sync(4, 128);
        if (tt == 0) {
          if (rank == 0) mbar_arrive(&aready[as]);
          else mbar_arrive_remote(&aready[as], 0);
        }
        if (++as == C::A_STAGES) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (both CTAs, 128 threads each)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool issuer = (threadIdx.x == 64);
    constexpr int NCHUNK = BN / 64;
    constexpr int CPB = C::TILES * NCHUNK;       // 64-channel output chunks per block
    float ssum[NCHUNK][2], ssq[NCHUNK][2];
#pragma unroll
    for (int q = 0; q < NCHUNK; ++q) ssum[q][0] = ssum[q][1] = ssq[q][0] = ssq[q][1] = 0.f;

    // BNB covers the channel chunks q >= q0 of this CTA's N tile (q0 = 0 unless the launch also produces columns
    // that only need plain statistics, e.g. the [d_up | d_skip] gradient of a decoder block's concat input)
    int q0 = NCHUNK;
    if (BNB) {
      q0 = (p.bnb_col0 - n_tile * BN) / 64;
      q0 = q0 < 0 ? 0 : (q0 > NCHUNK ? NCHUNK : q0);
    }
    const int nqb = NCHUNK - q0;                 // BNB chunks per tile
    // this thread's two columns of every BNB chunk (channel of the BN = output column - bnb_col0)
    float csc[NCHUNK][2], csh[NCHUNK][2], cmu[NCHUNK][2];
    if (BNB) {
#pragma unroll
      for (int q = 0; q < NCHUNK; ++q) {
        const int c = n_tile * BN + q * 64 + lane * 2 - p.bnb_col0;
        const bool on = q >= q0;
        csc[q][0] = on ? __ldg(p.bnb_scale + c) : 0.f;
        csc[q][1] = on ? __ldg(p.bnb_scale + c + 1) : 0.f;
        csh[q][0] = on ? __ldg(p.bnb_shift + c) : 0.f;
        csh[q][1] = on ? __ldg(p.bnb_shift + c + 1) : 0.f;
        cmu[q][0] = on ? __ldg(p.bnb_mean + c) : 0.f;
        cmu[q][1] = on ? __ldg(p.bnb_mean + c + 1) : 0.f;
      }
    }
    // BNB y-tile pipeline, driven by lane 0 of the second epilogue warp (`y_issuer`; the first warp's lane 0 issues
    // the output stores).  BNB chunks are visited in the consumer's order (block, tile, channel chunk q0..NCHUNK-1);
    // chunk k uses slot k % YS.  Loads run YS-1 chunks ahead of the consumer (the slot of chunk k-1 is free once
    // every thread has passed a staging barrier after its stats loop), L2 prefetches kPfAhead chunks further.
    // Coordinates advance incrementally: the first version recomputed them with integer divisions per chunk, which
    // made this one thread the slowest of the epilogue (0.18 ms of a 0.55 ms launch, gpurun_out/bnb_timing*).
    constexpr uint32_t YS = C::Y_SLOTS > 0 ? C::Y_SLOTS : 1;
    constexpr uint32_t YS_LOG = (YS == 4) ? 2 : ((YS == 2) ? 1 : 0);
    constexpr int kPfAhead = 2;
    const bool y_issuer = (threadIdx.x == 96);
    struct YIter {
      int mp, t, q, x0, y0, n;
      bool valid;
    };
    auto y_block = [&](YIter& it) {              // block coordinates of pair item it.mp (this CTA's half)
      const int mb_ = 2 * it.mp + (int)rank;
      it.x0 = (mb_ % p.blocks_x) * (8 * C::TILES);
      it.y0 = ((mb_ / p.blocks_x) % p.blocks_y) * 16;
      it.n = mb_ / (p.blocks_x * p.blocks_y);
    };
    auto y_next = [&](YIter& it) {
      if (++it.q == NCHUNK) {
        it.q = q0;
        if (++it.t == C::TILES) {
          it.t = 0;
          it.mp += m_step;
          it.valid = it.mp < m_pairs;
          if (it.valid) y_block(it);
        }
      }
    };
    YIter yld = {m_first, 0, q0, 0, 0, 0, BNB && nqb > 0 && m_first < m_pairs}, ypf = yld;
    uint32_t y_issued = 0;
    auto y_load_one = [&]() {
      if (!yld.valid) return;
      uint64_t* bar = &yfull[y_issued & (YS - 1)];
      mbar_arrive_expect_tx(bar, C::STG_BYTES);
      tma_load_5d(sY + (y_issued & (YS - 1)) * C::STG_BYTES, &mapY, bar, n_tile * BN + yld.q * 64 - p.bnb_col0,
                  yld.x0 + 8 * yld.t, yld.y0, yld.n, 0);
      ++y_issued;
      y_next(yld);
    };
    auto y_prefetch_one = [&]() {
      if (!ypf.valid) return;
      tma_prefetch_5d(&mapY, n_tile * BN + ypf.q * 64 - p.bnb_col0, ypf.x0 + 8 * ypf.t, ypf.y0, ypf.n, 0);
      y_next(ypf);
    };
    if (BNB && y_issuer && yld.valid) {
      y_block(yld);
      for (uint32_t i = 0; i < YS; ++i) y_load_one();
      ypf = yld;
      for (int i = 0; i < kPfAhead; ++i) y_prefetch_one();
    }
    uint32_t yctr = 0;                           // BNB chunks whose statistics loop this thread has finished

    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int mp = m_first; mp < m_pairs; mp += m_step, ++it) {
      const int acs = it & 1;
      const uint32_t acph = (it >> 1) & 1;
      const int mb = 2 * mp + (int)rank;
      const int bx = mb % p.blocks_x;
      const int by = (mb / p.blocks_x) % p.blocks_y;
      const int n = mb / (p.blocks_x * p.blocks_y);
      mbar_wait(&tfull[acs], acph);
      tc_fence_after_sync();
#pragma unroll
      for (int t = 0; t < C::TILES; ++t) {
#pragma unroll
        for (int q = 0; q < NCHUNK; ++q, ++chunk_ctr) {
          uint32_t v[64];
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acs * (C::TILES * BN) + t * BN + q * 64;
          tmem_ld_32x32b_x32(taddr, v);
          tmem_ld_32x32b_x32(taddr + 32, v + 32);
          tmem_ld_wait();
          const int ncol0 = n_tile * BN + q * 64;
          if (!BNB && p.ep_scale != nullptr) {      // eval-mode BatchNorm + ReLU folded into the store
            const float4* es = reinterpret_cast<const float4*>(sEp + q * 64);
            const float4* eh = reinterpret_cast<const float4*>(sEp + BN + q * 64);
#pragma unroll
            for (int j = 0; j < 16; ++j) {         // broadcast 16-byte reads: every thread needs all 64 columns
              const float4 s4 = es[j], h4 = eh[j];
              v[4 * j + 0] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * j + 0]), s4.x, h4.x), 0.f));
              v[4 * j + 1] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * j + 1]), s4.y, h4.y), 0.f));
              v[4 * j + 2] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * j + 2]), s4.z, h4.z), 0.f));
              v[4 * j + 3] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[4 * j + 3]), s4.w, h4.w), 0.f));
            }
          }
          uint8_t* stg = sStg + (chunk_ctr & 1) * C::STG_BYTES;
          uint4* rowp = reinterpret_cast<uint4*>(stg + row * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            w.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            w.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            w.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            rowp[j ^ (row & 7)] = w;
          }
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1, 128);
          if (issuer) {
            tma_store_5d(&mapD, stg, ncol0, bx * (8 * C::TILES) + 8 * t, by * 16, n, 0);   // out-of-range n: clipped
            tma_store_commit();
          }
          if (BNB && y_issuer) {
            // everyone is past the stats loop of BNB chunk yctr-1: its slot is free for chunk yctr-1+YS
            while (y_issued < yctr + YS && yld.valid) {
              y_load_one();
              y_prefetch_one();
            }
          }
          if (BNB && q >= q0) {
            // (sum g, sum g*(y-mean)) over this quad's 32 rows for this thread's two columns; g = dA where
            // relu(bn(y)) > 0.  dA itself is stored unmasked: the apply kernel masks with the same expression.
            mbar_wait(&yfull[yctr & (YS - 1)], (yctr >> YS_LOG) & 1);
            const uint32_t* words = reinterpret_cast<const uint32_t*>(stg);
            const uint32_t* ywords = reinterpret_cast<const uint32_t*>(sY + (yctr & (YS - 1)) * C::STG_BYTES);
            const float sc0 = csc[q][0], sc1 = csc[q][1], sh0 = csh[q][0], sh1 = csh[q][1];
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;     // sum g, sum g*y (mean folded in after the loop)
#pragma unroll 16
            for (int r = 0; r < 32; ++r) {
              const int rr = quad * 32 + r;
              const int idx = rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3));
              const uint32_t w = words[idx], yw = ywords[idx];
              const float ya = bf16lo(yw), yb = bf16hi(yw);
              const float ga = fmaf(ya, sc0, sh0) > 0.f ? bf16lo(w) : 0.f;
              const float gb = fmaf(yb, sc1, sh1) > 0.f ? bf16hi(w) : 0.f;
              s0 += ga;
              s1 += gb;
              q0 = fmaf(ga, ya, q0);
              q1 = fmaf(gb, yb, q1);
            }
            ssum[q][0] += s0;
            ssum[q][1] += s1;
            ssq[q][0] += fmaf(-cmu[q][0], s0, q0);             // 32-row partial of sum g*(y-mean)
            ssq[q][1] += fmaf(-cmu[q][1], s1, q1);
            ++yctr;
          } else if (p.stats != nullptr) {
            const uint32_t* words = reinterpret_cast<const uint32_t*>(stg);
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int rr = quad * 32 + r;
              const uint32_t w = words[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))];
              const float a = bf16lo(w), b = bf16hi(w);
              s0 += a;
              s1 += b;
              q0 = fmaf(a, a, q0);
              q1 = fmaf(b, b, q1);
            }
            ssum[q][0] += s0;
            ssum[q][1] += s1;
            ssq[q][0] += q0;
            ssq[q][1] += q1;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty[acs]);
        else mbar_arrive_remote(&tempty[acs], 0);
      }
    }
    if (issuer) tma_store_wait_all<0>();
    if (p.stats != nullptr) {
      named_bar_sync(1, 128);
      float* red = reinterpret_cast<float*>(sStg);
#pragma unroll
      for (int q = 0; q < NCHUNK; ++q) {
        const int c = q * 64 + lane * 2;
        red[(quad * BN + c) * 2 + 0] = ssum[q][0];
        red[(quad * BN + c) * 2 + 1] = ssq[q][0];
        red[(quad * BN + c + 1) * 2 + 0] = ssum[q][1];
        red[(quad * BN + c + 1) * 2 + 1] = ssq[q][1];
      }
      named_bar_sync(1, 128);
      const int t = threadIdx.x - 64;
      const int srow = m_first * 2 + (int)rank;      // one partial row per CTA
      float* dst = p.stats + (static_cast<size_t>(srow) * p.n_total + n_tile * BN) * 2;
      for (int i = t; i < BN * 2; i += 128) {
        float v = red[i] + red[BN * 2 + i] + red[2 * BN * 2 + i] + red[3 * BN * 2 + i];
        const int col = n_tile * BN + (i >> 1);
        if (BNB && (i & 1) && col >= p.bnb_col0) v *= __ldg(p.bnb_invstd + col - p.bnb_col0);   // -> sum g*xhat
        dst[i] = v;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();               // nobody leaves (or frees TMEM) while the pair may still touch it
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

static int halo2_map(CUtensorMap* m, const void* base, int C, int S, int B, int H, int W, int bw, int bh) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, 1};
  uint64_t str[4] = {(uint64_t)S * 2, (uint64_t)W * S * 2, (uint64_t)H * W * S * 2, (uint64_t)B * H * W * S * 2};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, 1, 1};
  return make_tmap_bf16_5d(m, base, dims, str, box);
}

template <int BN, bool BNB, bool PRO>
static int halo2_launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d,
                          const CUtensorMap& y, const Halo2Params& p, int grid, cudaStream_t stream) {
  using C = H2Cfg<BN, BNB, PRO>;
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(conv3_halo2_kernel<BN, BNB, PRO>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM),
                       "cudaFuncSetAttribute(conv3_halo2)");
    if (e) return e;
    attr_set = true;
  }
  launch_k(conv3_halo2_kernel<BN, BNB, PRO>, dim3(grid), dim3(C::THREADS), C::SMEM, stream, a0, a1, b, d, y, p);
  return check_launch("conv3_halo2_kernel");
}

bool conv3_halo2_eligible(const sunet_conv_gemm_args* a) {
  static const int mode = [] {
    const char* e = getenv("SUNET_HALO_2CTA");     // default on; SUNET_HALO_2CTA=0 falls back to conv3_halo_kernel
    return e ? atoi(e) : 1;
  }();
  if (!mode) return false;
  const bool ok_n = (a->n_total % 64 == 0) && (a->n_total % 256 != 0 || getenv("SUNET_HALO2_NO256") == nullptr);
  return ok_n && a->a_mode == SUNET_A_CONV3X3 && a->d_mode == SUNET_D_NHWC && a->bias == nullptr &&
         a->height % 16 == 0 && a->width % 16 == 0;
}

static int halo2_bn(int n_total) {
  if (n_total % 256 == 0) return 256;
  return (n_total % 128 == 0) ? 128 : 64;
}

static int halo2_slots(int batch, int height, int width, int n_tiles, int bw) {
  const int m_pairs = (batch * (height / 16) * (width / bw) + 1) / 2;
  int slots = (num_sms() / 2) / n_tiles;
  if (slots < 1) slots = 1;
  if (slots > m_pairs) slots = m_pairs;
  return slots;
}

int conv3_halo2_stat_rows(int batch, int height, int width, int n_total) {
  const int bn = halo2_bn(n_total);
  return 2 * halo2_slots(batch, height, width, n_total / bn, bn == 256 ? 8 : 16);
}

int conv3_halo2_launch(const sunet_conv_gemm_args* a, cudaStream_t stream) {
  const int B = a->batch, H = a->height, W = a->width;
  const int bn = halo2_bn(a->n_total);
  const int bw = (bn == 256) ? 8 : 16;
  CUtensorMap mA0, mA1, mB, mD, mY;
  int e;
  const bool bnb = a->bnb_y != nullptr;
  if (bnb && (!a->stats || !a->bnb_scale || !a->bnb_shift || !a->bnb_mean || !a->bnb_invstd))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: bnb_y needs stats and the four bnb_* vectors");
  if ((e = halo2_map(&mA0, a->src0, a->src0_channels, a->src0_pix_stride, B, H, W, bw + 2, 18))) return e;
  if (a->src1) {
    if ((e = halo2_map(&mA1, a->src1, a->src1_channels, a->src1_pix_stride, B, H, W, bw + 2, 18))) return e;
  } else {
    mA1 = mA0;
  }
  if ((e = make_tmap_bf16_2d(&mB, a->weights, (uint64_t)a->k_total, (uint64_t)a->n_total, (uint64_t)a->k_total * 2,
                             (uint32_t)(bn / 2))))
    return e;
  if ((e = halo2_map(&mD, a->dst, a->n_total, a->dst_pix_stride, B, H, W, 8, 16))) return e;
  if (bnb) {
    const int yc = a->n_total - a->bnb_col0;
    if (a->bnb_col0 < 0 || a->bnb_col0 % 64 || yc <= 0)
      return set_error(SUNET_ERR_INVALID, "conv_gemm: bnb_col0 %d must be a multiple of 64 below n_total", a->bnb_col0);
    if (a->bnb_y_pix_stride < yc || (a->bnb_y_pix_stride % 8))
      return set_error(SUNET_ERR_INVALID, "conv_gemm: bad bnb_y_pix_stride %d", a->bnb_y_pix_stride);
    if ((e = halo2_map(&mY, a->bnb_y, yc, a->bnb_y_pix_stride, B, H, W, 8, 16))) return e;
  } else {
    mY = mD;
  }
  Halo2Params p;
  p.cpt0 = a->src0_channels / 64;
  p.cpt1 = a->src1 ? a->src1_channels / 64 : 0;
  p.blocks_x = W / bw;
  p.blocks_y = H / 16;
  p.m_blocks = B * p.blocks_x * p.blocks_y;
  p.n_tiles = a->n_total / bn;
  p.stats = a->stats;
  p.n_total = a->n_total;
  p.bnb_scale = a->bnb_scale;
  p.bnb_shift = a->bnb_shift;
  p.bnb_mean = a->bnb_mean;
  p.bnb_invstd = a->bnb_invstd;
  p.bnb_col0 = bnb ? a->bnb_col0 : 0;
  p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift;
  p.pro_scale = a->pro_scale;
  p.pro_shift = a->pro_shift;
  p.pro_mask = a->pro_mask;
  const int grid = halo2_slots(B, H, W, p.n_tiles, bw) * p.n_tiles * 2;
  const bool pro = a->pro_scale != nullptr;
  if (pro) {
    if (bnb || a->ep_scale || !a->pro_shift || !(a->pro_mask & 3))
      return set_error(SUNET_ERR_INVALID, "conv_gemm: pro_* needs both vectors, a source mask, and no bnb_* / ep_* epilogue");
    if (bn == 256) return halo2_launch_t<256, false, true>(mA0, mA1, mB, mD, mY, p, grid, stream);
    if (bn == 128) return halo2_launch_t<128, false, true>(mA0, mA1, mB, mD, mY, p, grid, stream);
    return halo2_launch_t<64, false, true>(mA0, mA1, mB, mD, mY, p, grid, stream);
  }
  if (bnb) {
    if (bn == 256) return halo2_launch_t<256, true, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
    if (bn == 128) return halo2_launch_t<128, true, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
    return halo2_launch_t<64, true, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
  }
  if (bn == 256) return halo2_launch_t<256, false, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
  if (bn == 128) return halo2_launch_t<128, false, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
  return halo2_launch_t<64, false, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
}

}  // namespace sunet
