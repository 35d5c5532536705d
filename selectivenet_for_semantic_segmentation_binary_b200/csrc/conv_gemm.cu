// G1 — the tensor-core implicit GEMM behind every "pixels x channels" contraction of
// SUNet_B: conv3x3 forward and dgrad, ConvTranspose2d(k2,s2) forward and dgrad.
//
//   D[pixel, n] = sum_{tap, c} A_tap[pixel, c] * Wp[n, tap*C + c]      (+ bias[n])
//
// * A is never materialised: for every (tap, 64-channel chunk) the producer warp issues one
//   TMA box load straight out of the NHWC bf16 activation tensor.  The 3x3 halo is a
//   coordinate offset, zero padding is TMA out-of-bounds fill, the skip concat is a second
//   tensor map (the decoder conv reads "up" and "skip" from their own buffers), and the 2x2
//   gather of the transposed-conv dgrad is a rank-5 view of the high-resolution tensor.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16; bf16 x bf16 -> fp32 in TMEM).
//   Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// * Epilogue warps: tcgen05.ld -> (+bias) -> bf16 -> 128B-swizzled staging tile -> TMA store
//   (NHWC, or the 2x2 pixel-shuffle scatter of the transposed conv), and the per-channel
//   sum / sum-of-squares BatchNorm needs, accumulated per CTA and written as one
//   deterministic partial row per CTA (no atomics).
//
// Replaces: nn.Conv2d / nn.ConvTranspose2d forward+backward-data as dispatched by
// /root/reference/model.py:11,44-45,51-52,57-58 (cuDNN in the reference).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

struct ConvGemmParams {
  int gs_load;      // 0: conv-type coords (c, x+dx, y+dy, n, 0); 1: gather coords (c, b, x, a, n*H+y)
  int taps;         // 1, 4 or 9
  int cpt0, cpt1;   // 64-channel chunks per tap coming from source 0 / source 1
  int tw, th, nb;   // tile = nb images x th rows x tw pixels (= 128 GEMM rows)
  int tiles_x, tiles_y;
  int H;            // rows per image of the M grid
  int m_tiles, n_tiles;
  int gs_store;     // 0: NHWC store; 1: 2x2 scatter store
  int out_cpt;      // scatter store: 64-channel chunks per (a,b) tap
  const float* bias;
  float* stats;     // [gridDim.x / n_tiles][n_total][2] or nullptr
  int n_total;
  // BNB: fused BatchNorm-backward reduction (see conv3_halo2.cu); NHWC store only
  const float* bnb_scale;
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  const float* ep_scale;     // inference epilogue: dst = relu(acc * ep_scale[n] + ep_shift[n]) (nullptr = off)
  const float* ep_shift;
  int dbg_shift, dbg_boff;  // SUNET_DBG_SHIFT / SUNET_DBG_BOFF: descriptor-swizzle experiment (scripts/gpu_probe.py)
};

template <int BN, bool BNB = false>
struct Cfg {
  // KSP = 64-wide K steps per pipeline stage.  The narrow tiles finish a K step in 128 (BN=64) or 256
  // (BN=128) tensor-core cycles, which is less than one mbarrier round trip of the single issuing thread;
  // two K steps per stage halve the number of round trips.
  static constexpr int KSP = (BN == 256) ? 1 : 2;
  // BNB: two 16 KB slots for y tiles are paid for with one pipeline stage
  static constexpr int STAGES = BNB ? ((BN == 256) ? 3 : 2) : ((BN == 256) ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int Y_SLOTS = BNB ? 2 : 0;
  static constexpr int A_BYTES = 128 * 128;      // 128 pixels x 64 bf16
  static constexpr int B_BYTES = BN * 128;       // BN rows x 64 bf16
  static constexpr int STG_BYTES = 128 * 128;    // one 64-column output chunk
  static constexpr int SMEM =
      STAGES * KSP * (A_BYTES + B_BYTES) + (2 + Y_SLOTS) * STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr int TMEM_COLS = 2 * BN;
};

constexpr int kThreads = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue

template <int BN, bool BNB>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapD,
                 const __grid_constant__ CUtensorMap mapY, const ConvGemmParams p) {
  pdl_wait();
  pdl_trigger();
  using C = Cfg<BN, BNB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::STAGES * C::KSP * C::A_BYTES;
  uint8_t* sStg = sB + C::STAGES * C::KSP * C::B_BYTES;
  uint8_t* sY = sStg + 2 * C::STG_BYTES;          // BNB: 2 slots of y tiles (same box / swizzle as mapD)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + C::Y_SLOTS * C::STG_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;         // [STAGES]
  uint64_t* tfull_bar = bars + 2 * C::STAGES;     // [2]
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
  uint64_t* yfull = bars + 2 * C::STAGES + 4;     // [2], BNB only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
      mbar_init(&yfull[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapD);
    if (BNB) tma_prefetch_desc(&mapY);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Static persistent schedule: CTA b owns N tile (b % n_tiles) and every
  // (gridDim.x / n_tiles)-th M tile, so its BN-statistics accumulators stay on one
  // channel range for the whole kernel.
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles;
  const int m_step = gridDim.x / p.n_tiles;
  const int cpt = p.cpt0 + p.cpt1;
  const int ksteps = p.taps * cpt;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (warp-uniform, elected issue)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = m_first; mt < p.m_tiles; mt += m_step) {
        const int xt = mt % p.tiles_x;
        const int yt = (mt / p.tiles_x) % p.tiles_y;
        const int nt = mt / (p.tiles_x * p.tiles_y);
        const int x0 = xt * p.tw, y0 = yt * p.th, n0 = nt * p.nb;
        for (int ks0 = 0; ks0 < ksteps; ks0 += C::KSP) {
          const int nks = min(C::KSP, ksteps - ks0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const bool leader = elect_one();
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], nks * (C::A_BYTES + C::B_BYTES));
          for (int j = 0; j < nks; ++j) {
            const int ks = ks0 + j;
            const int tap = ks / cpt;
            int cc = ks - tap * cpt;
            const CUtensorMap* mapA = &mapA0;
            if (cc >= p.cpt0) {
              cc -= p.cpt0;
              mapA = &mapA1;
            }
            uint8_t* dA = sA + (stage * C::KSP + j) * C::A_BYTES;
            int dx = 0, dy = 0;
            if (p.taps == 9) {
              dy = tap / 3 - 1;
              dx = tap - (tap / 3) * 3 - 1;
            }
            if (leader) {
              if (p.gs_load) {
                tma_load_5d(dA, mapA, &full_bar[stage], cc * 64, tap & 1, x0, tap >> 1, n0 * p.H + y0);
              } else {
                tma_load_5d(dA, mapA, &full_bar[stage], cc * 64, x0 + dx, y0 + dy, n0, 0);
              }
              tma_load_2d(sB + (stage * C::KSP + j) * C::B_BYTES, &mapB, &full_bar[stage], ks * 64, n_tile * BN);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (warp-uniform, elected issue)
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int mt = m_first; mt < p.m_tiles; mt += m_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int ks0 = 0; ks0 < ksteps; ks0 += C::KSP) {
          const int nks = min(C::KSP, ksteps - ks0);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (elect_one()) {
            for (int j = 0; j < nks; ++j) {
              uint64_t adesc =
                  make_smem_desc_sw128(smem_u32(sA + (stage * C::KSP + j) * C::A_BYTES) + p.dbg_shift * 128, 16, 1024);
              adesc |= static_cast<uint64_t>(p.dbg_boff & 7) << 49;   // experiment hook, 0 in production
              const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + (stage * C::KSP + j) * C::B_BYTES), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // +32 bytes along K inside the 128B swizzle span = +2 in the (addr >> 4) field
                umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (ks0 | j | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[as]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;      // GEMM row (pixel) inside the tile
    const bool issuer = (threadIdx.x == 64);
    constexpr int NCHUNK = BN / 64;
    float ssum[NCHUNK][2], ssq[NCHUNK][2];
#pragma unroll
    for (int q = 0; q < NCHUNK; ++q) ssum[q][0] = ssum[q][1] = ssq[q][0] = ssq[q][1] = 0.f;

    // BNB (fused BN-backward reduction, as in conv3_halo2.cu): this thread's two columns of every chunk, and the
    // y-tile pipeline driven by the issuer thread (chunk k of this CTA = tile k / NCHUNK, channel chunk k % NCHUNK;
    // slot k & 1; loads one chunk ahead, L2 prefetches four more)
    float csc[NCHUNK][2], csh[NCHUNK][2], cmu[NCHUNK][2];
    if (BNB) {
#pragma unroll
      for (int q = 0; q < NCHUNK; ++q) {
        const int c = n_tile * BN + q * 64 + lane * 2;
        csc[q][0] = __ldg(p.bnb_scale + c);
        csc[q][1] = __ldg(p.bnb_scale + c + 1);
        csh[q][0] = __ldg(p.bnb_shift + c);
        csh[q][1] = __ldg(p.bnb_shift + c + 1);
        cmu[q][0] = __ldg(p.bnb_mean + c);
        cmu[q][1] = __ldg(p.bnb_mean + c + 1);
      }
    }
    constexpr int kPfAhead = 4;
    uint32_t y_issued = 0, y_prefetched = 0;
    const uint32_t my_tiles = (m_first < p.m_tiles) ? (uint32_t)((p.m_tiles - 1 - m_first) / m_step + 1) : 0u;
    const uint32_t y_total = my_tiles * NCHUNK;
    auto y_coords = [&](uint32_t k, int& c0, int& xx, int& yy, int& nn) {
      const int mt_ = m_first + (int)(k / NCHUNK) * m_step;
      c0 = n_tile * BN + (int)(k % NCHUNK) * 64;
      xx = (mt_ % p.tiles_x) * p.tw;
      yy = ((mt_ / p.tiles_x) % p.tiles_y) * p.th;
      nn = (mt_ / (p.tiles_x * p.tiles_y)) * p.nb;
    };
    auto y_pump = [&](uint32_t upto) {
      while (y_issued < upto && y_issued < y_total) {
        int c0, xx, yy, nn;
        y_coords(y_issued, c0, xx, yy, nn);
        uint64_t* bar = &yfull[y_issued & 1];
        mbar_arrive_expect_tx(bar, C::STG_BYTES);
        tma_load_5d(sY + (y_issued & 1) * C::STG_BYTES, &mapY, bar, c0, xx, yy, nn, 0);
        ++y_issued;
      }
      if (y_prefetched < y_issued) y_prefetched = y_issued;
      while (y_prefetched < y_issued + kPfAhead && y_prefetched < y_total) {
        int c0, xx, yy, nn;
        y_coords(y_prefetched, c0, xx, yy, nn);
        tma_prefetch_5d(&mapY, c0, xx, yy, nn, 0);
        ++y_prefetched;
      }
    };
    if (BNB && issuer) y_pump(2);

    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int mt = m_first; mt < p.m_tiles; mt += m_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int xt = mt % p.tiles_x;
      const int yt = (mt / p.tiles_x) % p.tiles_y;
      const int nt = mt / (p.tiles_x * p.tiles_y);
      const int x0 = xt * p.tw, y0 = yt * p.th, n0 = nt * p.nb;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
#pragma unroll
      for (int q = 0; q < NCHUNK; ++q, ++chunk_ctr) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + q * 64;
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_32x32b_x32(taddr + 32, v + 32);
        tmem_ld_wait();
        const int ncol0 = n_tile * BN + q * 64;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + ncol0 + j));
        }
        if (!BNB && p.ep_scale != nullptr) {      // eval-mode BatchNorm + ReLU folded into the store
#pragma unroll
          for (int j = 0; j < 64; ++j)
            v[j] = __float_as_uint(fmaxf(
                fmaf(__uint_as_float(v[j]), __ldg(p.ep_scale + ncol0 + j), __ldg(p.ep_shift + ncol0 + j)), 0.f));
        }
        uint8_t* stg = sStg + (chunk_ctr & 1) * C::STG_BYTES;
        // 128B-swizzled staging row (matches the TMA store map): 16B chunk j lands at j ^ (row & 7)
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
        if (issuer) tma_store_wait_read<0>();  // previous chunk's store no longer reads the other buffer
        named_bar_sync(1, 128);
        if (issuer) {
          if (p.gs_store) {
            const int oc = (ncol0 >> 6);
            const int tap = oc / p.out_cpt;
            const int c0 = (oc - tap * p.out_cpt) * 64;
            tma_store_5d(&mapD, stg, c0, tap & 1, x0, tap >> 1, n0 * p.H + y0);
          } else {
            tma_store_5d(&mapD, stg, ncol0, x0, y0, n0, 0);
          }
          tma_store_commit();
          if (BNB) y_pump(chunk_ctr + 2);        // everyone is past the stats loop of chunk_ctr-1: its slot is free
        }
        if (BNB) {
          mbar_wait(&yfull[chunk_ctr & 1], (chunk_ctr >> 1) & 1);
          const uint32_t* words = reinterpret_cast<const uint32_t*>(stg);
          const uint32_t* ywords = reinterpret_cast<const uint32_t*>(sY + (chunk_ctr & 1) * C::STG_BYTES);
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
          ssq[q][0] += fmaf(-cmu[q][0], s0, q0);
          ssq[q][1] += fmaf(-cmu[q][1], s1, q1);
        } else if (p.stats != nullptr) {
          // Column statistics over this warp's own 32 rows (rows it wrote itself -> the
          // named barrier above already ordered the writes).  Lane = channel pair.
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
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (issuer) tma_store_wait_all<0>();
    if (p.stats != nullptr) {
      // combine the 4 row-quadrants through (now idle) staging memory; one partial row per CTA
      named_bar_sync(1, 128);
      float* red = reinterpret_cast<float*>(sStg);  // [4][BN][2]
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
      float* dst = p.stats + (static_cast<size_t>(m_first) * p.n_total + n_tile * BN) * 2;
      for (int i = t; i < BN * 2; i += 128) {
        float v = red[i] + red[BN * 2 + i] + red[2 * BN * 2 + i] + red[3 * BN * 2 + i];
        if (BNB && (i & 1)) v *= __ldg(p.bnb_invstd + n_tile * BN + (i >> 1));     // sum g*(y-mean) -> sum g*xhat
        dst[i] = v;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side

static int pow2_floor_div(int v, int cap) {
  // largest power of two that divides v, capped at cap
  int t = 1;
  while (t < cap && (v % (t * 2)) == 0) t *= 2;
  return t;
}

struct TileGeom {
  int tw, th, nb, tiles_x, tiles_y, tiles_n, m_tiles;
};
static TileGeom tile_geom(int B, int H, int W, int rows_per_tile) {
  TileGeom g;
  g.tw = pow2_floor_div(W, rows_per_tile);
  g.th = pow2_floor_div(H, rows_per_tile / g.tw);
  g.nb = rows_per_tile / (g.tw * g.th);
  if (g.nb > 1 && g.th != H) {
    // images may only be stacked when a tile covers whole images
    g.nb = 1;
  }
  g.tiles_x = W / g.tw;
  g.tiles_y = H / g.th;
  g.tiles_n = (B + g.nb - 1) / g.nb;
  g.m_tiles = g.tiles_x * g.tiles_y * g.tiles_n;
  return g;
}

static int make_act_map(CUtensorMap* m, const void* base, int gs, int C, int S, int B, int H, int W, const TileGeom& g,
                        int rows_per_tile) {
  // gs == 0: NHWC tensor [B][H][W][S] viewed as (C, W, H, B, 1), box (64, tw, th, nb, 1)
  // gs == 1: high-res tensor [B][2H][2W][S] viewed as (C, 2, W, 2, B*H), box (64, 1, tw, 1, th*nb)
  uint64_t dims[5], str[4];
  uint32_t box[5];
  const uint64_t e = 2;
  if (!gs) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = B; dims[4] = 1;
    str[0] = (uint64_t)S * e;
    str[1] = (uint64_t)W * S * e;
    str[2] = (uint64_t)H * W * S * e;
    str[3] = (uint64_t)B * H * W * S * e;
    box[0] = 64; box[1] = g.tw; box[2] = g.th; box[3] = g.nb; box[4] = 1;
  } else {
    dims[0] = C; dims[1] = 2; dims[2] = W; dims[3] = 2; dims[4] = (uint64_t)B * H;
    str[0] = (uint64_t)S * e;
    str[1] = (uint64_t)2 * S * e;
    str[2] = (uint64_t)2 * W * S * e;
    str[3] = (uint64_t)4 * W * S * e;
    box[0] = 64; box[1] = 1; box[2] = g.tw; box[3] = 1; box[4] = g.th * g.nb;
  }
  (void)rows_per_tile;
  return make_tmap_bf16_5d(m, base, dims, str, box);
}

static int pick_bn(int n_total) {
  if (n_total % 256 == 0) return 256;
  if (n_total % 128 == 0) return 128;
  return 64;
}

static int conv_gemm_grid(int m_tiles, int n_tiles) {
  int slots = num_sms() / n_tiles;
  if (slots < 1) slots = 1;
  if (slots > m_tiles) slots = m_tiles;
  return slots * n_tiles;
}

template <int BN, bool BNB>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d,
                  const CUtensorMap& y, const ConvGemmParams& p, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(conv_gemm_kernel<BN, BNB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg<BN, BNB>::SMEM),
                       "cudaFuncSetAttribute(conv_gemm)");
    if (e) return e;
    attr_set = true;
  }
  launch_k(conv_gemm_kernel<BN, BNB>, dim3(grid), dim3(kThreads), Cfg<BN, BNB>::SMEM, stream, a0, a1, b, d, y, p);
  return check_launch("conv_gemm_kernel");
}

}  // namespace sunet

using namespace sunet;

namespace sunet {
bool conv3_halo_eligible(const sunet_conv_gemm_args* a);
int conv3_halo_stat_rows(int batch, int height, int width, int n_total);
int conv3_halo_launch(const sunet_conv_gemm_args* a, cudaStream_t stream);
bool conv3_halo2_eligible(const sunet_conv_gemm_args* a);
int conv3_halo2_stat_rows(int batch, int height, int width, int n_total);
int conv3_halo2_launch(const sunet_conv_gemm_args* a, cudaStream_t stream);
}  // namespace sunet

extern "C" int sunet_conv_gemm_stat_rows(const sunet_conv_gemm_args* a) {
  if (!a) return -1;
  const int batch = a->batch, height = a->height, width = a->width, n_total = a->n_total;
  if (batch <= 0 || height <= 0 || width <= 0 || n_total <= 0 || n_total % 64) return -1;
  if (conv3_halo2_eligible(a)) return conv3_halo2_stat_rows(batch, height, width, n_total);
  if (conv3_halo_eligible(a)) return conv3_halo_stat_rows(batch, height, width, n_total);
  TileGeom g = tile_geom(batch, height, width, 128);
  const int bn = pick_bn(n_total);
  const int n_tiles = n_total / bn;
  return conv_gemm_grid(g.m_tiles, n_tiles) / n_tiles;
}

extern "C" int sunet_conv_gemm_bnb_supported(const sunet_conv_gemm_args* a) {
  if (!a || a->batch <= 0 || a->height <= 0 || a->width <= 0 || a->n_total <= 0 || a->n_total % 64) return 0;
  if (conv3_halo2_eligible(a)) return 1;
  // the per-tap kernel implements it for its 128- and 256-wide tiles with a plain NHWC store (ConvT backward-data)
  return (!conv3_halo_eligible(a) && a->d_mode == SUNET_D_NHWC && a->bias == nullptr && a->n_total % 128 == 0) ? 1 : 0;
}

extern "C" int sunet_conv_gemm_pro_supported(const sunet_conv_gemm_args* a) {
  if (!a || a->batch <= 0 || a->height <= 0 || a->width <= 0 || a->n_total <= 0 || a->n_total % 64) return 0;
  return conv3_halo2_eligible(a) ? 1 : 0;          // the prologue transform lives in the CTA-pair halo kernel
}

extern "C" int sunet_conv_gemm(const sunet_conv_gemm_args* a, sunet_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a) return set_error(SUNET_ERR_INVALID, "conv_gemm: null args");
  const int B = a->batch, H = a->height, W = a->width;
  if (B <= 0 || H <= 0 || W <= 0) return set_error(SUNET_ERR_INVALID, "conv_gemm: bad grid %d x %d x %d", B, H, W);
  if (!a->src0 || !a->weights || !a->dst) return set_error(SUNET_ERR_INVALID, "conv_gemm: null tensor");
  if (a->src0_channels <= 0 || a->src0_channels % 64 || a->src1_channels % 64 || a->n_total <= 0 || a->n_total % 64)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: channel counts must be multiples of 64 (got %d,%d -> %d)",
                     a->src0_channels, a->src1_channels, a->n_total);
  if (a->src1 == nullptr && a->src1_channels != 0)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: src1_channels without src1");
  int taps;
  switch (a->a_mode) {
    case SUNET_A_CONV3X3: taps = 9; break;
    case SUNET_A_PLAIN: taps = 1; break;
    case SUNET_A_GATHER2X2: taps = 4; break;
    default: return set_error(SUNET_ERR_INVALID, "conv_gemm: bad a_mode %d", a->a_mode);
  }
  if (a->a_mode == SUNET_A_GATHER2X2 && a->src1)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: gather mode takes one source");
  const int ctot = a->src0_channels + (a->src1 ? a->src1_channels : 0);
  if (a->k_total != taps * ctot)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: k_total %d != taps %d x channels %d", a->k_total, taps, ctot);
  if (a->src0_pix_stride < a->src0_channels || (a->src0_pix_stride % 8) ||
      (a->src1 && (a->src1_pix_stride < a->src1_channels || (a->src1_pix_stride % 8))) || (a->dst_pix_stride % 8))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: pixel strides must be >= channels and multiples of 8");
  if (a->d_mode == SUNET_D_SCATTER2X2 && (a->n_total % 4 || (a->n_total / 4) % 64))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: scatter store needs n_total = 4 x (multiple of 64)");

  if ((a->ep_scale == nullptr) != (a->ep_shift == nullptr))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: ep_scale and ep_shift go together");
  if (a->ep_scale && (a->stats || a->bias || a->bnb_y || a->d_mode != SUNET_D_NHWC))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: the inference epilogue excludes stats / bias / bnb / scatter store");
  if (conv3_halo2_eligible(a)) return conv3_halo2_launch(a, stream);
  if (a->pro_scale != nullptr)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: the training prologue (pro_scale) is not available for this "
                                        "shape/mode; check sunet_conv_gemm_pro_supported()");
  const bool bnb = a->bnb_y != nullptr;
  if (bnb && !sunet_conv_gemm_bnb_supported(a))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: the fused BN-backward epilogue (bnb_y) is not available for this "
                                        "shape/mode; check sunet_conv_gemm_bnb_supported()");
  if (bnb && (!a->stats || !a->bnb_scale || !a->bnb_shift || !a->bnb_mean || !a->bnb_invstd))
    return set_error(SUNET_ERR_INVALID, "conv_gemm: bnb_y needs stats and the four bnb_* vectors");
  if (bnb && a->bnb_col0 != 0)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: bnb_col0 != 0 is implemented by the CTA-pair halo kernel only");
  if (conv3_halo_eligible(a)) return conv3_halo_launch(a, stream);

  TileGeom g = tile_geom(B, H, W, 128);
  if (g.tw * g.th * g.nb != 128)
    return set_error(SUNET_ERR_INVALID, "conv_gemm: cannot tile %d x %d x %d into 128-pixel tiles", B, H, W);
  const int bn = pick_bn(a->n_total);
  const int n_tiles = a->n_total / bn;

  CUtensorMap mA0, mA1, mB, mD, mY;
  int e;
  const int gs_load = (a->a_mode == SUNET_A_GATHER2X2);
  if ((e = make_act_map(&mA0, a->src0, gs_load, a->src0_channels, a->src0_pix_stride, B, H, W, g, 128))) return e;
  if (a->src1) {
    if ((e = make_act_map(&mA1, a->src1, 0, a->src1_channels, a->src1_pix_stride, B, H, W, g, 128))) return e;
  } else {
    mA1 = mA0;
  }
  if ((e = make_tmap_bf16_2d(&mB, a->weights, (uint64_t)a->k_total, (uint64_t)a->n_total, (uint64_t)a->k_total * 2,
                             (uint32_t)bn)))
    return e;
  const int gs_store = (a->d_mode == SUNET_D_SCATTER2X2);
  const int dst_c = gs_store ? a->n_total / 4 : a->n_total;
  if (a->dst_pix_stride < dst_c) return set_error(SUNET_ERR_INVALID, "conv_gemm: dst_pix_stride < channels");
  if ((e = make_act_map(&mD, a->dst, gs_store, dst_c, a->dst_pix_stride, B, H, W, g, 128))) return e;
  if (bnb) {
    if (a->bnb_y_pix_stride < a->n_total || (a->bnb_y_pix_stride % 8))
      return set_error(SUNET_ERR_INVALID, "conv_gemm: bad bnb_y_pix_stride %d", a->bnb_y_pix_stride);
    if ((e = make_act_map(&mY, a->bnb_y, 0, a->n_total, a->bnb_y_pix_stride, B, H, W, g, 128))) return e;
  } else {
    mY = mD;
  }

  ConvGemmParams p;
  p.gs_load = gs_load;
  p.taps = taps;
  p.cpt0 = a->src0_channels / 64;
  p.cpt1 = a->src1 ? a->src1_channels / 64 : 0;
  p.tw = g.tw; p.th = g.th; p.nb = g.nb;
  p.tiles_x = g.tiles_x; p.tiles_y = g.tiles_y;
  p.H = H;
  p.m_tiles = g.m_tiles;
  p.n_tiles = n_tiles;
  p.gs_store = gs_store;
  p.out_cpt = dst_c / 64;
  p.bias = a->bias;
  p.stats = a->stats;
  p.n_total = a->n_total;
  p.bnb_scale = a->bnb_scale;
  p.bnb_shift = a->bnb_shift;
  p.bnb_mean = a->bnb_mean;
  p.bnb_invstd = a->bnb_invstd;
  p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift;
  {
    const char* s = getenv("SUNET_DBG_SHIFT");
    const char* b = getenv("SUNET_DBG_BOFF");
    p.dbg_shift = s ? atoi(s) : 0;
    p.dbg_boff = b ? atoi(b) : 0;
  }
  const int grid = conv_gemm_grid(g.m_tiles, n_tiles);
  if (bnb) return (bn == 256) ? launch<256, true>(mA0, mA1, mB, mD, mY, p, grid, stream)
                              : launch<128, true>(mA0, mA1, mB, mD, mY, p, grid, stream);
  switch (bn) {
    case 256: return launch<256, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
    case 128: return launch<128, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
    default: return launch<64, false>(mA0, mA1, mB, mD, mY, p, grid, stream);
  }
}
