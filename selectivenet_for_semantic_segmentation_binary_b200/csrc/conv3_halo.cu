// G1h — conv3x3 (forward and dgrad) with ONE halo tile per 64-channel chunk shared by all nine taps.
//
// conv_gemm.cu loads a fresh 128-pixel A tile for every (tap, chunk): nine L2->SMEM transfers of
// the same pixels, which makes every conv3x3 launch bound by the ~64 B/clk/SM the SM can pull from L2
// (profiles/r01: 510 / 1030 / 1450 TFLOP/s for N = 64 / 128 / 256 follow that model exactly).  Here a CTA
// owns a 16x16 output block (two M=128 tiles: left and right 8 columns), TMA-loads its 18x18 halo once
// per chunk, and addresses tap (r,s) of tile t as a *shifted window* of that halo:
//     descriptor start = halo + ((r*18) + 8t + s) * 128 B,   stride between 8-row groups = 18 * 128 B
// (tcgen05.mma swizzles on absolute shared-memory address bits, so any 128-byte-aligned start works —
// scripts/gpu_probe.py swizzle_exp).  A traffic drops 9x -> 1.27x of the block, and each weight tile
// (one tap x 64 channels x BN outputs) now feeds 256 pixels instead of 128.
//
// Pipelines: A ring (2 halo slots), B ring (weight tiles), TMEM double buffer (2 tiles x BN x 2).
// Epilogue, BN statistics and scheduling are the same as conv_gemm.cu.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

struct HaloParams {
  int cpt0, cpt1;            // 64-channel chunks from source 0 / 1
  int blocks_x, blocks_y;    // 16x16 blocks per image
  int m_blocks, n_tiles;
  float* stats;              // [gridDim.x / n_tiles][n_total][2] or nullptr
  int n_total;
  const float* ep_scale;     // inference epilogue: dst = relu(acc * ep_scale[n] + ep_shift[n]) (nullptr = off)
  const float* ep_shift;
};

// TILES = M=128 tiles per block: 2 (16x16 block, BN <= 128) or 1 (8 wide x 16 tall block, BN = 256 —
// TMEM holds TILES x BN x 2 accumulator stages = 512 columns either way).
template <int BN>
struct HCfg {
  static constexpr int TILES = (BN == 256) ? 1 : 2;
  static constexpr int PITCH = 8 * TILES + 2;                  // halo row pitch in pixels (18 or 10)
  static constexpr int A_TX = 18 * PITCH * 128;                // 18 halo rows
  static constexpr int A_SLOT = (A_TX + 1023) / 1024 * 1024;   // padded to the 1024-byte swizzle atom
  static constexpr int A_STAGES = 2;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int B_STAGES = (BN == 256) ? 4 : ((BN == 128) ? 5 : 8);
  static constexpr int STG_BYTES = 128 * 128;
  static constexpr int SMEM = A_STAGES * A_SLOT + B_STAGES * B_BYTES + 2 * STG_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * TILES * BN;
};

constexpr int kHThreads = 192;

template <int BN>
__global__ void __launch_bounds__(kHThreads, 1)
conv3_halo_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapD,
                  const HaloParams p) {
  pdl_wait();
  pdl_trigger();
  using C = HCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + C::A_STAGES * C::A_SLOT;
  uint8_t* sStg = sB + C::B_STAGES * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + 2 * C::STG_BYTES);
  uint64_t* afull = bars;                                 // [A_STAGES]
  uint64_t* aempty = afull + C::A_STAGES;
  uint64_t* bfull = aempty + C::A_STAGES;                 // [B_STAGES]
  uint64_t* bempty = bfull + C::B_STAGES;
  uint64_t* tfull = bempty + C::B_STAGES;                 // [2]
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::A_STAGES; ++i) {
      mbar_init(&afull[i], 1);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < C::B_STAGES; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapD);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles;
  const int m_step = gridDim.x / p.n_tiles;
  const int cpt = p.cpt0 + p.cpt1;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int mb = m_first; mb < p.m_blocks; mb += m_step) {
        const int bx = mb % p.blocks_x;
        const int by = (mb / p.blocks_x) % p.blocks_y;
        const int n = mb / (p.blocks_x * p.blocks_y);
        for (int cc = 0; cc < cpt; ++cc) {
          const CUtensorMap* mapA = (cc < p.cpt0) ? &mapA0 : &mapA1;
          const int c0 = ((cc < p.cpt0) ? cc : cc - p.cpt0) * 64;
          mbar_wait(&aempty[as], aph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&afull[as], C::A_TX);
            tma_load_5d(sA + as * C::A_SLOT, mapA, &afull[as], c0, bx * (8 * C::TILES) - 1, by * 16 - 1, n, 0);
          }
          __syncwarp();
          if (++as == C::A_STAGES) {
            as = 0;
            aph ^= 1;
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&bempty[bs], bph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&bfull[bs], C::B_BYTES);
              tma_load_2d(sB + bs * C::B_BYTES, &mapB, &bfull[bs], (tap * cpt + cc) * 64, n_tile * BN);
            }
            __syncwarp();
            if (++bs == C::B_STAGES) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int mb = m_first; mb < p.m_blocks; mb += m_step, ++it) {
        const int acs = it & 1;
        const uint32_t acph = (it >> 1) & 1;
        mbar_wait(&tempty[acs], acph ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acs * (C::TILES * BN);
        for (int cc = 0; cc < cpt; ++cc) {
          mbar_wait(&afull[as], aph);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + as * C::A_SLOT);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&bfull[bs], bph);
            tc_fence_after_sync();
            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + bs * C::B_BYTES), 16, 1024);
            const int r = tap / 3, s = tap - r * 3;
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < C::TILES; ++t) {
                const uint64_t adesc =
                    make_smem_desc_sw128(a_base + (r * C::PITCH + 8 * t + s) * 128, 16, C::PITCH * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem_d + t * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (cc | tap | k) != 0 ? 1u : 0u);
              }
              umma_commit(&bempty[bs]);
            }
            __syncwarp();
            if (++bs == C::B_STAGES) {
              bs = 0;
              bph ^= 1;
            }
          }
          if (elect_one()) umma_commit(&aempty[as]);
          __syncwarp();
          if (++as == C::A_STAGES) {
            as = 0;
            aph ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull[acs]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (128 threads)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // row of the 128-pixel tile: (ty = row / 8, tx = row % 8)
    const bool issuer = (threadIdx.x == 64);
    constexpr int NCHUNK = BN / 64;
    float ssum[NCHUNK][2], ssq[NCHUNK][2];
#pragma unroll
    for (int q = 0; q < NCHUNK; ++q) ssum[q][0] = ssum[q][1] = ssq[q][0] = ssq[q][1] = 0.f;

    int it = 0;
    uint32_t chunk_ctr = 0;
    for (int mb = m_first; mb < p.m_blocks; mb += m_step, ++it) {
      const int acs = it & 1;
      const uint32_t acph = (it >> 1) & 1;
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
          if (p.ep_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              v[j] = __float_as_uint(fmaxf(
                  fmaf(__uint_as_float(v[j]), __ldg(p.ep_scale + ncol0 + j), __ldg(p.ep_shift + ncol0 + j)), 0.f));
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
            tma_store_5d(&mapD, stg, ncol0, bx * (8 * C::TILES) + 8 * t, by * 16, n, 0);
            tma_store_commit();
          }
          if (p.stats != nullptr) {
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
      if (lane == 0) mbar_arrive(&tempty[acs]);
    }
    if (issuer) tma_store_wait_all<0>();
    if (p.stats != nullptr) {
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
      for (int i = t; i < BN * 2; i += 128) dst[i] = red[i] + red[BN * 2 + i] + red[2 * BN * 2 + i] + red[3 * BN * 2 + i];
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

static int halo_map(CUtensorMap* m, const void* base, int C, int S, int B, int H, int W, int bw, int bh) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B, 1};
  uint64_t str[4] = {(uint64_t)S * 2, (uint64_t)W * S * 2, (uint64_t)H * W * S * 2, (uint64_t)B * H * W * S * 2};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, 1, 1};
  return make_tmap_bf16_5d(m, base, dims, str, box);
}

template <int BN>
static int halo_launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& d,
                         const HaloParams& p, int grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    int e = check_cuda(cudaFuncSetAttribute(conv3_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            HCfg<BN>::SMEM),
                       "cudaFuncSetAttribute(conv3_halo)");
    if (e) return e;
    attr_set = true;
  }
  launch_k(conv3_halo_kernel<BN>, dim3(grid), dim3(kHThreads), HCfg<BN>::SMEM, stream, a0, a1, b, d, p);
  return check_launch("conv3_halo_kernel");
}

static int halo_bn(int n_total) {
  // N % 256 == 0: one 128-pixel tile x 256 channels per block (two 128-wide tiles measured slower than
  // the per-tap 128x256 kernel; the single-tile halo variant keeps the wide N and still loads A once)
  if (n_total % 256 == 0 && getenv("SUNET_HALO_NO256") == nullptr) return 256;
  return (n_total % 128 == 0) ? 128 : 64;
}

bool conv3_halo_eligible(const sunet_conv_gemm_args* a) {
  if (a->n_total % 256 == 0 && getenv("SUNET_HALO_NO256") != nullptr) return false;   // fall back to conv_gemm<256>
  return a->a_mode == SUNET_A_CONV3X3 && a->d_mode == SUNET_D_NHWC && a->bias == nullptr && a->height % 16 == 0 &&
         a->width % 16 == 0 && getenv("SUNET_NO_HALO") == nullptr;
}

int conv3_halo_stat_rows(int batch, int height, int width, int n_total) {
  const int bn = halo_bn(n_total);
  const int n_tiles = n_total / bn;
  const int m_blocks = batch * (height / 16) * (width / (bn == 256 ? 8 : 16));
  int slots = num_sms() / n_tiles;
  if (slots < 1) slots = 1;
  if (slots > m_blocks) slots = m_blocks;
  return slots;
}

int conv3_halo_launch(const sunet_conv_gemm_args* a, cudaStream_t stream) {
  const int B = a->batch, H = a->height, W = a->width;
  const int bn = halo_bn(a->n_total);
  const int bw = (bn == 256) ? 8 : 16;           // block width in pixels
  CUtensorMap mA0, mA1, mB, mD;
  int e;
  if ((e = halo_map(&mA0, a->src0, a->src0_channels, a->src0_pix_stride, B, H, W, bw + 2, 18))) return e;
  if (a->src1) {
    if ((e = halo_map(&mA1, a->src1, a->src1_channels, a->src1_pix_stride, B, H, W, bw + 2, 18))) return e;
  } else {
    mA1 = mA0;
  }
  if ((e = make_tmap_bf16_2d(&mB, a->weights, (uint64_t)a->k_total, (uint64_t)a->n_total, (uint64_t)a->k_total * 2,
                             (uint32_t)bn)))
    return e;
  if ((e = halo_map(&mD, a->dst, a->n_total, a->dst_pix_stride, B, H, W, 8, 16))) return e;
  HaloParams p;
  p.cpt0 = a->src0_channels / 64;
  p.cpt1 = a->src1 ? a->src1_channels / 64 : 0;
  p.blocks_x = W / bw;
  p.blocks_y = H / 16;
  p.m_blocks = B * p.blocks_x * p.blocks_y;
  p.n_tiles = a->n_total / bn;
  p.stats = a->stats;
  p.n_total = a->n_total;
  p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift;
  const int slots = conv3_halo_stat_rows(B, H, W, a->n_total);
  const int grid = slots * p.n_tiles;
  if (bn == 256) return halo_launch_t<256>(mA0, mA1, mB, mD, p, grid, stream);
  if (bn == 128) return halo_launch_t<128>(mA0, mA1, mB, mD, p, grid, stream);
  return halo_launch_t<64>(mA0, mA1, mB, mD, p, grid, stream);
}

}  // namespace sunet
