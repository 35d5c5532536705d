// Heads (three 1x1 convs 64 -> 1), the SelectiveNet losses, the thresholding + confusion
// matrix counters, and the Adam step.  HBM-bound or tiny; warp-shuffle reductions, two-stage
// deterministic sums for floating point, integer atomics only for exact counts.
//
// Reference semantics:
//   heads   /root/reference/model.py:62,65-66,96,99-101
//   losses  /root/reference/selective_loss.py:58-85, train.py:78,195-201 (BCEWithLogitsLoss)
//   metrics /root/reference/train.py:211-239, eval.py:228-251, utils/compute_metric.py:10-26
//   Adam    /root/reference/train.py:88-92,209 (torch.optim.Adam defaults)
#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

static inline int grid_for(long long items, int threads, int per_sm) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static inline bool aligned16(const void* a, const void* b, const void* c, const void* d) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

struct alignas(16) bf16x8 {
  uint32_t w[4];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// stable BCE-with-logits: max(x,0) - x*t + log1p(exp(-|x|))   (== torch.nn.BCEWithLogitsLoss)
__device__ __forceinline__ float bce_logits(float x, float t) {
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }
// Phase-1 reductions (sums over millions of pixels, consumed at 1e-5 relative): MUFU-based forms, ~28 instructions
// per pixel instead of ~95 — the precise expf / log1pf / division made loss_sums ALU-bound at 58 % of the HBM
// roofline (profiles/r02/ew_bw_stream_kernels.log).  Per-pixel absolute error <= ~2e-7 (ex2.approx / lg2.approx / rcp),
// i.e. <= 4e-7 of a mean loss of O(0.5); the per-pixel GRADIENTS (loss_bwd) keep the precise forms.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float bce_logits_fast(float x, float t) {
  const float z = __expf(-fabsf(x));
  const float l1p = z < 1e-4f ? z * fmaf(-0.5f, z, 1.f) : __logf(1.f + z);
  return fmaxf(x, 0.f) - x * t + l1p;
}

// ------------------------------------------------------------------ heads forward
struct HeadW {
  const float* w[3];
  const float* b[3];
};

__global__ void __launch_bounds__(256)
heads_fwd_kernel(const __nv_bfloat16* __restrict__ a, int as, HeadW hw, int nheads, float* __restrict__ logits,
                 long long P) {
  pdl_wait();
  pdl_trigger();
  // 8 lanes per pixel, 8 channels per lane
  const int sub = threadIdx.x & 7;
  float w[3][8];
  float bias[3];
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    bias[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) w[h][j] = 0.f;
    if (h < nheads) {
      bias[h] = __ldg(hw.b[h]);
#pragma unroll
      for (int j = 0; j < 8; ++j) w[h][j] = __ldg(hw.w[h] + sub * 8 + j);
    }
  }
  const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
  // loop bound is warp-uniform (4 pixels per warp per trip) so the full-mask shuffles are legal
  for (long long p0 = blockIdx.x * (long long)(blockDim.x >> 3) + ((threadIdx.x >> 5) << 2); p0 < P; p0 += stride) {
    const long long p = p0 + ((threadIdx.x & 31) >> 3);
    const bool valid = p < P;
    bf16x8 v = {{0u, 0u, 0u, 0u}};
    if (valid) v = *reinterpret_cast<const bf16x8*>(a + p * as + sub * 8);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = bf16lo(v.w[i]), hi = bf16hi(v.w[i]);
#pragma unroll
      for (int h = 0; h < 3; ++h) acc[h] = fmaf(lo, w[h][2 * i], fmaf(hi, w[h][2 * i + 1], acc[h]));
    }
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 1);
      acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 2);
      acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 4);
    }
    if (valid && sub < nheads) {
      const float r = sub == 0 ? acc[0] + bias[0] : (sub == 1 ? acc[1] + bias[1] : acc[2] + bias[2]);
      logits[(long long)sub * P + p] = r;
    }
  }
}

// ------------------------------------------------------------------ BN + ReLU of the last block fused with the heads
// a = relu(y*scale + shift) is written once (backward needs it) and the three 64->1 dots are taken from
// the registers that just produced it: saves re-reading the 8.4 MB/patch activation (heads_fwd_kernel).
__global__ void __launch_bounds__(256)
bn_relu_heads_kernel(const __nv_bfloat16* __restrict__ y, int ys, const float* __restrict__ scale,
                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ a, int as, HeadW hw, int nheads,
                     float* __restrict__ logits, long long P) {
  pdl_wait();
  pdl_trigger();
  const int sub = threadIdx.x & 7;
  float w[3][8], bias[3], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + sub * 8 + j);
    sh[j] = __ldg(shift + sub * 8 + j);
  }
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    bias[h] = (h < nheads) ? __ldg(hw.b[h]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) w[h][j] = (h < nheads) ? __ldg(hw.w[h] + sub * 8 + j) : 0.f;
  }
  constexpr int U = 4;
  const long long ppb = (long long)(blockDim.x >> 3);       // pixels per block per sub-trip (32)
  // trip count is uniform across the block (bounds depend on blockIdx only), so full-mask shuffles are legal
  for (long long base = blockIdx.x * ppb * U; base < P; base += (long long)gridDim.x * ppb * U) {
    bf16x8 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = base + u * ppb + (threadIdx.x >> 3);
      v[u] = bf16x8{{0u, 0u, 0u, 0u}};
      if (p < P) v[u] = *reinterpret_cast<const bf16x8*>(y + p * ys + sub * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = base + u * ppb + (threadIdx.x >> 3);
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[2 * i] = fmaxf(fmaf(bf16lo(v[u].w[i]), sc[2 * i], sh[2 * i]), 0.f);
        f[2 * i + 1] = fmaxf(fmaf(bf16hi(v[u].w[i]), sc[2 * i + 1], sh[2 * i + 1]), 0.f);
      }
      bf16x8 o;
#pragma unroll
      for (int i = 0; i < 4; ++i) o.w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      if (a != nullptr && p < P) *reinterpret_cast<bf16x8*>(a + p * as + sub * 8) = o;
      float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // the heads read the bf16-rounded activation, exactly as the un-fused path does
        const float lo = bf16lo(o.w[i]), hi = bf16hi(o.w[i]);
#pragma unroll
        for (int h = 0; h < 3; ++h) acc[h] = fmaf(lo, w[h][2 * i], fmaf(hi, w[h][2 * i + 1], acc[h]));
      }
#pragma unroll
      for (int h = 0; h < 3; ++h) {
        acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 1);
        acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 2);
        acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 4);
      }
      if (p < P && sub < nheads) {
        const float r = sub == 0 ? acc[0] + bias[0] : (sub == 1 ? acc[1] + bias[1] : acc[2] + bias[2]);
        logits[(long long)sub * P + p] = r;
      }
    }
  }
}

// ------------------------------------------------------------------ heads backward
// BN = true: `a` holds y, the raw conv output of the last block; the activation relu(bn(y)) is recomputed (bit-
// identical to what bn_relu_heads produced, which then need not be stored at all), and the BatchNorm-backward
// reduction of that block is accumulated on the way: bnb[block][64][2] = (sum g, sum g*xhat) with
// g = dA (bf16-rounded, as stored) where the activation is > 0.
struct HeadBN {
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  float* partials;     // [gridDim.x][64][2]
  const __nv_bfloat16* addend;   // optional: gradient of the same activation from heads handled by an earlier call
  int adds;                      // its pixel stride (may alias dA: each element is read and then written by one thread)
};

// Thread mapping: 16 lanes per pixel, 4 channels per lane (one 8-byte load), 16 pixels per block and trip, U trips in
// flight.  With 4 channels per lane the 24 per-channel constants (3 head weights, scale, shift, mean) and the 23
// accumulators fit in registers.  The first version used 8 lanes x 8 channels and re-read the constants from shared
// memory for every pixel because 48 constants + 43 accumulators did not fit: 24 broadcast loads per lane and pixel put it
// on the shared-memory port (ncu, profiles/r02/ncu_full_small_kernels.md: LSU wavefronts 82-83 % of peak, DRAM 31 %,
// 0.85-0.89 ms for 2.25 GB), and wider loads do not help — a broadcast only merges lanes of one 128-byte wavefront.
// Now 0.75 ms, LSU wavefronts 4 %, issue slots 59 % busy at 16 resident warps; 2 pixels in flight with 24 resident warps
// (80 registers, small spills) measured slower (0.98 ms).
template <bool BN>
__global__ void __launch_bounds__(256, 2)
heads_bwd_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ a, int as, HeadW hw, int nheads,
                 __nv_bfloat16* __restrict__ dA, int das, float* __restrict__ partials, long long P, HeadBN bn) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[256][16];
  const int sub = threadIdx.x & 15;           // channels [4*sub, 4*sub + 4)
  const int pslot = threadIdx.x >> 4;         // pixel slot of this thread within a 16-pixel trip
  float w[3][4], sc[4], sh[4], mu[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 3; ++h) w[h][j] = (h < nheads) ? __ldg(hw.w[h] + sub * 4 + j) : 0.f;
    sc[j] = BN ? __ldg(bn.scale + sub * 4 + j) : 0.f;
    sh[j] = BN ? __ldg(bn.shift + sub * 4 + j) : 0.f;
    mu[j] = BN ? __ldg(bn.mean + sub * 4 + j) : 0.f;
  }
  float sg[4], sgy[4], dw[3][4];
  float db[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sg[j] = sgy[j] = 0.f;
#pragma unroll
    for (int h = 0; h < 3; ++h) dw[h][j] = 0.f;
  }
  // U pixels per thread per trip: all loads are issued before any is consumed
  constexpr int U = 4;
  constexpr long long ppb = 16;                                       // pixels per block per sub-trip
  for (long long p0 = blockIdx.x * ppb * U + pslot; p0 < P; p0 += (long long)gridDim.x * ppb * U) {
    float g[U][3];
    uint2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * ppb;
      if (p < P) {
#pragma unroll
        for (int h = 0; h < 3; ++h) g[u][h] = (h < nheads) ? __ldg(dl + (long long)h * P + p) : 0.f;
        v[u] = *reinterpret_cast<const uint2*>(a + p * as + sub * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + u * ppb;
      if (p < P) {
        uint32_t ow[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {           // channel pair i of this thread's 4 channels
          const int c = sub * 4 + 2 * i;
          const uint32_t vw = i ? v[u].y : v[u].x;
          float a0 = bf16lo(vw), a1 = bf16hi(vw);
          float y0 = 0.f, y1 = 0.f;
          if (BN) {
            y0 = a0;
            y1 = a1;
            // the activation bn_relu_heads fed to the heads: relu(bn(y)) rounded to bf16
            const uint32_t r = pack_bf16x2(fmaxf(fmaf(y0, sc[2 * i], sh[2 * i]), 0.f),
                                           fmaxf(fmaf(y1, sc[2 * i + 1], sh[2 * i + 1]), 0.f));
            a0 = bf16lo(r);
            a1 = bf16hi(r);
          }
          float o0 = g[u][0] * w[0][2 * i] + g[u][1] * w[1][2 * i] + g[u][2] * w[2][2 * i];
          float o1 = g[u][0] * w[0][2 * i + 1] + g[u][1] * w[1][2 * i + 1] + g[u][2] * w[2][2 * i + 1];
          if (BN && bn.addend != nullptr) {     // more than three heads (UNet, n_cls = 2): second call adds the first
            const uint32_t aw = *reinterpret_cast<const uint32_t*>(bn.addend + p * bn.adds + c);
            o0 += bf16lo(aw);
            o1 += bf16hi(aw);
          }
#pragma unroll
          for (int h = 0; h < 3; ++h) {
            dw[h][2 * i] = fmaf(g[u][h], a0, dw[h][2 * i]);
            dw[h][2 * i + 1] = fmaf(g[u][h], a1, dw[h][2 * i + 1]);
          }
          ow[i] = pack_bf16x2(o0, o1);
          if (BN) {
            const float g0 = a0 > 0.f ? bf16lo(ow[i]) : 0.f;
            const float g1 = a1 > 0.f ? bf16hi(ow[i]) : 0.f;
            sg[2 * i] += g0;
            sg[2 * i + 1] += g1;
            sgy[2 * i] = fmaf(g0, y0 - mu[2 * i], sgy[2 * i]);
            sgy[2 * i + 1] = fmaf(g1, y1 - mu[2 * i + 1], sgy[2 * i + 1]);
          }
        }
        *reinterpret_cast<uint2*>(dA + p * das + sub * 4) = make_uint2(ow[0], ow[1]);
        if (sub == 0) {
#pragma unroll
          for (int h = 0; h < 3; ++h) db[h] += g[u][h];
        }
      }
    }
  }
  // block reduction over the 16 pixel slots (fixed order): red[t][h*4 + j] = dw, red[t][12 + h] = db
#pragma unroll
  for (int h = 0; h < 3; ++h) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][h * 4 + j] = dw[h][j];
    red[threadIdx.x][12 + h] = db[h];
  }
  __syncthreads();
  // partial row layout per block: [3][65] = 64 weight grads + bias grad
  for (int o = threadIdx.x; o < 3 * 65; o += 256) {
    const int h = o / 65, c = o % 65;
    float s = 0.f;
    if (c < 64) {
      const int sb = c >> 2, j = c & 3;
      for (int r = 0; r < 16; ++r) s += red[r * 16 + sb][h * 4 + j];
    } else {
      for (int r = 0; r < 16; ++r) s += red[r * 16][12 + h];
    }
    partials[(size_t)blockIdx.x * 195 + o] = s;
  }
  if (BN) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[threadIdx.x][j] = sg[j];
      red[threadIdx.x][4 + j] = sgy[j];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int c = threadIdx.x >> 1, which = threadIdx.x & 1;       // channel, (sum g | sum g*(y-mean))
      const int sb = c >> 2, j = c & 3;
      float s = 0.f;
      for (int r = 0; r < 16; ++r) s += red[r * 16 + sb][which * 4 + j];
      if (which) s *= __ldg(bn.invstd + c);
      bn.partials[((size_t)blockIdx.x * 64 + c) * 2 + which] = s;
    }
  }
}

struct HeadG {
  float* dw[3];
  float* db[3];
};
// one block per output (3 heads x (64 weight grads + 1 bias grad)); the per-block partial rows are summed in a
// fixed order (strided per thread, then a shared-memory tree), so the result does not depend on scheduling
__global__ void __launch_bounds__(128)
heads_bwd_reduce_kernel(const float* __restrict__ partials, int blocks, int nheads, HeadG hg) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[128];
  const int o = blockIdx.x;
  const int h = o / 65, c = o % 65;
  if (h >= nheads) return;
  double s = 0.0;
  for (int b = threadIdx.x; b < blocks; b += 128) s += (double)partials[(size_t)b * 195 + o];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (c < 64) {
      if (hg.dw[h]) hg.dw[h][c] = (float)red[0];
    } else {
      if (hg.db[h]) hg.db[h][0] = (float)red[0];
    }
  }
}

// ------------------------------------------------------------------ losses
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streaming kernels below read 16 bytes per array per thread per trip (float4 / uchar4 / 2 x longlong2) when every
// pointer is 16-byte aligned (VEC = 4; the launcher falls back to VEC = 1 otherwise); the last pixels % 4 are taken
// by the first threads of block 0 as scalars.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int VEC>
__global__ void __launch_bounds__(256)
loss_sums_kernel(const float* __restrict__ out, const float* __restrict__ sel, const float* __restrict__ aux,
                 const float* __restrict__ tgt, long long P, double* __restrict__ partials) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][3];
  float S = 0.f, R = 0.f, A = 0.f;
  auto pixel = [&](float t, float xs, float xo, float xa) {
    if (sel) {
      const float s = sigmoid_fast(xs);
      S += s;
      if (out) R = fmaf(bce_logits_fast(xo, t), s, R);
    }
    if (aux) A += bce_logits_fast(xa, t);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (VEC == 4) {
    const long long nq = P >> 2;
    for (long long q = tid; q < nq; q += stride) {
      const float4 t = ldg4(tgt + 4 * q);
      float4 xs = make_float4(0.f, 0.f, 0.f, 0.f), xo = xs, xa = xs;
      if (sel) xs = ldg4(sel + 4 * q);
      if (out) xo = ldg4(out + 4 * q);
      if (aux) xa = ldg4(aux + 4 * q);
      pixel(t.x, xs.x, xo.x, xa.x);
      pixel(t.y, xs.y, xo.y, xa.y);
      pixel(t.z, xs.z, xo.z, xa.z);
      pixel(t.w, xs.w, xo.w, xa.w);
    }
    const long long p = (nq << 2) + tid;          // tail (at most 3 pixels)
    if (p < P) pixel(__ldg(tgt + p), sel ? __ldg(sel + p) : 0.f, out ? __ldg(out + p) : 0.f, aux ? __ldg(aux + p) : 0.f);
  } else {
    for (long long p = tid; p < P; p += stride)
      pixel(__ldg(tgt + p), sel ? __ldg(sel + p) : 0.f, out ? __ldg(out + p) : 0.f, aux ? __ldg(aux + p) : 0.f);
  }
  S = warp_sum(S);
  R = warp_sum(R);
  A = warp_sum(A);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = S;
    red[warp][1] = R;
    red[warp][2] = A;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
  }
}
// fixed-order fold of the per-block partials: thread t sums rows t, t+256, ... then a shared-memory tree
__global__ void __launch_bounds__(256) loss_sums_final_kernel(const double* __restrict__ partials, int blocks,
                                                              double* sums, double* pixels_out, double pixels) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[3][256];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int b = threadIdx.x; b < blocks; b += 256) {
    s0 += partials[(size_t)b * 3 + 0];
    s1 += partials[(size_t)b * 3 + 1];
    s2 += partials[(size_t)b * 3 + 2];
  }
  red[0][threadIdx.x] = s0;
  red[1][threadIdx.x] = s1;
  red[2][threadIdx.x] = s2;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      red[0][threadIdx.x] += red[0][threadIdx.x + w];
      red[1][threadIdx.x] += red[1][threadIdx.x + w];
      red[2][threadIdx.x] += red[2][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) sums[threadIdx.x] = red[threadIdx.x][0];
  if (threadIdx.x == 3 && pixels_out) *pixels_out = pixels;
}
__global__ void loss_finalize_kernel(const double* sums, double P, const double* P_dev, float lamb, float tc,
                                     float* results) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (P_dev) P = *P_dev;
    const double S = sums[0], R = sums[1], A = sums[2];
    const double c = S / P;
    const double risk = (S != 0.0) ? R / S : 0.0 / 0.0;
    const double d = ((double)tc - c) > 0.0 ? ((double)tc - c) : 0.0;
    const double sl = risk + (double)lamb * d * d;
    const double al = A / P;
    results[0] = (float)sl;
    results[1] = (float)c;
    results[2] = (float)al;
    results[3] = (float)(sl + al);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ out, const float* __restrict__ sel, const float* __restrict__ aux,
                const float* __restrict__ tgt, long long P, const double* __restrict__ sums, double Pg,
                const double* __restrict__ Pg_dev, float lamb, float tc, const float* __restrict__ g_sel,
                const float* __restrict__ g_aux, float* __restrict__ d_out, float* __restrict__ d_sel,
                float* __restrict__ d_aux) {
  pdl_wait();
  pdl_trigger();
  if (Pg_dev) Pg = *Pg_dev;
  const double S = sums[0], R = sums[1];
  const float gs = g_sel ? *g_sel : 1.f;
  const float ga = g_aux ? *g_aux : 1.f;
  const double c = S / Pg;
  const double d = ((double)tc - c) > 0.0 ? ((double)tc - c) : 0.0;
  const float invS = (float)(1.0 / S);
  const float k_const = (float)(-R / (S * S) - 2.0 * (double)lamb * d / Pg);  // dL/ds_i minus the l_i/S term
  const float invP = (float)(1.0 / Pg);
  const bool sel_path = d_out || d_sel;
  auto f_aux = [&](float xa, float t) { return ga * (sigmoid_acc(xa) - t) * invP; };
  auto f_sel = [&](float xs, float xo, float t, float& go, float& gsl) {
    const float s = sigmoid_acc(xs);
    go = gs * s * (sigmoid_acc(xo) - t) * invS;
    gsl = gs * s * (1.f - s) * fmaf(bce_logits(xo, t), invS, k_const);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long p0 = tid;
  if (VEC == 4) {
    const long long nq = P >> 2;
    for (long long q = tid; q < nq; q += stride) {
      const float4 t = ldg4(tgt + 4 * q);
      if (d_aux) {
        const float4 xa = ldg4(aux + 4 * q);
        *reinterpret_cast<float4*>(d_aux + 4 * q) = make_float4(f_aux(xa.x, t.x), f_aux(xa.y, t.y), f_aux(xa.z, t.z),
                                                                f_aux(xa.w, t.w));
      }
      if (sel_path) {
        const float4 xs = ldg4(sel + 4 * q), xo = ldg4(out + 4 * q);
        float4 go, gl;
        f_sel(xs.x, xo.x, t.x, go.x, gl.x);
        f_sel(xs.y, xo.y, t.y, go.y, gl.y);
        f_sel(xs.z, xo.z, t.z, go.z, gl.z);
        f_sel(xs.w, xo.w, t.w, go.w, gl.w);
        if (d_out) *reinterpret_cast<float4*>(d_out + 4 * q) = go;
        if (d_sel) *reinterpret_cast<float4*>(d_sel + 4 * q) = gl;
      }
    }
    p0 = (nq << 2) + tid;                         // tail (at most 3 pixels)
  }
  for (long long p = p0; p < P; p += stride) {
    const float t = __ldg(tgt + p);
    if (d_aux) d_aux[p] = f_aux(__ldg(aux + p), t);
    if (sel_path) {
      float go, gl;
      f_sel(__ldg(sel + p), __ldg(out + p), t, go, gl);
      if (d_out) d_out[p] = go;
      if (d_sel) d_sel[p] = gl;
    }
  }
}

// ------------------------------------------------------------------ metrics
template <int LT>
__device__ __forceinline__ int load_label(const void* label, long long p) {
  if (LT == 0) return (int)reinterpret_cast<const uint8_t*>(label)[p];
  if (LT == 1) return (int)(uint8_t)(int)reinterpret_cast<const float*>(label)[p];  // .astype('uint8')
  const long long v = reinterpret_cast<const long long*>(label)[p];
  return (v >= 0 && v < 256) ? (int)v : 255;
}
template <int LT>
__device__ __forceinline__ void load_label4(const void* label, long long q, int lab[4]) {
  if (LT == 0) {
    const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(label) + q);
    lab[0] = v.x; lab[1] = v.y; lab[2] = v.z; lab[3] = v.w;
  } else if (LT == 1) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(label) + q);
    lab[0] = (int)(uint8_t)(int)v.x; lab[1] = (int)(uint8_t)(int)v.y;
    lab[2] = (int)(uint8_t)(int)v.z; lab[3] = (int)(uint8_t)(int)v.w;
  } else {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(label) + 2 * q);
    const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(label) + 2 * q + 1);
    lab[0] = (a.x >= 0 && a.x < 256) ? (int)a.x : 255; lab[1] = (a.y >= 0 && a.y < 256) ? (int)a.y : 255;
    lab[2] = (b.x >= 0 && b.x < 256) ? (int)b.x : 255; lab[3] = (b.y >= 0 && b.y < 256) ? (int)b.y : 255;
  }
}

// Coalesced, vectorised integer histogram: 4 pixels per thread per load (two groups in flight), six register
// counters per thread, one warp reduction + six 64-bit atomics per block at the end (integers: order-independent).
template <int LT, int VEC>
__global__ void __launch_bounds__(256)
metric_hist_kernel(const float* __restrict__ out, const float* __restrict__ sel, const void* __restrict__ label,
                   long long P, float thr_out, float thr_sel, int masked, unsigned long long* __restrict__ counts) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned int red[8][6];
  unsigned int c00 = 0, c01 = 0, c10 = 0, c11 = 0, csel = 0, ctot = 0;
  auto pixel = [&](float xo, float xs, int lab) {
    const bool pred = xo >= thr_out;
    const bool selected = sel ? (xs >= thr_sel) : true;
    ctot += 1;
    csel += selected ? 1u : 0u;
    const bool use = (!masked || selected);
    c00 += (use && lab == 0 && !pred) ? 1u : 0u;
    c01 += (use && lab == 0 && pred) ? 1u : 0u;
    c10 += (use && lab == 1 && !pred) ? 1u : 0u;
    c11 += (use && lab == 1 && pred) ? 1u : 0u;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long p0 = tid;
  if (VEC == 4) {
    const long long nq = P >> 2;
    for (long long q = tid; q < nq; q += 2 * stride) {
      const long long q2 = q + stride;
      const bool two = q2 < nq;
      int la[4], lb[4] = {0, 0, 0, 0};
      const float4 oa = ldg4(out + 4 * q);
      float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), ob = sa, sb = sa;
      if (sel) sa = ldg4(sel + 4 * q);
      load_label4<LT>(label, q, la);
      if (two) {
        ob = ldg4(out + 4 * q2);
        if (sel) sb = ldg4(sel + 4 * q2);
        load_label4<LT>(label, q2, lb);
      }
      pixel(oa.x, sa.x, la[0]);
      pixel(oa.y, sa.y, la[1]);
      pixel(oa.z, sa.z, la[2]);
      pixel(oa.w, sa.w, la[3]);
      if (two) {
        pixel(ob.x, sb.x, lb[0]);
        pixel(ob.y, sb.y, lb[1]);
        pixel(ob.z, sb.z, lb[2]);
        pixel(ob.w, sb.w, lb[3]);
      }
    }
    p0 = (nq << 2) + tid;                         // tail (at most 3 pixels)
  }
  for (long long p = p0; p < P; p += stride)
    pixel(__ldg(out + p), sel ? __ldg(sel + p) : 0.f, load_label<LT>(label, p));
  const unsigned int c[6] = {c00, c01, c10, c11, csel, ctot};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    unsigned int v = __reduce_add_sync(0xffffffffu, c[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (s) atomicAdd(counts + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256)
adam_kernel(const sunet_adam_tensor* __restrict__ table, float lr, float b1, float b2, float eps, float wd,
            int step, const float* __restrict__ lr_dev, const int* __restrict__ step_dev) {
  pdl_wait();
  pdl_trigger();
  const sunet_adam_tensor t = table[blockIdx.y];
  // hyper-parameters that change per step may live on the device so that a captured CUDA graph
  // of the whole training step stays valid across replays
  if (lr_dev) lr = *lr_dev;
  if (step_dev) step = *step_dev;
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = lr / bc1;
  auto upd = [&](float g, float& p, float& m, float& v) {
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long i0 = tid;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                     reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15) == 0;
  if (vec) {
    const long long nq = t.numel >> 2;
    for (long long q = tid; q < nq; q += stride) {
      const float4 g = *reinterpret_cast<const float4*>(t.grad + 4 * q);
      float4 p = *reinterpret_cast<float4*>(t.param + 4 * q);
      float4 m = *reinterpret_cast<float4*>(t.exp_avg + 4 * q);
      float4 v = *reinterpret_cast<float4*>(t.exp_avg_sq + 4 * q);
      upd(g.x, p.x, m.x, v.x);
      upd(g.y, p.y, m.y, v.y);
      upd(g.z, p.z, m.z, v.z);
      upd(g.w, p.w, m.w, v.w);
      *reinterpret_cast<float4*>(t.param + 4 * q) = p;
      *reinterpret_cast<float4*>(t.exp_avg + 4 * q) = m;
      *reinterpret_cast<float4*>(t.exp_avg_sq + 4 * q) = v;
    }
    i0 = (nq << 2) + tid;
  }
  for (long long i = i0; i < t.numel; i += stride) {
    float p = t.param[i], m = t.exp_avg[i], v = t.exp_avg_sq[i];
    upd(t.grad[i], p, m, v);
    t.param[i] = p;
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
  }
}

}  // namespace sunet

using namespace sunet;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int sunet_heads_fwd(const void* a, int a_pix_stride, const float* w0, const float* b0, const float* w1,
                               const float* b1, const float* w2, const float* b2, int nheads, float* logits,
                               long long pixels, sunet_stream_t stream_) {
  if (!a || !logits || pixels <= 0 || (nheads != 1 && nheads != 3) || a_pix_stride < 64 || a_pix_stride % 8)
    return set_error(SUNET_ERR_INVALID, "heads_fwd: bad arguments");
  HeadW hw = {{w0, w1, w2}, {b0, b1, b2}};
  for (int h = 0; h < nheads; ++h)
    if (!hw.w[h] || !hw.b[h]) return set_error(SUNET_ERR_INVALID, "heads_fwd: missing head %d parameters", h);
  launch_k(heads_fwd_kernel, dim3(grid_for(pixels * 8, 256, 8)), dim3(256), 0, STREAM, reinterpret_cast<const __nv_bfloat16*>(a),
                                                                      a_pix_stride, hw, nheads, logits, pixels);
  return check_launch("heads_fwd");
}

extern "C" int sunet_bn_relu_heads(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                                   int a_pix_stride, const float* w0, const float* b0, const float* w1,
                                   const float* b1, const float* w2, const float* b2, int nheads, float* logits,
                                   long long pixels, sunet_stream_t stream_) {
  if (!y || !scale || !shift || !logits || pixels <= 0 || (nheads != 1 && nheads != 3) || y_pix_stride < 64 ||
      y_pix_stride % 8 || (a && (a_pix_stride < 64 || a_pix_stride % 8)))
    return set_error(SUNET_ERR_INVALID, "bn_relu_heads: bad arguments");
  HeadW hw = {{w0, w1, w2}, {b0, b1, b2}};
  for (int h = 0; h < nheads; ++h)
    if (!hw.w[h] || !hw.b[h]) return set_error(SUNET_ERR_INVALID, "bn_relu_heads: missing head %d parameters", h);
  launch_k(bn_relu_heads_kernel, dim3(grid_for(pixels * 2, 256, 8)), dim3(256), 0, STREAM, 
      reinterpret_cast<const __nv_bfloat16*>(y), y_pix_stride, scale, shift, reinterpret_cast<__nv_bfloat16*>(a),
      a_pix_stride, hw, nheads, logits, pixels);
  return check_launch("bn_relu_heads");
}

static int heads_bwd_blocks(long long pixels, bool bn) {
  (void)bn;
  return grid_for(pixels * 16, 256, 2);      // 16 lanes per pixel; 2 resident blocks per SM (128 registers)
}

extern "C" int sunet_heads_bwd(const float* dlogits, const void* a, int a_pix_stride, const float* w0, const float* w1,
                               const float* w2, int nheads, void* dA, int dA_pix_stride, float* dw0, float* db0,
                               float* dw1, float* db1, float* dw2, float* db2, long long pixels, void* workspace,
                               size_t workspace_bytes, sunet_stream_t stream_) {
  if (!dlogits || !a || !dA || !workspace || pixels <= 0 || (nheads != 1 && nheads != 3) || a_pix_stride < 64 ||
      a_pix_stride % 8 || dA_pix_stride < 64 || dA_pix_stride % 8)
    return set_error(SUNET_ERR_INVALID, "heads_bwd: bad arguments");
  HeadW hw = {{w0, w1, w2}, {nullptr, nullptr, nullptr}};
  for (int h = 0; h < nheads; ++h)
    if (!hw.w[h]) return set_error(SUNET_ERR_INVALID, "heads_bwd: missing head %d weights", h);
  int blocks = heads_bwd_blocks(pixels, false);
  const size_t need = (size_t)blocks * 195 * sizeof(float);
  if (workspace_bytes < need) return set_error(SUNET_ERR_WORKSPACE, "heads_bwd: workspace %zu < %zu", workspace_bytes, need);
  float* partials = reinterpret_cast<float*>(workspace);
  HeadBN nobn = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  launch_k(heads_bwd_kernel<false>, dim3(blocks), dim3(256), 0, STREAM, dlogits,
           reinterpret_cast<const __nv_bfloat16*>(a), a_pix_stride, hw, nheads, reinterpret_cast<__nv_bfloat16*>(dA),
           dA_pix_stride, partials, pixels, nobn);
  int e = check_launch("heads_bwd");
  if (e) return e;
  HeadG hg = {{dw0, dw1, dw2}, {db0, db1, db2}};
  launch_k(heads_bwd_reduce_kernel, dim3(3 * 65), dim3(128), 0, STREAM, partials, blocks, nheads, hg);
  return check_launch("heads_bwd_reduce");
}

extern "C" int sunet_heads_bwd_bn_rows(long long pixels) {
  return pixels > 0 ? heads_bwd_blocks(pixels, true) : -1;
}

extern "C" int sunet_heads_bwd_bn(const float* dlogits, const void* y, int y_pix_stride, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, const float* w0,
                                  const float* w1, const float* w2, int nheads, void* dA, int dA_pix_stride,
                                  float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2,
                                  float* bn_partials, const void* addend, int addend_pix_stride, long long pixels,
                                  void* workspace, size_t workspace_bytes, sunet_stream_t stream_) {
  if (!dlogits || !y || !dA || !workspace || !scale || !shift || !mean || !invstd || !bn_partials || pixels <= 0 ||
      (nheads != 1 && nheads != 3) || y_pix_stride < 64 || y_pix_stride % 8 || dA_pix_stride < 64 ||
      dA_pix_stride % 8)
    return set_error(SUNET_ERR_INVALID, "heads_bwd_bn: bad arguments");
  HeadW hw = {{w0, w1, w2}, {nullptr, nullptr, nullptr}};
  for (int h = 0; h < nheads; ++h)
    if (!hw.w[h]) return set_error(SUNET_ERR_INVALID, "heads_bwd_bn: missing head %d weights", h);
  int blocks = heads_bwd_blocks(pixels, true);
  const size_t need = (size_t)blocks * 195 * sizeof(float);
  if (workspace_bytes < need)
    return set_error(SUNET_ERR_WORKSPACE, "heads_bwd_bn: workspace %zu < %zu", workspace_bytes, need);
  float* partials = reinterpret_cast<float*>(workspace);
  if (addend && (addend_pix_stride < 64 || addend_pix_stride % 8))
    return set_error(SUNET_ERR_INVALID, "heads_bwd_bn: bad addend stride");
  HeadBN bn = {scale, shift, mean, invstd, bn_partials, reinterpret_cast<const __nv_bfloat16*>(addend),
               addend_pix_stride};
  launch_k(heads_bwd_kernel<true>, dim3(blocks), dim3(256), 0, STREAM, dlogits,
           reinterpret_cast<const __nv_bfloat16*>(y), y_pix_stride, hw, nheads, reinterpret_cast<__nv_bfloat16*>(dA),
           dA_pix_stride, partials, pixels, bn);
  int e = check_launch("heads_bwd_bn");
  if (e) return e;
  HeadG hg = {{dw0, dw1, dw2}, {db0, db1, db2}};
  launch_k(heads_bwd_reduce_kernel, dim3(3 * 65), dim3(128), 0, STREAM, partials, blocks, nheads, hg);
  return check_launch("heads_bwd_reduce");
}

extern "C" int sunet_loss_sums(const float* out, const float* sel, const float* aux, const float* target,
                               long long pixels, double* sums, double* pixels_out, void* workspace,
                               size_t workspace_bytes, sunet_stream_t stream_) {
  if (!target || !sums || !workspace || pixels <= 0) return set_error(SUNET_ERR_INVALID, "loss_sums: bad arguments");
  if (out && !sel) return set_error(SUNET_ERR_INVALID, "loss_sums: out given without sel");
  const bool vec = aligned16(out, sel, aux, target);
  const int blocks = grid_for(vec ? (pixels + 3) / 4 : pixels, 256, 8);
  if (workspace_bytes < (size_t)blocks * 3 * sizeof(double))
    return set_error(SUNET_ERR_WORKSPACE, "loss_sums: workspace too small");
  double* partials = reinterpret_cast<double*>(workspace);
  if (vec) launch_k(loss_sums_kernel<4>, dim3(blocks), dim3(256), 0, STREAM, out, sel, aux, target, pixels, partials);
  else launch_k(loss_sums_kernel<1>, dim3(blocks), dim3(256), 0, STREAM, out, sel, aux, target, pixels, partials);
  int e = check_launch("loss_sums");
  if (e) return e;
  launch_k(loss_sums_final_kernel, dim3(1), dim3(256), 0, STREAM, (const double*)partials, blocks, sums, pixels_out,
           (double)pixels);
  return check_launch("loss_sums_final");
}

extern "C" int sunet_loss_finalize(const double* sums, long long global_pixels, const double* global_pixels_dev,
                                   float lamb, float target_coverage, float* results, sunet_stream_t stream_) {
  if (!sums || !results || (global_pixels <= 0 && !global_pixels_dev))
    return set_error(SUNET_ERR_INVALID, "loss_finalize: bad arguments");
  launch_k(loss_finalize_kernel, dim3(1), dim3(32), 0, STREAM, sums, (double)global_pixels, global_pixels_dev, lamb,
           target_coverage, results);
  return check_launch("loss_finalize");
}

extern "C" int sunet_loss_bwd(const float* out, const float* sel, const float* aux, const float* target,
                              long long pixels, const double* sums, long long global_pixels,
                              const double* global_pixels_dev, float lamb, float target_coverage, const float* g_sel,
                              const float* g_aux, float* d_out, float* d_sel, float* d_aux, sunet_stream_t stream_) {
  if (!target || !sums || pixels <= 0 || (global_pixels <= 0 && !global_pixels_dev))
    return set_error(SUNET_ERR_INVALID, "loss_bwd: bad arguments");
  if ((d_out || d_sel) && (!out || !sel)) return set_error(SUNET_ERR_INVALID, "loss_bwd: d_out/d_sel need out and sel");
  if (d_aux && !aux) return set_error(SUNET_ERR_INVALID, "loss_bwd: d_aux needs aux");
  const bool vec = aligned16(out, sel, aux, target) && aligned16(d_out, d_sel, d_aux, nullptr);
  const int blocks = grid_for(vec ? (pixels + 3) / 4 : pixels, 256, 8);
  if (vec)
    launch_k(loss_bwd_kernel<4>, dim3(blocks), dim3(256), 0, STREAM, out, sel, aux, target, pixels, sums,
             (double)global_pixels, global_pixels_dev, lamb, target_coverage, g_sel, g_aux, d_out, d_sel, d_aux);
  else
    launch_k(loss_bwd_kernel<1>, dim3(blocks), dim3(256), 0, STREAM, out, sel, aux, target, pixels, sums,
             (double)global_pixels, global_pixels_dev, lamb, target_coverage, g_sel, g_aux, d_out, d_sel, d_aux);
  return check_launch("loss_bwd");
}

extern "C" int sunet_metric_hist(const float* out, const float* sel, const void* label, int label_dtype,
                                 long long pixels, float thr_out, float thr_sel, int masked,
                                 unsigned long long* counts, sunet_stream_t stream_) {
  if (!out || !label || !counts || pixels <= 0) return set_error(SUNET_ERR_INVALID, "metric_hist: bad arguments");
  if (masked && !sel) return set_error(SUNET_ERR_INVALID, "metric_hist: masked counting needs a selection map");
  const bool vec = aligned16(out, sel, nullptr, nullptr) && (reinterpret_cast<uintptr_t>(label) & 15) == 0;
  const int blocks = grid_for(vec ? (pixels + 7) / 8 : pixels, 256, 8);
#define SUNET_HIST(LT)                                                                                             \
  do {                                                                                                             \
    if (vec)                                                                                                       \
      launch_k(metric_hist_kernel<LT, 4>, dim3(blocks), dim3(256), 0, STREAM, out, sel, label, pixels, thr_out,    \
               thr_sel, masked, counts);                                                                           \
    else                                                                                                           \
      launch_k(metric_hist_kernel<LT, 1>, dim3(blocks), dim3(256), 0, STREAM, out, sel, label, pixels, thr_out,    \
               thr_sel, masked, counts);                                                                           \
  } while (0)
  switch (label_dtype) {
    case 0: SUNET_HIST(0); break;
    case 1: SUNET_HIST(1); break;
    case 2: SUNET_HIST(2); break;
    default:
      return set_error(SUNET_ERR_INVALID, "metric_hist: bad label dtype %d", label_dtype);
  }
#undef SUNET_HIST
  return check_launch("metric_hist");
}

extern "C" int sunet_adam_step(const sunet_adam_tensor* table, int n_tensors, long long max_numel, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step, const float* lr_dev,
                               const int* step_dev, sunet_stream_t stream_) {
  if (!table || n_tensors <= 0 || max_numel <= 0 || (step <= 0 && !step_dev))
    return set_error(SUNET_ERR_INVALID, "adam_step: bad arguments");
  long long bx = (max_numel + 256 * 16 - 1) / (256 * 16);      // float4 per thread, ~4 trips
  const long long cap = (long long)num_sms() * 8 / (n_tensors < 8 ? 1 : 8);     // few tensors: fill the SMs from x alone
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_tensors);
  launch_k(adam_kernel, dim3(grid), dim3(256), 0, STREAM, table, lr, beta1, beta2, eps, weight_decay, step, lr_dev, step_dev);
  return check_launch("adam_step");
}

extern "C" int sunet_abi_version(void) { return SUNET_ABI_VERSION; }
extern "C" long long sunet_launch_count(void) { return sunet::launch_count(); }
extern "C" const char* sunet_last_error(void) { return sunet::last_error(); }
