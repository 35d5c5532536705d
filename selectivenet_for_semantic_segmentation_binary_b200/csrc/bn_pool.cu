// HBM-bound kernels around the convolutions: layout packing, BatchNorm finalize,
// BN+ReLU(+MaxPool) apply and its backward.  All activation traffic is 16-byte vectors of
// 8 bf16 channels; reductions are two-stage and deterministic (per-block partial rows, then a
// fixed-order sum) — no floating-point atomics anywhere.
//
// Reference semantics: /root/reference/model.py:9-15 (CBR_2D = Conv2d -> BatchNorm2d -> ReLU),
// :31,35,39 (MaxPool2d(2), first-max tie break on backward).
#include "common.h"
#include "ptx.cuh"
#include "../../include/sunet_b200.h"

namespace sunet {

static inline int ew_grid(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

struct alignas(16) bf16x8 {
  uint32_t w[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16lo(v.w[i]);
    f[2 * i + 1] = bf16hi(v.w[i]);
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float bf16_elem(const bf16x8& v, int j) {
  return (j & 1) ? bf16hi(v.w[j >> 1]) : bf16lo(v.w[j >> 1]);
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------ packing kernels
__global__ void pack_input_im2col_kernel(const float* __restrict__ x, bf16x8* __restrict__ out, int B, int cin, int H,
                                         int W) {
  pdl_wait();
  pdl_trigger();
  // one thread = one pixel x 8 output channels
  const long long total = (long long)B * H * W * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    const long long pix = i >> 3;
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float v = 0.f;
      if (k < 9 * cin) {
        const int tap = k / cin, ci = k - tap * cin;
        const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = __ldg(x + (((long long)n * cin + ci) * H + sy) * W + sx);
      }
      f[j] = v;
    }
    out[i] = pack8(f);
  }
}

// one thread = one pixel: 9*CIN coalesced fp32 loads (lanes = consecutive x), eight 16-byte stores
template <int CIN>
__global__ void __launch_bounds__(256)
pack_input_im2col_t_kernel(const float* __restrict__ x, bf16x8* __restrict__ out, int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)B * H * W;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float f[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) f[k] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      const bool in = sy >= 0 && sy < H && sx >= 0 && sx < W;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        if (in) f[tap * CIN + ci] = __ldg(x + (((long long)n * CIN + ci) * H + sy) * W + sx);
    }
    bf16x8* o = out + pix * 8;
    bf16x8 z = {{0u, 0u, 0u, 0u}};
#pragma unroll
    for (int g = 0; g < 4; ++g) o[g] = pack8(f + g * 8);
#pragma unroll
    for (int g = 4; g < 8; ++g) o[g] = z;
  }
}

// Paired-pixel first layer (see sunet_pack_input_im2col32): 32 channels per pixel (9*CIN real, rest zero), so two
// horizontally adjacent pixels share one 128-byte row.  One thread = one pixel: 9*CIN coalesced fp32 loads,
// four 16-byte stores.
template <int CIN>
__global__ void __launch_bounds__(256)
pack_input_im2col32_kernel(const float* __restrict__ x, bf16x8* __restrict__ out, int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  static_assert(9 * CIN <= 32, "im2col32 holds at most 32 channels per pixel");
  const uint32_t total = (uint32_t)B * H * W;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const uint32_t rowi = pix / (uint32_t)W;
    const int xx = (int)(pix - rowi * (uint32_t)W);
    const int n = (int)(rowi / (uint32_t)H);
    const int yy = (int)(rowi - (uint32_t)n * (uint32_t)H);
    float f[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) f[k] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      const bool in = sy >= 0 && sy < H && sx >= 0 && sx < W;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        if (in) f[tap * CIN + ci] = __ldg(x + (((long long)n * CIN + ci) * H + sy) * W + sx);
    }
    bf16x8* o = out + (size_t)pix * 4;
#pragma unroll
    for (int g = 0; g < 4; ++g) o[g] = pack8(f + g * 8);
  }
}

// Device-side input pipeline (SURVEY.md §8(f) "next" #2; /root/reference/utils/data_utils.py:94-126,159-168,216-219):
// uint8 HWC patch -> /255 -> Normalization(mean, std) -> RandomFlip -> CHW float32, fused with the first layer's
// im2col.  The byte -> float32 map is a 256-entry table built on the host with numpy's own arithmetic, so the values
// are bit-identical to the reference's; flips are index arithmetic (bit 0: left-right, bit 1: up-down, per image).
__global__ void __launch_bounds__(256)
pack_input_u8_im2col32_kernel(const uint8_t* __restrict__ img, const float* __restrict__ lut,
                              const uint8_t* __restrict__ flip, bf16x8* __restrict__ out, int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_lut[256];
  s_lut[threadIdx.x] = __ldg(lut + threadIdx.x);
  __syncthreads();
  const uint32_t total = (uint32_t)B * H * W;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const uint32_t rowi = pix / (uint32_t)W;
    const int xx = (int)(pix - rowi * (uint32_t)W);
    const int n = (int)(rowi / (uint32_t)H);
    const int yy = (int)(rowi - (uint32_t)n * (uint32_t)H);
    const int f = flip ? flip[n] : 0;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const int iy = (f & 2) ? H - 1 - sy : sy, ix = (f & 1) ? W - 1 - sx : sx;
        const uint8_t* src = img + (((size_t)n * H + iy) * W + ix) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) v[tap * 3 + ci] = s_lut[src[ci]];
      }
    }
    bf16x8* o = out + (size_t)pix * 4;
#pragma unroll
    for (int g = 0; g < 4; ++g) o[g] = pack8(v + g * 8);
  }
}

// label bytes -> {0,1} float32 with the same flips: (label / 255.0).astype(uint8) is 1 only for 255
__global__ void __launch_bounds__(256)
pack_label_u8_kernel(const uint8_t* __restrict__ label, const uint8_t* __restrict__ flip, float* __restrict__ out,
                     int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  const uint32_t total = (uint32_t)B * H * W;
  for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const uint32_t rowi = pix / (uint32_t)W;
    const int xx = (int)(pix - rowi * (uint32_t)W);
    const int n = (int)(rowi / (uint32_t)H);
    const int yy = (int)(rowi - (uint32_t)n * (uint32_t)H);
    const int f = flip ? flip[n] : 0;
    const int iy = (f & 2) ? H - 1 - yy : yy, ix = (f & 1) ? W - 1 - xx : xx;
    out[pix] = label[((size_t)n * H + iy) * W + ix] == 255 ? 1.f : 0.f;
  }
}

// weights of the paired-pixel first layer: wf[128][64]; row n < 64: [w_n (k = tap*cin+ci, 32 wide) | 0],
// row 64 + n: [0 | w_n] — a plain GEMM over pixel PAIRS then yields both pixels' 64 outputs side by side.
__device__ __forceinline__ float conv1_pair_weight(const float* __restrict__ w, int cin, int i) {
  const int k = i & 63, nrow = i >> 6;
  const int half = nrow >> 6, co = nrow & 63;
  const int kk = k - 32 * half;
  if (kk < 0 || kk >= 9 * cin) return 0.f;
  const int tap = kk / cin, ci = kk - tap * cin;
  return w[((long long)co * cin + ci) * 9 + tap];
}
__global__ void pack_conv1_pair_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int cin) {
  pdl_wait();
  pdl_trigger();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 * 64; i += gridDim.x * blockDim.x)
    wf[i] = __float2bfloat16_rn(conv1_pair_weight(w, cin, i));
}

__global__ void pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                    __nv_bfloat16* __restrict__ wd, int co_n, int ci_n) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)co_n * ci_n * 9;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the packed forward layout [co][tap][ci] so the bf16 writes are coalesced
    const int ci = (int)(i % ci_n);
    const int tap = (int)((i / ci_n) % 9);
    const int co = (int)(i / ((long long)ci_n * 9));
    const float v = w[((long long)co * ci_n + ci) * 9 + tap];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    wf[i] = b;
    if (wd) wd[((long long)ci * 9 + (8 - tap)) * co_n + co] = b;
  }
}

__global__ void pack_conv1_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int co_n, int cin) {
  pdl_wait();
  pdl_trigger();
  const int total = co_n * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i & 63, co = i >> 6;
    float v = 0.f;
    if (k < 9 * cin) {
      const int tap = k / cin, ci = k - tap * cin;
      v = w[((long long)co * cin + ci) * 9 + tap];
    }
    wf[i] = __float2bfloat16_rn(v);
  }
}

__global__ void pack_convT_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                  __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                  float* __restrict__ bias4, int ci_n, int co_n) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)ci_n * co_n * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes wf [(tap*co_n + co)][ci]
    const int ci = (int)(i % ci_n);
    const int row = (int)(i / ci_n);
    const int co = row % co_n, tap = row / co_n;
    const __nv_bfloat16 b = __float2bfloat16_rn(w[((long long)ci * co_n + co) * 4 + tap]);
    wf[i] = b;
    if (wd) wd[(long long)ci * (4 * co_n) + row] = b;
    if (ci == 0 && bias4) bias4[row] = bias ? bias[co] : 0.f;
  }
}

// ------------------------------------------------------------------ all weight packs in one launch
// blockIdx.y = job (one per layer), blockIdx.x strides over that layer's 32x32 channel tiles.  Replaces 17
// tiny launches per forward pass: at 16 patches per GPU the step is launch-latency sensitive.
// Tiled repack (channel counts are multiples of 32): src[a][b][t] fp32 (t = filter tap, fastest) -> two bf16 operands,
//   out_b[(a*T + t)*Bn + b]         (b fastest)      conv3x3: forward pack   convT: dgrad pack
//   out_a[rowA(b, t)*An + a]        (a fastest)      conv3x3: dgrad pack (rowA = b*9 + 8 - t, the flipped tap)
//                                                     convT: forward pack (rowA = t*Bn + b)
// A block moves a 32(a) x 32(b) x T tile through shared memory so that the fp32 reads and both bf16 writes are
// contiguous runs (the element-wise form of this kernel read with stride T and wrote with stride T*An).
template <int T, bool CONVT>
__device__ __forceinline__ void pack_tile(const sunet_pack_job& j, float (*tile)[32 * 9 + 1], int tix) {
  const int An = j.a, Bn = j.b;
  const int tiles_b = Bn >> 5;
  __nv_bfloat16* out_b = reinterpret_cast<__nv_bfloat16*>(CONVT ? j.wd : j.wf);
  __nv_bfloat16* out_a = reinterpret_cast<__nv_bfloat16*>(CONVT ? j.wf : j.wd);
  {
    const int a0 = (tix / tiles_b) << 5, b0 = (tix % tiles_b) << 5;
    for (int i = threadIdx.x; i < 32 * 32 * T; i += 256) {
      const int al = i / (32 * T), r = i - al * (32 * T);
      tile[al][r] = j.w[((size_t)(a0 + al) * Bn + b0) * T + r];
    }
    __syncthreads();
    if (out_b)
      for (int i = threadIdx.x; i < 32 * 32 * T; i += 256) {
        const int bl = i & 31, t = (i >> 5) % T, al = (i >> 5) / T;
        out_b[((size_t)(a0 + al) * T + t) * Bn + b0 + bl] = __float2bfloat16_rn(tile[al][bl * T + t]);
      }
    if (out_a)
      for (int i = threadIdx.x; i < 32 * 32 * T; i += 256) {
        const int al = i & 31, t = (i >> 5) % T, bl = (i >> 5) / T;
        const size_t row = CONVT ? ((size_t)t * Bn + b0 + bl) : ((size_t)(b0 + bl) * 9 + (8 - t));
        out_a[row * An + a0 + al] = __float2bfloat16_rn(tile[al][bl * T + t]);
      }
  }
}

// One launch packs every weight tensor of the network (table of jobs): 17 launches -> 1.  The grid is FLAT over the
// 32x32 channel tiles of all jobs (job j owns blocks [tile_start_j, tile_start_{j+1})): with one block column per job
// the 512x512 layer's 256 tiles were moved by 96 blocks while 1 500 others idled (84 us per step, 0.4 TB/s).
__device__ __forceinline__ int pack_job_tiles(const sunet_pack_job& j) {
  return (j.kind == 0 || j.kind == 2) ? (j.a >> 5) * (j.b >> 5) : 1;
}
__global__ void __launch_bounds__(256)
pack_table_kernel(const sunet_pack_job* __restrict__ jobs, int n_jobs) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][32 * 9 + 1];
  __shared__ int s_job, s_tix;
  if (threadIdx.x == 0) {
    int jb = 0;
    while (jb + 1 < n_jobs && (int)blockIdx.x >= jobs[jb + 1].tile_start) ++jb;
    s_job = jb;
    s_tix = (int)blockIdx.x - jobs[jb].tile_start;
  }
  __syncthreads();
  const sunet_pack_job j = jobs[s_job];
  const int tix = s_tix;
  if (tix >= pack_job_tiles(j)) return;
  if (j.kind == 0) {            // conv3x3: a = cout, b = cin
    pack_tile<9, false>(j, tile, tix);
  } else if (j.kind == 1) {     // first conv: a = cout, b = cin, K padded to 64
    const int co_n = j.a, cin = j.b;
    __nv_bfloat16* wf = reinterpret_cast<__nv_bfloat16*>(j.wf);
    for (int i = threadIdx.x; i < co_n * 64; i += blockDim.x) {
      const int k = i & 63, co = i >> 6;
      float v = 0.f;
      if (k < 9 * cin) {
        const int tap = k / cin, ci = k - tap * cin;
        v = j.w[((long long)co * cin + ci) * 9 + tap];
      }
      wf[i] = __float2bfloat16_rn(v);
    }
  } else if (j.kind == 3) {     // first conv, paired-pixel form: a = 64 (cout), b = cin
    __nv_bfloat16* wf = reinterpret_cast<__nv_bfloat16*>(j.wf);
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) wf[i] = __float2bfloat16_rn(conv1_pair_weight(j.w, j.b, i));
  } else {                      // ConvTranspose2d: a = cin, b = cout
    pack_tile<4, true>(j, tile, tix);
    if (j.bias4 && tix == 0)
      for (int i = threadIdx.x; i < 4 * j.b; i += blockDim.x) j.bias4[i] = j.bias ? j.bias[i % j.b] : 0.f;
  }
}

// ------------------------------------------------------------------ BN finalize
// one warp per channel: lanes stride over the per-CTA partial rows, fp64 tree reduction
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ stats, int rows, int C, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ conv_bias, float* running_mean, float* running_var,
                   long long* nbt, float momentum, float eps, float* scale, float* shift, float* mean,
                   float* invstd) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c == 0 && lane == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int r = lane; r < rows; r += 32) {
    const float2 v = *reinterpret_cast<const float2*>(stats + ((size_t)r * C + c) * 2);
    s += (double)v.x;
    q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  const double m = s / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  const float istd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * istd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)m * sc;
  mean[c] = (float)m;
  invstd[c] = istd;
  if (running_mean) {
    const float b = conv_bias ? conv_bias[c] : 0.f;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)m + b);
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_eval_affine_kernel(const float* gamma, const float* beta, const float* conv_bias,
                                      const float* rm, const float* rv, float eps, float* scale, float* shift, int C) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
}

// one warp per channel (lanes stride over the per-CTA rows, fixed-order fp64 tree): the serial per-thread loop over
// up to 296 rows took 14 us per launch
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const float* stats, int rows, int n_total, int col_offset, int C, float* out) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int r = lane; r < rows; r += 32) s += (double)stats[((size_t)r * n_total + col_offset + c) * 2];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = (float)s;
}

// ------------------------------------------------------------------ BN + ReLU (+ pool) forward
// POOL: two 2x2 windows x 8 channels per thread and trip, all eight 16-byte loads issued before the first is used
// (one window per trip left the kernel latency-bound at 80 % of the HBM rate: 24 warps per SM with 64 B in flight each).
template <bool POOL>
__global__ void __launch_bounds__(256)
bn_relu_pool_kernel(const __nv_bfloat16* __restrict__ y, int ys, const float* __restrict__ scale,
                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ a, int as,
                    __nv_bfloat16* __restrict__ pooled, int ps, __nv_bfloat16* __restrict__ ywin, int yws, int B,
                    int H, int W, int C) {
  pdl_wait();
  pdl_trigger();
  const int G = C >> 3;
  // POOL: one item = one 2x2 window x 8 channels; else one pixel x 8 channels
  const int HW = POOL ? (H >> 1) : H, WW = POOL ? (W >> 1) : W;
  const long long total = (long long)B * HW * WW * G;
  if (POOL) {
    constexpr int U = 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
      bf16x8 v[U][4];
      long long p00[U], ppix[U];
      int g[U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        on[u] = i < total;
        // item counts fit 32 bits (checked by the launcher): unsigned 32-bit divisions, not 64-bit ones
        const uint32_t i32 = on[u] ? (uint32_t)i : 0u;
        const uint32_t wpix32 = i32 / (uint32_t)G;
        g[u] = (int)(i32 - wpix32 * (uint32_t)G);
        const uint32_t rowi = wpix32 / (uint32_t)WW;
        const int wx = (int)(wpix32 - rowi * (uint32_t)WW);
        const int n = (int)(rowi / (uint32_t)HW);
        const int wy = (int)(rowi - (uint32_t)n * (uint32_t)HW);
        p00[u] = ((long long)n * H + wy * 2) * W + wx * 2;
        ppix[u] = wpix32;
        if (on[u]) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            v[u][k] = *reinterpret_cast<const bf16x8*>(y + (p00[u] + (k >> 1) * W + (k & 1)) * ys + g[u] * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!on[u]) continue;
        float sc[8], sh[8];
        *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + g[u] * 8));
        *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + g[u] * 8 + 4));
        *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + g[u] * 8));
        *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + g[u] * 8 + 4));
        // (best, yb) = fp32 pre-ReLU activation and conv output y of the window's FIRST maximum — the element
        // MaxPool2d's backward routes the gradient to (same rule as pool_win_chan); the pooled activation is
        // relu(best): max over the window of relu(raw) = relu(max raw), exactly
        float best[8], yb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long pix = p00[u] + (k >> 1) * W + (k & 1);
          float f[8];
          unpack8(v[u][k], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float raw = fmaf(f[j], sc[j], sh[j]);
            if (k == 0 || raw > best[j]) {
              best[j] = raw;
              yb[j] = f[j];
            }
            f[j] = fmaxf(raw, 0.f);
          }
          *reinterpret_cast<bf16x8*>(a + pix * as + g[u] * 8) = pack8(f);
        }
        if (ywin != nullptr) *reinterpret_cast<bf16x8*>(ywin + ppix[u] * yws + g[u] * 8) = pack8(yb);   // exact: y is bf16
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], 0.f);
        *reinterpret_cast<bf16x8*>(pooled + ppix[u] * ps + g[u] * 8) = pack8(best);
      }
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t i32 = (uint32_t)i;
    const uint32_t wpix32 = i32 / (uint32_t)G;
    const int g = (int)(i32 - wpix32 * (uint32_t)G);
    const long long pix = wpix32;
    float sc[8], sh[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale + g * 8));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift + g * 8));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(y + pix * ys + g * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
    *reinterpret_cast<bf16x8*>(a + pix * as + g * 8) = pack8(f);
  }
}

// plain 2x2 / stride-2 max-pool of an activation (inference path: the conv epilogue already applied BN + ReLU)
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const __nv_bfloat16* __restrict__ a, int as, __nv_bfloat16* __restrict__ pooled, int ps, int B, int H,
                  int W, int C) {
  pdl_wait();
  pdl_trigger();
  const uint32_t G = (uint32_t)C >> 3, HW = (uint32_t)H >> 1, WW = (uint32_t)W >> 1;
  const uint32_t total = (uint32_t)B * HW * WW * G;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t win = i / G, g = i - win * G;
    const uint32_t rowi = win / WW, wx = win - rowi * WW;
    const uint32_t n = rowi / HW, wy = rowi - n * HW;
    const size_t p00 = ((size_t)n * H + wy * 2) * W + wx * 2;
    bf16x8 v[4];
    v[0] = *reinterpret_cast<const bf16x8*>(a + p00 * as + g * 8);
    v[1] = *reinterpret_cast<const bf16x8*>(a + (p00 + 1) * as + g * 8);
    v[2] = *reinterpret_cast<const bf16x8*>(a + (p00 + W) * as + g * 8);
    v[3] = *reinterpret_cast<const bf16x8*>(a + (p00 + W + 1) * as + g * 8);
    float m[8];
    unpack8(v[0], m);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      float f[8];
      unpack8(v[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
    *reinterpret_cast<bf16x8*>(pooled + (size_t)win * ps + g * 8) = pack8(m);
  }
}

// ------------------------------------------------------------------ BN + ReLU (+ pool) backward
//   g = (dA [+ dPool routed to the first maximum of each 2x2 window]) * [relu(bn(y)) > 0]
//   dgamma = sum g*xhat, dbeta = sum g, dy = kg*g + k1*y + k0  (the three per-channel coefficients fold scale,
//   invstd, mean and the two sums; bn_bwd_finalize_par_kernel)
// Non-pooled blocks (11 of 14): per-channel vectors hoisted out of the loop and FOUR independent 16-byte loads per
// operand in flight per thread — these kernels are pure HBM streams and latency, not arithmetic, limits them.
constexpr int EW_U = 4;

struct ChanVec {
  float v[8];
};
__device__ __forceinline__ ChanVec load_chan(const float* __restrict__ p, int g) {
  ChanVec c;
  *reinterpret_cast<float4*>(c.v) = __ldg(reinterpret_cast<const float4*>(p + g * 8));
  *reinterpret_cast<float4*>(c.v + 4) = __ldg(reinterpret_cast<const float4*>(p + g * 8 + 4));
  return c;
}

__global__ void __launch_bounds__(256)
bn_relu_flat_kernel(const __nv_bfloat16* __restrict__ y, int ys, const float* __restrict__ scale,
                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ a, int as, long long total, int G) {
  pdl_wait();
  pdl_trigger();
  const int g = threadIdx.x % G;                 // 256 % G == 0 and every stride below is a multiple of 256
  const int lg = __ffs(G) - 1;                   // G divides 256, so it is a power of two: i / G is a shift
  const ChanVec sc = load_chan(scale, g), sh = load_chan(shift, g);
  for (long long i0 = blockIdx.x * (long long)(256 * EW_U) + threadIdx.x; i0 < total;
       i0 += (long long)gridDim.x * 256 * EW_U) {
    bf16x8 v[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) v[u] = *reinterpret_cast<const bf16x8*>(y + (i >> lg) * ys + g * 8);
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc.v[j], sh.v[j]), 0.f);
        *reinterpret_cast<bf16x8*>(a + (i >> lg) * as + g * 8) = pack8(f);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_flat_kernel(const __nv_bfloat16* __restrict__ dA, int das, const __nv_bfloat16* __restrict__ y, int ys,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ partials, long long total, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[256][17];
  const int G = C >> 3;
  const int g = threadIdx.x % G;
  const int lg = __ffs(G) - 1;
  const ChanVec sc = load_chan(scale, g), sh = load_chan(shift, g), mu = load_chan(mean, g);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (long long i0 = blockIdx.x * (long long)(256 * EW_U) + threadIdx.x; i0 < total;
       i0 += (long long)gridDim.x * 256 * EW_U) {
    bf16x8 vy[EW_U], vg[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) {
        vy[u] = *reinterpret_cast<const bf16x8*>(y + (i >> lg) * ys + g * 8);
        vg[u] = *reinterpret_cast<const bf16x8*>(dA + (i >> lg) * das + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) {
        float fy[8], fg[8];
        unpack8(vy[u], fy);
        unpack8(vg[u], fg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = fmaf(fy[j], sc.v[j], sh.v[j]) > 0.f ? fg[j] : 0.f;
          acc[j] += gg;
          acc[8 + j] = fmaf(gg, fy[j] - mu.v[j], acc[8 + j]);     // x invstd after the loop
        }
      }
    }
  }
  {
    const ChanVec is = load_chan(invstd, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[8 + j] *= is.v[j];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  const int reps = 256 / G;
  for (int o = threadIdx.x; o < G * 16; o += 256) {
    const int og = o >> 4, oj = o & 15;
    float s = 0.f;
    for (int r = 0; r < reps; ++r) s += red[r * G + og][oj];
    const int c = og * 8 + (oj & 7);
    partials[((size_t)blockIdx.x * C + c) * 2 + (oj >> 3)] = s;
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_flat_kernel(const __nv_bfloat16* __restrict__ dA, int das, const __nv_bfloat16* __restrict__ y, int ys,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy, int dys, long long total,
                         int C) {
  pdl_wait();
  pdl_trigger();
  const int G = C >> 3;
  const int g = threadIdx.x % G;
  const int lg = __ffs(G) - 1;
  const ChanVec sc = load_chan(scale, g), sh = load_chan(shift, g);
  const ChanVec kg = load_chan(coef, g), k1 = load_chan(coef + C, g), k0 = load_chan(coef + 2 * C, g);
  for (long long i0 = blockIdx.x * (long long)(256 * EW_U) + threadIdx.x; i0 < total;
       i0 += (long long)gridDim.x * 256 * EW_U) {
    bf16x8 vy[EW_U], vg[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) {
        vy[u] = *reinterpret_cast<const bf16x8*>(y + (i >> lg) * ys + g * 8);
        vg[u] = *reinterpret_cast<const bf16x8*>(dA + (i >> lg) * das + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long i = i0 + u * 256;
      if (i < total) {
        float fy[8], fg[8], o[8];
        unpack8(vy[u], fy);
        unpack8(vg[u], fg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = fmaf(fy[j], sc.v[j], sh.v[j]) > 0.f ? fg[j] : 0.f;
          o[j] = fmaf(kg.v[j], gg, fmaf(k1.v[j], fy[j], k0.v[j]));
        }
        *reinterpret_cast<bf16x8*>(dy + (i >> lg) * dys + g * 8) = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------ pooled-layer backward, lean form
// The three encoder outputs feed both a skip connection and a 2x2 max-pool.  One thread = one 2x2 window
// x 8 channels.  Operands stay packed (bf16x8) and are unpacked one channel at a time, so only ~12 fp32
// values are live; per channel and window the work is 4 FFMA (activations), 3 FMNMX (window max), 4+4
// compares and 8 selects (first-max routing + ReLU mask).  The first maximum is taken on the fp32
// activations fmaf(y, scale, shift): identical to the bf16-pooled forward wherever the window maximum is
// unique after rounding, and the reference's own (fp32) choice where bf16 rounding made a tie.  Earlier
// forms of this kernel (bit-mask decisions, 64-bit index math) were instruction-bound at ~45 % of HBM peak.
struct PoolWin {
  bf16x8 vy[4], vg[4], vp;
  long long p00;      // top-left pixel of the window; pixel k = p00 + (k >> 1) * W + (k & 1)
};
__device__ __forceinline__ void pool_win_load(PoolWin& w, long long win, int g, const __nv_bfloat16* __restrict__ dA,
                                              int das, const __nv_bfloat16* __restrict__ dP, int dps,
                                              const __nv_bfloat16* __restrict__ y, int ys, int H, int W) {
  // window counts fit 32 bits (checked by the launcher): unsigned 32-bit divisions instead of 64-bit ones
  const uint32_t HW = (uint32_t)H >> 1, WW = (uint32_t)W >> 1;
  const uint32_t w32 = (uint32_t)win;
  const uint32_t rowi = w32 / WW;
  const int wx = (int)(w32 - rowi * WW);
  const int n = (int)(rowi / HW);
  const int wy = (int)(rowi - (uint32_t)n * HW);
  const long long p00 = ((long long)n * H + wy * 2) * W + wx * 2;
  w.p00 = p00;
#pragma unroll
  for (int k = 0; k < 4; ++k) w.vy[k] = *reinterpret_cast<const bf16x8*>(y + (p00 + (k >> 1) * W + (k & 1)) * ys + g * 8);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (dA) w.vg[k] = *reinterpret_cast<const bf16x8*>(dA + (p00 + (k >> 1) * W + (k & 1)) * das + g * 8);
    else w.vg[k] = bf16x8{{0u, 0u, 0u, 0u}};
  }
  w.vp = *reinterpret_cast<const bf16x8*>(dP + win * dps + g * 8);
}
// channel j of the window: yv[k] = conv outputs, gv[k] = gradient reaching y's activation (skip + routed pool
// gradient, ReLU-masked)
__device__ __forceinline__ void pool_win_chan(const PoolWin& w, int j, float sc, float sh, float* yv, float* gv) {
  float a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    yv[k] = bf16_elem(w.vy[k], j);
    gv[k] = bf16_elem(w.vg[k], j);
    a[k] = fmaf(yv[k], sc, sh);
  }
  const float dp = bf16_elem(w.vp, j);
  const float m = fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
  const bool w0 = a[0] == m;
  const bool w1 = !w0 && a[1] == m;
  const bool w2 = !w0 && !w1 && a[2] == m;
  const bool w3 = !w0 && !w1 && !w2;
  gv[0] = a[0] > 0.f ? gv[0] + (w0 ? dp : 0.f) : 0.f;
  gv[1] = a[1] > 0.f ? gv[1] + (w1 ? dp : 0.f) : 0.f;
  gv[2] = a[2] > 0.f ? gv[2] + (w2 ? dp : 0.f) : 0.f;
  gv[3] = a[3] > 0.f ? gv[3] + (w3 ? dp : 0.f) : 0.f;
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_pool_reduce_kernel(const __nv_bfloat16* __restrict__ dA, int das, const __nv_bfloat16* __restrict__ dP, int dps,
                          const __nv_bfloat16* __restrict__ y, int ys, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ invstd, float* __restrict__ partials, int B, int H, int W, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[256][17];
  const int G = C >> 3;
  const int g = threadIdx.x % G;
  const int lg = __ffs(G) - 1;
  const long long total = (long long)B * (H >> 1) * (W >> 1) * G;
  const ChanVec sc = load_chan(scale, g), sh = load_chan(shift, g), mu = load_chan(mean, g);
  float acc[16];        // [0,8): sum g      [8,16): sum g * (y - mean)   (x invstd at the end)
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    PoolWin w;
    pool_win_load(w, i >> lg, g, dA, das, dP, dps, y, ys, H, W);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float yv[4], gv[4];
      pool_win_chan(w, j, sc.v[j], sh.v[j], yv, gv);
      acc[j] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
      float t = gv[0] * (yv[0] - mu.v[j]);
      t = fmaf(gv[1], yv[1] - mu.v[j], t);
      t = fmaf(gv[2], yv[2] - mu.v[j], t);
      t = fmaf(gv[3], yv[3] - mu.v[j], t);
      acc[8 + j] += t;
    }
  }
  {
    const ChanVec is = load_chan(invstd, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[8 + j] *= is.v[j];
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  const int reps = 256 / G;
  for (int o = threadIdx.x; o < G * 16; o += 256) {
    const int og = o >> 4, oj = o & 15;
    float s = 0.f;
    for (int r = 0; r < reps; ++r) s += red[r * G + og][oj];
    const int c = og * 8 + (oj & 7);
    partials[((size_t)blockIdx.x * C + c) * 2 + (oj >> 3)] = s;
  }
}

// The five per-channel constant vectors live in shared memory, not in 40 registers per thread: with 9 sixteen-byte loads
// in flight per thread the kernel is bound by how many threads fit (bytes in flight per SM), and 128 registers allowed
// only two blocks (82 % of the HBM rate).
constexpr int kPoolApplyMaxC = 512;
// volatile: re-read at every use — hoisting the 40 loop-invariant values into registers is exactly what is being avoided
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
  return r;
}
__global__ void __launch_bounds__(256, 3)
bn_bwd_pool_apply_kernel(const __nv_bfloat16* __restrict__ dA, int das, const __nv_bfloat16* __restrict__ dP, int dps,
                         const __nv_bfloat16* __restrict__ y, int ys, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ coef,
                         __nv_bfloat16* __restrict__ dy, int dys, int B, int H, int W, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float cst[5][kPoolApplyMaxC];       // scale, shift, kg, k1, k0
  for (int c = threadIdx.x; c < C; c += 256) {
    cst[0][c] = scale[c];
    cst[1][c] = shift[c];
    cst[2][c] = coef[c];
    cst[3][c] = coef[C + c];
    cst[4][c] = coef[2 * C + c];
  }
  __syncthreads();
  const int G = C >> 3;
  const int g = threadIdx.x % G;
  const int lg = __ffs(G) - 1;
  const long long total = (long long)B * (H >> 1) * (W >> 1) * G;
  const uint32_t cbase = smem_u32(&cst[0][g * 8]);
  constexpr uint32_t CS = kPoolApplyMaxC * 4;      // bytes between the constant vectors
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    PoolWin w;
    pool_win_load(w, i >> lg, g, dA, das, dP, dps, y, ys, H, W);
    bf16x8 o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float2 sc = lds_f2(cbase + jj * 8), sh = lds_f2(cbase + CS + jj * 8), kg = lds_f2(cbase + 2 * CS + jj * 8);
      const float2 k1 = lds_f2(cbase + 3 * CS + jj * 8), k0 = lds_f2(cbase + 4 * CS + jj * 8);
      float lo[4], hi[4];
      {
        float yv[4], gv[4];
        pool_win_chan(w, 2 * jj, sc.x, sh.x, yv, gv);
#pragma unroll
        for (int k = 0; k < 4; ++k) lo[k] = fmaf(kg.x, gv[k], fmaf(k1.x, yv[k], k0.x));
      }
      {
        float yv[4], gv[4];
        pool_win_chan(w, 2 * jj + 1, sc.y, sh.y, yv, gv);
#pragma unroll
        for (int k = 0; k < 4; ++k) hi[k] = fmaf(kg.y, gv[k], fmaf(k1.y, yv[k], k0.y));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k].w[jj] = pack_bf16x2(lo[k], hi[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<bf16x8*>(dy + (w.p00 + (k >> 1) * W + (k & 1)) * dys + g * 8) = o[k];
  }
}

// Fold of reduction rows that come from up to two producers with their own row strides / column offsets
// (pooled blocks: skip-gradient rows from the decoder dgrad epilogue, pool-gradient rows from the encoder dgrad
// epilogue).  One warp per channel.
struct PartialSrc {
  const float* p;
  int rows, stride, col0;     // element (r, c, k) at p[(r * stride + col0 + c) * 2 + k]
};
__global__ void __launch_bounds__(256)
bn_bwd_finalize2_kernel(PartialSrc s0, PartialSrc s1, int C, double count, const float* __restrict__ scale,
                        const float* __restrict__ mean, const float* __restrict__ invstd, float* dgamma, float* dbeta,
                        float* coef) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double sg = 0.0, sgx = 0.0;
  for (int b = lane; b < s0.rows; b += 32) {
    const float2 v = *reinterpret_cast<const float2*>(s0.p + ((size_t)b * s0.stride + s0.col0 + c) * 2);
    sg += (double)v.x;
    sgx += (double)v.y;
  }
  for (int b = lane; b < s1.rows; b += 32) {
    const float2 v = *reinterpret_cast<const float2*>(s1.p + ((size_t)b * s1.stride + s1.col0 + c) * 2);
    sg += (double)v.x;
    sgx += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
  }
  if (lane == 0) {
    if (dbeta) dbeta[c] = (float)sg;
    if (dgamma) dgamma[c] = (float)sgx;
    const double scv = scale[c];
    const double k1 = -scv * (double)invstd[c] * sgx / count;
    const double k0 = -scv * sg / count - k1 * (double)mean[c];
    coef[c] = (float)scv;
    coef[C + c] = (float)k1;
    coef[2 * C + c] = (float)k0;
  }
}

// parallel fold of the per-block partial rows: one warp per channel, lanes stride over the rows
__global__ void __launch_bounds__(256)
bn_bwd_finalize_par_kernel(const float* __restrict__ partials, int blocks, int C, double count,
                           const float* __restrict__ scale, const float* __restrict__ mean,
                           const float* __restrict__ invstd, float* dgamma, float* dbeta, float* coef) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double sg = 0.0, sgx = 0.0;
  for (int b = lane; b < blocks; b += 32) {
    const float2 v = *reinterpret_cast<const float2*>(partials + ((size_t)b * C + c) * 2);
    sg += (double)v.x;
    sgx += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
  }
  if (lane == 0) {
    if (dbeta) dbeta[c] = (float)sg;
    if (dgamma) dgamma[c] = (float)sgx;
    const double scv = scale[c];
    const double k1 = -scv * (double)invstd[c] * sgx / count;
    const double k0 = -scv * sg / count - k1 * (double)mean[c];
    coef[c] = (float)scv;
    coef[C + c] = (float)k1;
    coef[2 * C + c] = (float)k0;
  }
}

static inline int flat_grid(long long total) {
  long long b = (total + 256 * EW_U - 1) / (256 * EW_U);
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace sunet

using namespace sunet;

#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int sunet_pack_input_im2col(const float* x, void* out, int batch, int cin, int height, int width,
                                       sunet_stream_t stream_) {
  if (!x || !out || batch <= 0 || cin <= 0 || cin * 9 > 64 || height <= 0 || width <= 0)
    return set_error(SUNET_ERR_INVALID, "pack_input_im2col: bad arguments (cin=%d)", cin);
  const long long total = (long long)batch * height * width * 8;
  if (cin == 3)
    launch_k(pack_input_im2col_t_kernel<3>, dim3(ew_grid(total / 8, 256)), dim3(256), 0, STREAM, x, reinterpret_cast<bf16x8*>(out), batch,
                                                                               height, width);
  else if (cin == 2)
    launch_k(pack_input_im2col_t_kernel<2>, dim3(ew_grid(total / 8, 256)), dim3(256), 0, STREAM, x, reinterpret_cast<bf16x8*>(out), batch,
                                                                               height, width);
  else
    launch_k(pack_input_im2col_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, x, reinterpret_cast<bf16x8*>(out), batch, cin,
                                                                       height, width);
  return check_launch("pack_input_im2col");
}

extern "C" int sunet_pack_input_im2col32(const float* x, void* out, int batch, int cin, int height, int width,
                                         sunet_stream_t stream_) {
  if (!x || !out || batch <= 0 || (cin != 2 && cin != 3) || height <= 0 || width <= 0 || (width & 1))
    return set_error(SUNET_ERR_INVALID, "pack_input_im2col32: bad arguments (cin=%d, width=%d)", cin, width);
  const long long total = (long long)batch * height * width;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "pack_input_im2col32: tensor too large");
  if (cin == 3)
    launch_k(pack_input_im2col32_kernel<3>, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, x,
             reinterpret_cast<bf16x8*>(out), batch, height, width);
  else
    launch_k(pack_input_im2col32_kernel<2>, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, x,
             reinterpret_cast<bf16x8*>(out), batch, height, width);
  return check_launch("pack_input_im2col32");
}

extern "C" int sunet_pack_input_u8_im2col32(const void* img, const float* lut, const void* flip, void* out, int batch,
                                            int height, int width, sunet_stream_t stream_) {
  if (!img || !lut || !out || batch <= 0 || height <= 0 || width <= 0 || (width & 1))
    return set_error(SUNET_ERR_INVALID, "pack_input_u8_im2col32: bad arguments (width=%d)", width);
  const long long total = (long long)batch * height * width;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "pack_input_u8_im2col32: tensor too large");
  launch_k(pack_input_u8_im2col32_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM,
           reinterpret_cast<const uint8_t*>(img), lut, reinterpret_cast<const uint8_t*>(flip),
           reinterpret_cast<bf16x8*>(out), batch, height, width);
  return check_launch("pack_input_u8_im2col32");
}

extern "C" int sunet_pack_label_u8(const void* label, const void* flip, float* out, int batch, int height, int width,
                                   sunet_stream_t stream_) {
  if (!label || !out || batch <= 0 || height <= 0 || width <= 0)
    return set_error(SUNET_ERR_INVALID, "pack_label_u8: bad arguments");
  const long long total = (long long)batch * height * width;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "pack_label_u8: tensor too large");
  launch_k(pack_label_u8_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM,
           reinterpret_cast<const uint8_t*>(label), reinterpret_cast<const uint8_t*>(flip), out, batch, height, width);
  return check_launch("pack_label_u8");
}

extern "C" int sunet_pack_conv1_pair_weights(const float* w, void* wf, int cout, int cin, sunet_stream_t stream_) {
  if (!w || !wf || cout != 64 || cin <= 0 || cin * 9 > 32)
    return set_error(SUNET_ERR_INVALID, "pack_conv1_pair_weights: needs cout = 64 and 9*cin <= 32");
  launch_k(pack_conv1_pair_kernel, dim3(32), dim3(256), 0, STREAM, w, reinterpret_cast<__nv_bfloat16*>(wf), cin);
  return check_launch("pack_conv1_pair_weights");
}

extern "C" int sunet_pack_conv3x3_weights(const float* w, void* wf, void* wd, int cout, int cin,
                                          sunet_stream_t stream_) {
  if (!w || !wf || cout <= 0 || cin <= 0) return set_error(SUNET_ERR_INVALID, "pack_conv3x3_weights: bad arguments");
  const long long total = (long long)cout * cin * 9;
  launch_k(pack_conv3x3_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, w, reinterpret_cast<__nv_bfloat16*>(wf),
                                                                reinterpret_cast<__nv_bfloat16*>(wd), cout, cin);
  return check_launch("pack_conv3x3_weights");
}

extern "C" int sunet_pack_conv1_weights(const float* w, void* wf, int cout, int cin, sunet_stream_t stream_) {
  if (!w || !wf || cout <= 0 || cin <= 0 || cin * 9 > 64)
    return set_error(SUNET_ERR_INVALID, "pack_conv1_weights: bad arguments");
  launch_k(pack_conv1_kernel, dim3(ew_grid(cout * 64, 256)), dim3(256), 0, STREAM, w, reinterpret_cast<__nv_bfloat16*>(wf), cout, cin);
  return check_launch("pack_conv1_weights");
}

extern "C" int sunet_pack_convT_weights(const float* w, const float* bias, void* wf, void* wd, float* bias4, int cin,
                                        int cout, sunet_stream_t stream_) {
  if (!w || !wf || cin <= 0 || cout <= 0) return set_error(SUNET_ERR_INVALID, "pack_convT_weights: bad arguments");
  const long long total = (long long)cin * cout * 4;
  launch_k(pack_convT_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, w, bias, reinterpret_cast<__nv_bfloat16*>(wf),
                                                              reinterpret_cast<__nv_bfloat16*>(wd), bias4, cin, cout);
  return check_launch("pack_convT_weights");
}

extern "C" int sunet_pack_weights_table(const sunet_pack_job* jobs_dev, int n_jobs, int total_tiles,
                                        sunet_stream_t stream_) {
  if (!jobs_dev || n_jobs <= 0 || total_tiles <= 0)
    return set_error(SUNET_ERR_INVALID, "pack_weights_table: bad arguments");
  launch_k(pack_table_kernel, dim3((unsigned)total_tiles), dim3(256), 0, STREAM, jobs_dev, n_jobs);
  return check_launch("pack_weights_table");
}

extern "C" int sunet_bn_finalize(const float* stats, int rows, int channels, long long count, const float* gamma,
                                 const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                                 long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                                 float* mean, float* invstd, sunet_stream_t stream_) {
  if (!stats || rows <= 0 || channels <= 0 || count <= 0 || !gamma || !beta || !scale || !shift || !mean || !invstd)
    return set_error(SUNET_ERR_INVALID, "bn_finalize: bad arguments");
  if ((running_mean == nullptr) != (running_var == nullptr))
    return set_error(SUNET_ERR_INVALID, "bn_finalize: running_mean/var must both be given or both NULL");
  launch_k(bn_finalize_kernel, dim3((channels + 7) / 8), dim3(256), 0, STREAM, stats, rows, channels, (double)count, gamma, beta,
                                                                  conv_bias, running_mean, running_var,
                                                                  num_batches_tracked, momentum, eps, scale, shift,
                                                                  mean, invstd);
  return check_launch("bn_finalize");
}

extern "C" int sunet_bn_eval_affine(const float* gamma, const float* beta, const float* conv_bias,
                                    const float* running_mean, const float* running_var, float eps, float* scale,
                                    float* shift, int channels, sunet_stream_t stream_) {
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift || channels <= 0)
    return set_error(SUNET_ERR_INVALID, "bn_eval_affine: bad arguments");
  launch_k(bn_eval_affine_kernel, dim3((channels + 127) / 128), dim3(128), 0, STREAM, gamma, beta, conv_bias, running_mean, running_var,
                                                                     eps, scale, shift, channels);
  return check_launch("bn_eval_affine");
}

extern "C" int sunet_colsum_finalize(const float* stats, int rows, int n_total, int col_offset, int channels,
                                     float* out, sunet_stream_t stream_) {
  if (!stats || !out || rows <= 0 || channels <= 0 || col_offset < 0 || col_offset + channels > n_total)
    return set_error(SUNET_ERR_INVALID, "colsum_finalize: bad arguments");
  launch_k(colsum_finalize_kernel, dim3((channels + 7) / 8), dim3(256), 0, STREAM, stats, rows, n_total, col_offset, channels, out);
  return check_launch("colsum_finalize");
}

static int check_act(const char* what, int stride, int channels) {
  if (channels <= 0 || channels % 8 || stride < channels || stride % 8)
    return set_error(SUNET_ERR_INVALID, "%s: channels %d / pixel stride %d must be multiples of 8, stride >= channels",
                     what, channels, stride);
  return SUNET_OK;
}

static int bn_relu_pool_impl(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                             int a_pix_stride, void* pooled, int pooled_pix_stride, void* ywin, int ywin_pix_stride,
                             int batch, int height, int width, int channels, sunet_stream_t stream_);

extern "C" int sunet_bn_relu_pool(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                                  int a_pix_stride, void* pooled, int pooled_pix_stride, int batch, int height,
                                  int width, int channels, sunet_stream_t stream_) {
  return bn_relu_pool_impl(y, y_pix_stride, scale, shift, a, a_pix_stride, pooled, pooled_pix_stride, nullptr, 0, batch,
                           height, width, channels, stream_);
}

extern "C" int sunet_bn_relu_pool_ywin(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                                       int a_pix_stride, void* pooled, int pooled_pix_stride, void* ywin,
                                       int ywin_pix_stride, int batch, int height, int width, int channels,
                                       sunet_stream_t stream_) {
  if (!pooled || !ywin) return set_error(SUNET_ERR_INVALID, "bn_relu_pool_ywin: pooled and ywin are required");
  int e = check_act("bn_relu_pool_ywin(ywin)", ywin_pix_stride, channels);
  if (e) return e;
  return bn_relu_pool_impl(y, y_pix_stride, scale, shift, a, a_pix_stride, pooled, pooled_pix_stride, ywin,
                           ywin_pix_stride, batch, height, width, channels, stream_);
}

static int bn_relu_pool_impl(const void* y, int y_pix_stride, const float* scale, const float* shift, void* a,
                             int a_pix_stride, void* pooled, int pooled_pix_stride, void* ywin, int ywin_pix_stride,
                             int batch, int height, int width, int channels, sunet_stream_t stream_) {
  if (!y || !scale || !shift || !a || batch <= 0 || height <= 0 || width <= 0)
    return set_error(SUNET_ERR_INVALID, "bn_relu_pool: bad arguments");
  int e;
  if ((e = check_act("bn_relu_pool(y)", y_pix_stride, channels))) return e;
  if ((e = check_act("bn_relu_pool(a)", a_pix_stride, channels))) return e;
  const __nv_bfloat16* yp = reinterpret_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(a);
  if (pooled) {
    if ((height | width) & 1) return set_error(SUNET_ERR_INVALID, "bn_relu_pool: odd size %d x %d", height, width);
    if ((e = check_act("bn_relu_pool(pooled)", pooled_pix_stride, channels))) return e;
    const long long total = (long long)batch * (height / 2) * (width / 2) * (channels / 8);
    if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "bn_relu_pool: tensor too large (%lld items)", total);
    launch_k(bn_relu_pool_kernel<true>, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM,
        yp, y_pix_stride, scale, shift, ap, a_pix_stride, reinterpret_cast<__nv_bfloat16*>(pooled), pooled_pix_stride,
        reinterpret_cast<__nv_bfloat16*>(ywin), ywin_pix_stride, batch, height, width, channels);
  } else {
    const long long total = (long long)batch * height * width * (channels / 8);
    if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "bn_relu_pool: tensor too large (%lld items)", total);
    if (256 % (channels / 8) == 0)
      launch_k(bn_relu_flat_kernel, dim3(flat_grid(total)), dim3(256), 0, STREAM, yp, y_pix_stride, scale, shift, ap, a_pix_stride, total,
                                                                channels / 8);
    else
      launch_k(bn_relu_pool_kernel<false>, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM, yp, y_pix_stride, scale, shift, ap,
               a_pix_stride, nullptr, 0, nullptr, 0, batch, height, width, channels);
  }
  return check_launch("bn_relu_pool");
}

extern "C" int sunet_bn_relu_pool_bwd(const void* dA, int dA_pix_stride, const void* dPool, int dPool_pix_stride,
                                      const void* y, int y_pix_stride, const float* scale, const float* shift,
                                      const float* mean, const float* invstd, const float* gamma, float* dgamma,
                                      float* dbeta, void* dy, int dy_pix_stride, int batch, int height, int width,
                                      int channels, void* workspace, size_t workspace_bytes, sunet_stream_t stream_) {
  (void)gamma;
  if ((!dA && !dPool) || !y || !scale || !shift || !mean || !invstd || !dy || !workspace)
    return set_error(SUNET_ERR_INVALID, "bn_relu_pool_bwd: bad arguments");
  int e;
  if ((e = check_act("bn_relu_pool_bwd(y)", y_pix_stride, channels))) return e;
  if ((e = check_act("bn_relu_pool_bwd(dy)", dy_pix_stride, channels))) return e;
  if (dA && (e = check_act("bn_relu_pool_bwd(dA)", dA_pix_stride, channels))) return e;
  if (dPool && (e = check_act("bn_relu_pool_bwd(dPool)", dPool_pix_stride, channels))) return e;
  const int G = channels / 8;
  if (256 % G) return set_error(SUNET_ERR_INVALID, "bn_relu_pool_bwd: channels %d unsupported", channels);
  const bool pool = dPool != nullptr;
  if (pool && ((height | width) & 1)) return set_error(SUNET_ERR_INVALID, "bn_relu_pool_bwd: odd size");
  const long long total = (long long)batch * (pool ? height / 2 : height) * (pool ? width / 2 : width) * G;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "bn_relu_pool_bwd: tensor too large (%lld items)", total);
  int blocks = ew_grid(total, 256);
  const int max_blocks = num_sms() * 2;
  if (blocks > max_blocks) blocks = max_blocks;
  const size_t need = ((size_t)blocks * channels * 2 + 3 * (size_t)channels) * sizeof(float);
  if (workspace_bytes < need)
    return set_error(SUNET_ERR_WORKSPACE, "bn_relu_pool_bwd: workspace %zu < %zu", workspace_bytes, need);
  float* partials = reinterpret_cast<float*>(workspace);
  float* coef = partials + (size_t)blocks * channels * 2;
  const __nv_bfloat16* dAp = reinterpret_cast<const __nv_bfloat16*>(dA);
  const __nv_bfloat16* dPp = reinterpret_cast<const __nv_bfloat16*>(dPool);
  const __nv_bfloat16* yp = reinterpret_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* dyp = reinterpret_cast<__nv_bfloat16*>(dy);
  const double count = (double)batch * height * width;
  if (pool)
    launch_k(bn_bwd_pool_reduce_kernel, dim3(blocks), dim3(256), 0, STREAM, dAp, dA_pix_stride, dPp, dPool_pix_stride, yp,
                                                            y_pix_stride, scale, shift, mean, invstd, partials, batch,
                                                            height, width, channels);
  else
    launch_k(bn_bwd_reduce_flat_kernel, dim3(blocks), dim3(256), 0, STREAM, dAp, dA_pix_stride, yp, y_pix_stride, scale, shift, mean,
                                                           invstd, partials, total, channels);
  if ((e = check_launch("bn_bwd_reduce"))) return e;
  launch_k(bn_bwd_finalize_par_kernel, dim3((channels + 7) / 8), dim3(256), 0, STREAM, partials, blocks, channels, count, scale, mean,
                                                                      invstd, dgamma, dbeta, coef);
  if ((e = check_launch("bn_bwd_finalize"))) return e;
  const int ablocks = ew_grid(total, 256);
  if (pool && channels > kPoolApplyMaxC)
    return set_error(SUNET_ERR_INVALID, "bn_relu_pool_bwd: pooled layers take at most %d channels", kPoolApplyMaxC);
  if (pool)
    launch_k(bn_bwd_pool_apply_kernel, dim3(ablocks), dim3(256), 0, STREAM, dAp, dA_pix_stride, dPp, dPool_pix_stride, yp, y_pix_stride,
                                                            scale, shift, coef, dyp, dy_pix_stride, batch, height,
                                                            width, channels);
  else
    launch_k(bn_bwd_apply_flat_kernel, dim3(flat_grid(total)), dim3(256), 0, STREAM, dAp, dA_pix_stride, yp, y_pix_stride, scale, shift,
                                                                    coef, dyp, dy_pix_stride, total, channels);
  return check_launch("bn_bwd_apply");
}

extern "C" int sunet_bn_bwd_apply(const void* dA, int dA_pix_stride, const void* y, int y_pix_stride,
                                  const float* scale, const float* shift, const float* mean, const float* invstd,
                                  const float* partials, int partial_rows, float* dgamma, float* dbeta, void* dy,
                                  int dy_pix_stride, int batch, int height, int width, int channels, void* workspace,
                                  size_t workspace_bytes, sunet_stream_t stream_) {
  if (!dA || !y || !scale || !shift || !mean || !invstd || !partials || partial_rows <= 0 || !dy || !workspace)
    return set_error(SUNET_ERR_INVALID, "bn_bwd_apply: bad arguments");
  int e;
  if ((e = check_act("bn_bwd_apply(y)", y_pix_stride, channels))) return e;
  if ((e = check_act("bn_bwd_apply(dy)", dy_pix_stride, channels))) return e;
  if ((e = check_act("bn_bwd_apply(dA)", dA_pix_stride, channels))) return e;
  const int G = channels / 8;
  if (256 % G) return set_error(SUNET_ERR_INVALID, "bn_bwd_apply: channels %d unsupported", channels);
  const long long total = (long long)batch * height * width * G;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "bn_bwd_apply: tensor too large (%lld items)", total);
  const size_t need = 3 * (size_t)channels * sizeof(float);
  if (workspace_bytes < need)
    return set_error(SUNET_ERR_WORKSPACE, "bn_bwd_apply: workspace %zu < %zu", workspace_bytes, need);
  float* coef = reinterpret_cast<float*>(workspace);
  const double count = (double)batch * height * width;
  launch_k(bn_bwd_finalize_par_kernel, dim3((channels + 7) / 8), dim3(256), 0, STREAM, partials, partial_rows, channels, count, scale,
                                                                      mean, invstd, dgamma, dbeta, coef);
  if ((e = check_launch("bn_bwd_finalize"))) return e;
  launch_k(bn_bwd_apply_flat_kernel, dim3(flat_grid(total)), dim3(256), 0, STREAM, 
      reinterpret_cast<const __nv_bfloat16*>(dA), dA_pix_stride, reinterpret_cast<const __nv_bfloat16*>(y),
      y_pix_stride, scale, shift, coef, reinterpret_cast<__nv_bfloat16*>(dy), dy_pix_stride, total, channels);
  return check_launch("bn_bwd_apply");
}

extern "C" int sunet_bn_pool_bwd_apply(const void* dA, int dA_pix_stride, const void* dPool, int dPool_pix_stride,
                                       const void* y, int y_pix_stride, const float* scale, const float* shift,
                                       const float* mean, const float* invstd, const float* partials0, int rows0,
                                       int stride0, int col0, const float* partials1, int rows1, int stride1,
                                       int col1, float* dgamma, float* dbeta, void* dy, int dy_pix_stride, int batch,
                                       int height, int width, int channels, void* workspace, size_t workspace_bytes,
                                       sunet_stream_t stream_) {
  if (!dA || !dPool || !y || !scale || !shift || !mean || !invstd || !partials0 || !partials1 || rows0 <= 0 ||
      rows1 <= 0 || !dy || !workspace || stride0 < col0 + channels || stride1 < col1 + channels || col0 < 0 || col1 < 0)
    return set_error(SUNET_ERR_INVALID, "bn_pool_bwd_apply: bad arguments");
  int e;
  if ((e = check_act("bn_pool_bwd_apply(y)", y_pix_stride, channels))) return e;
  if ((e = check_act("bn_pool_bwd_apply(dy)", dy_pix_stride, channels))) return e;
  if ((e = check_act("bn_pool_bwd_apply(dA)", dA_pix_stride, channels))) return e;
  if ((e = check_act("bn_pool_bwd_apply(dPool)", dPool_pix_stride, channels))) return e;
  const int G = channels / 8;
  if (256 % G) return set_error(SUNET_ERR_INVALID, "bn_pool_bwd_apply: channels %d unsupported", channels);
  if ((height | width) & 1) return set_error(SUNET_ERR_INVALID, "bn_pool_bwd_apply: odd size");
  const long long total = (long long)batch * (height / 2) * (width / 2) * G;
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "bn_pool_bwd_apply: tensor too large");
  if (workspace_bytes < 3 * (size_t)channels * sizeof(float))
    return set_error(SUNET_ERR_WORKSPACE, "bn_pool_bwd_apply: workspace too small");
  float* coef = reinterpret_cast<float*>(workspace);
  const double count = (double)batch * height * width;
  PartialSrc s0 = {partials0, rows0, stride0, col0}, s1 = {partials1, rows1, stride1, col1};
  launch_k(bn_bwd_finalize2_kernel, dim3((channels + 7) / 8), dim3(256), 0, STREAM, s0, s1, channels, count, scale,
           mean, invstd, dgamma, dbeta, coef);
  if ((e = check_launch("bn_bwd_finalize2"))) return e;
  if (channels > kPoolApplyMaxC)
    return set_error(SUNET_ERR_INVALID, "bn_pool_bwd_apply: pooled layers take at most %d channels", kPoolApplyMaxC);
  launch_k(bn_bwd_pool_apply_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM,
           reinterpret_cast<const __nv_bfloat16*>(dA), dA_pix_stride, reinterpret_cast<const __nv_bfloat16*>(dPool),
           dPool_pix_stride, reinterpret_cast<const __nv_bfloat16*>(y), y_pix_stride, scale, shift, coef,
           reinterpret_cast<__nv_bfloat16*>(dy), dy_pix_stride, batch, height, width, channels);
  return check_launch("bn_pool_bwd_apply");
}

extern "C" int sunet_maxpool2x2(const void* a, int a_pix_stride, void* pooled, int pooled_pix_stride, int batch,
                                int height, int width, int channels, sunet_stream_t stream_) {
  if (!a || !pooled || batch <= 0 || height <= 0 || width <= 0 || ((height | width) & 1))
    return set_error(SUNET_ERR_INVALID, "maxpool2x2: bad arguments");
  int e;
  if ((e = check_act("maxpool2x2(a)", a_pix_stride, channels))) return e;
  if ((e = check_act("maxpool2x2(pooled)", pooled_pix_stride, channels))) return e;
  const long long total = (long long)batch * (height / 2) * (width / 2) * (channels / 8);
  if (total >= (1ll << 31)) return set_error(SUNET_ERR_INVALID, "maxpool2x2: tensor too large");
  launch_k(maxpool2x2_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, STREAM,
           reinterpret_cast<const __nv_bfloat16*>(a), a_pix_stride, reinterpret_cast<__nv_bfloat16*>(pooled),
           pooled_pix_stride, batch, height, width, channels);
  return check_launch("maxpool2x2");
}
