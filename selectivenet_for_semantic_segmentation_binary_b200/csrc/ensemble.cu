// Evaluation ensemble: mean of the optionally re-scaled logit maps of several checkpoints,
// taken before the threshold.
//
// Reference semantics: /root/reference/eval.py:209-222
//     outputs.append(fn_tonumpy(scale(net(input))))  for every net;   output = np.mean(np.asarray(outputs), axis=0)
// with scale in {None, clip = np.clip(x, 0, 1), minmax = (x - x.min()) / (x.max() - x.min()) over the whole batch
// tensor, sigmoid = 1 / (1 + exp(-x))} (eval.py:162-179).  np.mean over the leading axis of a float32 array is a
// sequential float32 sum ((a0 + a1) + a2 ...) followed by one float32 division by the count; the kernel performs the
// same IEEE operations in the same order (__fadd_rn / __fdiv_rn, no FMA contraction), so the mean map — and with it
// every thresholded mask — is bit-identical for None / clip / minmax.  `sigmoid` uses CUDA's expf, which is not
// numpy's float32 exp: not bit-pinned (the reference's own sigmoid / clip branches raise on CUDA tensors).
#include <float.h>

#include "common.h"
#include "../../include/sunet_b200.h"

namespace sunet {

static inline int ens_grid(long long items, int threads, int per_sm) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ batch min / max of one fp32 map
__global__ void __launch_bounds__(256)
minmax_partial_kernel(const float* __restrict__ x, long long n, float* __restrict__ partials) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[2][8];
  float mn = FLT_MAX, mx = -FLT_MAX;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long i0 = tid;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long nq = n >> 2;
    for (long long q = tid; q < nq; q += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
      mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
    i0 = (nq << 2) + tid;
  }
  for (long long i = i0; i < n; i += stride) {
    const float v = __ldg(x + i);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = mn;
    red[1][warp] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      mn = fminf(mn, red[0][w]);
      mx = fmaxf(mx, red[1][w]);
    }
    partials[2 * blockIdx.x] = mn;
    partials[2 * blockIdx.x + 1] = mx;
  }
}
__global__ void __launch_bounds__(256) minmax_final_kernel(const float* __restrict__ partials, int blocks,
                                                           float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[2][256];
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (int b = threadIdx.x; b < blocks; b += 256) {
    mn = fminf(mn, partials[2 * b]);
    mx = fmaxf(mx, partials[2 * b + 1]);
  }
  red[0][threadIdx.x] = mn;
  red[1][threadIdx.x] = mx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      red[0][threadIdx.x] = fminf(red[0][threadIdx.x], red[0][threadIdx.x + w]);
      red[1][threadIdx.x] = fmaxf(red[1][threadIdx.x], red[1][threadIdx.x + w]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = red[0][0];
    out[1] = red[1][0];
  }
}

// ------------------------------------------------------------------ ensemble mean
constexpr int kMaxModels = 16;
struct EnsMaps {
  const float* map[kMaxModels];
  const float* minmax[kMaxModels];    // device (min, max) of each map, minmax mode only
};

template <int MODE>
__device__ __forceinline__ float ens_scale(float x, float mn, float range) {
  if (MODE == SUNET_SCALE_CLIP) return fminf(fmaxf(x, 0.f), 1.f);
  if (MODE == SUNET_SCALE_MINMAX) return __fdiv_rn(__fsub_rn(x, mn), range);
  if (MODE == SUNET_SCALE_SIGMOID) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(256)
ensemble_mean_kernel(EnsMaps maps, int M, long long P, float* __restrict__ mean) {
  pdl_wait();
  pdl_trigger();
  float mn[kMaxModels], rg[kMaxModels];
#pragma unroll
  for (int m = 0; m < kMaxModels; ++m) {
    mn[m] = rg[m] = 0.f;
    if (MODE == SUNET_SCALE_MINMAX && m < M) {
      mn[m] = __ldg(maps.minmax[m]);
      rg[m] = __fsub_rn(__ldg(maps.minmax[m] + 1), mn[m]);
    }
  }
  const float count = (float)M;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nq = P >> 2;
  for (long long q = tid; q < nq; q += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < kMaxModels; ++m) {
      if (m < M) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(maps.map[m]) + q);
        const float a = mn[m], r = rg[m];
        const float4 s = make_float4(ens_scale<MODE>(v.x, a, r), ens_scale<MODE>(v.y, a, r),
                                     ens_scale<MODE>(v.z, a, r), ens_scale<MODE>(v.w, a, r));
        if (m == 0) {
          acc = s;
        } else {
          acc.x = __fadd_rn(acc.x, s.x);
          acc.y = __fadd_rn(acc.y, s.y);
          acc.z = __fadd_rn(acc.z, s.z);
          acc.w = __fadd_rn(acc.w, s.w);
        }
      }
    }
    *reinterpret_cast<float4*>(mean + 4 * q) = make_float4(__fdiv_rn(acc.x, count), __fdiv_rn(acc.y, count),
                                                           __fdiv_rn(acc.z, count), __fdiv_rn(acc.w, count));
  }
  for (long long p = (nq << 2) + tid; p < P; p += stride) {      // tail (at most 3 pixels)
    float acc = 0.f;
#pragma unroll
    for (int m = 0; m < kMaxModels; ++m) {
      if (m < M) {
        const float s = ens_scale<MODE>(__ldg(maps.map[m] + p), mn[m], rg[m]);
        acc = (m == 0) ? s : __fadd_rn(acc, s);
      }
    }
    mean[p] = __fdiv_rn(acc, count);
  }
}

}  // namespace sunet

using namespace sunet;

extern "C" int sunet_minmax_f32(const float* x, long long n, float* out, void* workspace, size_t workspace_bytes,
                                sunet_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !out || !workspace || n <= 0) return set_error(SUNET_ERR_INVALID, "minmax_f32: bad arguments");
  const int blocks = ens_grid((n + 3) / 4, 256, 8);
  if (workspace_bytes < (size_t)blocks * 2 * sizeof(float))
    return set_error(SUNET_ERR_WORKSPACE, "minmax_f32: workspace too small");
  float* partials = reinterpret_cast<float*>(workspace);
  launch_k(minmax_partial_kernel, dim3(blocks), dim3(256), 0, stream, x, n, partials);
  int e = check_launch("minmax_partial");
  if (e) return e;
  launch_k(minmax_final_kernel, dim3(1), dim3(256), 0, stream, (const float*)partials, blocks, out);
  return check_launch("minmax_final");
}

extern "C" int sunet_ensemble_mean(const float* const* maps, const float* const* minmax, int n_models,
                                   long long pixels, int scale, float* mean, sunet_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!maps || !mean || pixels <= 0 || n_models < 1 || n_models > kMaxModels)
    return set_error(SUNET_ERR_INVALID, "ensemble_mean: 1..%d maps expected", kMaxModels);
  if (scale == SUNET_SCALE_MINMAX && !minmax)
    return set_error(SUNET_ERR_INVALID, "ensemble_mean: minmax scaling needs the per-map (min, max) pointers");
  EnsMaps em;
  for (int m = 0; m < kMaxModels; ++m) {
    em.map[m] = m < n_models ? maps[m] : nullptr;
    em.minmax[m] = (m < n_models && minmax) ? minmax[m] : nullptr;
    if (m < n_models && (!em.map[m] || (reinterpret_cast<uintptr_t>(em.map[m]) & 15)))
      return set_error(SUNET_ERR_INVALID, "ensemble_mean: map %d is null or not 16-byte aligned", m);
    if (m < n_models && scale == SUNET_SCALE_MINMAX && !em.minmax[m])
      return set_error(SUNET_ERR_INVALID, "ensemble_mean: map %d has no (min, max)", m);
  }
  if (reinterpret_cast<uintptr_t>(mean) & 15) return set_error(SUNET_ERR_INVALID, "ensemble_mean: mean not aligned");
  const int blocks = ens_grid((pixels + 3) / 4, 256, 8);
  switch (scale) {
    case SUNET_SCALE_NONE:
      launch_k(ensemble_mean_kernel<SUNET_SCALE_NONE>, dim3(blocks), dim3(256), 0, stream, em, n_models, pixels, mean);
      break;
    case SUNET_SCALE_CLIP:
      launch_k(ensemble_mean_kernel<SUNET_SCALE_CLIP>, dim3(blocks), dim3(256), 0, stream, em, n_models, pixels, mean);
      break;
    case SUNET_SCALE_MINMAX:
      launch_k(ensemble_mean_kernel<SUNET_SCALE_MINMAX>, dim3(blocks), dim3(256), 0, stream, em, n_models, pixels, mean);
      break;
    case SUNET_SCALE_SIGMOID:
      launch_k(ensemble_mean_kernel<SUNET_SCALE_SIGMOID>, dim3(blocks), dim3(256), 0, stream, em, n_models, pixels, mean);
      break;
    default:
      return set_error(SUNET_ERR_INVALID, "ensemble_mean: bad scale mode %d", scale);
  }
  return check_launch("ensemble_mean");
}
