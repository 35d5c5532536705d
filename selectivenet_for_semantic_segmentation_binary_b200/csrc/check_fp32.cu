// fp32 CHECK MODE (BASELINE.json north_star: "logits and loss must agree within 2e-2 relative in bf16, or 1e-4 in an
// fp32 check mode").  Slow, obviously-correct SIMT fp32 twins of every op of the hot path: NHWC fp32 activations,
// parameters read in the reference's own layouts (no repack), fp64 accumulation inside every contraction and every
// reduction over pixels (values are STORED in fp32: the only roundings are one per stored element).  Selected with
// SUNET_CHECK_FP32=1 (engine_fp32.SUNetPlanF32); the losses, the counting kernel and Adam are already fp32 and are
// shared with the fast path.  Test infrastructure of
// the product path's ORCHESTRATION (layer order, concat order, BatchNorm bookkeeping, pool routing, loss plumbing):
// with fp32 numerics a wrong tap, a swapped concat half or a missed term shows up at 1e-1, not inside bf16 noise.
// Nothing here is tuned; it is never used by bench.py.
//
// Reference semantics: model.py:9-15 (Conv3x3 + BN + ReLU), :31 (MaxPool2d(2), first maximum), :44-45
// (ConvTranspose2d k2 s2), :83 (concat [up | skip]), :96-101 (1x1 heads); backward = autograd of the same (train.py:208).
#include "common.h"
#include "../../include/sunet_b200.h"

namespace sunet {

static inline int f32_grid(long long items, int threads = 256) {
  long long b = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ conv3x3 (pad 1), two-source input
__global__ void __launch_bounds__(256)
f32_conv3x3_fwd_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
                       const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y, int B, int H,
                       int W, int Cout) {
  const int Cin = c0 + c1;
  const long long total = (long long)B * H * W * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long p = i / Cout;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    double acc = bias ? (double)bias[co] : 0.0;
    for (int r = 0; r < 3; ++r) {
      const int iy = yy + r - 1;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = xx + s - 1;
        if (ix < 0 || ix >= W) continue;
        const long long q = ((long long)n * H + iy) * W + ix;
        const float* wp = w + ((long long)co * Cin) * 9 + r * 3 + s;
        const float* a = x0 + q * c0;
        for (int ci = 0; ci < c0; ++ci) acc += (double)a[ci] * (double)wp[(long long)ci * 9];
        if (c1) {
          const float* b = x1 + q * c1;
          for (int ci = 0; ci < c1; ++ci) acc += (double)b[ci] * (double)wp[(long long)(c0 + ci) * 9];
        }
      }
    }
    y[i] = (float)acc;
  }
}

// dx[p][ci] = sum_{r,s,co} dy[p - (r-1, s-1)][co] * w[co][ci][r][s]; channels [0,c0) -> dx0, [c0,c0+c1) -> dx1
__global__ void __launch_bounds__(256)
f32_conv3x3_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx0, int c0,
                         float* __restrict__ dx1, int c1, int B, int H, int W, int Cout) {
  const int Cin = c0 + c1;
  const long long total = (long long)B * H * W * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long p = i / Cin;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    double acc = 0.0;
    for (int r = 0; r < 3; ++r) {
      const int oy = yy - (r - 1);
      if (oy < 0 || oy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ox = xx - (s - 1);
        if (ox < 0 || ox >= W) continue;
        const float* g = dy + (((long long)n * H + oy) * W + ox) * Cout;
        const float* wp = w + (long long)ci * 9 + r * 3 + s;
        for (int co = 0; co < Cout; ++co) acc += (double)g[co] * (double)wp[(long long)co * Cin * 9];
      }
    }
    if (ci < c0) dx0[p * c0 + ci] = (float)acc;
    else dx1[p * c1 + (ci - c0)] = (float)acc;
  }
}

// dw[co][ci][r][s] = sum_p dy[p][co] * x[p + (r-1, s-1)][ci]: one block per (co, ci), fp64 accumulation over pixels
__global__ void __launch_bounds__(128)
f32_conv3x3_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x0, int c0,
                         const float* __restrict__ x1, int c1, float* __restrict__ dw, int B, int H, int W, int Cout) {
  const int Cin = c0 + c1;
  const int co = blockIdx.x / Cin, ci = blockIdx.x % Cin;
  const float* x = ci < c0 ? x0 : x1;
  const int cs = ci < c0 ? c0 : c1, cc = ci < c0 ? ci : ci - c0;
  double acc[9];
  for (int t = 0; t < 9; ++t) acc[t] = 0.0;
  const long long P = (long long)B * H * W;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const float g = dy[p * Cout + co];
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    for (int r = 0; r < 3; ++r) {
      const int iy = yy + r - 1;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = xx + s - 1;
        if (ix < 0 || ix >= W) continue;
        acc[r * 3 + s] += (double)g * (double)x[(p + (long long)(r - 1) * W + (s - 1)) * cs + cc];
      }
    }
  }
  __shared__ double red[128];
  for (int t = 0; t < 9; ++t) {
    red[threadIdx.x] = acc[t];
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
      if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
      __syncthreads();
    }
    if (threadIdx.x == 0) dw[((long long)co * Cin + ci) * 9 + t] = (float)red[0];
    __syncthreads();
  }
}

// ------------------------------------------------------------------ ConvTranspose2d k2 s2: w [Cin][Cout][2][2]
__global__ void __launch_bounds__(256)
f32_convT_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ y, int B, int h, int wd, int Cin, int Cout) {
  const long long total = (long long)B * 2 * h * 2 * wd * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long p = i / Cout;
    const int X = (int)(p % (2 * wd)), Y = (int)((p / (2 * wd)) % (2 * h)), n = (int)(p / ((long long)4 * wd * h));
    const float* a = x + (((long long)n * h + (Y >> 1)) * wd + (X >> 1)) * Cin;
    const float* wp = w + (long long)co * 4 + (Y & 1) * 2 + (X & 1);
    double acc = (double)bias[co];
    for (int ci = 0; ci < Cin; ++ci) acc += (double)a[ci] * (double)wp[(long long)ci * Cout * 4];
    y[i] = (float)acc;
  }
}
__global__ void __launch_bounds__(256)
f32_convT_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int h,
                       int wd, int Cin, int Cout) {
  const long long total = (long long)B * h * wd * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long p = i / Cin;
    const int xx = (int)(p % wd), yy = (int)((p / wd) % h), n = (int)(p / ((long long)wd * h));
    double acc = 0.0;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const float* g = dy + (((long long)n * 2 * h + 2 * yy + a) * 2 * wd + 2 * xx + b) * Cout;
        const float* wp = w + (long long)ci * Cout * 4 + a * 2 + b;
        for (int co = 0; co < Cout; ++co) acc += (double)g[co] * (double)wp[(long long)co * 4];
      }
    dx[i] = (float)acc;
  }
}
// dw[ci][co][a][b] = sum_p x[p][ci] * dy[2p + (a,b)][co]; dbias[co] = sum over all output pixels of dy (block co when ci == 0)
__global__ void __launch_bounds__(128)
f32_convT_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                       float* __restrict__ dbias, int B, int h, int wd, int Cin, int Cout) {
  const int ci = blockIdx.x / Cout, co = blockIdx.x % Cout;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const long long P = (long long)B * h * wd;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const int xx = (int)(p % wd), yy = (int)((p / wd) % h), n = (int)(p / ((long long)wd * h));
    const double v = (double)x[p * Cin + ci];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        acc[a * 2 + b] += v * (double)dy[(((long long)n * 2 * h + 2 * yy + a) * 2 * wd + 2 * xx + b) * Cout + co];
  }
  __shared__ double red[128];
  for (int t = 0; t < 4; ++t) {
    red[threadIdx.x] = acc[t];
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
      if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
      __syncthreads();
    }
    if (threadIdx.x == 0) dw[((long long)ci * Cout + co) * 4 + t] = (float)red[0];
    __syncthreads();
  }
  if (ci == 0 && dbias) {
    double s = 0.0;
    const long long PO = (long long)B * 4 * h * wd;
    for (long long p = threadIdx.x; p < PO; p += blockDim.x) s += (double)dy[p * Cout + co];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
      if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
      __syncthreads();
    }
    if (threadIdx.x == 0) dbias[co] = (float)red[0];
  }
}

// ------------------------------------------------------------------ BatchNorm (training) + ReLU (+ pool)
// one block per channel: batch statistics in fp64, then scale/shift/mean/invstd and the running-stat update
// (y is the BIAS-FREE conv output: the conv bias only shifts the mean, model.py:11-12)
__global__ void __launch_bounds__(256)
f32_bn_stats_kernel(const float* __restrict__ y, long long P, int C, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ conv_bias, float* running_mean,
                    float* running_var, long long* nbt, float momentum, float eps, float* scale, float* shift,
                    float* mean, float* invstd) {
  const int c = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const double v = (double)y[p * C + c];
    s += v;
    q += v * v;
  }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      rs[threadIdx.x] += rs[threadIdx.x + k];
      rq[threadIdx.x] += rq[threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = (double)P;
    const double mu = rs[0] / n;
    double var = rq[0] / n - mu * mu;
    if (var < 0.0) var = 0.0;
    const double istd = 1.0 / sqrt(var + (double)eps);
    const double sc = (double)gamma[c] * istd;
    scale[c] = (float)sc;
    shift[c] = (float)((double)beta[c] - mu * sc);
    mean[c] = (float)mu;
    invstd[c] = (float)istd;
    if (running_mean) {
      const double b = conv_bias ? (double)conv_bias[c] : 0.0;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * (mu + b));
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * var * n / (n - 1.0));
    }
    if (c == 0 && nbt) *nbt += 1;
  }
}
__global__ void __launch_bounds__(256)
f32_bn_relu_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                   float* __restrict__ a, long long total, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    a[i] = fmaxf((float)((double)y[i] * (double)scale[c] + (double)shift[c]), 0.f);
  }
}
__global__ void __launch_bounds__(256)
f32_maxpool_kernel(const float* __restrict__ a, float* __restrict__ pooled, int B, int H, int W, int C) {
  const int h2 = H >> 1, w2 = W >> 1;
  const long long total = (long long)B * h2 * w2 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int xx = (int)(p % w2), yy = (int)((p / w2) % h2), n = (int)(p / ((long long)w2 * h2));
    const float* base = a + (((long long)n * H + 2 * yy) * W + 2 * xx) * C + c;
    float m = base[0];
    m = fmaxf(m, base[C]);
    m = fmaxf(m, base[(long long)W * C]);
    m = fmaxf(m, base[(long long)W * C + C]);
    pooled[i] = m;
  }
}
// g = (dA + dPool routed to the FIRST maximum of its 2x2 window of a) * (a > 0), written to gbuf; both inputs optional
__global__ void __launch_bounds__(256)
f32_relu_pool_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ dPool, const float* __restrict__ a,
                         float* __restrict__ g, int B, int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int xx = (int)(p % W), yy = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    float v = dA ? dA[i] : 0.f;
    if (dPool) {
      const int wy = yy >> 1, wx = xx >> 1;
      const float* base = a + (((long long)n * H + 2 * wy) * W + 2 * wx) * C + c;
      const float v00 = base[0], v01 = base[C], v10 = base[(long long)W * C], v11 = base[(long long)W * C + C];
      int win = 0;
      float m = v00;
      if (v01 > m) { m = v01; win = 1; }
      if (v10 > m) { m = v10; win = 2; }
      if (v11 > m) { m = v11; win = 3; }
      if (win == ((yy & 1) * 2 + (xx & 1))) v += dPool[(((long long)n * (H >> 1) + wy) * (W >> 1) + wx) * C + c];
    }
    g[i] = a[i] > 0.f ? v : 0.f;
  }
}
// per channel: dbeta = sum g, dgamma = sum g*xhat (fp64), then coefficients for the apply pass
__global__ void __launch_bounds__(256)
f32_bn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y, long long P, int C,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float* dgamma, float* dbeta,
                         double* __restrict__ sums) {
  const int c = blockIdx.x;
  const double mu = (double)mean[c], istd = (double)invstd[c];
  double s = 0.0, q = 0.0;
  for (long long p = threadIdx.x; p < P; p += blockDim.x) {
    const double gv = (double)g[p * C + c];
    s += gv;
    q += gv * ((double)y[p * C + c] - mu) * istd;
  }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      rs[threadIdx.x] += rs[threadIdx.x + k];
      rq[threadIdx.x] += rq[threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dbeta[c] = (float)rs[0];
    dgamma[c] = (float)rq[0];
    sums[2 * c] = rs[0];
    sums[2 * c + 1] = rq[0];
  }
}
// dy = scale * (g - sum g / n - xhat * sum g*xhat / n), in place over g
__global__ void __launch_bounds__(256)
f32_bn_bwd_apply_kernel(float* __restrict__ g, const float* __restrict__ y, long long total, int C, double n,
                        const float* __restrict__ scale, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const double* __restrict__ sums) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const double xhat = ((double)y[i] - (double)mean[c]) * (double)invstd[c];
    g[i] = (float)((double)scale[c] * ((double)g[i] - sums[2 * c] / n - xhat * sums[2 * c + 1] / n));
  }
}

// ------------------------------------------------------------------ heads (nheads x 64 -> 1)
__global__ void __launch_bounds__(256)
f32_heads_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ b, int nheads,
                     float* __restrict__ logits, long long P, int C) {
  const long long total = P * nheads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i / P);
    const long long p = i % P;
    double acc = (double)b[h];
    for (int c = 0; c < C; ++c) acc += (double)a[p * C + c] * (double)w[h * C + c];
    logits[i] = (float)acc;
  }
}
__global__ void __launch_bounds__(256)
f32_heads_bwd_dA_kernel(const float* __restrict__ dl, const float* __restrict__ w, int nheads, float* __restrict__ dA,
                        long long P, int C, int accumulate) {
  const long long total = P * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    double acc = accumulate ? (double)dA[i] : 0.0;
    for (int h = 0; h < nheads; ++h) acc += (double)dl[(long long)h * P + p] * (double)w[h * C + c];
    dA[i] = (float)acc;
  }
}
// one block per (head, channel | bias): dw[h][c] = sum_p dl[h][p] * a[p][c]; db[h] = sum_p dl[h][p]
__global__ void __launch_bounds__(256)
f32_heads_bwd_dw_kernel(const float* __restrict__ dl, const float* __restrict__ a, float* __restrict__ dw,
                        float* __restrict__ db, long long P, int C) {
  const int h = blockIdx.x / (C + 1), c = blockIdx.x % (C + 1);
  double s = 0.0;
  for (long long p = threadIdx.x; p < P; p += blockDim.x)
    s += (double)dl[(long long)h * P + p] * (c < C ? (double)a[p * C + c] : 1.0);
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (c < C) dw[h * C + c] = (float)red[0];
    else db[h] = (float)red[0];
  }
}

}  // namespace sunet

using namespace sunet;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

extern "C" int sunet_f32_conv3x3_fwd(const float* x0, int c0, const float* x1, int c1, const float* w, const float* bias,
                                     float* y, int batch, int height, int width, int cout, sunet_stream_t stream_) {
  if (!x0 || c0 <= 0 || (c1 > 0 && !x1) || !w || !y || batch <= 0 || height <= 0 || width <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_conv3x3_fwd: bad arguments");
  launch_k(f32_conv3x3_fwd_kernel, dim3(f32_grid((long long)batch * height * width * cout)), dim3(256), 0, STREAM, x0, c0,
           x1, c1, w, bias, y, batch, height, width, cout);
  return check_launch("f32_conv3x3_fwd");
}
extern "C" int sunet_f32_conv3x3_dgrad(const float* dy, const float* w, float* dx0, int c0, float* dx1, int c1, int batch,
                                       int height, int width, int cout, sunet_stream_t stream_) {
  if (!dy || !w || !dx0 || c0 <= 0 || (c1 > 0 && !dx1) || batch <= 0 || height <= 0 || width <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_conv3x3_dgrad: bad arguments");
  launch_k(f32_conv3x3_dgrad_kernel, dim3(f32_grid((long long)batch * height * width * (c0 + c1))), dim3(256), 0, STREAM,
           dy, w, dx0, c0, dx1, c1, batch, height, width, cout);
  return check_launch("f32_conv3x3_dgrad");
}
extern "C" int sunet_f32_conv3x3_wgrad(const float* dy, const float* x0, int c0, const float* x1, int c1, float* dw,
                                       int batch, int height, int width, int cout, sunet_stream_t stream_) {
  if (!dy || !x0 || c0 <= 0 || (c1 > 0 && !x1) || !dw || batch <= 0 || height <= 0 || width <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_conv3x3_wgrad: bad arguments");
  launch_k(f32_conv3x3_wgrad_kernel, dim3(cout * (c0 + c1)), dim3(128), 0, STREAM, dy, x0, c0, x1, c1, dw, batch, height,
           width, cout);
  return check_launch("f32_conv3x3_wgrad");
}
extern "C" int sunet_f32_convT_fwd(const float* x, const float* w, const float* bias, float* y, int batch, int height,
                                   int width, int cin, int cout, sunet_stream_t stream_) {
  if (!x || !w || !bias || !y || batch <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_convT_fwd: bad arguments");
  launch_k(f32_convT_fwd_kernel, dim3(f32_grid((long long)batch * 4 * height * width * cout)), dim3(256), 0, STREAM, x, w,
           bias, y, batch, height, width, cin, cout);
  return check_launch("f32_convT_fwd");
}
extern "C" int sunet_f32_convT_dgrad(const float* dy, const float* w, float* dx, int batch, int height, int width,
                                     int cin, int cout, sunet_stream_t stream_) {
  if (!dy || !w || !dx || batch <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_convT_dgrad: bad arguments");
  launch_k(f32_convT_dgrad_kernel, dim3(f32_grid((long long)batch * height * width * cin)), dim3(256), 0, STREAM, dy, w,
           dx, batch, height, width, cin, cout);
  return check_launch("f32_convT_dgrad");
}
extern "C" int sunet_f32_convT_wgrad(const float* dy, const float* x, float* dw, float* dbias, int batch, int height,
                                     int width, int cin, int cout, sunet_stream_t stream_) {
  if (!dy || !x || !dw || batch <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_convT_wgrad: bad arguments");
  launch_k(f32_convT_wgrad_kernel, dim3(cin * cout), dim3(128), 0, STREAM, dy, x, dw, dbias, batch, height, width, cin,
           cout);
  return check_launch("f32_convT_wgrad");
}
extern "C" int sunet_f32_bn_stats(const float* y, long long pixels, int channels, const float* gamma, const float* beta,
                                  const float* conv_bias, float* running_mean, float* running_var,
                                  long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                                  float* mean, float* invstd, sunet_stream_t stream_) {
  if (!y || pixels <= 1 || channels <= 0 || !gamma || !beta || !scale || !shift || !mean || !invstd)
    return set_error(SUNET_ERR_INVALID, "f32_bn_stats: bad arguments");
  launch_k(f32_bn_stats_kernel, dim3(channels), dim3(256), 0, STREAM, y, pixels, channels, gamma, beta, conv_bias,
           running_mean, running_var, num_batches_tracked, momentum, eps, scale, shift, mean, invstd);
  return check_launch("f32_bn_stats");
}
extern "C" int sunet_f32_bn_relu_pool(const float* y, const float* scale, const float* shift, float* a, float* pooled,
                                      int batch, int height, int width, int channels, sunet_stream_t stream_) {
  if (!y || !scale || !shift || !a || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 ||
      (pooled && ((height | width) & 1)))
    return set_error(SUNET_ERR_INVALID, "f32_bn_relu_pool: bad arguments");
  const long long total = (long long)batch * height * width * channels;
  launch_k(f32_bn_relu_kernel, dim3(f32_grid(total)), dim3(256), 0, STREAM, y, scale, shift, a, total, channels);
  int e = check_launch("f32_bn_relu");
  if (e || !pooled) return e;
  launch_k(f32_maxpool_kernel, dim3(f32_grid(total / 4)), dim3(256), 0, STREAM, (const float*)a, pooled, batch, height,
           width, channels);
  return check_launch("f32_maxpool");
}
extern "C" int sunet_f32_bn_relu_pool_bwd(const float* dA, const float* dPool, const float* y, const float* a,
                                          const float* scale, const float* mean, const float* invstd, float* dgamma,
                                          float* dbeta, float* dy, int batch, int height, int width, int channels,
                                          void* workspace, size_t workspace_bytes, sunet_stream_t stream_) {
  if ((!dA && !dPool) || !y || !a || !scale || !mean || !invstd || !dgamma || !dbeta || !dy || !workspace || batch <= 0 ||
      height <= 0 || width <= 0 || channels <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_bn_relu_pool_bwd: bad arguments");
  if (workspace_bytes < (size_t)channels * 2 * sizeof(double))
    return set_error(SUNET_ERR_WORKSPACE, "f32_bn_relu_pool_bwd: workspace too small");
  const long long P = (long long)batch * height * width, total = P * channels;
  double* sums = reinterpret_cast<double*>(workspace);
  launch_k(f32_relu_pool_bwd_kernel, dim3(f32_grid(total)), dim3(256), 0, STREAM, dA, dPool, a, dy, batch, height, width,
           channels);
  int e = check_launch("f32_relu_pool_bwd");
  if (e) return e;
  launch_k(f32_bn_bwd_reduce_kernel, dim3(channels), dim3(256), 0, STREAM, (const float*)dy, y, P, channels, mean, invstd,
           dgamma, dbeta, sums);
  if ((e = check_launch("f32_bn_bwd_reduce"))) return e;
  launch_k(f32_bn_bwd_apply_kernel, dim3(f32_grid(total)), dim3(256), 0, STREAM, dy, y, total, channels, (double)P, scale,
           mean, invstd, (const double*)sums);
  return check_launch("f32_bn_bwd_apply");
}
extern "C" int sunet_f32_heads_fwd(const float* a, const float* w, const float* b, int nheads, float* logits,
                                   long long pixels, int channels, sunet_stream_t stream_) {
  if (!a || !w || !b || !logits || nheads <= 0 || pixels <= 0 || channels <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_heads_fwd: bad arguments");
  launch_k(f32_heads_fwd_kernel, dim3(f32_grid(pixels * nheads)), dim3(256), 0, STREAM, a, w, b, nheads, logits, pixels,
           channels);
  return check_launch("f32_heads_fwd");
}
extern "C" int sunet_f32_heads_bwd(const float* dlogits, const float* a, const float* w, int nheads, float* dA,
                                   int accumulate, float* dw, float* db, long long pixels, int channels,
                                   sunet_stream_t stream_) {
  if (!dlogits || !a || !w || !dA || !dw || !db || nheads <= 0 || pixels <= 0 || channels <= 0)
    return set_error(SUNET_ERR_INVALID, "f32_heads_bwd: bad arguments");
  launch_k(f32_heads_bwd_dA_kernel, dim3(f32_grid(pixels * channels)), dim3(256), 0, STREAM, dlogits, w, nheads, dA,
           pixels, channels, accumulate);
  int e = check_launch("f32_heads_bwd_dA");
  if (e) return e;
  launch_k(f32_heads_bwd_dw_kernel, dim3(nheads * (channels + 1)), dim3(256), 0, STREAM, dlogits, a, dw, db, pixels,
           channels);
  return check_launch("f32_heads_bwd_dw");
}
