// Development probe (not part of the public C ABI): sustained tcgen05.mma rate for one shape with
// shared-memory-resident operands, for each combination of K-major / MN-major A and B.  No TMA, no
// epilogue: it isolates "tensor pipe + shared-memory operand fetch".  Used by scripts/mma_rate.py to
// decide operand layouts.
#include "../../selectivenet_for_semantic_segmentation_binary_b200/csrc/common.h"
#include "../../selectivenet_for_semantic_segmentation_binary_b200/csrc/ptx.cuh"

namespace sunet {

__global__ void __launch_bounds__(128, 1)
mma_probe_kernel(int n, int a_mn, int b_mn, int iters, int lbo_a, int lbo_b, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // deterministic finite operand data
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  fence_proxy_async_smem();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n, a_mn != 0, b_mn != 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // K-major: +32 B per K step inside the 128 B row; MN-major: +16 rows of 128 B per K step
        const uint64_t ad = a_mn ? make_smem_desc_sw128(sa + k * 2048, lbo_a, 1024)
                                 : make_smem_desc_sw128(sa, 16, 1024) + 2 * k;
        const uint64_t bd = b_mn ? make_smem_desc_sw128(sb + k * 2048, lbo_b, 1024)
                                 : make_smem_desc_sw128(sb, 16, 1024) + 2 * k;
        umma_bf16(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}

// CTA-pair variant: M = 256 (128 rows per CTA), each CTA supplies n/2 rows of B; leader issues.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_probe_pair_kernel(int n, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_alloc_pair(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  fence_proxy_async_smem();
  cluster_sync_all();
  const uint32_t tmem = tmem_slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, n, false, false);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc_sw128(sa, 16, 1024) + 2 * k;
        const uint64_t bd = make_smem_desc_sw128(sb, 16, 1024) + 2 * k;
        umma_bf16_pair(tmem + (it & 1) * 256, ad, bd, idesc, 1u);
      }
    }
    umma_commit_pair(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x >> 1] = t1 - t0;
  }
  if (threadIdx.x == 0 && rank == 1) mbar_wait(&bar, 0);   // the multicast commit also lands here
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem, 512);
  }
}

}  // namespace sunet

extern "C" int sunet_dbg_mma_probe_pair(int n, int iters, int pairs, long long* out, void* stream_) {
  using namespace sunet;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n < 32 || n > 256 || n % 16 || iters <= 0 || pairs <= 0 || !out)
    return set_error(SUNET_ERR_INVALID, "mma_probe_pair: bad arguments");
  static bool attr = false;
  if (!attr) {
    int e = check_cuda(cudaFuncSetAttribute(mma_probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            100 * 1024), "cudaFuncSetAttribute(mma_probe_pair)");
    if (e) return e;
    attr = true;
  }
  mma_probe_pair_kernel<<<pairs * 2, 128, 100 * 1024, stream>>>(n, iters, out);
  return check_launch("mma_probe_pair_kernel");
}

// cycles per CTA for `iters` x 4 MMAs of shape 128 x n x 16 written to out[grid]
extern "C" int sunet_dbg_mma_probe(int n, int a_mn, int b_mn, int iters, int grid, long long* out, void* stream_) {
  using namespace sunet;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n < 16 || n > 256 || n % 16 || iters <= 0 || grid <= 0 || !out)
    return set_error(SUNET_ERR_INVALID, "mma_probe: bad arguments");
  static bool attr = false;
  if (!attr) {
    int e = check_cuda(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024),
                       "cudaFuncSetAttribute(mma_probe)");
    if (e) return e;
    attr = true;
  }
  // MN-major blocks of 64 elements are 64 K-rows x 128 B = 8 KB apart
  mma_probe_kernel<<<grid, 128, 100 * 1024, stream>>>(n, a_mn, b_mn, iters, 8192, 8192, out);
  return check_launch("mma_probe_kernel");
}
