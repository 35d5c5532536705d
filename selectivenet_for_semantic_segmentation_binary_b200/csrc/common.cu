#include "common.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace sunet {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SUNET_OK;
  return set_error(SUNET_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
static std::atomic<long long> g_launches{0};
int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SUNET_PDL");
    return e ? (atoi(e) != 0) : false;
  }();
  return on;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_bf16_5d(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5]) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(SUNET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < 5; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < 4; ++i) gstr[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(SUNET_ERR_INVALID, "tensor base not 16B aligned");
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SUNET_ERR_CUDA,
                     "cuTensorMapEncodeTiled(5d) failed: %d dims=(%llu,%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu,%llu) "
                     "box=(%u,%u,%u,%u,%u)",
                     (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                     (unsigned long long)dims[3], (unsigned long long)dims[4], (unsigned long long)strides_bytes[0],
                     (unsigned long long)strides_bytes[1], (unsigned long long)strides_bytes[2],
                     (unsigned long long)strides_bytes[3], box[0], box[1], box[2], box[3], box[4]);
  return SUNET_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                      uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(SUNET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t bdim[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SUNET_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d cols=%llu rows=%llu stride=%llu box_rows=%u",
                     (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes,
                     box_rows);
  return SUNET_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    n = v;
  }
  return n;
}

}  // namespace sunet
