// Host-side helpers shared by every translation unit of libsunet_b200.so:
// error reporting (thread-local last-error string, integer return codes; nothing
// here throws or exits) and TMA tensor-map construction through the driver entry
// point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

enum {
  SUNET_OK = 0,
  SUNET_ERR_INVALID = 1,   // bad argument / unsupported shape
  SUNET_ERR_CUDA = 2,      // a CUDA runtime / driver call failed
  SUNET_ERR_WORKSPACE = 3  // caller-provided workspace too small
};

int set_error(int code, const char* fmt, ...);
const char* last_error();
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);   // also counts the launch (see sunet_launch_count)
long long launch_count();

// Rank-5 bf16 tensor map with 128-byte swizzle.  dims/strides are in elements /
// bytes, innermost first; strides[0] is implied (2 bytes).  box[0] must be 64
// (128 bytes = the swizzle span).
int make_tmap_bf16_5d(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5]);
// Rank-2 K-major bf16 matrix [rows][cols], box = 64 cols x box_rows rows.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                      uint32_t box_rows);

int num_sms();

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the library starts with pdl_wait() (griddepcontrol.wait: block until the preceding kernel of
// the stream has completed and its writes are visible) and signals pdl_trigger() right away; every launch goes
// through launch_k(), which sets cudaLaunchAttributeProgrammaticStreamSerialization when SUNET_PDL=1.
// OFF by default: measured on B200 inside the CUDA graph of a 16-patch step, 5.02 ms without vs 5.11 ms with
// early triggers and 5.04 ms with implicit triggers (gpurun_out/bench32_*): the graph already hides the launch
// latency and early-resident dependents only take slots.  Without the attribute both instructions are no-ops.
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
#ifndef SUNET_PDL_NO_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace sunet
