// Shared helpers for the YourMT3 B200 hot-path library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#if defined(__CUDACC__)
#define YMT3_HD __host__ __device__ __forceinline__
#else
#define YMT3_HD inline
#endif

#define YMT3_OK 0
#define YMT3_ERR_INVALID 1
#define YMT3_ERR_CUDA 2
#define YMT3_ERR_UNSUPPORTED 3

// thread-local last-error string (never throw across the C ABI)
void ymt3_set_error(const char* fmt, ...);

#define YMT3_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ymt3_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                     cudaGetErrorString(_e));                                        \
      return YMT3_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define YMT3_REQUIRE(cond, ...)                                                      \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ymt3_set_error(__VA_ARGS__);                                                   \
      return YMT3_ERR_INVALID;                                                       \
    }                                                                                \
  } while (0)

static inline int ymt3_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// number of SMs of the current device (cached)
int ymt3_num_sms();

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// The decode step is ~100 short dependent kernels replayed from a CUDA graph; with PDL the next kernel's CTAs are
// scheduled (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel
// drains, instead of after it has fully retired.  Contract: a kernel launched through ymt3_launch_pdl must call
// pdl_wait() before its first access to global memory (griddepcontrol.wait returns once every prerequisite grid has
// completed and its writes are visible, so RAW and WAR hazards are both covered) and pdl_launch_dependents() as
// early as it likes.  Both are no-ops for ordinary launches.  MEASURED (A/B on B200, CUDA-graph decode loop): no gain for
// YPTF.MoE+Multi at 3328-6656 sequences (619.6 vs 614.5 ms) and 8 % SLOWER for T5-small at 256 sequences (1093 vs
// 1003 ms per batch) - early-scheduled dependents sit on SM resources while they wait - so the attribute is OFF by
// default and only set when YMT3_PDL=1.
bool ymt3_pdl_enabled();
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t ymt3_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ymt3_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
