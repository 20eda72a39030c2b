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
