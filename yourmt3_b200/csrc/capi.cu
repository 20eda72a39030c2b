// C-ABI plumbing: thread-local error string, version and device info.
#include "common.cuh"
#include "../../include/ymt3_b200.h"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void ymt3_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ymt3_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool ymt3_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YMT3_PDL");   // opt-in: measured slower inside the decode graph (common.cuh)
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

extern "C" const char* ymt3_last_error(void) { return g_err; }
extern "C" int ymt3_abi_version(void) { return YMT3_ABI_VERSION; }

extern "C" int ymt3_device_info(char* name, int cap, int* num_sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  YMT3_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  YMT3_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
  if (name && cap > 0) {
    strncpy(name, p.name, cap - 1);
    name[cap - 1] = 0;
  }
  if (num_sms) *num_sms = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return YMT3_OK;
}
