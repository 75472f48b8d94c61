// libb200rec: error reporting and device queries shared by every entry point.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void b200rec_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int b200rec_set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
  return B200REC_ERR_CUDA;
}

int b200rec_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int b200rec_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

long long b200rec_l2_bytes() {
  static long long cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 126ll << 20;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrL2CacheSize, dev) != cudaSuccess || n <= 0) n = 126 << 20;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" const char* b200rec_last_error(void) { return g_err; }
extern "C" int b200rec_version(void) { return B200REC_VERSION; }
extern "C" int b200rec_sm_count(void) { return b200rec_num_sms(); }
extern "C" int64_t b200rec_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
