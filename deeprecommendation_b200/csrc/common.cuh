// Shared device/host helpers for libb200rec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <atomic>

#include "../../include/b200rec.h"

// placed after EVERY kernel launch: checks the launch and counts it (b200rec_launch_count, bench.py's gpu_launches)
#define B200REC_CHECK_LAUNCH()                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return b200rec_set_cuda_error(e__);  \
    b200rec_count_launch();                                      \
  } while (0)

#define B200REC_CUDA(call)                                       \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return b200rec_set_cuda_error(e__);  \
  } while (0)

int b200rec_set_cuda_error(cudaError_t e);          // api.cu: records the message, returns B200REC_ERR_CUDA
int b200rec_fail(int code, const char* msg);        // api.cu: records msg, returns code
void b200rec_count_launch();                        // api.cu

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that launches on a second GPU must opt that
// device in too.  One bit per device ordinal, set once the attribute call succeeded there (thread-safe).
struct B200recSmemOptIn {
  std::atomic<unsigned long long> done{0ull};
};
template <typename K>
static inline cudaError_t b200rec_opt_in_smem(B200recSmemOptIn& s, K kernel, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (s.done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) s.done.fetch_or(bit, std::memory_order_release);
  return e;
}

// number of SMs of the current device (cached)
int b200rec_num_sms();
// L2 cache size of the current device in bytes (cached)
long long b200rec_l2_bytes();

namespace b200rec {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// 128-bit read-only loads of 4 consecutive table elements as fp32 (tables are fp32 or bf16).
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  float4 o;
  o.x = __uint_as_float(r.x << 16);
  o.y = __uint_as_float(r.x & 0xffff0000u);
  o.z = __uint_as_float(r.y << 16);
  o.w = __uint_as_float(r.y & 0xffff0000u);
  return o;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<unsigned*>(&a);
  r.y = *reinterpret_cast<unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- counter-based random bits for training-time dropout masks: Philox2x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3";
// round: (hi, lo) = 0xD256D193 * c0;  c0' = hi ^ key ^ c1;  c1' = lo;  key += 0x9E3779B9).  64 bits per (counter, key); the forward and the
// backward kernel regenerate the same mask from the same (seed, element index) — nothing is stored.
__host__ __device__ __forceinline__ uint2 philox2x32_10(unsigned c0, unsigned c1, unsigned key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long prod = 0xD256D193ull * (unsigned long long)c0;
    const unsigned hi = (unsigned)(prod >> 32), lo = (unsigned)prod;
    c0 = hi ^ key ^ c1;
    c1 = lo;
    key += 0x9E3779B9u;
  }
  return make_uint2(c0, c1);
}

// Dropout multipliers of 4 consecutive hidden units of one (candidate row b, rated item c) pair of AttentionNet (attention_ncf.py:112-117:
// Linear -> ReLU -> Dropout(p) -> Linear): unit group `g` = (hidden index / 4).  keep iff 16 random bits < thr16 = round((1-p)·65536);
// the multiplier of a kept unit is 65536 / thr16 (exactly unbiased for the keep probability actually used).
__device__ __forceinline__ void att_dropout_mult(unsigned key, unsigned thr16, float scale, int b, int c, int g, float (&m)[4]) {
  const uint2 r = philox2x32_10((unsigned)c, ((unsigned)b << 6) | (unsigned)g, key);
  m[0] = (r.x & 0xffffu) < thr16 ? scale : 0.f;
  m[1] = (r.x >> 16) < thr16 ? scale : 0.f;
  m[2] = (r.y & 0xffffu) < thr16 ? scale : 0.f;
  m[3] = (r.y >> 16) < thr16 ? scale : 0.f;
}

}  // namespace b200rec
