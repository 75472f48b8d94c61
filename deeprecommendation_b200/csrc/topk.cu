// Row-wise top-k of a score matrix (k <= 64) — the selection step of the reference's only serving path
// (src/webapp/backend.py:113-121: `sort_values(by='score', ascending=False).iloc[:k]` over one user's scores for every
// candidate item) and of per-user top-K over all items (BASELINE config 4; NDCG cut-offs 5/10/20, eval.py:167-169).
// One CTA per row: every thread keeps the best k of its strided share in a small sorted register/local list (one pass
// over the scores, coalesced), the 256 lists are merged by k rounds of a block arg-max over shared memory.
// Ties are broken towards the lower column index, NaNs are never selected.
#include <math.h>

#include "common.cuh"

namespace b200rec {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_MAX = 64;

template <int KMAX>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_rows_kernel(const float* __restrict__ scores, long long ld, int C, int k, float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
  extern __shared__ unsigned char topk_smem[];
  float* s_val = reinterpret_cast<float*>(topk_smem);                    // [TOPK_THREADS * k]
  int* s_idx = reinterpret_cast<int*>(s_val + (size_t)TOPK_THREADS * k);  // [TOPK_THREADS * k]
  __shared__ float r_val[TOPK_THREADS / 32];
  __shared__ int r_idx[TOPK_THREADS / 32], r_own[TOPK_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ row = scores + (long long)blockIdx.x * ld;

  float v[KMAX];
  int ix[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { v[j] = -INFINITY; ix[j] = -1; }
  for (int c = tid; c < C; c += TOPK_THREADS) {
    const float x = __ldg(row + c);
    if (x > v[k - 1] || (ix[k - 1] < 0 && x == x)) {      // better than the current k-th (or list not full); NaN never enters
      // insertion into the descending list (k is small)
      int j = k - 1;
      while (j > 0 && (x > v[j - 1] || ix[j - 1] < 0)) { v[j] = v[j - 1]; ix[j] = ix[j - 1]; --j; }
      v[j] = x; ix[j] = c;
    }
  }
  for (int j = 0; j < k; ++j) { s_val[(size_t)tid * k + j] = v[j]; s_idx[(size_t)tid * k + j] = ix[j]; }
  __syncthreads();
  // k rounds: every thread offers the head of its list; block arg-max (ties -> lower column); the winner pops its head
  int head = 0;
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY;
    int bi = -1, bo = tid;
    if (head < k) {
      bi = s_idx[(size_t)tid * k + head];
      if (bi >= 0) bv = s_val[(size_t)tid * k + head];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL, bv, o);
      const int oi = __shfl_xor_sync(FULL, bi, o), oo = __shfl_xor_sync(FULL, bo, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; bo = oo; }
    }
    if (lane == 0) { r_val[warp] = bv; r_idx[warp] = bi; r_own[warp] = bo; }
    __syncthreads();
    float wv = -INFINITY;
    int wi = -1, wo = -1;
#pragma unroll
    for (int w = 0; w < TOPK_THREADS / 32; ++w) {
      const float cv = r_val[w];
      const int ci = r_idx[w];
      if (ci >= 0 && (wi < 0 || cv > wv || (cv == wv && ci < wi))) { wv = cv; wi = ci; wo = r_own[w]; }
    }
    if (wo == tid) ++head;
    if (tid == 0) {
      out_val[(long long)blockIdx.x * k + round] = wi >= 0 ? wv : -INFINITY;     // fewer than k valid scores -> (-inf, -1)
      out_idx[(long long)blockIdx.x * k + round] = wi;
    }
    __syncthreads();
  }
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_topk_rows(const float* scores, int64_t rows, int64_t cols, int64_t ld, int k, float* out_val, int64_t* out_idx,
                                 b200rec_stream_t stream) {
  if (rows < 0 || cols <= 0 || k <= 0 || k > TOPK_MAX || ld < cols || !scores || !out_val || !out_idx)
    return b200rec_fail(B200REC_ERR_BAD_ARG, "topk_rows: need 1 <= k <= 64 and valid buffers");
  if (rows == 0) return B200REC_OK;
  if (cols > INT32_MAX || rows > INT32_MAX) return b200rec_fail(B200REC_ERR_UNSUPPORTED, "topk_rows: dims > int32");
  const size_t smem = (size_t)TOPK_THREADS * k * (sizeof(float) + sizeof(int));
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 16) topk_rows_kernel<16><<<(unsigned)rows, TOPK_THREADS, smem, st>>>(scores, ld, (int)cols, k, out_val, out_idx);
  else {
    B200REC_CUDA(cudaFuncSetAttribute(topk_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_rows_kernel<64><<<(unsigned)rows, TOPK_THREADS, smem, st>>>(scores, ld, (int)cols, k, out_val, out_idx);
  }
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}
