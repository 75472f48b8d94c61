// K7: negative sampling of the ranking datasets ON THE DEVICE.
//
// Reference (src/neural_collaborative_filtering/datasets/base.py:57-78, per SAMPLE, on the host): every ranking sample carries the list of
// the user's candidate negatives with their ratings; `__getitem__` draws one with probability p_k = r_k^w / sum_j r_j^w ('sum_dynamic'; the
// shipped w = 0.0 makes it uniform) through np.random.choice(p=...), i.e. the inverse CDF of ONE uniform: cdf = cumsum(p), pick the first k
// with cdf_k > u.  Here the lists are resident (CSR over samples) and a batch of sample rows is drawn in one launch, warp per sample:
//   u_b    = 53-bit uniform in [0, 1) from Philox2x32-10 (counter = offset + b, key = seed) — numpy's construction of a double from two words
//   choice = first k with cumsum_k(r^w) > u_b * sum(r^w)        (float64; elements of probability 0 are never chosen)
// torch / numpy random streams cannot be bit-matched, so parity is defined GIVEN the uniforms: the kernel returns them, and the test feeds
// the same uniforms to the reference's inverse-CDF rule (tests/test_neg_sample_gpu.py).
#include <math.h>

#include "common.cuh"

namespace b200rec {

constexpr int NEG_WARPS = 8;

__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__global__ void __launch_bounds__(NEG_WARPS * 32)
sample_negatives_kernel(const int64_t* __restrict__ sample_rows, int B, const int64_t* __restrict__ neg_ptr, const int64_t* __restrict__ neg_item,
                        const float* __restrict__ neg_rating, double w, unsigned key, unsigned key_hi, unsigned long long offset,
                        int64_t* __restrict__ out_item, int* __restrict__ out_pos, double* __restrict__ out_u) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * NEG_WARPS + (threadIdx.x >> 5);
  if (b >= B) return;
  const int64_t row = sample_rows[b];
  const int64_t s = neg_ptr[row], e = neg_ptr[row + 1];
  const unsigned long long ctr = offset + (unsigned long long)b;
  const uint2 r = philox2x32_10((unsigned)ctr, (unsigned)(ctr >> 32) ^ key_hi, key);
  const double u = ((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6)) * (1.0 / 9007199254740992.0);
  if (lane == 0 && out_u) out_u[b] = u;
  if (e <= s) {                                    // no candidate negatives: nothing to draw
    if (lane == 0) { out_item[b] = -1; if (out_pos) out_pos[b] = -1; }
    return;
  }
  const bool uniform = (w == 0.0) || neg_rating == nullptr;
  double total = 0.0;
  if (uniform) {
    total = (double)(e - s);
  } else {
    for (int64_t j0 = s; j0 < e; j0 += 32) {
      const int64_t j = j0 + lane;
      const double p = j < e ? pow((double)__ldg(neg_rating + j), w) : 0.0;
      total += __shfl_sync(FULL, warp_inclusive_scan(p, lane), 31);
    }
  }
  const double target = u * total;
  int64_t pick = -1, last_pos = -1;                // last_pos: last element of non-zero probability (the guard for target rounding up to total)
  double carry = 0.0;
  for (int64_t j0 = s; j0 < e && pick < 0; j0 += 32) {
    const int64_t j = j0 + lane;
    const double p = j < e ? (uniform ? 1.0 : pow((double)__ldg(neg_rating + j), w)) : 0.0;
    const double cum = carry + warp_inclusive_scan(p, lane);
    const unsigned hit = __ballot_sync(FULL, j < e && cum > target);
    const unsigned nz = __ballot_sync(FULL, j < e && p > 0.0);
    if (hit) pick = j0 + (__ffs(hit) - 1);
    if (nz) last_pos = j0 + (31 - __clz(nz));
    carry = __shfl_sync(FULL, cum, 31);
  }
  if (pick < 0) pick = last_pos >= 0 ? last_pos : e - 1;
  if (lane == 0) {
    out_item[b] = neg_item[pick];
    if (out_pos) out_pos[b] = (int)(pick - s);
  }
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_sample_negatives(const int64_t* sample_rows, int64_t B, const int64_t* neg_ptr, const int64_t* neg_item, const float* neg_rating,
                                        double w, uint64_t seed, uint64_t offset, int64_t* out_item, int* out_pos, double* out_u,
                                        b200rec_stream_t stream) {
  if (B < 0 || B >= (1ll << 31) - NEG_WARPS) return b200rec_fail(B200REC_ERR_BAD_ARG, "sample_negatives: B out of range");
  if (B == 0) return B200REC_OK;
  if (!sample_rows || !neg_ptr || !neg_item || !out_item) return b200rec_fail(B200REC_ERR_BAD_ARG, "sample_negatives: null argument");
  if (!(w == w)) return b200rec_fail(B200REC_ERR_BAD_ARG, "sample_negatives: w is NaN");
  sample_negatives_kernel<<<ceil_div_i(B, NEG_WARPS), NEG_WARPS * 32, 0, (cudaStream_t)stream>>>(
      sample_rows, (int)B, neg_ptr, neg_item, neg_rating, w, (unsigned)seed, (unsigned)(seed >> 32), (unsigned long long)offset, out_item, out_pos, out_u);
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}
