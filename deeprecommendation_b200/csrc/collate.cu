// K6: the dynamic-profile collate ON THE DEVICE.
//
// Reference (src/content_providers/dynamic_profiles_provider.py:30-73, per batch, on the host with pandas + sklearn):
//   rated_items_ids = sort(unique(concat(movieId lists of the batch's users)))                              (:59)
//   user_matrix     = multi_hot(lists, rated_items_ids), ones replaced in row-major order by
//                     rating - (meanRating + 2.5) / 2                                                      (:62-66)
// Here the provider keeps every user's list resident in HBM (CSR over users: item numbers ascending per user, the centred rating
// already rounded to fp32 exactly as the reference rounds it: float64 arithmetic, one cast), and a batch of user rows becomes
//   rated (I,) ascending item numbers, and the CSR of user_matrix (row_ptr, col = rank of the item in `rated`, val)
// in three launches of integer work: mark a bitmap of the catalogue, rank it (popcount scan), fill.  Entries whose centred rating is
// exactly 0.0 are dropped from the CSR — in the dense matrix they ARE the "unrated" value (attention_ncf.py:158-159,192) — but their
// items stay in `rated`, as in the reference (the union is taken over the movieId lists).
// Everything is bit-exact against the host collate (tests/test_collate_gpu.py); bytes moved: 8 B per list entry read twice.
#include "common.cuh"

namespace b200rec {

constexpr int COLLATE_WARPS = 8;
constexpr int COLLATE_THREADS = COLLATE_WARPS * 32;
constexpr int COLLATE_SCAN_THREADS = 1024;
constexpr int COLLATE_SMEM_WORDS = 8192;        // catalogue bitmaps up to 262,144 items are first built in shared memory (32 KB)

// warp per batch row: mark the row's items in the catalogue bitmap, count the row's non-zero entries
template <bool SMEM>
__global__ void __launch_bounds__(COLLATE_THREADS)
collate_mark_kernel(const int64_t* __restrict__ user_rows, int B, const int64_t* __restrict__ list_ptr, const int* __restrict__ list_item,
                    const float* __restrict__ list_val, int n_words, unsigned* __restrict__ bits, int* __restrict__ row_nz) {
  extern __shared__ unsigned s_bits[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (SMEM) {
    for (int w = threadIdx.x; w < n_words; w += COLLATE_THREADS) s_bits[w] = 0u;
    __syncthreads();
  }
  const int b = blockIdx.x * COLLATE_WARPS + warp;
  if (b < B) {
    const int64_t u = user_rows[b];
    const int64_t s = list_ptr[u], e = list_ptr[u + 1];
    int nz = 0;
    for (int64_t j = s + lane; j < e; j += 32) {
      const int it = __ldg(list_item + j);
      const unsigned m = 1u << (it & 31);
      if (SMEM) {
        atomicOr(&s_bits[it >> 5], m);
      } else if (!(__ldcg(bits + (it >> 5)) & m)) {        // (a stale read only costs a redundant atomic)
        atomicOr(bits + (it >> 5), m);
      }
      nz += list_val == nullptr ? 1 : (__ldg(list_val + j) != 0.f ? 1 : 0);
    }
    nz = warp_sum_i(nz);
    if (lane == 0) row_nz[b] = nz;
  }
  if (SMEM) {
    __syncthreads();
    for (int w = threadIdx.x; w < n_words; w += COLLATE_THREADS) {
      const unsigned v = s_bits[w];
      if (v) atomicOr(bits + w, v);
    }
  }
}

// exclusive scan of one value per thread over the CTA; returns the thread's prefix, `total` = the CTA's sum (all threads)
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                                 // (s_warp of the previous call has been read by everyone)
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int wsum = lane < (COLLATE_SCAN_THREADS / 32) ? s_warp[lane] : 0, winc = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, winc, o);
    if (lane >= o) winc += t;
  }
  total = __shfl_sync(FULL, winc, 31);
  const int wbase = __shfl_sync(FULL, winc - wsum, warp);
  return wbase + inc - v;
}

// one CTA: rank of every bitmap word (popcount scan) + the ascending list of marked items; exclusive scan of the row counts
__global__ void __launch_bounds__(COLLATE_SCAN_THREADS)
collate_rank_kernel(const unsigned* __restrict__ bits, int n_words, int* __restrict__ word_rank, int64_t* __restrict__ rated,
                    int* __restrict__ row_ptr /* in: counts at [0, B), out: exclusive prefix at [0, B] */, int B, int* __restrict__ counts) {
  __shared__ int s_warp[COLLATE_SCAN_THREADS / 32];
  int carry = 0;
  for (int base = 0; base < n_words; base += COLLATE_SCAN_THREADS) {
    const int w = base + threadIdx.x;
    unsigned m = w < n_words ? bits[w] : 0u;
    int total;
    int p = carry + block_exclusive_scan(__popc(m), s_warp, total);
    if (w < n_words) word_rank[w] = p;
    while (m) {
      rated[p++] = (int64_t)w * 32 + (__ffs(m) - 1);
      m &= m - 1;
    }
    carry += total;
  }
  if (threadIdx.x == 0) counts[0] = carry;
  carry = 0;
  for (int base = 0; base < B; base += COLLATE_SCAN_THREADS) {
    const int b = base + threadIdx.x;
    const int c = b < B ? row_ptr[b] : 0;
    int total;
    const int p = carry + block_exclusive_scan(c, s_warp, total);
    if (b < B) row_ptr[b] = p;
    carry += total;
  }
  if (threadIdx.x == 0) { row_ptr[B] = carry; counts[1] = carry; }
}

// warp per batch row: the kept entries of the row, in list order, with the column = rank of the item among the marked ones
__global__ void __launch_bounds__(COLLATE_THREADS)
collate_fill_kernel(const int64_t* __restrict__ user_rows, int B, const int64_t* __restrict__ list_ptr, const int* __restrict__ list_item,
                    const float* __restrict__ list_val, const unsigned* __restrict__ bits, const int* __restrict__ word_rank,
                    const int* __restrict__ row_ptr, int* __restrict__ um_col, float* __restrict__ um_val) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * COLLATE_WARPS + (threadIdx.x >> 5);
  if (b >= B) return;
  const int64_t u = user_rows[b];
  const int64_t s = list_ptr[u], e = list_ptr[u + 1];
  int out = row_ptr[b];
  for (int64_t j0 = s; j0 < e; j0 += 32) {
    const int64_t j = j0 + lane;
    const bool valid = j < e;
    const int it = valid ? __ldg(list_item + j) : 0;
    const float v = !valid ? 0.f : (list_val == nullptr ? 1.f : __ldg(list_val + j));
    const bool keep = valid && v != 0.f;
    const unsigned mask = __ballot_sync(FULL, keep);
    if (keep) {
      const int pos = out + __popc(mask & ((1u << lane) - 1u));
      const int w = it >> 5;
      um_col[pos] = __ldg(word_rank + w) + __popc(__ldg(bits + w) & ((1u << (it & 31)) - 1u));
      um_val[pos] = v;
    }
    out += __popc(mask);
  }
}

}  // namespace b200rec

using namespace b200rec;

static inline size_t collate_align(size_t n) { return (n + 255) & ~(size_t)255; }

extern "C" size_t b200rec_collate_workspace(int64_t B, int64_t n_items) {
  (void)B;
  const size_t words = (size_t)((n_items + 31) / 32);
  return collate_align(words * 4) * 2;             // bitmap + rank of every bitmap word
}

extern "C" int b200rec_collate_interacted(const int64_t* user_rows, int64_t B, const int64_t* list_ptr, const int* list_item, const float* list_val,
                                          int64_t n_items, int64_t* rated, int* um_row_ptr, int* um_col, float* um_val, int* counts,
                                          void* workspace, size_t workspace_bytes, b200rec_stream_t stream) {
  if (B < 0 || n_items < 0 || B >= (1ll << 31) - COLLATE_SCAN_THREADS || n_items >= (1ll << 31) - 32 * COLLATE_SCAN_THREADS)
    return b200rec_fail(B200REC_ERR_BAD_ARG, "collate_interacted: B / n_items out of range");
  if (!um_row_ptr || !counts || (B > 0 && (!user_rows || !list_ptr)))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "collate_interacted: null argument");
  if (workspace_bytes < b200rec_collate_workspace(B, n_items) || (n_items > 0 && !workspace))
    return b200rec_fail(B200REC_ERR_WORKSPACE, "collate_interacted: workspace too small (b200rec_collate_workspace)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_words = (int)((n_items + 31) / 32);
  unsigned* bits = reinterpret_cast<unsigned*>(workspace);
  int* word_rank = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(workspace) + collate_align((size_t)n_words * 4));
  if (n_words > 0) B200REC_CUDA(cudaMemsetAsync(bits, 0, (size_t)n_words * 4, st));
  const int grid = ceil_div_i(B, COLLATE_WARPS);
  if (B > 0) {
    if (n_words <= COLLATE_SMEM_WORDS) {
      collate_mark_kernel<true><<<grid, COLLATE_THREADS, (size_t)n_words * 4, st>>>(user_rows, (int)B, list_ptr, list_item, list_val, n_words, bits,
                                                                                    um_row_ptr);
    } else {
      collate_mark_kernel<false><<<grid, COLLATE_THREADS, 0, st>>>(user_rows, (int)B, list_ptr, list_item, list_val, n_words, bits, um_row_ptr);
    }
    B200REC_CHECK_LAUNCH();
  }
  collate_rank_kernel<<<1, COLLATE_SCAN_THREADS, 0, st>>>(bits, n_words, word_rank, rated, um_row_ptr, (int)B, counts);
  B200REC_CHECK_LAUNCH();
  if (B > 0) {
    collate_fill_kernel<<<grid, COLLATE_THREADS, 0, st>>>(user_rows, (int)B, list_ptr, list_item, list_val, bits, word_rank, um_row_ptr, um_col,
                                                          um_val);
    B200REC_CHECK_LAUNCH();
  }
  return B200REC_OK;
}
