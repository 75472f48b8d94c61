// K1s — projection of SPARSE profile rows: warp-vectorised gather-sum over the transposed weight ("EmbeddingBag").
//
// The reference feeds one-hot rows (src/content_providers/one_hot_provider.py:17-21, fixed_profiles_provider.py:49-50) and the
// multi-hot columns of the item profiles (21 genre + 945 personnel columns, ~1 % dense) through the same
// nn.Linear as dense columns (models/basic_ncf.py:38-39, attention_ncf.py:150-151, gnn_ncf.py:300-301): a (B, K) x (K, N) GEMM
// that multiplies by zeros.  Here a row is its list of non-zero columns:
//
//     Y[m] (+)= bias + sum_{k in row m} val[k] * Wt[col[k]]          Wt = W^T, (K, N) row-major: one 128-bit load per lane and non-zero
//
// `accumulate` adds onto a Y that the dense-column GEMM (K1a over the dense slice of a mixed profile) has already written, so a
// mixed row costs one GEMM over its dense columns + one gather-sum over its non-zeros.  One-hot rows are the degenerate case
// (`ids`: one table row per output row) — a pure embedding lookup.  G = N/4 (rounded up to a power of two <= 32) lanes own one
// output row, 32/G rows advance per warp, 8 table rows are in flight per lane group; sums run in list order: deterministic.
// Bound: HBM gather — (N*s + 4 [+4]) bytes per non-zero ("K1 projection (one-hot / multi-hot input)").
#include <algorithm>

#include "common.cuh"

namespace b200rec {

struct GatherSumParams {
  const int* row_ptr;        // (M + 1) or null (one-hot: ids)
  const int* col;            // (nnz)
  const float* val;          // (nnz) or null = 1
  const long long* ids;      // (M) one-hot column per row (row_ptr == null); negative = empty row
  long long M;
  const void* Wt;            // (K, N) fp32 or bf16
  long long ldwt;
  int N;
  const float* bias;         // (N) or null
  float* Y;
  long long ldy;
  int accumulate;
};

constexpr int GS_WARPS = 8;

template <int G, int NV, typename T>
__global__ void __launch_bounds__(GS_WARPS * 32) gather_sum_kernel(const GatherSumParams p) {
  constexpr int RPW = 32 / G;                                  // rows per warp
  const int lane = threadIdx.x & 31, g = lane / G, sl = lane % G;
  const long long warp0 = (long long)blockIdx.x * GS_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * GS_WARPS * RPW;
  for (long long row = warp0 * RPW + g; row < p.M; row += stride) {
    float acc[NV][4];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
      const int c = (sl + nv * G) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < p.N) {
        if (p.accumulate) a = *reinterpret_cast<const float4*>(p.Y + row * p.ldy + c);
        else if (p.bias) a = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      }
      acc[nv][0] = a.x; acc[nv][1] = a.y; acc[nv][2] = a.z; acc[nv][3] = a.w;
    }
    if (p.row_ptr == nullptr) {                                // one-hot: a single table row
      const long long id = __ldg(p.ids + row);
      if (id >= 0) {
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
          const int c = (sl + nv * G) * 4;
          if (c < p.N) {
            const float4 x = ld4(reinterpret_cast<const T*>(p.Wt) + id * p.ldwt + c);
            acc[nv][0] += x.x; acc[nv][1] += x.y; acc[nv][2] += x.z; acc[nv][3] += x.w;
          }
        }
      }
    } else {
      const int s = __ldg(p.row_ptr + row), e = __ldg(p.row_ptr + row + 1);
      for (int k0 = s; k0 < e; k0 += 8) {
        int c8[8];
        float v8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {                          // the G lanes of a row read the same index word (one broadcast sector)
          const bool ok = k0 + u < e;
          c8[u] = ok ? __ldg(p.col + k0 + u) : 0;
          v8[u] = ok ? (p.val ? __ldg(p.val + k0 + u) : 1.f) : 0.f;
        }
        float4 x[8][NV];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int nv = 0; nv < NV; ++nv) {
            const int c = (sl + nv * G) * 4;
            x[u][nv] = ld4(reinterpret_cast<const T*>(p.Wt) + (long long)c8[u] * p.ldwt + (c < p.N ? c : 0));
          }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int nv = 0; nv < NV; ++nv) {
            acc[nv][0] = fmaf(v8[u], x[u][nv].x, acc[nv][0]); acc[nv][1] = fmaf(v8[u], x[u][nv].y, acc[nv][1]);
            acc[nv][2] = fmaf(v8[u], x[u][nv].z, acc[nv][2]); acc[nv][3] = fmaf(v8[u], x[u][nv].w, acc[nv][3]);
          }
      }
    }
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
      const int c = (sl + nv * G) * 4;
      if (c < p.N) *reinterpret_cast<float4*>(p.Y + row * p.ldy + c) = make_float4(acc[nv][0], acc[nv][1], acc[nv][2], acc[nv][3]);
    }
  }
}

template <int G, int NV, typename T>
static int gs_launch(const GatherSumParams& p, cudaStream_t st) {
  const long long rows_per_cta = (long long)GS_WARPS * (32 / G);
  const long long want = (p.M + rows_per_cta - 1) / rows_per_cta;
  const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)b200rec_num_sms() * 32));   // grid-stride beyond 32 CTAs per SM
  gather_sum_kernel<G, NV, T><<<grid, GS_WARPS * 32, 0, st>>>(p);
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}

template <typename T>
static int gs_dispatch(const GatherSumParams& p, cudaStream_t st) {
  const int q = (p.N + 3) / 4;                                  // float4 columns of a row
  if (q <= 8) return gs_launch<8, 1, T>(p, st);
  if (q <= 16) return gs_launch<16, 1, T>(p, st);
  if (q <= 32) return gs_launch<32, 1, T>(p, st);
  if (q <= 64) return gs_launch<32, 2, T>(p, st);
  if (q <= 128) return gs_launch<32, 4, T>(p, st);
  return b200rec_fail(B200REC_ERR_UNSUPPORTED, "linear_sparse: N > 512");
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_linear_sparse(const int* row_ptr, const int* col, const float* val, const int64_t* ids, int64_t M, const void* Wt,
                                     int64_t K, int64_t N, int64_t ldwt, int wt_dtype, const float* bias, float* Y, int64_t ldy, int accumulate,
                                     b200rec_stream_t stream) {
  if (M == 0) return B200REC_OK;
  if (M < 0 || K <= 0 || N <= 0 || !Wt || !Y) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear_sparse: null / empty operand");
  if ((row_ptr == nullptr) == (ids == nullptr)) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear_sparse: give either (row_ptr, col[, val]) or ids");
  if (row_ptr && !col) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear_sparse: row_ptr without col");
  const int esz = wt_dtype == B200REC_BF16 ? 2 : 4;
  if (wt_dtype != B200REC_F32 && wt_dtype != B200REC_BF16) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear_sparse: bad wt_dtype");
  if ((N % 4) || (ldwt % 4) || (ldy % 4) || ldwt < N || ldy < N || ((uintptr_t)Wt % 16) || ((uintptr_t)Y % 16) || (bias && ((uintptr_t)bias % 16)) ||
      ((ldwt * esz) % 8))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "linear_sparse: N, ldwt, ldy must be multiples of 4 and Wt / Y / bias 16-byte aligned");
  GatherSumParams p;
  p.row_ptr = row_ptr; p.col = col; p.val = val; p.ids = reinterpret_cast<const long long*>(ids); p.M = M; p.Wt = Wt; p.ldwt = ldwt; p.N = (int)N;
  p.bias = bias; p.Y = Y; p.ldy = ldy; p.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  return wt_dtype == B200REC_F32 ? gs_dispatch<float>(p, st) : gs_dispatch<__nv_bfloat16>(p, st);
}

// dense rows -> CSR of their non-zero entries in columns [c0, c1) (column numbers relative to c0), for callers that only hold the
// dense (B, K) profile matrix: count pass, exclusive scan by the caller (b200rec_exclusive_scan_i32), fill pass.  Entry order inside
// a row = column order (what the GEMM's k-order would be).
__global__ void __launch_bounds__(256) dense_nnz_count_kernel(const float* __restrict__ X, long long ldx, long long M, int c0, int c1, int* __restrict__ cnt) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  int n = 0;
  for (int c = c0 + lane; c < c1; c += 32) n += __ldg(X + row * ldx + c) != 0.f;
  n = warp_sum_i(n);
  if (lane == 0) cnt[row] = n;
}

__global__ void __launch_bounds__(256) dense_nnz_fill_kernel(const float* __restrict__ X, long long ldx, long long M, int c0, int c1,
                                                             const int* __restrict__ row_ptr, int* __restrict__ col, float* __restrict__ val) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  int base = __ldg(row_ptr + row);
  for (int cb = c0; cb < c1; cb += 32) {
    const int c = cb + lane;
    const float v = c < c1 ? __ldg(X + row * ldx + c) : 0.f;
    const unsigned m = __ballot_sync(FULL, v != 0.f);
    if (v != 0.f) {
      const int o = base + __popc(m & ((1u << lane) - 1u));
      col[o] = c - c0;
      val[o] = v;
    }
    base += __popc(m);
  }
}

extern "C" int b200rec_dense_nnz_count(const float* X, int64_t M, int64_t ldx, int64_t c0, int64_t c1, int* counts, b200rec_stream_t stream) {
  if (M == 0) return B200REC_OK;
  if (!X || !counts || M < 0 || c0 < 0 || c1 < c0 || c1 > ldx) return b200rec_fail(B200REC_ERR_BAD_ARG, "dense_nnz_count: bad argument");
  dense_nnz_count_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(X, ldx, M, (int)c0, (int)c1, counts);
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}

extern "C" int b200rec_dense_nnz_fill(const float* X, int64_t M, int64_t ldx, int64_t c0, int64_t c1, const int* row_ptr, int* col, float* val,
                                      b200rec_stream_t stream) {
  if (M == 0) return B200REC_OK;
  if (!X || !row_ptr || !col || !val || M < 0 || c0 < 0 || c1 < c0 || c1 > ldx) return b200rec_fail(B200REC_ERR_BAD_ARG, "dense_nnz_fill: bad argument");
  dense_nnz_fill_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(X, ldx, M, (int)c0, (int)c1, row_ptr, col, val);
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}
