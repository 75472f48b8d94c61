// K1a — fp32 "parity mode" linear layer:  Y = act((X · Wᵀ + bias) ∘ row_scale)
//
// Replaces every nn.Linear on the reference's hot path (basic_ncf.py:38-39, attention_ncf.py:150-151,216,
// gnn_ncf.py:300-301 and — transform-before-gather — the per-edge Linear of gnn_ncf.py:91-93).
// CUDA-core FFMA with fp32 accumulation: the reference tolerance in fp32 mode is rel <= 1e-5, which rules
// out TF32/bf16 tensor-core products; the bf16-mode tcgen05 path lives in gemm_tc.cu.
//
// Both operands are K-contiguous (X row-major [M,K], W row-major [N,K] exactly as nn.Linear stores it).
// Tiling: BM x BN x 16 per CTA, 256 threads, TM x TN register tile, double-buffered shared memory, split-K
// (deterministic: partials to workspace, reduced in fixed order) when M*N alone cannot fill 148 SMs.
// The reference profile width F = 2094 is even but not a multiple of 4, so rows are only 8-byte aligned:
// the loader is templated on the vector width (4 / 2 / 1 floats) and picked per call from K, ld and the
// base pointers.
#include "common.cuh"

namespace b200rec {

constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;

template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void ldg_vec(float (&dst)[VEC], const float* p, bool ok) {
  if (!ok) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = 0.f;
    return;
  }
  if constexpr (VEC == 4) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  } else if constexpr (VEC == 2) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    dst[0] = v.x; dst[1] = v.y;
  } else {
    dst[0] = __ldg(p);
  }
}

struct Epilogue {
  const float* bias;       // [N] or null
  const float* row_scale;  // [M] or null: applied AFTER the bias (t = dinv * (xW^T + b))
  int relu;
};

template <typename OutT>
__device__ __forceinline__ void epilogue_store(OutT* Y, long long ldy, int row, int col0, const float (&v)[4], int N,
                                               const Epilogue& ep, bool vec_ok) {
  float o[4];
  float rs = ep.row_scale ? __ldg(ep.row_scale + row) : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = v[j];
    if (col0 + j < N) {
      if (ep.bias) x += __ldg(ep.bias + col0 + j);
      x *= rs;
      if (ep.relu) x = fmaxf(x, 0.f);
    }
    o[j] = x;
  }
  OutT* dst = Y + (long long)row * ldy + col0;
  if (vec_ok && col0 + 3 < N) {
    st4(dst, make_float4(o[0], o[1], o[2], o[3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j < N) st1(dst + j, o[j]);
  }
}

// RM/RN = number of 4-wide row/column groups per thread (TM = 4*RM, TN = 4*RN).
template <int BM, int BN, int RM, int RN, int VEC, typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_tn_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ W, long long ldw, int M, int N, int K,
               int k_per_split, OutT* __restrict__ Y, long long ldy, float* __restrict__ partial, Epilogue ep, int out_vec_ok) {
  static_assert((BM / (4 * RM)) * (BN / (4 * RN)) == GEMM_THREADS, "thread grid must be 256");
  constexpr int TX = BN / (4 * RN);           // threads along N
  constexpr int PAD = 4;
  constexpr int LA = BM * BK / VEC / GEMM_THREADS;   // vector loads per thread for the X tile
  constexpr int LB = BN * BK / VEC / GEMM_THREADS;
  static_assert(LA >= 1 && LB >= 1, "tile too small for this vector width");
  constexpr int KV = BK / VEC;                // vectors per tile row

  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);

  float acc[4 * RM][4 * RN];
#pragma unroll
  for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
    for (int j = 0; j < 4 * RN; ++j) acc[i][j] = 0.f;

  float ra[LA][VEC], rb[LB][VEC];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int v = tid + i * GEMM_THREADS;
      int r = v / KV, kk = (v % KV) * VEC;
      int gr = m0 + r, gk = k0 + kk;
      ldg_vec<VEC>(ra[i], X + (long long)gr * ldx + gk, gr < M && gk < k_end);
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int v = tid + i * GEMM_THREADS;
      int r = v / KV, kk = (v % KV) * VEC;
      int gr = n0 + r, gk = k0 + kk;
      ldg_vec<VEC>(rb[i], W + (long long)gr * ldw + gk, gr < N && gk < k_end);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int v = tid + i * GEMM_THREADS;
      int r = v / KV, kk = (v % KV) * VEC;
#pragma unroll
      for (int e = 0; e < VEC; ++e) As[buf][kk + e][r] = ra[i][e];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int v = tid + i * GEMM_THREADS;
      int r = v / KV, kk = (v % KV) * VEC;
#pragma unroll
      for (int e = 0; e < VEC; ++e) Bs[buf][kk + e][r] = rb[i][e];
    }
  };

  // NOTE: with VEC > 1 the split boundaries (k_begin, k_per_split) are multiples of BK, and K % VEC == 0, so a
  // vector is either fully inside [k_begin, k_end) or fully outside.
  int buf = 0;
  if (k_begin < k_end) {
    load_tiles(k_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool has_next = (k0 + BK) < k_end;
    if (has_next) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4 * RM], b[4 * RN];
#pragma unroll
      for (int g = 0; g < RM; ++g) {
        float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / RM) + ty * 4]);
        a[4 * g + 0] = t.x; a[4 * g + 1] = t.y; a[4 * g + 2] = t.z; a[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < RN; ++g) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / RN) + tx * 4]);
        b[4 * g + 0] = t.x; b[4 * g + 1] = t.y; b[4 * g + 2] = t.z; b[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4 * RM; ++i)
#pragma unroll
        for (int j = 0; j < 4 * RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue
#pragma unroll
  for (int gi = 0; gi < RM; ++gi)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = m0 + gi * (BM / RM) + ty * 4 + i;
      if (row >= M) continue;
#pragma unroll
      for (int gj = 0; gj < RN; ++gj) {
        int col0 = n0 + gj * (BN / RN) + tx * 4;
        if (col0 >= N) continue;
        float v[4] = {acc[4 * gi + i][4 * gj + 0], acc[4 * gi + i][4 * gj + 1], acc[4 * gi + i][4 * gj + 2],
                      acc[4 * gi + i][4 * gj + 3]};
        if (partial) {
          float* dst = partial + ((long long)blockIdx.z * M + row) * N + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col0 + j < N) dst[j] = v[j];
        } else {
          epilogue_store<OutT>(Y, ldy, row, col0, v, N, ep, out_vec_ok != 0);
        }
      }
    }
}

// Fixed-order reduction of the split-K partials + the epilogue.
template <typename OutT>
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, OutT* __restrict__ Y,
                                     long long ldy, Epilogue ep) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)M * N;
  if (idx >= total) return;
  int row = (int)(idx / N), col = (int)(idx % N);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + idx];
  if (ep.bias) s += __ldg(ep.bias + col);
  if (ep.row_scale) s *= __ldg(ep.row_scale + row);
  if (ep.relu) s = fmaxf(s, 0.f);
  st1(Y + (long long)row * ldy + col, s);
}

struct GemmPlan {
  int big;          // 1: 128x128 tile, 0: 64x64 tile
  int splits;
  int k_per_split;
};

static GemmPlan plan_gemm(long long M, long long N, long long K) {
  GemmPlan p;
  const int sms = b200rec_num_sms();
  long long big_tiles = ((M + 127) / 128) * ((N + 127) / 128);
  p.big = big_tiles >= sms ? 1 : 0;
  long long tiles = p.big ? big_tiles : ((M + 63) / 64) * ((N + 63) / 64);
  int splits = 1;
  if (tiles < 2LL * sms) {
    long long want = (2LL * sms + tiles - 1) / tiles;
    long long max_by_k = K / (4 * BK);                   // at least 4 k-tiles per split
    if (max_by_k < 1) max_by_k = 1;
    splits = (int)(want < max_by_k ? want : max_by_k);
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
  }
  long long kps = (K + splits - 1) / splits;
  kps = ((kps + BK - 1) / BK) * BK;
  p.k_per_split = (int)kps;
  p.splits = (int)((K + kps - 1) / kps);
  if (p.splits < 1) p.splits = 1;
  return p;
}

template <int VEC, typename OutT>
static int launch_gemm(const float* X, long long ldx, const float* W, long long ldw, int M, int N, int K, OutT* Y,
                       long long ldy, const Epilogue& ep, const GemmPlan& p, float* partial, cudaStream_t st) {
  const bool out_vec = ((uintptr_t)Y % 16 == 0) && (ldy % 4 == 0);
  if (p.big) {
    dim3 grid(ceil_div_i(N, 128), ceil_div_i(M, 128), p.splits);
    gemm_tn_kernel<128, 128, 2, 2, VEC, OutT><<<grid, GEMM_THREADS, 0, st>>>(X, ldx, W, ldw, M, N, K, p.k_per_split, Y, ldy,
                                                                             partial, ep, out_vec);
  } else {
    dim3 grid(ceil_div_i(N, 64), ceil_div_i(M, 64), p.splits);
    gemm_tn_kernel<64, 64, 1, 1, VEC, OutT><<<grid, GEMM_THREADS, 0, st>>>(X, ldx, W, ldw, M, N, K, p.k_per_split, Y, ldy,
                                                                           partial, ep, out_vec);
  }
  B200REC_CHECK_LAUNCH();
  if (partial) {
    long long total = (long long)M * N;
    splitk_reduce_kernel<OutT><<<ceil_div_i(total, 256), 256, 0, st>>>(partial, p.splits, M, N, Y, ldy, ep);
    B200REC_CHECK_LAUNCH();
  }
  return B200REC_OK;
}

}  // namespace b200rec

using namespace b200rec;

extern "C" size_t b200rec_linear_workspace(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmPlan p = plan_gemm(M, N, K);
  return p.splits > 1 ? (size_t)p.splits * (size_t)M * (size_t)N * sizeof(float) : 0;
}

extern "C" int b200rec_linear(const float* X, int64_t M, int64_t K, int64_t ldx, const float* W, int64_t N, int64_t ldw,
                              const float* bias, const float* row_scale, int relu, void* Y, int64_t ldy, int y_dtype,
                              void* workspace, size_t workspace_bytes, b200rec_stream_t stream) {
  if (M == 0) return B200REC_OK;                     // an empty batch is not an error (its buffers may be null)
  if (M < 0 || N <= 0 || K <= 0 || !W || !Y || !X) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear: bad argument");
  if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return b200rec_fail(B200REC_ERR_UNSUPPORTED, "linear: dim > int32");
  if (ldx < K || ldw < K || ldy < N) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear: leading dimension too small");
  if (y_dtype != B200REC_F32 && y_dtype != B200REC_BF16) return b200rec_fail(B200REC_ERR_BAD_ARG, "linear: bad y_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  GemmPlan p = plan_gemm(M, N, K);
  float* partial = nullptr;
  if (p.splits > 1) {
    size_t need = (size_t)p.splits * (size_t)M * (size_t)N * sizeof(float);
    if (!workspace || workspace_bytes < need) return b200rec_fail(B200REC_ERR_WORKSPACE, "linear: workspace too small");
    partial = (float*)workspace;
  }
  Epilogue ep{bias, row_scale, relu};
  int vec = 1;
  if (K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)W % 16 == 0) vec = 4;
  else if (K % 2 == 0 && ldx % 2 == 0 && ldw % 2 == 0 && (uintptr_t)X % 8 == 0 && (uintptr_t)W % 8 == 0) vec = 2;
#define B200REC_GEMM_DISPATCH(V)                                                                                        \
  (y_dtype == B200REC_F32                                                                                               \
       ? launch_gemm<V, float>(X, ldx, W, ldw, (int)M, (int)N, (int)K, (float*)Y, ldy, ep, p, partial, st)              \
       : launch_gemm<V, __nv_bfloat16>(X, ldx, W, ldw, (int)M, (int)N, (int)K, (__nv_bfloat16*)Y, ldy, ep, p, partial, st))
  if (vec == 4) return B200REC_GEMM_DISPATCH(4);
  if (vec == 2) return B200REC_GEMM_DISPATCH(2);
  return B200REC_GEMM_DISPATCH(1);
#undef B200REC_GEMM_DISPATCH
}
