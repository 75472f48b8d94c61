// K2 — AttentionNCF dynamic user profile: ragged item-item attention + segmented softmax + weighted pooling.
//
// Replaces attention_ncf.py:154-216 of the reference (the (B*I, E) repeat_interleave/repeat materialisation, the
// boolean-mask gather, AttentionNet on the valid pairs, the -inf filled (B, I) softmax, `scores * user_matrix`,
// the (B,I)x(I,F) pooling GEMM and the user-embedding Linear) by ONE memory-bound kernel over the non-zeros of
// `user_matrix`, using the exact algebraic factorisation:
//
//   AttentionNet.0 = [A1c | A1r]           P_c = Ec·A1cᵀ + a1   (B, H)      P_r = Er·A1rᵀ   (I, H)
//   s_bi = a2 · ReLU(P_c[b] + P_r[i]) + a20                      (score; MODE_NET)
//   Q    = R · W_Uᵀ                                              (I, U)   pooling in embedding space
//   user_emb[b] = Σ_i softmax_i(s_b·)·um_bi · Q[i] + b_U
//
// (cosine variant / att_dense=None variant: s_bi = <P_c[b], P_r[i]>, MODE_DOT, with the host passing
// normalised embeddings resp. 4-wide [sc,1,0,0]/[1,sr,0,0] rows.)
//
// One CTA owns one candidate row b and each of its 8 warps a contiguous slice of the row (round 1 used one warp per
// row: 4.6 % warps active, 560 us at B=512 — profiles/r01).  Per warp, non-zeros are consumed 32 at a time: for each of the 32 the warp loads the
// P_r row with one coalesced 128-bit load per lane, every lane forms its partial dot product, and a 31-shuffle
// butterfly reduce-scatter leaves lane j with the full score of non-zero j (1 shuffle per score instead of 5).
// The softmax is online (running max / denominator, FlashAttention-style rescale), so each segment is read once;
// pooling re-walks the 32 non-zeros with one coalesced Q-row load each; the per-warp (max, denominator, pooled vector)
// states are merged through shared memory at the end.  Nothing of size (B, I, *) exists.
// Two front-ends feed the same core: a CSR reader (ragged-native entry) and a dense-row scanner that compacts
// the reference's dense (B, I) `user_matrix` on the fly through a per-warp shared-memory queue (drop-in entry;
// exact 0.0 means "unrated", attention_ncf.py:158,192).
#include <math.h>

#include "common.cuh"

namespace b200rec {

constexpr int ATT_WARPS = 8;      // one CTA per candidate row; the row's non-zeros are split over the warps
enum { MODE_NET = 0, MODE_DOT = 1 };

struct AttParams {
  const float* Pc;     // (B, H) fp32, bias a1 folded in
  const void* Pr;      // (I, H) table
  const void* Q;       // (I, U) table
  const float* a2;     // (H) fp32 (MODE_NET)
  const float* a20;    // device scalar (MODE_NET) or null
  const float* bU;     // (U) or null
  float* out;          // (B, U)
  long long ldo;
  float* att;          // (B, I) or null; must be zero-filled by the caller
  int B, I, H, U;
  // training-mode extras (attention_ncf.py:185-203)
  const float* Ec;     // (B, E) or null: candidate embeddings for the isclose() target mask
  const float* Er;     // (I, E)
  int E;
  float atol, rtol;
  int drop_zero_scores;   // message_dropout is not None: scores == 0.0 -> -inf (:189)
  float score_scale;      // message dropout: kept scores are scaled by 1/(1-p) (:187)
  long long ldPr, ldQ;
};

template <int HV, int UV, int MODE, typename T>
struct RowCore {
  const AttParams& p;
  const int lane;
  const int b;
  float pc[HV][4], a2[HV][4];
  float acc[UV][4];
  float m, l, a20;

  __device__ RowCore(const AttParams& p_, int lane_, int b_) : p(p_), lane(lane_), b(b_) {
#pragma unroll
    for (int hv = 0; hv < HV; ++hv) {
      const int h = lane * 4 + hv * 128;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f), w = c;
      if (h < p.H) {
        c = ld4(p.Pc + (long long)b * p.H + h);
        if (MODE == MODE_NET) w = ld4(p.a2 + h);
      }
      pc[hv][0] = c.x; pc[hv][1] = c.y; pc[hv][2] = c.z; pc[hv][3] = c.w;
      a2[hv][0] = w.x; a2[hv][1] = w.y; a2[hv][2] = w.z; a2[hv][3] = w.w;
    }
#pragma unroll
    for (int uv = 0; uv < UV; ++uv)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[uv][e] = 0.f;
    m = -INFINITY;
    l = 0.f;
    a20 = (MODE == MODE_NET && p.a20) ? __ldg(p.a20) : 0.f;
  }

  // lane j holds non-zero j of this batch (col < 0 beyond `count`)
  __device__ void batch(int my_col, float my_val, int count) {
    const T* __restrict__ Pr = reinterpret_cast<const T*>(p.Pr);
    const T* __restrict__ Q = reinterpret_cast<const T*>(p.Q);
    float v[32];
#pragma unroll
    constexpr int LB = (HV == 1 && UV == 1) ? 16 : 8;     // row gathers in flight per warp
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += LB) {
      if (j0 < count) {
        float4 pr[LB][HV];
#pragma unroll
        for (int jj = 0; jj < LB; ++jj) {
          const int c = __shfl_sync(FULL, my_col, j0 + jj);
#pragma unroll
          for (int hv = 0; hv < HV; ++hv) {
            const int h = lane * 4 + hv * 128;
            pr[jj][hv] = (c >= 0 && h < p.H) ? ld4(Pr + (long long)c * p.ldPr + h) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int jj = 0; jj < LB; ++jj) {
          float s = 0.f;
#pragma unroll
          for (int hv = 0; hv < HV; ++hv) {
            const float r[4] = {pr[jj][hv].x, pr[jj][hv].y, pr[jj][hv].z, pr[jj][hv].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (MODE == MODE_NET) s = fmaf(a2[hv][e], fmaxf(pc[hv][e] + r[e], 0.f), s);
              else s = fmaf(pc[hv][e], r[e], s);
            }
          }
          v[j0 + jj] = s;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < LB; ++jj) v[j0 + jj] = 0.f;
      }
    }
    // butterfly reduce-scatter: after the last step lane j holds sum over lanes of v[j]
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
      const bool up = (lane & k) != 0;
#pragma unroll
      for (int i = 0; i < k; ++i) {
        const float send = up ? v[i] : v[i + k];
        const float keep = up ? v[i + k] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, k);
      }
    }
    float s = (v[0] + a20) * p.score_scale;
    if (my_col < 0) s = -INFINITY;                    // beyond `count`, or an explicit zero in a CSR
    if (p.drop_zero_scores && s == 0.f) s = -INFINITY;
    if (p.Ec != nullptr) {     // training: mask rated items whose EMBEDDING is close to the candidate's (:199)
      unsigned masked = 0u;
      for (int j = 0; j < count; ++j) {
        const int c = __shfl_sync(FULL, my_col, j);
        if (c < 0) continue;
        bool ok = true;
        for (int e = lane; e < p.E; e += 32) {
          const float x = __ldg(p.Ec + (long long)b * p.E + e), y = __ldg(p.Er + (long long)c * p.E + e);
          ok = ok && (fabsf(x - y) <= p.atol + p.rtol * fabsf(y));
        }
        if (__all_sync(FULL, ok)) masked |= (1u << j);
      }
      if ((masked >> lane) & 1u) s = -INFINITY;
    }
    if (p.att != nullptr && my_col >= 0) p.att[(long long)b * p.I + my_col] = s;    // raw score; normalised in finish()

    const float m_new = fmaxf(m, warp_max(s));
    float pj = 0.f;
    if (m_new != -INFINITY) {
      const float scale = __expf(m - m_new);          // m == -inf -> 0
      pj = (s == -INFINITY) ? 0.f : __expf(s - m_new);
      l = l * scale + warp_sum(pj);
#pragma unroll
      for (int uv = 0; uv < UV; ++uv)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[uv][e] *= scale;
      m = m_new;
    }
    const float wgt = pj * my_val;                    // attention x centred rating (:212)
#pragma unroll
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += LB) {
      if (j0 < count) {
        float4 q[LB][UV];
        float w[LB];
#pragma unroll
        for (int jj = 0; jj < LB; ++jj) {
          const int c = __shfl_sync(FULL, my_col, j0 + jj);
          w[jj] = __shfl_sync(FULL, wgt, j0 + jj);
#pragma unroll
          for (int uv = 0; uv < UV; ++uv) {
            const int u = lane * 4 + uv * 128;
            q[jj][uv] = (c >= 0 && u < p.U) ? ld4(Q + (long long)c * p.ldQ + u) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int jj = 0; jj < LB; ++jj)
#pragma unroll
          for (int uv = 0; uv < UV; ++uv) {
            acc[uv][0] = fmaf(w[jj], q[jj][uv].x, acc[uv][0]);
            acc[uv][1] = fmaf(w[jj], q[jj][uv].y, acc[uv][1]);
            acc[uv][2] = fmaf(w[jj], q[jj][uv].z, acc[uv][2]);
            acc[uv][3] = fmaf(w[jj], q[jj][uv].w, acc[uv][3]);
          }
      }
    }
  }

  // ---- cross-warp merge (one CTA = one candidate row, every warp owns a slice of the row's non-zeros) ----------
  __device__ void export_state(float* s_m, float* s_l, float* s_acc, int warp) const {
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
    for (int uv = 0; uv < UV; ++uv)
      *reinterpret_cast<float4*>(s_acc + (size_t)warp * (UV * 128) + lane * 4 + uv * 128) =
          make_float4(acc[uv][0], acc[uv][1], acc[uv][2], acc[uv][3]);
  }
};

// after __syncthreads(): global running max / denominator of the row and the pooled output
template <int UV>
__device__ __forceinline__ void merge_and_write(const AttParams& p, int b, const float* s_m, const float* s_l, const float* s_acc,
                                                float& M, float& Lsum) {
  M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) M = fmaxf(M, s_m[w]);
  float scale[ATT_WARPS];
  Lsum = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) {
    scale[w] = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - M);
    Lsum += s_l[w] * scale[w];
  }
  const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;      // no valid rated item -> weights 0 -> user_emb = b_U (:208-209)
  for (int u = threadIdx.x; u < p.U; u += ATT_WARPS * 32) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) a = fmaf(s_acc[(size_t)w * (UV * 128) + u], scale[w], a);
    p.out[(long long)b * p.ldo + u] = fmaf(a, inv, p.bU ? __ldg(p.bU + u) : 0.f);
  }
}

__device__ __forceinline__ float normalise_score(float s, float M, float Lsum) {
  return (Lsum > 0.f && s != -INFINITY) ? __expf(s - M) / Lsum : 0.f;
}

// ---- ragged-native front-end: CSR of user_matrix ------------------------------------------------------------
template <int HV, int UV, int MODE, typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32, (HV == 1 && UV == 1) ? 2 : 1)
attention_pool_csr_kernel(AttParams p, const int* __restrict__ row_ptr, const int* __restrict__ col,
                          const float* __restrict__ val, const int* __restrict__ row_nnz, long long padded_stride) {
  __shared__ float s_m[ATT_WARPS], s_l[ATT_WARPS];
  __shared__ __align__(16) float s_acc[ATT_WARPS * UV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  RowCore<HV, UV, MODE, T> core(p, lane, b);
  // either a true CSR (row_ptr) or the row-padded form written by um_compact_kernel (row b starts at b*stride)
  const long long start = row_nnz ? (long long)b * padded_stride : (long long)__ldg(row_ptr + b);
  const long long end = row_nnz ? start + __ldg(row_nnz + b) : (long long)__ldg(row_ptr + b + 1);
  const int blocks = (int)((end - start + 31) >> 5);
  const int bpw = (blocks + ATT_WARPS - 1) / ATT_WARPS;
  const long long my_start = start + (long long)warp * bpw * 32, my_end = min(end, my_start + (long long)bpw * 32);
  for (long long k0 = my_start; k0 < my_end; k0 += 32) {
    const long long k = k0 + lane;
    int c = -1;
    float v = 0.f;
    if (k < my_end) { c = __ldg(col + k); v = __ldg(val + k); }
    const bool valid = (k < my_end) && (v != 0.f);     // an explicit 0.0 in the CSR is "unrated", like the dense form
    core.batch(valid ? c : -1, valid ? v : 0.f, (int)min(32LL, my_end - k0));
  }
  core.export_state(s_m, s_l, s_acc, warp);
  __syncthreads();
  float M, Lsum;
  merge_and_write<UV>(p, b, s_m, s_l, s_acc, M, Lsum);
  if (p.att != nullptr) {
    for (long long k = my_start + lane; k < my_end; k += 32) {
      if (__ldg(val + k) != 0.f) {
        float* a = p.att + (long long)b * p.I + __ldg(col + k);
        *a = normalise_score(*a, M, Lsum);
      }
    }
  }
}

// ---- segment-parallel ragged kernel -------------------------------------------------------------------------
// Rows are heavy-tailed (mean ~550 non-zeros per candidate at config 2, max ~2.7k): with one CTA per row the kernel
// lasts as long as its longest row.  Here a CTA owns ONE SEGMENT of at most ATT_SEG non-zeros of one row (grid =
// B x ceil(I / ATT_SEG), CTAs beyond a row's length exit at once), writes its (max, denominator, pooled vector) to a
// partial slot, and a small second kernel merges a row's partials, adds b_U and normalises the attention weights.
constexpr int ATT_SEG = 512;

template <int HV, int UV, int MODE, typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32, (HV == 1 && UV == 1) ? 2 : 1)
attention_seg_kernel(AttParams p, const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val,
                     const int* __restrict__ row_nnz, long long padded_stride, float* __restrict__ partials, int nseg_max) {
  __shared__ float s_m[ATT_WARPS], s_l[ATT_WARPS];
  __shared__ __align__(16) float s_acc[ATT_WARPS * UV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x, seg = blockIdx.y;
  const long long start = row_nnz ? (long long)b * padded_stride : (long long)__ldg(row_ptr + b);
  const long long end = row_nnz ? start + __ldg(row_nnz + b) : (long long)__ldg(row_ptr + b + 1);
  const long long seg_start = start + (long long)seg * ATT_SEG;
  if (seg_start >= end) return;                                    // (whole CTA) this row has fewer segments
  const long long seg_end = min(end, seg_start + ATT_SEG);
  RowCore<HV, UV, MODE, T> core(p, lane, b);
  const int blocks = (int)((seg_end - seg_start + 31) >> 5);
  const int bpw = (blocks + ATT_WARPS - 1) / ATT_WARPS;
  const long long my_start = seg_start + (long long)warp * bpw * 32, my_end = min(seg_end, my_start + (long long)bpw * 32);
  for (long long k0 = my_start; k0 < my_end; k0 += 32) {
    const long long k = k0 + lane;
    int c = -1;
    float v = 0.f;
    if (k < my_end) { c = __ldg(col + k); v = __ldg(val + k); }
    const bool valid = (k < my_end) && (v != 0.f);
    core.batch(valid ? c : -1, valid ? v : 0.f, (int)min(32LL, my_end - k0));
  }
  core.export_state(s_m, s_l, s_acc, warp);
  __syncthreads();
  // merge the 8 warps -> one partial (m, l, acc[U]) for this segment
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) M = fmaxf(M, s_m[w]);
  float scale[ATT_WARPS], Lsum = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) {
    scale[w] = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - M);
    Lsum += s_l[w] * scale[w];
  }
  float* slot = partials + ((long long)b * nseg_max + seg) * (p.U + 2);
  if (threadIdx.x == 0) { slot[0] = M; slot[1] = Lsum; }
  for (int u = threadIdx.x; u < p.U; u += ATT_WARPS * 32) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) a = fmaf(s_acc[(size_t)w * (UV * 128) + u], scale[w], a);
    slot[2 + u] = a;
  }
}

__global__ void __launch_bounds__(128)
attention_merge_kernel(AttParams p, const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val,
                       const int* __restrict__ row_nnz, long long padded_stride, const float* __restrict__ partials, int nseg_max) {
  const int b = blockIdx.x;
  const long long start = row_nnz ? (long long)b * padded_stride : (long long)__ldg(row_ptr + b);
  const long long end = row_nnz ? start + __ldg(row_nnz + b) : (long long)__ldg(row_ptr + b + 1);
  const int nseg = (int)((end - start + ATT_SEG - 1) / ATT_SEG);
  const float* base = partials + (long long)b * nseg_max * (p.U + 2);
  float M = -INFINITY;
  for (int sgi = 0; sgi < nseg; ++sgi) M = fmaxf(M, base[(long long)sgi * (p.U + 2)]);
  float Lsum = 0.f;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const float ms = base[(long long)sgi * (p.U + 2)];
    Lsum += base[(long long)sgi * (p.U + 2) + 1] * ((ms == -INFINITY) ? 0.f : __expf(ms - M));
  }
  const float inv = Lsum > 0.f ? 1.f / Lsum : 0.f;      // no valid rated item -> weights 0 -> user_emb = b_U (:208-209)
  for (int u = threadIdx.x; u < p.U; u += blockDim.x) {
    float a = 0.f;
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const float ms = base[(long long)sgi * (p.U + 2)];
      a = fmaf(base[(long long)sgi * (p.U + 2) + 2 + u], (ms == -INFINITY) ? 0.f : __expf(ms - M), a);
    }
    p.out[(long long)b * p.ldo + u] = fmaf(a, inv, p.bU ? __ldg(p.bU + u) : 0.f);
  }
  if (p.att != nullptr) {
    for (long long k = start + threadIdx.x; k < end; k += blockDim.x) {
      if (__ldg(val + k) != 0.f) {
        float* a = p.att + (long long)b * p.I + __ldg(col + k);
        *a = normalise_score(*a, M, Lsum);
      }
    }
  }
}

// ---- dense (B, I) user_matrix -> row-padded (col, val) lists + per-row counts ---------------------------------
// One CTA per row; pure streaming (four independent coalesced loads per thread and iteration, no dependent work), so
// the dense scan costs a few microseconds instead of sitting on the critical path of the pooling kernel, and the
// pooling kernel can split every row's non-zeros EVENLY over its warps.
__global__ void __launch_bounds__(ATT_WARPS * 32)
um_compact_kernel(const float* __restrict__ um, long long ld_um, int I, int* __restrict__ col, float* __restrict__ val,
                  int* __restrict__ row_nnz) {
  __shared__ int s_cnt[ATT_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const float* __restrict__ row = um + (long long)b * ld_um;
  const int chunks = (I + 31) >> 5;
  const int cpw = (chunks + ATT_WARPS - 1) / ATT_WARPS;
  const int i_begin = warp * cpw * 32, i_end = min(I, i_begin + cpw * 32);
  int cnt = 0;
  for (int i0 = i_begin; i0 < i_end; i0 += 128) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 32 + lane;
      v[u] = (i < i_end) ? __ldg(row + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cnt += __popc(__ballot_sync(FULL, v[u] != 0.f));
  }
  if (lane == 0) s_cnt[warp] = cnt;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; ++w) {
    if (w < warp) base += s_cnt[w];
    total += s_cnt[w];
  }
  if (threadIdx.x == 0) row_nnz[b] = total;
  long long out = (long long)b * I + base;
  for (int i0 = i_begin; i0 < i_end; i0 += 128) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 32 + lane;
      v[u] = (i < i_end) ? __ldg(row + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool nz = v[u] != 0.f;
      const unsigned mask = __ballot_sync(FULL, nz);
      if (nz) {
        const long long pos = out + __popc(mask & ((1u << lane) - 1u));
        col[pos] = i0 + u * 32 + lane;
        val[pos] = v[u];
      }
      out += __popc(mask);
    }
  }
}

// ---- drop-in front-end: the reference's dense (B, I) user_matrix --------------------------------------------
template <int HV, int UV, int MODE, typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_pool_dense_kernel(AttParams p, const float* __restrict__ um, long long ld_um) {
  __shared__ int qcol[ATT_WARPS][64];
  __shared__ float qval[ATT_WARPS][64];
  __shared__ float s_m[ATT_WARPS], s_l[ATT_WARPS];
  __shared__ __align__(16) float s_acc[ATT_WARPS * UV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  RowCore<HV, UV, MODE, T> core(p, lane, b);
  const float* __restrict__ row = um + (long long)b * ld_um;
  // every warp scans a contiguous slice of the row (multiples of 32 columns) and compacts its non-zeros on the fly
  const int chunks = (p.I + 31) >> 5;
  const int cpw = (chunks + ATT_WARPS - 1) / ATT_WARPS;
  const int i_begin = warp * cpw * 32, i_end = min(p.I, i_begin + cpw * 32);
  int qn = 0;
  for (int i0 = i_begin; i0 < i_end; i0 += 32) {
    const int i = i0 + lane;
    const float v = (i < i_end) ? __ldg(row + i) : 0.f;
    const bool nz = v != 0.f;
    const unsigned mask = __ballot_sync(FULL, nz);
    if (nz) {
      const int pos = qn + __popc(mask & ((1u << lane) - 1u));
      qcol[warp][pos] = i;
      qval[warp][pos] = v;
    }
    qn += __popc(mask);
    __syncwarp();
    if (qn >= 32) {
      core.batch(qcol[warp][lane], qval[warp][lane], 32);
      const int rem = qn - 32;
      int tc = 0; float tv = 0.f;
      if (lane < rem) { tc = qcol[warp][32 + lane]; tv = qval[warp][32 + lane]; }
      __syncwarp();
      if (lane < rem) { qcol[warp][lane] = tc; qval[warp][lane] = tv; }
      qn = rem;
      __syncwarp();
    }
  }
  if (qn > 0) core.batch(lane < qn ? qcol[warp][lane] : -1, lane < qn ? qval[warp][lane] : 0.f, qn);
  core.export_state(s_m, s_l, s_acc, warp);
  __syncthreads();
  float M, Lsum;
  merge_and_write<UV>(p, b, s_m, s_l, s_acc, M, Lsum);
  if (p.att != nullptr) {
    __syncwarp();
    for (int i = i_begin + lane; i < i_end; i += 32) {
      if (__ldg(row + i) != 0.f) {
        float* a = p.att + (long long)b * p.I + i;
        *a = normalise_score(*a, M, Lsum);
      }
    }
  }
}

struct AttInputs {
  const float* um; long long ld_um;
  const int* row_ptr; const int* col; const float* val;
  void* ws; size_t ws_bytes;
};

static size_t att_partials_bytes(long long B, long long I, int U) {
  const long long nseg = (I + ATT_SEG - 1) / ATT_SEG;
  return (size_t)(B * (nseg > 0 ? nseg : 1)) * (size_t)(U + 2) * sizeof(float);
}
static size_t att_compact_bytes(long long B, long long I) {
  return (size_t)(B * I) * (sizeof(int) + sizeof(float)) + (size_t)B * sizeof(int) + 64;
}
// dense form: compaction lists + partial slots; CSR form: partial slots only
static size_t att_workspace_bytes(long long B, long long I, int U, bool dense) {
  return att_partials_bytes(B, I, U) + 256 + (dense ? att_compact_bytes(B, I) : 0);
}

template <int HV, int UV, int MODE, typename T>
static int launch_att(const AttParams& p, const AttInputs& in, cudaStream_t st) {
  const bool dense = in.um != nullptr;
  const bool have_ws = in.ws && in.ws_bytes >= att_workspace_bytes(p.B, p.I, p.U, dense);
  if (!have_ws) {                      // no workspace: single fused kernel, one CTA per row
    if (dense) attention_pool_dense_kernel<HV, UV, MODE, T><<<p.B, ATT_WARPS * 32, 0, st>>>(p, in.um, in.ld_um);
    else attention_pool_csr_kernel<HV, UV, MODE, T><<<p.B, ATT_WARPS * 32, 0, st>>>(p, in.row_ptr, in.col, in.val, nullptr, 0);
    B200REC_CHECK_LAUNCH();
    return B200REC_OK;
  }
  float* partials = reinterpret_cast<float*>(in.ws);
  const int* ccol = in.col; const float* cval = in.val; const int* cnnz = nullptr; const int* rp = in.row_ptr;
  long long stride = 0;
  if (dense) {                         // streaming compaction of the dense matrix into row-padded lists
    unsigned char* q = reinterpret_cast<unsigned char*>(in.ws) + ((att_partials_bytes(p.B, p.I, p.U) + 255) & ~(size_t)255);
    int* wcol = reinterpret_cast<int*>(q);
    float* wval = reinterpret_cast<float*>(wcol + (size_t)p.B * p.I);
    int* wnnz = reinterpret_cast<int*>(wval + (size_t)p.B * p.I);
    um_compact_kernel<<<p.B, ATT_WARPS * 32, 0, st>>>(in.um, in.ld_um, p.I, wcol, wval, wnnz);
    B200REC_CHECK_LAUNCH();
    ccol = wcol; cval = wval; cnnz = wnnz; rp = nullptr; stride = p.I;
  }
  const int nseg_max = p.I > 0 ? (p.I + ATT_SEG - 1) / ATT_SEG : 1;
  dim3 grid(p.B, nseg_max);
  if (nseg_max > 65535) return b200rec_fail(B200REC_ERR_UNSUPPORTED, "attention_pool: more than 65535 segments per row");
  attention_seg_kernel<HV, UV, MODE, T><<<grid, ATT_WARPS * 32, 0, st>>>(p, rp, ccol, cval, cnnz, stride, partials, nseg_max);
  B200REC_CHECK_LAUNCH();
  attention_merge_kernel<<<p.B, 128, 0, st>>>(p, rp, ccol, cval, cnnz, stride, partials, nseg_max);
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}

template <int MODE, typename T>
static int dispatch_att(const AttParams& p, const AttInputs& in, cudaStream_t st) {
  const int w = p.H > p.U ? p.H : p.U;
  if (w <= 128) return launch_att<1, 1, MODE, T>(p, in, st);
  if (w <= 256) return launch_att<2, 2, MODE, T>(p, in, st);
  if (w <= 512) return launch_att<4, 4, MODE, T>(p, in, st);
  return b200rec_fail(B200REC_ERR_UNSUPPORTED, "attention_pool: att_dense / user_emb wider than 512");
}

}  // namespace b200rec

using namespace b200rec;

extern "C" size_t b200rec_attention_pool_workspace(int64_t B, int64_t I, int U, int dense) {
  return (B > 0 && I > 0 && U > 0) ? att_workspace_bytes(B, I, U, dense != 0) : 0;
}

extern "C" int b200rec_attention_pool(const b200rec_attention_t* a, b200rec_stream_t stream) {
  if (!a) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: null descriptor");
  if (a->B < 0 || a->I < 0 || a->H <= 0 || a->U <= 0 || (a->H % 4) || (a->U % 4))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: H and U must be positive multiples of 4");
  if (a->B == 0) return B200REC_OK;
  if (!a->Pc || !a->out || (a->I > 0 && (!a->Pr || !a->Q))) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: null table");
  if (!a->user_matrix && !(a->row_ptr && (a->col || a->I == 0)))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: need user_matrix or CSR");
  if (a->mode == MODE_NET && !a->a2) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: a2 missing");
  if (a->ldo < a->U) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: ldo too small");
  if ((uintptr_t)a->Pc % 16 || (uintptr_t)a->Pr % 8 || (uintptr_t)a->Q % 8 || (uintptr_t)a->out % 16 || (a->ldo % 4) ||
      (a->table_dtype == B200REC_F32 && ((uintptr_t)a->Pr % 16 || (uintptr_t)a->Q % 16)) || (a->bU && (uintptr_t)a->bU % 16) ||
      (a->a2 && (uintptr_t)a->a2 % 16))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: pointers must be 16-byte aligned");
  AttParams p;
  p.Pc = a->Pc; p.Pr = a->Pr; p.Q = a->Q; p.a2 = a->a2; p.a20 = a->a20; p.bU = a->bU;
  p.out = a->out; p.ldo = a->ldo; p.att = a->att_weights;
  p.B = (int)a->B; p.I = (int)a->I; p.H = a->H; p.U = a->U;
  p.Ec = a->train_cand_emb; p.Er = a->train_rated_emb; p.E = a->E; p.atol = a->atol; p.rtol = a->rtol;
  p.drop_zero_scores = a->drop_zero_scores;
  p.score_scale = a->score_scale == 0.f ? 1.f : a->score_scale;
  p.ldPr = a->ld_pr ? a->ld_pr : a->H;
  p.ldQ = a->ld_q ? a->ld_q : a->U;
  if (p.ldPr < a->H || p.ldQ < a->U || (p.ldPr % 4) || (p.ldQ % 4)) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: bad table leading dimension");
  if (p.Ec && (!p.Er || p.E <= 0)) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: training mask needs both embeddings");
  cudaStream_t st = (cudaStream_t)stream;
  const long long ld_um = a->user_matrix ? (a->ld_user_matrix ? a->ld_user_matrix : a->I) : 0;
  AttInputs in{a->user_matrix, ld_um, a->row_ptr, a->col, a->val, a->workspace, a->workspace_bytes};
  if (a->table_dtype == B200REC_F32) {
    if (a->mode == MODE_NET) return dispatch_att<MODE_NET, float>(p, in, st);
    if (a->mode == MODE_DOT) return dispatch_att<MODE_DOT, float>(p, in, st);
  } else if (a->table_dtype == B200REC_BF16) {
    if (a->mode == MODE_NET) return dispatch_att<MODE_NET, __nv_bfloat16>(p, in, st);
    if (a->mode == MODE_DOT) return dispatch_att<MODE_DOT, __nv_bfloat16>(p, in, st);
  }
  return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool: bad mode / table_dtype");
}
