// K2 backward — gradient of AttentionNCF's attention pooling (training step, reference NCF/train.py:99-105 through
// attention_ncf.py:154-216) for the factorised forward of csrc/attention_pool.cu:
//
//   s_bi  = score_scale · (a2·ReLU(Pc[b] + Pr[i]) + a20)      (MODE_NET)      or   score_scale · <Pc[b], Pr[i]>   (MODE_DOT)
//   α_b·  = softmax over the unmasked non-zeros of user_matrix[b]            (emitted by the forward kernel as `att`)
//   out[b] = Σ_i α_bi·um_bi·Q[i] + bU
//
// With g = dL/dout and pooled[b] = out[b] − bU:
//   dQ[i]   += α_bi·um_bi · g[b]
//   ds_bi    = score_scale · α_bi · (um_bi·<g[b], Q[i]> − <g[b], pooled[b]>)        (softmax backward; the row sum Σ_j α_bj·dα_bj
//                                                                                    is <g[b], pooled[b]> — no second pass)
//   NET:  t = ds_bi · a2 ∘ 1[Pc[b] + Pr[i] > 0];  dPc[b] += t;  dPr[i] += t;  da2 += ds_bi·ReLU(Pc[b] + Pr[i]);  da20 += ds_bi
//   DOT:  dPc[b] += ds_bi·Pr[i];  dPr[i] += ds_bi·Pc[b]
//
// Everything the forward masked out (unrated items, the isclose() target mask :199, zero scores under message dropout :189) has
// α = 0 and contributes nothing, so the masks are not re-evaluated here.  One CTA per (candidate row, column slice) — `n_slices` deals
// a row's non-zeros out over several CTAs, because a CTA-per-row grid ran 43 % busy behind its 2.7 k-non-zero rows (ncu, profiles/r01);
// the per-row outputs then come as one part per slice and the caller adds the parts.  The CTA's 8 warps scan their slice of the row's
// α (coalesced, 19 MB per 512 x 9.4k batch), and every non-zero is handled by a whole warp — one 128-bit load of the Q row and of
// the Pr row per lane, a 5-shuffle dot product, register accumulation of the per-row gradients (dPc, da2, da20: summed over the
// warps in warp order, bit-reproducible) and 128-bit vector reductions (`red.global.add.v4.f32`) into the per-item gradients
// dPr / dQ, which many rows share (their summation order is the hardware's, as with torch's index_add_ in the reference).
#include "common.cuh"

namespace b200rec {

constexpr int BWD_WARPS = 8;
enum { BWD_NET = 0, BWD_DOT = 1 };

struct AttBwdParams {
  const float *Pc, *Pr, *Q, *a2, *um, *att, *out, *bU, *g;
  long long ldPc, ldPr, ldQ, ld_um, ldo, ldg;
  int B, I, H, U;
  int n_slices, slice_cols;  // grid.y and the columns per slice (multiple of 32); a row's non-zeros are dealt out over gridDim.y CTAs
  float score_scale;
  float *dPc, *dPr, *dQ, *da2_rows, *da20_rows;     // dPc (B,H) and the per-row parts are written; dPr (I,H), dQ (I,U) are added to
  unsigned drop_key, drop_thr16;                    // AttentionNet's inner dropout: the forward's Philox mask is regenerated (65536 = off)
  float drop_scale;
  const unsigned long long* drop_seed_dev;          // the seed in device memory (graph replays), or nullptr
};

__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int HV, int UV, int MODE>
struct BwdRow {
  const AttBwdParams& p;
  const int lane;
  const int brow;
  float pc[HV][4], a2[HV][4], gb[UV][4];
  float dpc[HV][4], da2[HV][4];
  float da20, gdot;

  __device__ BwdRow(const AttBwdParams& p_, int lane_, int b) : p(p_), lane(lane_), brow(b) {
    da20 = 0.f;
#pragma unroll
    for (int hv = 0; hv < HV; ++hv) {
      const int h = lane * 4 + hv * 128;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f), w = c;
      if (h < p.H) {
        c = ld4(p.Pc + (long long)b * p.ldPc + h);
        if (MODE == BWD_NET) w = ld4(p.a2 + h);
      }
      pc[hv][0] = c.x; pc[hv][1] = c.y; pc[hv][2] = c.z; pc[hv][3] = c.w;
      a2[hv][0] = w.x; a2[hv][1] = w.y; a2[hv][2] = w.z; a2[hv][3] = w.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) dpc[hv][q] = da2[hv][q] = 0.f;
    }
    float part = 0.f;
#pragma unroll
    for (int uv = 0; uv < UV; ++uv) {
      const int u = lane * 4 + uv * 128;
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f), o = gg, bu = gg;
      if (u < p.U) {
        gg = ld4(p.g + (long long)b * p.ldg + u);
        o = ld4(p.out + (long long)b * p.ldo + u);
        if (p.bU) bu = ld4(p.bU + u);
      }
      gb[uv][0] = gg.x; gb[uv][1] = gg.y; gb[uv][2] = gg.z; gb[uv][3] = gg.w;
      part += gg.x * (o.x - bu.x) + gg.y * (o.y - bu.y) + gg.z * (o.z - bu.z) + gg.w * (o.w - bu.w);
    }
    gdot = warp_sum(part);
  }

  // two non-zeros of the row at once (their four row loads are independent); `live1` = the second one exists
  __device__ __forceinline__ void pair(int i0, float alpha0, float um0, int i1, float alpha1, float um1, bool live1) {
    const int ii[2] = {i0, i1};
    const float al[2] = {alpha0, alpha1}, uu[2] = {um0, um1};
    float4 q[2][UV], pr[2][HV];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int uv = 0; uv < UV; ++uv) {
        const int u = lane * 4 + uv * 128;
        q[k][uv] = (u < p.U) ? ld4(p.Q + (long long)ii[k] * p.ldQ + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int hv = 0; hv < HV; ++hv) {
        const int h = lane * 4 + hv * 128;
        pr[k][hv] = (h < p.H) ? ld4(p.Pr + (long long)ii[k] * p.ldPr + h) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !live1) break;                                   // warp-uniform
      float part = 0.f;
#pragma unroll
      for (int uv = 0; uv < UV; ++uv)
        part += gb[uv][0] * q[k][uv].x + gb[uv][1] * q[k][uv].y + gb[uv][2] * q[k][uv].z + gb[uv][3] * q[k][uv].w;
      const float dot = warp_sum(part);
      const float ds = p.score_scale * al[k] * (uu[k] * dot - gdot);
      const float cq = al[k] * uu[k];
#pragma unroll
      for (int uv = 0; uv < UV; ++uv) {
        const int u = lane * 4 + uv * 128;
        if (u < p.U) red_add4(p.dQ + (long long)ii[k] * p.U + u, make_float4(cq * gb[uv][0], cq * gb[uv][1], cq * gb[uv][2], cq * gb[uv][3]));
      }
#pragma unroll
      for (int hv = 0; hv < HV; ++hv) {
        const int h = lane * 4 + hv * 128;
        const float r[4] = {pr[k][hv].x, pr[k][hv].y, pr[k][hv].z, pr[k][hv].w};
        float t[4];
        float dm[4] = {1.f, 1.f, 1.f, 1.f};                         // hidden = ReLU(z) ∘ dm: the forward's dropout multipliers of this pair
        if (MODE == BWD_NET && p.drop_thr16 < 65536u) {
          unsigned dkey = p.drop_key;
          if (p.drop_seed_dev != nullptr) {
            const unsigned long long sd = __ldg(p.drop_seed_dev);
            dkey = (unsigned)(sd ^ (sd >> 32));
          }
          att_dropout_mult(dkey, p.drop_thr16, p.drop_scale, brow, ii[k], lane + 32 * hv, dm);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (MODE == BWD_NET) {
            const float z = pc[hv][e] + r[e];
            t[e] = z > 0.f ? ds * a2[hv][e] * dm[e] : 0.f;
            dpc[hv][e] += t[e];
            da2[hv][e] = fmaf(ds * dm[e], fmaxf(z, 0.f), da2[hv][e]);
          } else {
            dpc[hv][e] = fmaf(ds, r[e], dpc[hv][e]);
            t[e] = ds * pc[hv][e];
          }
        }
        if (h < p.H) red_add4(p.dPr + (long long)ii[k] * p.H + h, make_float4(t[0], t[1], t[2], t[3]));
      }
      da20 += ds;
    }
  }
};

template <int HV, int UV, int MODE>
__global__ void __launch_bounds__(BWD_WARPS * 32) attention_pool_bwd_kernel(AttBwdParams p) {
  __shared__ float s_dpc[BWD_WARPS][HV * 128];
  __shared__ float s_da2[MODE == BWD_NET ? BWD_WARPS : 1][MODE == BWD_NET ? HV * 128 : 1];
  __shared__ float s_da20[BWD_WARPS];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BwdRow<HV, UV, MODE> row(p, lane, b);
  const float* arow = p.att + (long long)b * p.I;
  const float* urow = p.um + (long long)b * p.ld_um;
  const int c_begin = (int)blockIdx.y * p.slice_cols, c_end = min(p.I, c_begin + p.slice_cols);
  const long long orow = (long long)blockIdx.y * p.B + b;        // row of the per-slice outputs (dPc, da2_rows, da20_rows)
  for (int base = c_begin + warp * 32; base < c_end; base += BWD_WARPS * 32) {
    const int i = base + lane;
    float a = 0.f, u = 0.f;
    if (i < c_end) { a = __ldcs(arow + i); u = __ldcs(urow + i); }
    unsigned mask = __ballot_sync(FULL, a != 0.f);
    while (mask) {
      const int j0 = __ffs(mask) - 1;
      mask &= mask - 1;
      const bool live1 = mask != 0u;
      const int j1 = live1 ? __ffs(mask) - 1 : j0;
      mask &= mask - 1;                                           // (0 & anything stays 0)
      row.pair(base + j0, __shfl_sync(FULL, a, j0), __shfl_sync(FULL, u, j0), base + j1, __shfl_sync(FULL, a, j1), __shfl_sync(FULL, u, j1), live1);
    }
  }
#pragma unroll
  for (int hv = 0; hv < HV; ++hv)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_dpc[warp][hv * 128 + lane * 4 + e] = row.dpc[hv][e];
      if (MODE == BWD_NET) s_da2[warp][hv * 128 + lane * 4 + e] = row.da2[hv][e];
    }
  if (lane == 0) s_da20[warp] = row.da20;
  __syncthreads();
  for (int h = threadIdx.x; h < p.H; h += BWD_WARPS * 32) {
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) {
      s += s_dpc[w][h];
      if (MODE == BWD_NET) s2 += s_da2[w][h];
    }
    p.dPc[orow * p.H + h] = s;
    if (MODE == BWD_NET) p.da2_rows[orow * p.H + h] = s2;
  }
  if (MODE == BWD_NET && threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += s_da20[w];
    p.da20_rows[orow] = s;
  }
}

template <int MODE>
static int launch_bwd(const AttBwdParams& p, cudaStream_t st) {
  const int w = p.H > p.U ? p.H : p.U;
  const dim3 grid(p.B, p.n_slices);               // every part is written, also one whose slice starts beyond I (zeros)
  if (w <= 128) attention_pool_bwd_kernel<1, 1, MODE><<<grid, BWD_WARPS * 32, 0, st>>>(p);
  else if (w <= 256) attention_pool_bwd_kernel<2, 2, MODE><<<grid, BWD_WARPS * 32, 0, st>>>(p);
  else if (w <= 512) attention_pool_bwd_kernel<4, 4, MODE><<<grid, BWD_WARPS * 32, 0, st>>>(p);
  else return b200rec_fail(B200REC_ERR_UNSUPPORTED, "attention_pool_backward: att_dense / user_emb wider than 512");
  B200REC_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec

using namespace b200rec;

extern "C" int b200rec_attention_pool_backward(const b200rec_attention_bwd_t* a, b200rec_stream_t stream) {
  if (!a) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: null descriptor");
  if (a->B < 0 || a->I < 0 || a->H <= 0 || a->U <= 0 || (a->H % 4) || (a->U % 4))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: H and U must be positive multiples of 4");
  if (a->B == 0 || a->I == 0) return B200REC_OK;
  if (!a->Pc || !a->Pr || !a->Q || !a->user_matrix || !a->att_weights || !a->out || !a->grad_out || !a->dPc || !a->dPr || !a->dQ)
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: null operand");
  if (a->mode != BWD_NET && a->mode != BWD_DOT) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: bad mode");
  if (a->mode == BWD_NET && (!a->a2 || !a->da2_rows || !a->da20_rows))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: MODE_NET needs a2, da2_rows and da20_rows");
  AttBwdParams p;
  p.Pc = a->Pc; p.Pr = a->Pr; p.Q = a->Q; p.a2 = a->a2; p.um = a->user_matrix; p.att = a->att_weights; p.out = a->out; p.bU = a->bU;
  p.g = a->grad_out;
  p.ldPc = a->ld_pc ? a->ld_pc : a->H; p.ldPr = a->ld_pr ? a->ld_pr : a->H; p.ldQ = a->ld_q ? a->ld_q : a->U;
  p.ld_um = a->ld_user_matrix ? a->ld_user_matrix : a->I; p.ldo = a->ldo ? a->ldo : a->U; p.ldg = a->ld_grad_out ? a->ld_grad_out : a->U;
  p.B = (int)a->B; p.I = (int)a->I; p.H = a->H; p.U = a->U;
  p.score_scale = a->score_scale == 0.f ? 1.f : a->score_scale;
  p.drop_key = 0u; p.drop_thr16 = 65536u; p.drop_scale = 1.f; p.drop_seed_dev = nullptr;
  if (a->dropout_p != 0.f) {                        // same derivation as b200rec_attention_pool_dropout
    if (!(a->dropout_p > 0.f && a->dropout_p < 1.f) || a->B >= (1 << 26)) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: bad dropout_p");
    unsigned thr = (unsigned)((1.0 - (double)a->dropout_p) * 65536.0 + 0.5);
    if (thr < 1u) thr = 1u;
    if (thr > 65536u) thr = 65536u;
    p.drop_thr16 = thr;
    p.drop_scale = 65536.f / (float)thr;
    p.drop_key = (unsigned)(a->dropout_seed ^ (a->dropout_seed >> 32));
    p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(a->dropout_seed_dev);
  }
  {
    const int n_slices = a->n_slices > 1 ? a->n_slices : 1;
    if (n_slices > 65535) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: too many slices");
    p.n_slices = n_slices;
    p.slice_cols = (int)(((a->I + n_slices - 1) / n_slices + 31) / 32 * 32);
  }
  p.dPc = a->dPc; p.dPr = a->dPr; p.dQ = a->dQ; p.da2_rows = a->da2_rows; p.da20_rows = a->da20_rows;
  if (p.ldPc < p.H || p.ldPr < p.H || p.ldQ < p.U || p.ldo < p.U || p.ldg < p.U || p.ld_um < p.I || (p.ldPc % 4) || (p.ldPr % 4) || (p.ldQ % 4) ||
      (p.ldo % 4) || (p.ldg % 4))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: bad leading dimension");
  const void* al[] = {p.Pc, p.Pr, p.Q, p.a2, p.out, p.bU, p.g, p.dPr, p.dQ};
  for (const void* q : al)
    if ((uintptr_t)q % 16) return b200rec_fail(B200REC_ERR_BAD_ARG, "attention_pool_backward: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  return a->mode == BWD_NET ? launch_bwd<BWD_NET>(p, st) : launch_bwd<BWD_DOT>(p, st);
}
