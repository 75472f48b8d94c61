// K3s — edge-balanced ("stream") form of the GraphNCF propagation SpMM for inference (gnn_ncf.py:39-94 + :351; see csrc/spmm.cu
// for the algebra).  Same result as K3, different work decomposition:
//
//   K3 (spmm_chunk_kernel) gives a warp ONE ROW (or a 256-edge chunk of a long row).  On a partitioned graph most rows are short
//   — an item row keeps only the edges of one rank's users, ~50 on the MovieLens-25M shape — so a warp spends its time in the
//   prologue / epilogue, chunk lengths inside a CTA differ by 100x and half of the resident warp slots idle (measured per rank of
//   an 8-way partition: 21-25 G edges/s against 34 G on the whole graph).
//
//   Here a warp owns a SEGMENT of `seg` consecutive CSR entries, whatever rows they belong to.  The last entry of every row is
//   flagged in the top bit of its column word; the warp accumulates and, at a flag, finishes the row (epilogue) and starts the
//   next one — the flags of 32 entries arrive as ONE ballot, batches without a flag take a branch-free path.  Rows cut by a
//   segment boundary leave a partial in the slot list of csrc/spmm.cu's fix-up kernel (at most two per segment), added in
//   segment order: deterministic, no atomics.  Every warp does the same amount of work; there is no per-row prologue.
//
//   The destination normalisation is folded into the entry value at plan time (wd[k] = deg[dst]^-1/2 * w[k]; the reference forms
//   exactly this per-edge product, gnn_ncf.py:54,91), so finishing a row needs no dependent load.  Rows without entries are
//   zero-filled by the tail CTAs of the same launch.  d <= 128 fp32 or bf16 features: a lane owns 4 consecutive columns.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace b200rec {

constexpr int SS_WARPS = 8;

struct StreamParams {
  const int* colf;             // (nnz) source node | last-entry-of-its-row flag in bit 31
  const float* wd;             // (nnz) deg[dst]^-1/2 * edge weight
  long long nnz;
  int seg;                     // entries per warp (multiple of 32)
  int n_segs;
  const int* seg_first_j;      // (n_segs) index into rows_ne of the row that holds the segment's first entry
  const int* seg_head_slot;    // (n_segs) partial slot of that row if it began in an earlier segment, else -1
  const int* seg_tail_slot;    // (n_segs) partial slot of the row that is still open at the segment's end, else -1
  const int* rows_ne;          // (n_ne) rows with at least one entry, ascending
  int n_ne;
  const int* rows_empty;       // (n_empty) rows without entries
  int n_empty;
  const void* t;
  long long ld_t;
  int d;
  float* partials;
  float* x_next; long long ld_x;
  const float* acc_in; float* acc_out; long long ld_acc; float acc_scale;
  void* push_dst[B200REC_PEER_MAX];
  int push_parts; int push_rpp; long long push_off; long long push_ld;
};

template <bool PUSH>
__device__ __forceinline__ void ss_epilogue(const StreamParams& p, int row, int c, float4 v) {
  if constexpr (PUSH) {
    const int o = row / p.push_rpp;
    st4(reinterpret_cast<float*>(p.push_dst[o]) + p.push_off + (long long)(row - o * p.push_rpp) * p.push_ld + c, v);
    return;
  }
  if (p.x_next) st4(p.x_next + (long long)row * p.ld_x + c, v);
  if (p.acc_out) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.acc_in) a = *reinterpret_cast<const float4*>(p.acc_in + (long long)row * p.ld_acc + c);
    st4(p.acc_out + (long long)row * p.ld_acc + c,
        make_float4((a.x + v.x) * p.acc_scale, (a.y + v.y) * p.acc_scale, (a.z + v.z) * p.acc_scale, (a.w + v.w) * p.acc_scale));
  }
}

template <typename T> struct SsRow;
template <> struct SsRow<float> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load(const unsigned char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void fma(float (&a)[4], float w, const Raw& r) {
    a[0] = fmaf(w, __uint_as_float(r.x), a[0]); a[1] = fmaf(w, __uint_as_float(r.y), a[1]);
    a[2] = fmaf(w, __uint_as_float(r.z), a[2]); a[3] = fmaf(w, __uint_as_float(r.w), a[3]);
  }
};
template <> struct SsRow<__nv_bfloat16> {
  typedef uint2 Raw;
  static __device__ __forceinline__ Raw load(const unsigned char* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ void fma(float (&a)[4], float w, const Raw& r) {
    a[0] = fmaf(w, __uint_as_float(r.x << 16), a[0]); a[1] = fmaf(w, __uint_as_float(r.x & 0xffff0000u), a[1]);
    a[2] = fmaf(w, __uint_as_float(r.y << 16), a[2]); a[3] = fmaf(w, __uint_as_float(r.y & 0xffff0000u), a[3]);
  }
};

template <typename T, bool PUSH>
__global__ void __launch_bounds__(SS_WARPS * 32, 4)
spmm_stream_kernel(const __grid_constant__ StreamParams p) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * SS_WARPS + (threadIdx.x >> 5);
  const int c = lane * 4;                                    // the 4 columns this lane owns
  const bool active = c < p.d;
  if (wid >= p.n_segs) {
    // ---- tail CTAs: rows without entries (x' = 0) ----
    if constexpr (!PUSH) {
      const int e = wid - ((p.n_segs + SS_WARPS - 1) / SS_WARPS) * SS_WARPS;
      if (e >= 0 && e < p.n_empty && active) ss_epilogue<false>(p, __ldg(p.rows_empty + e), c, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    return;
  }
  const long long k0 = (long long)wid * p.seg;
  const long long kend = min(k0 + (long long)p.seg, p.nnz);
  int j = __ldg(p.seg_first_j + wid);
  const int head_slot = __ldg(p.seg_head_slot + wid);
  bool open_is_head = head_slot >= 0;                        // the row open at the segment's start began earlier: its sum here is a partial
  int jbase = j;
  int myrow = __ldg(p.rows_ne + min(jbase + lane, p.n_ne - 1));
  const unsigned char* base = reinterpret_cast<const unsigned char*>(p.t) + (size_t)(active ? c : 0) * sizeof(T);
  const unsigned stride_bytes = (unsigned)(p.ld_t * (long long)sizeof(T));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};

  auto finish_row = [&]() {
    if (open_is_head) {
      if (active) st4(p.partials + (long long)head_slot * p.d + c, make_float4(acc[0], acc[1], acc[2], acc[3]));
      open_is_head = false;
    } else {
      if (j - jbase >= 32) {                                 // next 32 row numbers (one reload per 32 finished rows)
        jbase = j;
        myrow = __ldg(p.rows_ne + min(jbase + lane, p.n_ne - 1));
      }
      const int row = __shfl_sync(FULL, myrow, j - jbase);
      if (active) ss_epilogue<PUSH>(p, row, c, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    ++j;
  };

  int cf_nx = 0;
  float w_nx = 0.f;
  if (k0 + lane < kend) {
    cf_nx = __ldcs(p.colf + k0 + lane);
    w_nx = __ldcs(p.wd + k0 + lane);
  }
  bool open = false;                                         // a row is still open after the segment's last entry
  for (long long kb = k0; kb < kend; kb += 32) {
    const int cnt = (int)min((long long)32, kend - kb);
    const int cf = cf_nx;
    const float wv = w_nx;
    cf_nx = 0; w_nx = 0.f;
    if (kb + 32 + lane < kend) {
      cf_nx = __ldcs(p.colf + kb + 32 + lane);
      w_nx = __ldcs(p.wd + kb + 32 + lane);
    }
    const unsigned last = __ballot_sync(FULL, cf < 0);       // (padding lanes hold 0: never flagged)
    const unsigned col = (unsigned)cf & 0x7fffffffu;
#pragma unroll 1
    for (int st0 = 0; st0 < 32; st0 += 8) {
      if (st0 >= cnt) break;                                  // warp-uniform
      typename SsRow<T>::Raw x[8];
      float ww[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned cc = __shfl_sync(FULL, col, st0 + u);
        ww[u] = __shfl_sync(FULL, wv, st0 + u);
        x[u] = SsRow<T>::load(base + (unsigned long long)cc * stride_bytes);
      }
      const unsigned m8 = (last >> st0) & 0xffu;
      if (m8 == 0u) {                                         // no row ends inside these 8 entries
#pragma unroll
        for (int u = 0; u < 8; ++u) SsRow<T>::fma(acc, ww[u], x[u]);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) {                         // (padding entries after the last real one carry weight 0 and no flag)
          SsRow<T>::fma(acc, ww[u], x[u]);
          if ((m8 >> u) & 1u) finish_row();
        }
      }
    }
    open = ((last >> (cnt - 1)) & 1u) == 0u;                  // is the block's last real entry the end of its row?
  }
  if (open && active) {                                      // the row continues in the next segment
    const int slot = open_is_head ? head_slot : __ldg(p.seg_tail_slot + wid);
    st4(p.partials + (long long)slot * p.d + c, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

}  // namespace b200rec

using namespace b200rec;

// shared with csrc/spmm.cu: the fix-up pass over the slot lists (declared there)
int b200rec_spmm_fixup_launch(const float* partials, int d, const int* multi_row, const int* multi_first_slot, const int* multi_n_slots, int n_multi,
                              float* x_next, long long ld_x, const float* acc_in, float* acc_out, long long ld_acc, float acc_scale,
                              void* const* push_dst, int push_parts, int push_rpp, long long push_off, long long push_ld, cudaStream_t st);

extern "C" int b200rec_spmm_stream(const b200rec_spmm_stream_t* a, b200rec_stream_t stream) {
  if (!a) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: null descriptor");
  if (a->d <= 0 || a->d > 128 || (a->d % 4)) return b200rec_fail(B200REC_ERR_UNSUPPORTED, "spmm_stream: d must be a multiple of 4, <= 128");
  if (a->seg <= 0 || (a->seg % 32) || a->n_segs < 0 || a->nnz < 0 || a->n_ne < 0 || a->n_empty < 0 || a->n_multi < 0)
    return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: bad plan sizes");
  if ((long long)a->n_segs * a->seg < a->nnz) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: segments do not cover the entries");
  if (a->n_segs > 0 && (!a->colf || !a->wd || !a->seg_first_j || !a->seg_head_slot || !a->seg_tail_slot || !a->rows_ne || !a->t || a->n_ne <= 0))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: null plan / feature pointer");
  if (a->n_empty > 0 && !a->rows_empty) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: null empty-row list");
  if (a->n_multi > 0 && (!a->partials || !a->multi_row || !a->multi_first_slot || !a->multi_n_slots))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: rows cut by segment boundaries need partials + lists");
  const int esz = a->t_dtype == B200REC_BF16 ? 2 : 4;
  if (a->t_dtype != B200REC_F32 && a->t_dtype != B200REC_BF16) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: bad t_dtype");
  if ((uintptr_t)a->t % 16 || ((a->ld_t * esz) % (4 * esz)) || (a->ld_t % 4) || (a->x_next && ((uintptr_t)a->x_next % 16 || a->ld_x % 4)) ||
      (a->acc_out && ((uintptr_t)a->acc_out % 16 || a->ld_acc % 4)) || (a->acc_in && (uintptr_t)a->acc_in % 16) ||
      (a->partials && (uintptr_t)a->partials % 16))
    return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: feature pointers / leading dims must allow vector access");
  if (a->push_parts < 0 || a->push_parts > B200REC_PEER_MAX) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: push_parts out of range");
  StreamParams p;
  p.colf = a->colf; p.wd = a->wd; p.nnz = a->nnz; p.seg = a->seg; p.n_segs = a->n_segs;
  p.seg_first_j = a->seg_first_j; p.seg_head_slot = a->seg_head_slot; p.seg_tail_slot = a->seg_tail_slot;
  p.rows_ne = a->rows_ne; p.n_ne = a->n_ne; p.rows_empty = a->rows_empty; p.n_empty = a->n_empty;
  p.t = a->t; p.ld_t = a->ld_t; p.d = a->d; p.partials = a->partials;
  p.x_next = a->x_next; p.ld_x = a->ld_x; p.acc_in = a->acc_in; p.acc_out = a->acc_out; p.ld_acc = a->ld_acc; p.acc_scale = a->acc_scale;
  p.push_parts = a->push_parts; p.push_rpp = a->push_rows_per_part; p.push_off = a->push_offset; p.push_ld = a->push_ld;
  for (int q = 0; q < B200REC_PEER_MAX; ++q) p.push_dst[q] = a->push_dst[q];
  if (p.push_parts > 0) {
    if (p.push_rpp <= 0 || (p.push_ld % 4) || (p.push_off % 4) || p.push_ld < p.d) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: bad push arguments");
    for (int q = 0; q < p.push_parts; ++q)
      if (!p.push_dst[q] || ((uintptr_t)p.push_dst[q] % 16)) return b200rec_fail(B200REC_ERR_BAD_ARG, "spmm_stream: null / misaligned push destination");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int seg_ctas = (p.n_segs + SS_WARPS - 1) / SS_WARPS;
  const int empty_ctas = p.push_parts > 0 ? 0 : (p.n_empty + SS_WARPS - 1) / SS_WARPS;
  const int grid = seg_ctas + empty_ctas;
  if (grid > 0) {
    const bool bf = a->t_dtype == B200REC_BF16;
    if (p.push_parts > 0) {
      if (bf) spmm_stream_kernel<__nv_bfloat16, true><<<grid, SS_WARPS * 32, 0, st>>>(p);
      else spmm_stream_kernel<float, true><<<grid, SS_WARPS * 32, 0, st>>>(p);
    } else {
      if (bf) spmm_stream_kernel<__nv_bfloat16, false><<<grid, SS_WARPS * 32, 0, st>>>(p);
      else spmm_stream_kernel<float, false><<<grid, SS_WARPS * 32, 0, st>>>(p);
    }
    B200REC_CHECK_LAUNCH();
  }
  if (a->n_multi > 0)
    return b200rec_spmm_fixup_launch(a->partials, a->d, a->multi_row, a->multi_first_slot, a->multi_n_slots, a->n_multi, a->x_next, a->ld_x,
                                     a->acc_in, a->acc_out, a->ld_acc, a->acc_scale, a->push_dst, a->push_parts, a->push_rows_per_part,
                                     a->push_offset, a->push_ld, st);
  return B200REC_OK;
}
