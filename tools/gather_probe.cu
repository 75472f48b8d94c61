// Random-row gather ceiling of this GPU: what HBM delivers when every access is an independent `row_bytes`-wide row at a
// random place of a table much larger than L2 — the access pattern of K2 (rows of Pr / Q) and K3 (rows of t).  The copy
// bandwidth in MEASURED_PEAKS.json is a streaming figure; this probe measures the same silicon under the gather pattern
// with nothing else in the way (no index structure beyond one int per row, no arithmetic, results folded into one value per
// warp), so that K2's / K3's `roofline.frac` can be read against both.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/gather_probe tools/gather_probe.cu
//   tools/build/gather_probe [table_GiB=4] [gathers_M=64]      -> one JSON line on stdout
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void fill_idx(unsigned* idx, long long n, unsigned n_rows, unsigned seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    idx[i] = hash32((unsigned)i * 2654435761u + seed) % n_rows;
}

// G lanes per row (16 bytes each), 32/G rows per warp step, UN steps in flight
template <int G, int UN>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ table, const unsigned* __restrict__ idx, long long n,
                                                     unsigned* __restrict__ sink) {
  constexpr int RPS = 32 / G;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int g = lane / G, sl = lane % G;
  unsigned acc = 0;
  for (long long k0 = warp * 32; k0 < n; k0 += n_warps * 32) {
    const unsigned my = (k0 + lane < n) ? __ldcs(idx + k0 + lane) : 0u;
#pragma unroll
    for (int s0 = 0; s0 < G; s0 += UN) {
      uint4 x[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const unsigned r = __shfl_sync(0xffffffffu, my, (s0 + u) * RPS + g);
        x[u] = __ldg(table + (size_t)r * G + sl);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) acc += x[u].x ^ x[u].y ^ x[u].z ^ x[u].w;
    }
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) sink[warp] = acc;
}

__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}

template <int G, int UN>
static float run(const uint4* table, const unsigned* idx, long long n, unsigned* sink, int ctas) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    CK(cudaEventRecord(a));
    gather_kernel<G, UN><<<ctas, 256>>>(table, idx, n, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const double gib = argc > 1 ? atof(argv[1]) : 4.0;
  const long long n = (long long)((argc > 2 ? atof(argv[2]) : 64.0) * 1e6);
  const size_t bytes = (size_t)(gib * (1ull << 30));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  uint4 *table, *dst; unsigned *idx, *sink;
  CK(cudaMalloc(&table, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMemset(table, 1, bytes));
  CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&sink, 1 << 24));
  const int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"table_GiB\": %.2f, \"gathers\": %lld, \"l2_MB\": %.0f", prop.name, gib, n, prop.l2CacheSize / 1e6);
  {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      CK(cudaEventRecord(a)); copy_kernel<<<sms * 8, 256>>>(table, dst, bytes / 16); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (r > 0 && ms < best) best = ms;
    }
    printf(", \"copy_GBs\": %.1f", 2.0 * bytes / (best * 1e-3) / 1e9);
  }
  const int row_bytes[3] = {128, 256, 512};
  for (int rb = 0; rb < 3; ++rb) {
    const unsigned n_rows = (unsigned)(bytes / row_bytes[rb]);
    fill_idx<<<sms * 8, 256>>>(idx, n, n_rows, 17u + rb);
    CK(cudaDeviceSynchronize());
    float best = 1e30f; int best_cfg = 0;
    for (int occ = 4; occ <= 8; occ += 4) {
      float ms[2];
      const int ctas = sms * occ;
      switch (row_bytes[rb]) {
        case 128: ms[0] = run<8, 4>(table, idx, n, sink, ctas); ms[1] = run<8, 8>(table, idx, n, sink, ctas); break;
        case 256: ms[0] = run<16, 8>(table, idx, n, sink, ctas); ms[1] = run<16, 16>(table, idx, n, sink, ctas); break;
        default: ms[0] = run<32, 8>(table, idx, n, sink, ctas); ms[1] = run<32, 16>(table, idx, n, sink, ctas); break;
      }
      for (int v = 0; v < 2; ++v) if (ms[v] < best) { best = ms[v]; best_cfg = occ * 100 + v; }
    }
    const int eff_bytes = row_bytes[rb];
    printf(", \"gather_%dB_GBs\": %.1f, \"gather_%dB_cfg\": %d", row_bytes[rb], (double)n * (eff_bytes + 4) / (best * 1e-3) / 1e9, row_bytes[rb], best_cfg);
  }
  printf("}\n");
  return 0;
}
