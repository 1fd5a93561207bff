// Micro-benchmark: does the per-ADDRESS serialisation of the L2 atomic unit bound the SGD update's memory pattern,
// and do split accumulators lift it?  Same loop as tools/l2_atomic_peak.cu's `both` (gather one 512-B row, reduce one
// 512-B row of increments into it; no arithmetic), on the item ids of a real rating stream (argv[3]: int32 ids), with
// the H most frequent rows given R ACCUMULATOR rows each: a hot update reads base + R accumulators (R+1 row gathers)
// and reduces into accumulator (warp mod R) - R times fewer reductions per address, same sum.
// Run: tools/l2_hot_rows rows row_floats ids.bin [mega-ids]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                           \
    }                                                                                    \
  } while (0)

__device__ __forceinline__ void red4(float4* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// MODE 0: reductions only, 2: gather + reduction
template <int MODE>
__global__ void __launch_bounds__(256) rep_kernel(float4* mat, float4* acc, const uint8_t* __restrict__ slot_of,
                                                  const int* __restrict__ idx, int64_t n, int nvec, int R, float* sink) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int myrep = (int)(warp % R);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
    const int mine = (i0 + lane < n) ? idx[i0 + lane] : 0;
    const int myslot = (i0 + lane < n) ? slot_of[mine] : 0;
    const int cnt = (int)min((int64_t)32, n - i0);
#pragma unroll 4
    for (int j = 0; j < cnt; j++) {
      const int row = __shfl_sync(0xffffffffu, mine, j);
      const int slot = __shfl_sync(0xffffffffu, myslot, j);
      if (lane < nvec) {
        float4* p = mat + (int64_t)row * nvec + lane;
        if (MODE == 2) {
          const float4 v = ldcg4(p);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (slot == 0) {
          red4(p, z);
        } else {
          float4* q = acc + (int64_t)(slot - 1) * R * nvec + lane;
          if (MODE == 2)
            for (int r = 0; r < R; r++) {
              const float4 v = ldcg4(q + (int64_t)r * nvec);
              a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
          red4(q + (int64_t)myrep * nvec, z);
        }
      }
    }
  }
  if (a.x + a.y + a.z + a.w == 12345.678f) *sink = a.x;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s rows row_floats ids.bin [mega-ids]\n", argv[0]);
    return 2;
  }
  const int rows = atoi(argv[1]), row_floats = atoi(argv[2]), nvec = row_floats / 4;
  FILE* f = fopen(argv[3], "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END);
  int64_t n = ftell(f) / 4;
  fseek(f, 0, SEEK_SET);
  if (argc > 4) n = std::min<int64_t>(n, (int64_t)atoll(argv[4]) << 20);
  std::vector<int> h(n);
  if (fread(h.data(), 4, (size_t)n, f) != (size_t)n) return 1;
  fclose(f);
  std::vector<int64_t> cnt(rows, 0);
  for (int x : h) cnt[x]++;
  std::vector<int> order(rows);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  float *mat, *acc, *sink;
  int* idx;
  uint8_t* slot_of;
  const int HMAX = 255, RMAX = 16;
  CK(cudaMalloc(&mat, (size_t)rows * row_floats * 4));
  CK(cudaMemset(mat, 0, (size_t)rows * row_floats * 4));
  CK(cudaMalloc(&acc, (size_t)HMAX * RMAX * row_floats * 4));
  CK(cudaMemset(acc, 0, (size_t)HMAX * RMAX * row_floats * 4));
  CK(cudaMalloc(&idx, n * 4));
  CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&slot_of, rows));
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("{\"ids\": %lld, \"rows\": %d, \"top_share\": %.4f, \"top8_share\": %.4f, \"top64_share\": %.4f}\n", (long long)n, rows,
         (double)cnt[order[0]] / n, (double)std::accumulate(order.begin(), order.begin() + 8, (int64_t)0, [&](int64_t s, int r) { return s + cnt[r]; }) / n,
         (double)std::accumulate(order.begin(), order.begin() + 64, (int64_t)0, [&](int64_t s, int r) { return s + cnt[r]; }) / n);
  for (int mode : {0, 2}) {
    for (int H : {0, 8, 32, 128, 255}) {
      for (int R : {2, 4, 8, 16}) {
        if (H == 0 && R != 2) continue;
        std::vector<uint8_t> so(rows, 0);
        for (int i = 0; i < H; i++) so[order[i]] = (uint8_t)(i + 1);
        CK(cudaMemcpy(slot_of, so.data(), rows, cudaMemcpyHostToDevice));
        float best = 1e30f;
        int best_cps = 0;
        for (int cps : {2, 4, 8}) {
          const int grid = prop.multiProcessorCount * cps;
          for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            if (mode == 0) rep_kernel<0><<<grid, 256>>>((float4*)mat, (float4*)acc, slot_of, idx, n, nvec, R, sink);
            else rep_kernel<2><<<grid, 256>>>((float4*)mat, (float4*)acc, slot_of, idx, n, nvec, R, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) { best = ms; best_cps = cps; }
          }
        }
        printf("{\"pattern\": \"%s\", \"hot_rows\": %d, \"accumulators\": %d, \"ms\": %.3f, \"g_updates_per_s\": %.3f, \"ctas_per_sm\": %d}\n",
               mode == 0 ? "red" : "both", H, H ? R : 1, best, n / (best * 1e-3) / 1e9, best_cps);
        fflush(stdout);
      }
    }
  }
  return 0;
}
