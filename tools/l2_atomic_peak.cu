// Micro-benchmark: what can B200's L2 do with the memory pattern of the SGD epoch kernel?
//
// The epoch kernel (csrc/mfb_sgd_stream.cu) gathers one item row (k floats, 512 B at k=128) from an
// L2-resident matrix (Netflix shape: 17,770 rows = 9.1 MB), and sends one row of increments back as
// reductions.  DRAM is idle (profiles/r1_sgd_stream_placed.md: 2.8 % of HBM); the L2 is the bound.  This
// program measures that bound directly, with no arithmetic and no dependence between rows:
//   red      one 512-B row reduction per warp step, 32 x red.global.add.v4.f32       (the kernel's way)
//   gather   one 512-B row read per warp step, 32 x ld.global.cg.v4                   (L2 -> SM)
//   both     gather + red of the same row                                             (= one update)
//   bulkred  one cp.reduce.async.bulk.global.shared::cta.add.f32 of 512 B per row     (Blackwell/Hopper bulk mover)
//   bulkld   one cp.async.bulk.shared::cta.global of 512 B per row + mbarrier
// each with rows drawn uniformly (the ceiling an even spread over the L2 slices allows) and from the
// Zipf(1.0) popularity of the synthetic data (hot rows: the busiest slice decides).
// and, given a file of int32 item ids (argv[4]: the vid array of a real training file in file order), from that.
// (Zipf(1.0) sampled WITH replacement puts 9.6 % of the rows on one address and measures the per-address
// serialisation of the L2 atomic unit; the synthetic rating files cannot exceed nu/nnz = 0.48 % per item.)
// Output: one JSON object per line.   Build: make -C tools
// Run: tools/l2_atomic_peak [rows] [row_floats] [mega-rows per launch] [vid.bin]
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ void red4(float4* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// MODE 0 red, 1 gather, 2 both.  One warp step = one row of NVEC float4 (NVEC <= 32: lanes beyond idle).
template <int MODE>
__global__ void __launch_bounds__(256) rows_kernel(float4* mat, const int* __restrict__ idx, int64_t n, int nvec, float* sink) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
    const int mine = (i0 + lane < n) ? idx[i0 + lane] : 0;
    const int cnt = (int)min((int64_t)32, n - i0);
#pragma unroll 4
    for (int j = 0; j < cnt; j++) {
      const int row = __shfl_sync(0xffffffffu, mine, j);
      float4* p = mat + (int64_t)row * nvec + lane;
      if (lane < nvec) {
        if (MODE == 1 || MODE == 2) {
          const float4 v = ldcg4(p);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (MODE == 0 || MODE == 2) red4(p, z);
      }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}

// bulk reduction: every warp owns a zeroed row in shared memory; lane 0 sends it with one instruction per row
__global__ void __launch_bounds__(256) bulkred_kernel(float* mat, const int* __restrict__ idx, int64_t n, int row_floats) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* my = reinterpret_cast<float*>(smem) + (size_t)w * row_floats;
  for (int q = lane; q < row_floats; q += 32) my[q] = 0.f;
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(my);
  const uint32_t bytes = (uint32_t)row_floats * 4;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
    const int mine = (i0 + lane < n) ? idx[i0 + lane] : 0;
    const int cnt = (int)min((int64_t)32, n - i0);
    for (int j = 0; j < cnt; j++) {
      const int row = __shfl_sync(0xffffffffu, mine, j);
      if (lane == 0) {
        float* dst = mat + (int64_t)row * row_floats;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// bulk gather: DEPTH row buffers per warp, one mbarrier each; lane 0 issues, all lanes wait and read one float4
template <int DEPTH>
__global__ void __launch_bounds__(256) bulkld_kernel(const float* mat, const int* __restrict__ idx, int64_t n, int row_floats, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* rows = reinterpret_cast<float*>(smem) + (size_t)w * DEPTH * row_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * DEPTH * row_floats * 4) + w * DEPTH;
  const uint32_t bytes = (uint32_t)row_floats * 4;
  if (lane == 0)
    for (int d = 0; d < DEPTH; d++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bars + d)) : "memory");
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  auto issue = [&](int slot, int row) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + slot);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rows + (size_t)slot * row_floats);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(mat + (int64_t)row * row_floats), "r"(bytes), "r"(bar)
                 : "memory");
  };
  auto wait = [&](int slot, unsigned parity) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + slot);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
  };
  unsigned phase = 0;  // bit d = parity expected next on slot d
  int64_t step = 0;
  for (int64_t i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
    const int mine = (i0 + lane < n) ? idx[i0 + lane] : 0;
    const int cnt = (int)min((int64_t)32, n - i0);
    // software pipeline inside the chunk of 32 rows: DEPTH rows in flight
    for (int j = 0; j < cnt + DEPTH; j++) {
      if (j >= DEPTH) {
        const int slot = (int)((step + j - DEPTH) % DEPTH);
        wait(slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
        acc += rows[(size_t)slot * row_floats + lane * 4];
        __syncwarp();
      }
      if (j < cnt) {
        const int row = __shfl_sync(0xffffffffu, mine, j);
        if (lane == 0) issue((int)((step + j) % DEPTH), row);
      }
    }
    step += cnt;
  }
  if (acc == 12345.678f) *sink = acc;
}

static std::vector<int> make_indices(int rows, int64_t n, bool zipf, uint64_t seed) {
  std::mt19937_64 rng(seed);
  std::vector<int> out(n);
  if (!zipf) {
    for (auto& x : out) x = (int)(rng() % (uint64_t)rows);
    return out;
  }
  // Zipf(1.0) over a random permutation of the rows (the generator's item popularity)
  std::vector<double> cdf(rows);
  double s = 0;
  for (int i = 0; i < rows; i++) cdf[i] = (s += 1.0 / (i + 1));
  std::vector<int> perm(rows);
  for (int i = 0; i < rows; i++) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  std::uniform_real_distribution<double> u(0.0, s);
  for (auto& x : out) x = perm[std::lower_bound(cdf.begin(), cdf.end(), u(rng)) - cdf.begin()];
  return out;
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 17770;
  const int row_floats = argc > 2 ? atoi(argv[2]) : 128;
  const int64_t n = (int64_t)(argc > 3 ? atoll(argv[3]) : 64) << 20;  // rows touched per launch
  const int nvec = row_floats / 4;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  float* mat;
  int* idx;
  float* sink;
  const size_t mat_bytes = (size_t)rows * row_floats * 4;
  CK(cudaMalloc(&mat, mat_bytes));
  CK(cudaMemset(mat, 0, mat_bytes));
  CK(cudaMalloc(&idx, n * sizeof(int)));
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int smem_red = 8 * row_floats * 4;
  constexpr int DEPTH = 4;
  const int smem_ld = 8 * DEPTH * row_floats * 4 + 8 * DEPTH * 8;
  CK(cudaFuncSetAttribute(bulkld_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ld));
  const char* vid_path = argc > 4 ? argv[4] : nullptr;
  for (int zipf = 0; zipf < (vid_path ? 3 : 2); zipf++) {
    std::vector<int> h;
    if (zipf < 2) {
      h = make_indices(rows, n, zipf != 0, 0x4D46B200);
    } else {
      h.resize(n);
      FILE* f = fopen(vid_path, "rb");
      if (!f || fread(h.data(), sizeof(int), (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "cannot read %lld ids from %s\n", (long long)n, vid_path);
        return 1;
      }
      fclose(f);
      for (int x : h)
        if (x < 0 || x >= rows) {
          fprintf(stderr, "id %d outside [0,%d)\n", x, rows);
          return 1;
        }
    }
    CK(cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 5; mode++) {
      static const char* names[] = {"red", "gather", "both", "bulkred", "bulkld"};
      for (int cps : {2, 4, 8}) {  // CTAs of 256 threads per SM
        const int grid = prop.multiProcessorCount * cps;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
          CK(cudaEventRecord(e0));
          switch (mode) {
            case 0: rows_kernel<0><<<grid, 256>>>((float4*)mat, idx, n, nvec, sink); break;
            case 1: rows_kernel<1><<<grid, 256>>>((float4*)mat, idx, n, nvec, sink); break;
            case 2: rows_kernel<2><<<grid, 256>>>((float4*)mat, idx, n, nvec, sink); break;
            case 3: bulkred_kernel<<<grid, 256, smem_red>>>(mat, idx, n, row_floats); break;
            case 4: bulkld_kernel<DEPTH><<<grid, 256, smem_ld>>>(mat, idx, n, row_floats, sink); break;
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0) best = std::min(best, ms);
        }
        const double rps = (double)n / (best * 1e-3);
        printf("{\"pattern\": \"%s\", \"rows\": \"%s\", \"matrix_rows\": %d, \"row_bytes\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, "
               "\"grows_per_s\": %.3f, \"row_gbs\": %.1f}\n",
               names[mode], zipf == 0 ? "uniform" : (zipf == 1 ? "zipf1.0 with replacement" : "training file"), rows, row_floats * 4, cps, best, rps / 1e9, rps * row_floats * 4 / 1e9);
        fflush(stdout);
      }
    }
  }
  return 0;
}
