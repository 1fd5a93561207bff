// Internal declarations shared by the .cu/.cc files behind include/mf_b200.h.
#ifndef MFB_INTERNAL_H
#define MFB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/mf_b200.h"

namespace mfb {

void set_error(const char* fmt, ...);

#define MFB_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      mfb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MFB_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define MFB_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      mfb::set_error(__VA_ARGS__);  \
      return MFB_E_ARG;             \
    }                               \
  } while (0)

// A rating file in file order.  Host staging vectors are filled by append_block / load_file and
// released by finalize() once the SoA tiles are resident in HBM.
struct Dataset {
  bool used = false, finalized = false;
  // host staging (file order)
  std::vector<int32_t> h_run_uid;
  std::vector<int32_t> h_run_off;   // [nruns+1]
  std::vector<int32_t> h_vid;
  std::vector<float> h_rating;
  std::vector<int64_t> h_block_off; // [nblocks+1] first run of each block
  bool pinned = false;              // host arrays registered with cudaHostRegister
  // device SoA tiles
  int64_t nruns = 0, nratings = 0, nblocks = 0;
  int32_t* d_run_uid = nullptr;
  int32_t* d_run_off = nullptr;
  int32_t* d_vid = nullptr;
  float* d_rating = nullptr;
  // dpmf static logical clock (dpmf.h:61-66 evaluated on the file order): steps since the row
  // was last touched, per record; and per row the clock of its last touch (for finish_noise)
  int32_t* d_uc = nullptr;
  int32_t* d_vc = nullptr;
  int32_t* d_last_u = nullptr;  // [nu]
  int32_t* d_last_v = nullptr;  // [nv]
};

struct Context {
  int device = 0;
  int nu = 0, nv = 0, dim = 0, stride = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  float* arr[12] = {nullptr};
  // work-queue head + fp64 accumulators live in one small device block
  int* d_counter = nullptr;      // [4]
  double* d_accum = nullptr;     // [8]
  double* h_accum = nullptr;     // pinned [8]
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-streamed epochs: a second stream carries the H2D copies of the next chunk
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_events;
  bool timed = false;
  int64_t launches = 0;
  // options
  int opt_ctas_per_sm = 0, opt_threads = 256, opt_batch = 4;
  std::vector<Dataset> datasets;
};

int64_t array_rows(const Context* c, int which);
int array_cols(const Context* c, int which);      // logical columns (dim or 1)
int array_stride(const Context* c, int which);    // device row stride in floats

// kernels (mfb_sgd.cu)
// runs [run_begin, run_end) of the dataset, in the given schedule
int launch_sgd(Context* c, Dataset* d, float eta, float lambda, float gb, int mode,
               int64_t run_begin, int64_t run_end);
int launch_sse(Context* c, Dataset* d, float gb);
int launch_fill_normal(Context* c, uint64_t seed, float scale);

// wire decoder (proto_wire.cc): appends every block of a [u32][mf.Block] file to the dataset
int load_blocks_file(const char* path, Dataset* d);

}  // namespace mfb

struct mfb_ctx {
  mfb::Context c;
};

struct mfb_blocks;
namespace mfb {
// host_blocks.cc: an mfb_blocks is a Dataset that only ever uses the host staging arrays
const Dataset* blocks_data(const mfb_blocks* b);
Dataset* blocks_mut(mfb_blocks* b);
mfb_blocks* blocks_new();
}  // namespace mfb

#endif
