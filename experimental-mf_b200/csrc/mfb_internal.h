// Internal declarations shared by the .cu/.cc files behind include/mf_b200.h.
#ifndef MFB_INTERNAL_H
#define MFB_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string.h>

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
  // compact wire form of vid/rating for host -> device streaming (mfb_blocks_pin builds it when the
  // data allow: item ids below 65536 and at most 256 distinct rating values): 3 bytes per record
  // instead of 8, lossless
  uint16_t* p_vid = nullptr;   // cudaHostAlloc'ed (page-locked from the start): item id, low 16 bits
  uint8_t* p_code = nullptr;
  uint8_t* p_vhi = nullptr;    // item id, bits 16..23: only when some id needs them (4 bytes per record instead of 3)
  float p_dict[256] = {0};
  bool packed = false;
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
  std::vector<int32_t> h_ucount, h_vcount;  // records per user / item (dpmf weights)
  double max_item_share = 0.0;              // records of the most rated item / all records
  cudaEvent_t refreshed = nullptr;          // pending H2D refresh of the tiles (copy stream)
  bool refresh_pending = false;
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
  cudaStream_t stream2 = nullptr;   // second compute stream of the streamed epoch (alternating chunks)
  cudaEvent_t ev_s2 = nullptr;
  int counter_slot = 0;             // which int of d_counter the next epoch launch uses as its queue head
  int width_div = 1;                // the next launch takes 1/width_div of the width the bounds allow
  std::vector<cudaEvent_t> chunk_events;
  // ... packed chunks land in one of two staging buffers and are expanded on the device
  uint16_t* d_stage_vid[2] = {nullptr, nullptr};
  uint8_t* d_stage_code[2] = {nullptr, nullptr};
  uint8_t* d_stage_vhi[2] = {nullptr, nullptr};
  float* d_dict = nullptr;  // [256]
  int64_t stage_capacity = 0;
  int stage_flip = 0;               // which staging buffer the next packed refresh uses

  bool timed = false;
  int64_t launches = 0;
  int64_t h2d_bytes = 0;  // bytes copied host -> device by the streamed epochs since create
  // dpmf: optional emulation of the reference's noise_ table (ordered parity mode only)
  float* d_noise_table = nullptr;
  int64_t noise_table_size = 0;
  double* d_norms = nullptr;  // [2*(stride+1)]
  // admf: validation records, per-run draws, the four regularisers
  int32_t* d_val_u = nullptr;
  int32_t* d_val_v = nullptr;
  float* d_val_r = nullptr;
  int64_t nvalid = 0;
  int32_t* d_draws = nullptr;
  int64_t ndraws = 0;
  float* d_lams = nullptr;  // [4]
  int* d_version = nullptr;                 // staleness probe (mfb_probe_*): per-item update counters
  unsigned long long* d_probe = nullptr;    // [4]
  int probe_item = -1;
  void* comm = nullptr;     // mfb::Comm (mfb_comm.cu): NCCL communicator of the DSGD ring
  // options
  int opt_ctas_per_sm = 0, opt_threads = 256, opt_batch = 4, opt_memopt = 0;
  int opt_kernel = 0;           // SGD kernel: 3 = sub-warp streaming (mfb_sgd_stream.cu); 1 = warp per run,
                                // one record at a time; 2 = warp per run, 4 records batched; 4 = burst
                                // (mfb_sgd_burst.cu); 0 = choose between 3 and 4
  int use_kernel = 3;           // ... the one chosen for the most recent epoch
  double rate_stream = 1.2e6, rate_burst = 5.0e6;  // updates/s per run in flight (measured; for the choice)
  int opt_phi_planes = 0;       // stream kernel: 0 = rows; 1 = the item matrix as four 128-byte planes during whole-epoch
                                // launches (transposed into d_phi_planes and back): narrows the spread over placements
                                // (14.5..16.0 against 15.1..18.9 ms) but the searched rows are as fast (14.8..15.3 against
                                // 14.9..16.0 ms after the search), so it is off; 2 = experiment (sector addressing only)
  bool planes_allowed = false;  // set around launches that own the item matrix alone (mfb_sgd_epoch, calibration)
  float* d_phi_planes = nullptr;  // scratch copy in plane layout; lives in the placement arena once the search ran
  int opt_two_streams = 1;      // streamed epochs: alternate chunk kernels over two streams at half width
  int opt_ring_peer = 1;        // DSGD ring: use the peer-memory mapping once mfb_comm_ipc_import has set it up (0 = stay with
                                // ncclSend/ncclRecv: the host program clears it on EVERY rank when any rank failed to map)
  int opt_sgld_flat = 1;        // dpmf, parallel schedule: 1 = sub-warp kernel (sgld_flat_kernel), 0 = warp per run (round 1),
                                // 2 = sub-warp kernel with 8 lanes x 4 float4 at k = 128 (experiment)
  int opt_file_decode = 1;      // out-of-core epoch (mfb_file_epoch.cu): 1 = the raw bytes of the file go to the GPU and
                                // the records are decoded there (mfb_wire_decode.cu); 0 = decoded by the host cores
  void* file_pipe = nullptr;    // mfb::FilePipe: pinned / device buffers of that path, kept between epochs
  int opt_epoch_launches = 1;   // diagnostic: mfb_sgd_epoch as this many launches over equal run ranges
  int opt_span_runs = 0;        // burst kernel: runs per claim (0 = by file size: 8, 16 or 32)
  int opt_tail_runs = 2;        // stream/burst kernels: runs per group handed out one by one at the end of a launch
  int opt_depth = 1;            // burst kernel: batches requested ahead (1 or 2)
  int opt_ring = 0;             // streaming kernel: item rows in flight per sub-warp (1..4; 0 = choose the
                                // deepest ring the budget of the hottest row leaves room for)
  int opt_row_concurrency = 16; // bound on the stale updates of the hottest item row in flight at once,
                                // at eta = 0.02 (0 = none); see mfb_sgd_stream.cu launch_stream_t.  16 = eta * c <= 0.24
                                // with the measured 0.75 rows per sub-warp.  Round 2, full sizes: 0.47 is stable in
                                // epoch 1 (factors still small) but NOT later, when the bound has widened with 0.02/eta
                                // and the factor norms have grown: NaN at 0.45-0.48 in DSGD cells from epoch 2 and on
                                // the Yahoo shape on one GPU (37,888 sub-warps at epoch 5); 0.24 is stable in both
  int opt_packed_h2d = 1;       // streamed epochs send the compact 3-byte records when the blocks have them
  int opt_throttle = 0;         // streaming kernel: closed loop on the L2 reduction queue (mfb_sgd_stream.cu)
  int opt_eta_scaling = 1;      // scale that bound with 0.02/eta (the budget is on eta * count)
  int last_grid = 0, last_threads = 0, last_ring = 0;  // launch shape of the most recent epoch kernel
  int opt_placement_trials = 16; // candidate placements of the item matrix tried before the first parallel epoch
                                // on a file of >= placement_min_ratings records (<= 1: off); tune_placement()
  int64_t placement_min_ratings = 4000000;
  bool placement_done[2] = {false, false};  // [0] rows of phi, [1] plane-layout scratch (tune_placement)
  bool bv_placed = false;
  float placement_ms[2][64] = {{0}};  // diagnostic: stage-1 calibration time of every candidate (mfb_placement_report)
  int placement_tried[2] = {0, 0}, placement_best[2] = {-1, -1};
  char* placement_arena = nullptr;  // candidate slots 1..n-1 (phi, bv, plane scratch each); slot 0 = place0
  int placement_slots = 0;
  float* place0[3] = {nullptr, nullptr, nullptr};  // the original allocations of phi, bv, plane scratch
  int opt_admf_weight = 0;      // admf kernel: item rows a run counts for in the hot-row budget (0 = depth + 2)
  int opt_admf_prefetch = 1;    // admf kernel: item rows requested ahead of the record being worked on (0..3); more than
                                // one gains nothing per run (251 warp instructions per step bound it) and costs width
  int opt_max_groups = 0;       // explicit cap on concurrent sub-warps (0 = derive from the above)
  int opt_run_fraction_ppm = 3500;  // user-runs in flight / user-runs of the file, parts per million (0 = no bound)
  // The ORDER effect this bound is for is created in the first epoch, while the factors leave their random
  // initialisation (tools/order_study.py, serial oracle: 5.6 % of the runs interleaved in every epoch: final tRMSE
  // +1.7e-3; in every epoch but the first: +1e-4), so "run_bound_epochs" = n lifts it after n epochs.  It stays on
  // in every epoch by default all the same: lifted, the wide launches of later epochs are UNSTABLE where it would
  // matter (ML-1M shape: divergence from epoch 2 at any hot-row budget; DSGD cells: NaN above row_concurrency 16),
  // and where they are stable nothing is gained (Netflix shape: the hardware width is reached either way; DSGD cells:
  // 6.1 ms per rank and epoch with or without) - measured in round 2, gpurun_out/r2_cells2.log.
  // model_age counts the whole epochs applied since the factors were last set (init / upload).
  int model_age = 0;
  int opt_run_bound_epochs = 1 << 30;
  std::vector<Dataset> datasets;
};

int64_t array_rows(const Context* c, int which);
int array_cols(const Context* c, int which);      // logical columns (dim or 1)
int array_stride(const Context* c, int which);    // device row stride in floats

// How many sub-warps may work at once.  Every sub-warp holds `inflight` item rows between gather
// and write-back; the hottest item (share p of all records) is therefore hit by about
// W*inflight*p stale updates at once, each applied with step eta.  The reference runs at most
// --fly (default 8) SgdFilter calls at a time (main.cc:50,97); here the bound is on the product
//     eta * W * inflight * p  <=  0.02 * row_concurrency
// (0.02 = the reference's default eta, main.cc:97).  Without it thousands of sub-warps read the
// same stale row: plain stores lose all but one update (measured: no convergence at ML-1M shape)
// and atomic accumulation applies them all at once (measured: divergence between 1.3 and 1.8).
// Second bound: the user-runs in flight are a "mini-batch" whose members do not see each other's
// updates; measured test-RMSE error vs the serial oracle grows with W / (runs in the file)
// (0.09% -> 7e-5, 0.35% -> 2.6e-4, 0.7% -> 9e-4, 1.4% -> 3e-3), so W <= run_fraction * runs.
struct LaunchShape {
  int grid, threads;
};
int64_t bounded_groups(const Context* c, int64_t groups, double max_item_share, int64_t total_runs,
                       double inflight, float eta);
LaunchShape pick_launch(Context* c, const void* kernel, int lpr, int64_t groups_needed,
                        double max_item_share, int64_t total_runs, int inflight = 6, float eta = 0.f);

// Where the item matrix lives decides the epoch time: the L2 slice of a line is a hash of its PHYSICAL
// address, the hottest slice bounds the atomic throughput, and which slices the rows of the most rated
// items share is luck per allocation (measured on one GPU, same data: 15.2 .. 19.6 ms per epoch over 8
// allocations, tools/exp_placement.py).  tune_placement copies phi/bv to `placement_trials` candidate
// allocations, runs the parallel SGD kernel with eta = 0 (increments of exactly zero: the model is
// unchanged, the memory traffic is the real one) over the first fifth of every given dataset on each,
// and keeps the fastest.  Once per context.
int tune_placement(Context* c, Dataset* const* ds, int nds, float gb, int mode, bool planes);
// kernels (mfb_sgd.cu)
// runs [run_begin, run_end) of the dataset, in the given schedule
int launch_sgd(Context* c, Dataset* d, float eta, float lambda, float gb, int mode,
               int64_t run_begin, int64_t run_end);
int launch_sse(Context* c, Dataset* d, float gb, int link = 0);
int launch_fill_normal(Context* c, uint64_t seed, float scale);
// expand n packed records (u16 item id, u8 rating code) into the SoA tiles
int launch_unpack(Context* c, const uint16_t* vid16, const uint8_t* vhi, const uint8_t* code, const float* dict,
                  int32_t* vid, float* rating, int64_t n);  // vhi may be NULL (ids below 65536)
// kernels (mfb_sgld.cu)
int launch_sgld(Context* c, Dataset* d, const mfb_sgld_params* p, float gb, int mode);
int launch_flush(Context* c, Dataset* d, const mfb_sgld_params* p);
int launch_col_sqnorms(Context* c, double* d_out);
// kernels (mfb_admf.cu)
int launch_admf(Context* c, Dataset* d, float eta, float eta_reg, int loss, float gb, int mode);

// wire decoder (proto_wire.cc): appends every block of a [u32][mf.Block] file to the dataset
int load_blocks_file(const char* path, Dataset* d);
// mfb_file_epoch.cu: releases c->file_pipe
void free_file_pipe(Context* c);

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
