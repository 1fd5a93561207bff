// C-ABI entry points of include/mf_b200.h: context, factor access, datasets, epoch drivers.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <sys/mman.h>

#include "mfb_internal.h"

namespace mfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int64_t array_rows(const Context* c, int which) {
  switch (which) {
    case MFB_THETA: case MFB_THETA_OLD: case MFB_BU: case MFB_BU_OLD: case MFB_UR: return c->nu;
    case MFB_PHI: case MFB_PHI_OLD: case MFB_BV: case MFB_BV_OLD: case MFB_VR: return c->nv;
    case MFB_LAMBDA_U: case MFB_LAMBDA_V: return c->dim;
    default: return 0;
  }
}
int array_cols(const Context* c, int which) {
  switch (which) {
    case MFB_THETA: case MFB_THETA_OLD: case MFB_PHI: case MFB_PHI_OLD: return c->dim;
    default: return 1;
  }
}
int array_stride(const Context* c, int which) {
  switch (which) {
    case MFB_THETA: case MFB_THETA_OLD: case MFB_PHI: case MFB_PHI_OLD: return c->stride;
    default: return 1;
  }
}

static int alloc_array(Context* c, int which) {
  if (c->arr[which]) return MFB_OK;
  size_t n = (size_t)array_rows(c, which) * array_stride(c, which);
  if (which == MFB_LAMBDA_U || which == MFB_LAMBDA_V) n = (size_t)c->stride;  // padded with zeros
  n = (n + 15) / 16 * 16;  // kernels fetch the aligned 16 bytes around a bias element
  if (n == 0) n = 16;
  MFB_CUDA(cudaMalloc(&c->arr[which], n * sizeof(float)));
  MFB_CUDA(cudaMemsetAsync(c->arr[which], 0, n * sizeof(float), c->stream));
  return MFB_OK;
}

static int64_t bounded_groups_alone(const Context* c, int64_t groups, double max_item_share, int64_t total_runs,
                                    double inflight, float eta);

int64_t bounded_groups(const Context* c, int64_t groups, double max_item_share, int64_t total_runs,
                       double inflight, float eta) {
  const int64_t g = bounded_groups_alone(c, groups, max_item_share, total_runs, inflight, eta);
  // width_div > 1: this launch shares the machine and the bounds with concurrent ones
  return std::max<int64_t>(1, g / std::max(c->width_div, 1));
}

static int64_t bounded_groups_alone(const Context* c, int64_t groups, double max_item_share, int64_t total_runs,
                                    double inflight, float eta) {
  if (c->opt_max_groups > 0) return std::min<int64_t>(groups, c->opt_max_groups);
  // both budgets are on (step size) x (stale updates applied at once): they widen as eta decays
  const double widen = (c->opt_eta_scaling && eta > 0.f) ? 0.02 / (double)eta : 1.0;
  if (c->opt_row_concurrency > 0 && max_item_share > 0.0) {
    const double budget = (double)c->opt_row_concurrency * widen;
    groups = std::min<int64_t>(groups, std::max<int64_t>(1, (int64_t)(budget / (max_item_share * std::max(inflight, 0.1)))));
  }
  if (c->opt_run_fraction_ppm > 0 && total_runs > 0 && c->model_age < c->opt_run_bound_epochs) {
    const double frac = (double)c->opt_run_fraction_ppm * 1e-6 * (c->opt_eta_scaling >= 2 ? widen : 1.0);
    groups = std::min<int64_t>(groups, std::max<int64_t>(1, (int64_t)(total_runs * frac)));
  }
  return groups;
}

LaunchShape pick_launch(Context* c, const void* kernel, int lpr, int64_t groups_needed,
                        double max_item_share, int64_t total_runs, int inflight, float eta) {
  int max_threads = c->opt_threads;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, max_threads, 0);
  per_sm = std::max(per_sm, 1);
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  const int groups_per_warp = 32 / lpr;
  int64_t groups = (int64_t)c->sm_count * per_sm * (max_threads / 32) * groups_per_warp;  // hardware
  groups = std::min(groups, std::max<int64_t>(groups_needed, 1));
  groups = bounded_groups(c, groups, max_item_share, total_runs, inflight, eta);
  // spread the warps over all SMs: one CTA per SM with as many warps as needed, more CTAs beyond 8
  const int64_t warps = (groups + groups_per_warp - 1) / groups_per_warp;
  LaunchShape ls;
  if (warps <= c->sm_count) {
    ls.grid = (int)warps;
    ls.threads = 32;
  } else {
    const int warps_per_sm = (int)((warps + c->sm_count - 1) / c->sm_count);
    const int max_wpc = max_threads / 32;
    const int ctas = (warps_per_sm + max_wpc - 1) / max_wpc;
    ls.threads = 32 * ((warps_per_sm + ctas - 1) / ctas);
    ls.grid = c->sm_count * ctas;
  }
  c->last_grid = ls.grid;
  c->last_threads = ls.threads;
  return ls;
}

static Dataset* get_ds(Context* c, int ds) {
  if (ds < 0 || ds >= (int)c->datasets.size() || !c->datasets[ds].used) {
    set_error("bad dataset id %d", ds);
    return nullptr;
  }
  return &c->datasets[ds];
}

static void free_ds_device(Dataset* d) {
  if (d->refreshed) cudaEventDestroy(d->refreshed);
  d->refreshed = nullptr;
  cudaFree(d->d_run_uid); cudaFree(d->d_run_off); cudaFree(d->d_vid); cudaFree(d->d_rating);
  cudaFree(d->d_uc); cudaFree(d->d_vc); cudaFree(d->d_last_u); cudaFree(d->d_last_v);
  d->d_run_uid = d->d_run_off = d->d_vid = nullptr;
  d->d_rating = nullptr;
  d->d_uc = d->d_vc = d->d_last_u = d->d_last_v = nullptr;
}

int tune_placement(Context* c, Dataset* const* ds, int nds, float gb, int mode, bool planes) {
  // Two searches, each run once when first needed: the plane-layout scratch copy the stream kernel works
  // on during whole-epoch launches (planes), and the rows of phi themselves (chunked / multi-GPU epochs).
  // bv is placed by whichever runs first.
  const int which = planes ? 1 : 0;
  if (c->placement_done[which]) return MFB_OK;
  const int trials = std::min(c->opt_placement_trials, 64);
  if (trials <= 1 || (mode != MFB_MODE_ATOMIC && mode != MFB_MODE_HOGWILD)) return MFB_OK;
  int64_t total = 0;
  for (int i = 0; i < nds; i++) total += ds[i]->nratings;
  if (total < c->placement_min_ratings) return MFB_OK;  // (not marked done: a larger file may follow)
  c->placement_done[which] = true;
  const size_t phi_bytes = ((size_t)c->nv * c->stride + 15) / 16 * 16 * sizeof(float);
  // only a matrix that lives in the L2 is bound by its slices; a larger one (Yahoo shape: 320 MB) streams
  // from HBM and would pay hundreds of MB of copies per candidate for nothing
  if (phi_bytes > ((size_t)64 << 20)) return MFB_OK;
  const size_t bv_bytes = ((size_t)c->nv + 15) / 16 * 16 * sizeof(float);
  // a slot of the arena holds phi, bv and the plane-layout scratch copy; candidate 0 is the set of
  // original allocations (kept: the second search starts from them again)
  const size_t slot = (2 * phi_bytes + bv_bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  if (!c->d_phi_planes) MFB_CUDA(cudaMalloc(&c->d_phi_planes, phi_bytes));
  if (!c->placement_arena) {
    int n = trials;
    while (n > 1 && cudaMalloc(&c->placement_arena, (size_t)(n - 1) * slot) != cudaSuccess) {
      cudaGetLastError();
      c->placement_arena = nullptr;
      n = (n + 1) / 2;
    }
    if (!c->placement_arena) return MFB_OK;
    c->placement_slots = n;
    c->place0[0] = c->arr[MFB_PHI];
    c->place0[1] = c->arr[MFB_BV];
    c->place0[2] = c->d_phi_planes;
  }
  const int n = c->placement_slots;
  char* const arena = c->placement_arena;
  auto phi_of = [&](int i) { return i == 0 ? c->place0[0] : (float*)(arena + (size_t)(i - 1) * slot); };
  auto bv_of = [&](int i) { return i == 0 ? c->place0[1] : (float*)(arena + (size_t)(i - 1) * slot + phi_bytes); };
  auto planes_of = [&](int i) { return i == 0 ? c->place0[2] : (float*)(arena + (size_t)(i - 1) * slot + phi_bytes + bv_bytes); };
  float* const phi_cur = c->arr[MFB_PHI];
  float* const bv_cur = c->arr[MFB_BV];
  const bool place_bv = !c->bv_placed;
  cudaEvent_t e0, e1;
  MFB_CUDA(cudaEventCreate(&e0));
  MFB_CUDA(cudaEventCreate(&e1));
  // the launch of the steady state: as wide as the run bound allows, deepest ring; probe off
  const int save_kernel = c->opt_kernel, save_ring = c->opt_ring, save_groups = c->opt_max_groups;
  const bool save_timed = c->timed, save_allowed = c->planes_allowed;
  int* const save_version = c->d_version;
  c->d_version = nullptr;
  c->planes_allowed = planes;
  int rc = MFB_OK;
  auto calibrate = [&](int divisor) {
    for (int i = 0; i < nds && rc == MFB_OK; i++) {
      Dataset* d = ds[i];
      if (!d->nruns) continue;
      c->opt_kernel = 0;
      c->opt_ring = 0;
      c->opt_max_groups = (int)std::max<int64_t>(1, bounded_groups_alone(c, (int64_t)c->sm_count * 64, d->max_item_share,
                                                                         d->nruns, 4.0, 0.02f / 8));
      const int64_t prefix = std::min<int64_t>(d->nruns, std::max<int64_t>(d->nruns / divisor, 100000));
      rc = launch_sgd(c, d, 0.f, 0.f, gb, MFB_MODE_ATOMIC, 0, prefix);  // increments of exactly zero
    }
  };
  auto timed_run = [&](int i, int divisor, float* ms) -> int {
    if (planes) c->d_phi_planes = planes_of(i);
    else c->arr[MFB_PHI] = phi_of(i);
    if (place_bv) c->arr[MFB_BV] = bv_of(i);
    MFB_CUDA(cudaEventRecord(e0, c->stream));
    calibrate(divisor);
    MFB_CUDA(cudaEventRecord(e1, c->stream));
    MFB_CUDA(cudaEventSynchronize(e1));
    MFB_CUDA(cudaEventElapsedTime(ms, e0, e1));
    return rc;
  };
  calibrate(10);  // warm-up, not timed
  for (int i = 0; i < n && rc == MFB_OK; i++) {  // every candidate gets the (unchanged) values
    cudaError_t e = cudaSuccess;
    if (!planes && phi_of(i) != phi_cur) e = cudaMemcpyAsync(phi_of(i), phi_cur, phi_bytes, cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess && place_bv && bv_of(i) != bv_cur)
      e = cudaMemcpyAsync(bv_of(i), bv_cur, bv_bytes, cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) {
      set_error("placement search: device copy failed: %s", cudaGetErrorString(e));
      rc = MFB_E_CUDA;
    }
  }
  // stage 1: every candidate over the first tenth of the file(s); stage 2: the three fastest over the
  // first fifth, twice (a short prefix sees the launch tail and a narrower mix of items)
  std::vector<std::pair<float, int>> order;
  for (int i = 0; i < n && rc == MFB_OK; i++) {
    float ms = 0.f;
    rc = timed_run(i, 10, &ms);
    c->placement_ms[which][i] = ms;
    order.push_back({ms, i});
  }
  int best = 0;
  if (rc == MFB_OK) {
    std::sort(order.begin(), order.end());
    float best_ms = 0.f;
    for (size_t j = 0; j < order.size() && j < 3 && rc == MFB_OK; j++) {
      float ms = 0.f;
      rc = timed_run(order[j].second, 5, &ms);
      rc = rc ? rc : timed_run(order[j].second, 5, &ms);
      if (j == 0 || ms < best_ms) { best = order[j].second; best_ms = ms; }
    }
    c->placement_tried[which] = n;
    c->placement_best[which] = best;
  }
  c->opt_kernel = save_kernel;
  c->opt_ring = save_ring;
  c->opt_max_groups = save_groups;
  c->timed = save_timed;
  c->planes_allowed = save_allowed;
  c->d_version = save_version;
  cudaStreamSynchronize(c->stream);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // every candidate holds the same (unchanged) values: keep the fastest
  if (planes) c->d_phi_planes = planes_of(best);
  else c->arr[MFB_PHI] = phi_of(best);
  if (place_bv) {
    c->arr[MFB_BV] = bv_of(best);
    c->bv_placed = true;
  }
  return rc;
}

static void begin_timing(Context* c) {
  cudaEventRecord(c->ev0, c->stream);
}
static void end_timing(Context* c) {
  cudaEventRecord(c->ev1, c->stream);
  c->timed = true;
}

}  // namespace mfb

using namespace mfb;

// ---- host-side helpers of the ingest: the loops over all records run on all cores -----------------
template <typename F>
static void parallel_ranges(int64_t n, const F& fn) {  // fn(thread, begin, end) over [0, n) cut into equal parts
  const int64_t hw = std::max(1u, std::thread::hardware_concurrency());
  const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(hw, n / (1 << 16)));
  std::vector<std::thread> pool;
  for (int64_t t = 1; t < parts; t++) pool.emplace_back([&, t]() { fn((int)t, n * t / parts, n * (t + 1) / parts); });
  fn(0, 0, n / parts);
  for (auto& th : pool) th.join();
}
// first value outside [0, limit) or -1; *pos = its index
static int64_t find_out_of_range(const int32_t* v, int64_t n, int32_t limit, int32_t* value) {
  std::atomic<int64_t> bad(-1);
  parallel_ranges(n, [&](int, int64_t b, int64_t e) {
    for (int64_t i = b; i < e; i++)
      if ((uint32_t)v[i] >= (uint32_t)limit) {
        int64_t cur = bad.load();
        while ((cur < 0 || i < cur) && !bad.compare_exchange_weak(cur, i)) {}
        return;
      }
  });
  const int64_t at = bad.load();
  if (at >= 0) *value = v[at];
  return at;
}
template <typename T>
static void append_parallel(std::vector<T>& dst, const std::vector<T>& src) {
  const size_t base = dst.size();
  if (src.empty()) return;
  if (base != 0 || src.size() < ((size_t)1 << 20)) {
    dst.insert(dst.end(), src.begin(), src.end());
    return;
  }
  // an empty destination: first touch and copy by all threads (a serial insert() of 400 MB spends most
  // of its time in page faults).  T is trivially copyable, so the elements may be written before the
  // vector's size says they exist; resize() then value-initialises... which would zero them - so the
  // vector is sized first over pre-faulted pages and filled afterwards.
  dst.reserve(src.size());
#ifdef MADV_POPULATE_WRITE
  parallel_ranges((int64_t)src.size(), [&](int, int64_t b, int64_t e) {
    const uintptr_t lo = ((uintptr_t)(dst.data() + b) + 4095) & ~(uintptr_t)4095, hi = (uintptr_t)(dst.data() + e) & ~(uintptr_t)4095;
    if (hi > lo) madvise((void*)lo, hi - lo, MADV_POPULATE_WRITE);
  });
#endif
  dst.resize(src.size());
  parallel_ranges((int64_t)src.size(), [&](int, int64_t b, int64_t e) { memcpy(dst.data() + b, src.data() + b, (size_t)(e - b) * sizeof(T)); });
}

static int append_host(Context* c, Dataset* d, const Dataset& src) {
  if (d->h_run_off.empty()) d->h_run_off.push_back(0);
  if (d->h_block_off.empty()) d->h_block_off.push_back(0);
  const int64_t base = (int64_t)d->h_vid.size(), rbase = (int64_t)d->h_run_uid.size();
  MFB_REQUIRE(base + (int64_t)src.h_vid.size() < (int64_t)INT32_MAX, "dataset too large for int32 offsets");
  int32_t badv = 0;
  MFB_REQUIRE(find_out_of_range(src.h_run_uid.data(), (int64_t)src.h_run_uid.size(), c->nu, &badv) < 0,
              "uid %d outside [0,%d)", badv, c->nu);
  MFB_REQUIRE(find_out_of_range(src.h_vid.data(), (int64_t)src.h_vid.size(), c->nv, &badv) < 0,
              "vid %d outside [0,%d)", badv, c->nv);
  d->h_run_uid.insert(d->h_run_uid.end(), src.h_run_uid.begin(), src.h_run_uid.end());
  d->h_run_off.reserve(d->h_run_off.size() + src.h_run_off.size());
  for (size_t r = 1; r < src.h_run_off.size(); r++) d->h_run_off.push_back((int32_t)(base + src.h_run_off[r]));
  append_parallel(d->h_vid, src.h_vid);
  append_parallel(d->h_rating, src.h_rating);
  for (size_t b = 1; b < src.h_block_off.size(); b++) d->h_block_off.push_back(rbase + src.h_block_off[b]);
  return MFB_OK;
}

template <typename T>
static int to_device(Context* c, const std::vector<T>& hv, T** dp) {
  size_t bytes = std::max<size_t>(hv.size(), 1) * sizeof(T);
  MFB_CUDA(cudaMalloc(dp, bytes));
  if (!hv.empty())
    MFB_CUDA(cudaMemcpyAsync(*dp, hv.data(), hv.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return MFB_OK;
}

template <typename T>
static void release(std::vector<T>& v) {
  std::vector<T>().swap(v);
}

extern "C" {

const char* mfb_last_error(void) { return g_err; }
const char* mfb_version(void) { return "0.1 sm_100a"; }

// util.h:163-165 with CACHE_LINE_SIZE 64: round the row up to a multiple of 16 floats
int mfb_padding(int dim) {
  return (int)((((size_t)dim * sizeof(float) - 1) / 64 * 64 + 64) / sizeof(float));
}

float mfb_seteta(float eta0, int round, float gam) {  // model.cc:36-38
  return (float)((double)eta0 * 1.0 / pow((double)round, (double)gam));
}
float mfb_seteta_cutoff(float eta0, int round, float gam, float mineta) {  // model.cc:350-352
  const float e = mfb_seteta(eta0, round, gam);
  return mineta > e ? mineta : e;
}

int mfb_create(mfb_ctx** out, int device, int nu, int nv, int dim) {
  MFB_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  MFB_REQUIRE(nu > 0 && nv > 0 && dim > 0 && dim <= 2048, "bad shape nu=%d nv=%d dim=%d", nu, nv, dim);
  int ndev = 0;
  MFB_CUDA(cudaGetDeviceCount(&ndev));
  MFB_REQUIRE(device >= 0 && device < ndev, "device %d of %d", device, ndev);
  MFB_CUDA(cudaSetDevice(device));
  mfb_ctx* h = new mfb_ctx();
  Context* c = &h->c;
  c->device = device;
  c->nu = nu;
  c->nv = nv;
  c->dim = dim;
  c->stride = mfb_padding(dim);
  cudaDeviceProp prop;
  MFB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  MFB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  MFB_CUDA(cudaEventCreate(&c->ev0));
  MFB_CUDA(cudaEventCreate(&c->ev1));
  MFB_CUDA(cudaMalloc(&c->d_counter, 4 * sizeof(int)));
  MFB_CUDA(cudaMalloc(&c->d_accum, 8 * sizeof(double)));
  MFB_CUDA(cudaMallocHost(&c->h_accum, 8 * sizeof(double)));
  for (int w : {MFB_THETA, MFB_PHI, MFB_BU, MFB_BV}) {
    int rc = alloc_array(c, w);
    if (rc) return rc;
  }
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  *out = h;
  return MFB_OK;
}

void mfb_destroy(mfb_ctx* h) {
  if (!h) return;
  Context* c = &h->c;
  mfb_comm_destroy(h);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_file_pipe(c);
  for (auto& d : c->datasets) free_ds_device(&d);
  if (c->placement_arena) {  // phi/bv/plane scratch may live inside the arena of the placement search
    c->arr[MFB_PHI] = c->place0[0];
    c->arr[MFB_BV] = c->place0[1];
    c->d_phi_planes = c->place0[2];
    cudaFree(c->placement_arena);
  }
  cudaFree(c->d_phi_planes);
  for (auto& p : c->arr) cudaFree(p);
  cudaFree(c->d_counter);
  cudaFree(c->d_accum);
  cudaFree(c->d_noise_table);
  cudaFree(c->d_norms);
  cudaFree(c->d_val_u); cudaFree(c->d_val_v); cudaFree(c->d_val_r);
  cudaFree(c->d_draws); cudaFree(c->d_lams);
  cudaFree(c->d_version); cudaFree(c->d_probe);
  cudaFreeHost(c->h_accum);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_s2) cudaEventDestroy(c->ev_s2);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->d_stage_vid[b]);
    cudaFree(c->d_stage_code[b]);
    cudaFree(c->d_stage_vhi[b]);
  }
  cudaFree(c->d_dict);
  for (auto e : c->chunk_events) cudaEventDestroy(e);
  delete h;
}

int mfb_set_stream(mfb_ctx* h, void* s) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return MFB_OK;
}

int mfb_sync(mfb_ctx* h) {
  MFB_REQUIRE(h, "ctx is NULL");
  MFB_CUDA(cudaStreamSynchronize(h->c.stream));
  return MFB_OK;
}

int mfb_set_option(mfb_ctx* h, const char* name, int value) {
  MFB_REQUIRE(h && name, "NULL argument");
  Context* c = &h->c;
  if (!strcmp(name, "ctas_per_sm")) {
    MFB_REQUIRE(value >= 0 && value <= 32, "ctas_per_sm out of range");
    c->opt_ctas_per_sm = value;
  } else if (!strcmp(name, "threads")) {
    MFB_REQUIRE(value == 32 || value == 64 || value == 128 || value == 256, "threads must be 32/64/128/256");
    c->opt_threads = value;
  } else if (!strcmp(name, "row_concurrency")) {
    MFB_REQUIRE(value >= 0, "row_concurrency must be >= 0");
    c->opt_row_concurrency = value;
  } else if (!strcmp(name, "run_fraction_ppm")) {
    MFB_REQUIRE(value >= 0, "run_fraction_ppm must be >= 0");
    c->opt_run_fraction_ppm = value;
  } else if (!strcmp(name, "run_bound_epochs")) {
    MFB_REQUIRE(value >= 0, "run_bound_epochs must be >= 0");
    c->opt_run_bound_epochs = value;
  } else if (!strcmp(name, "model_age")) {
    MFB_REQUIRE(value >= 0, "model_age must be >= 0");
    c->model_age = value;
  } else if (!strcmp(name, "max_groups")) {
    MFB_REQUIRE(value >= 0, "max_groups must be >= 0");
    c->opt_max_groups = value;
  } else if (!strcmp(name, "kernel")) {
    MFB_REQUIRE(value == 0 || value == 1 || value == 3 || value == 4, "kernel must be 0 (choose), 1 (generic), 3 (stream) or 4 (burst)");
    c->opt_kernel = value;
  } else if (!strcmp(name, "batch")) {
    MFB_REQUIRE(value == 4 || value == 8, "batch must be 4 or 8");
    c->opt_batch = value;
  } else if (!strcmp(name, "phi_planes")) {
    MFB_REQUIRE(value >= 0 && value <= 2, "phi_planes must be 0 (rows), 1 (128-byte planes) or 2 (experiment: sector addressing)");
    c->opt_phi_planes = value;
  } else if (!strcmp(name, "two_streams")) {
    c->opt_two_streams = value != 0;
  } else if (!strcmp(name, "ring_peer")) {
    c->opt_ring_peer = value != 0;
  } else if (!strcmp(name, "sgld_flat")) {
    c->opt_sgld_flat = value;
  } else if (!strcmp(name, "file_decode")) {
    MFB_REQUIRE(value == 0 || value == 1, "file_decode must be 0 (host cores) or 1 (device)");
    c->opt_file_decode = value;
  } else if (!strcmp(name, "epoch_launches")) {
    MFB_REQUIRE(value >= 1 && value <= 4096, "epoch_launches out of range");
    c->opt_epoch_launches = value;
  } else if (!strcmp(name, "span_runs")) {
    MFB_REQUIRE(value >= 0 && value <= 32, "span_runs must be 0..32");
    c->opt_span_runs = value;
  } else if (!strcmp(name, "tail_runs")) {
    MFB_REQUIRE(value >= 0 && value <= 1024, "tail_runs out of range");
    c->opt_tail_runs = value;
  } else if (!strcmp(name, "depth")) {
    MFB_REQUIRE(value >= 0 && value <= 2, "depth must be 0 (choose), 1 or 2");
    c->opt_depth = value;
  } else if (!strcmp(name, "ring")) {
    MFB_REQUIRE(value >= 0 && value <= 4, "ring must be 0..4");
    c->opt_ring = value;
  } else if (!strcmp(name, "packed_h2d")) {
    c->opt_packed_h2d = value != 0;
  } else if (!strcmp(name, "throttle")) {
    c->opt_throttle = value != 0;
  } else if (!strcmp(name, "eta_scaling")) {
    MFB_REQUIRE(value >= 0 && value <= 2, "eta_scaling must be 0, 1 (row budget) or 2 (row and run budgets)");
    c->opt_eta_scaling = value;
  } else if (!strcmp(name, "placement_trials")) {
    MFB_REQUIRE(value >= 0 && value <= 64, "placement_trials must be 0..64");
    c->opt_placement_trials = value;
  } else if (!strcmp(name, "placement_min_ratings")) {
    MFB_REQUIRE(value >= 0, "placement_min_ratings must be >= 0");
    c->placement_min_ratings = value;
  } else if (!strcmp(name, "admf_weight")) {
    MFB_REQUIRE(value >= 0 && value <= 64, "admf_weight must be 0 (rows ahead + 2) or 1..64");
    c->opt_admf_weight = value;
  } else if (!strcmp(name, "admf_prefetch")) {
    MFB_REQUIRE(value >= 0 && value <= 3, "admf_prefetch must be 0..3 rows ahead");
    c->opt_admf_prefetch = value;
  } else if (!strcmp(name, "memopt")) {
    c->opt_memopt = value;
  } else {
    set_error("unknown option %s", name);
    return MFB_E_ARG;
  }
  return MFB_OK;
}

int mfb_enable(mfb_ctx* h, int group) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  if (group == 1) {
    for (int w : {MFB_THETA_OLD, MFB_PHI_OLD, MFB_BU_OLD, MFB_BV_OLD}) {
      int rc = alloc_array(c, w);
      if (rc) return rc;
    }
  } else if (group == 2) {
    for (int w : {MFB_UR, MFB_VR, MFB_LAMBDA_U, MFB_LAMBDA_V}) {
      int rc = alloc_array(c, w);
      if (rc) return rc;
    }
  } else {
    set_error("unknown array group %d", group);
    return MFB_E_ARG;
  }
  return MFB_OK;
}

static int xfer(mfb_ctx* h, int which, float* host, int64_t row0, int64_t nrows, int64_t hs,
                bool up) {
  MFB_REQUIRE(h && host, "NULL argument");
  Context* c = &h->c;
  MFB_REQUIRE(which >= 0 && which < 12 && c->arr[which], "array %d not allocated", which);
  const int64_t rows = array_rows(c, which);
  const int cols = array_cols(c, which), ds = array_stride(c, which);
  MFB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= rows, "rows [%lld,+%lld) outside 0..%lld",
              (long long)row0, (long long)nrows, (long long)rows);
  MFB_REQUIRE(hs >= cols, "host stride %lld < %d columns", (long long)hs, cols);
  if (nrows == 0) return MFB_OK;
  MFB_CUDA(cudaSetDevice(c->device));
  float* dev = c->arr[which] + row0 * ds;
  if (up && (which == MFB_THETA || which == MFB_PHI)) c->model_age = 0;  // a new model: young until told otherwise
  if (up)
    MFB_CUDA(cudaMemcpy2DAsync(dev, ds * sizeof(float), host, hs * sizeof(float),
                               cols * sizeof(float), nrows, cudaMemcpyHostToDevice, c->stream));
  else
    MFB_CUDA(cudaMemcpy2DAsync(host, hs * sizeof(float), dev, ds * sizeof(float),
                               cols * sizeof(float), nrows, cudaMemcpyDeviceToHost, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  return MFB_OK;
}

int mfb_upload(mfb_ctx* h, int which, const float* host, int64_t row0, int64_t nrows, int64_t hs) {
  return xfer(h, which, const_cast<float*>(host), row0, nrows, hs, true);
}
int mfb_download(mfb_ctx* h, int which, float* host, int64_t row0, int64_t nrows, int64_t hs) {
  return xfer(h, which, host, row0, nrows, hs, false);
}

int mfb_init_normal(mfb_ctx* h, uint64_t seed, float scale) {
  MFB_REQUIRE(h, "ctx is NULL");
  MFB_CUDA(cudaSetDevice(h->c.device));
  h->c.model_age = 0;
  return launch_fill_normal(&h->c, seed, scale);
}

int mfb_snapshot_old(mfb_ctx* h) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  MFB_REQUIRE(c->arr[MFB_THETA_OLD], "admf shadows not enabled");
  const int src[4] = {MFB_THETA, MFB_PHI, MFB_BU, MFB_BV};
  const int dst[4] = {MFB_THETA_OLD, MFB_PHI_OLD, MFB_BU_OLD, MFB_BV_OLD};
  for (int i = 0; i < 4; i++) {
    size_t n = (size_t)array_rows(c, src[i]) * array_stride(c, src[i]) * sizeof(float);
    MFB_CUDA(cudaMemcpyAsync(c->arr[dst[i]], c->arr[src[i]], n, cudaMemcpyDeviceToDevice, c->stream));
  }
  return MFB_OK;
}

void* mfb_device_ptr(mfb_ctx* h, int which) {
  if (!h || which < 0 || which >= 12) return nullptr;
  return h->c.arr[which];
}

// ---- datasets ---------------------------------------------------------------------------------
int mfb_dataset_create(mfb_ctx* h, int* ds) {
  MFB_REQUIRE(h && ds, "NULL argument");
  Context* c = &h->c;
  for (size_t i = 0; i < c->datasets.size(); i++)
    if (!c->datasets[i].used) {
      c->datasets[i] = Dataset();
      c->datasets[i].used = true;
      *ds = (int)i;
      return MFB_OK;
    }
  c->datasets.emplace_back();
  c->datasets.back().used = true;
  *ds = (int)c->datasets.size() - 1;
  return MFB_OK;
}

int mfb_dataset_append_block(mfb_ctx* h, int ds, int32_t nusers, const int32_t* uid,
                             const int32_t* rec_off, const int32_t* vid, const float* rating) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(!d->finalized, "dataset %d already finalized", ds);
  MFB_REQUIRE(nusers >= 0 && (nusers == 0 || (uid && rec_off)), "bad block");
  if (d->h_run_off.empty()) d->h_run_off.push_back(0);
  if (d->h_block_off.empty()) d->h_block_off.push_back(0);
  const int64_t base = (int64_t)d->h_vid.size();
  const int64_t n = nusers ? rec_off[nusers] - rec_off[0] : 0;
  MFB_REQUIRE(n >= 0 && base + n < (int64_t)INT32_MAX, "dataset too large for int32 offsets");
  MFB_REQUIRE(n == 0 || (vid && rating), "NULL records");
  for (int32_t i = 0; i < nusers; i++) {
    MFB_REQUIRE(uid[i] >= 0 && uid[i] < c->nu, "uid %d outside [0,%d)", uid[i], c->nu);
    MFB_REQUIRE(rec_off[i + 1] >= rec_off[i], "rec_off not monotone");
    d->h_run_uid.push_back(uid[i]);
    d->h_run_off.push_back((int32_t)(base + rec_off[i + 1] - rec_off[0]));
  }
  const int32_t* v0 = vid + (nusers ? rec_off[0] : 0);
  for (int64_t k = 0; k < n; k++)
    MFB_REQUIRE(v0[k] >= 0 && v0[k] < c->nv, "vid %d outside [0,%d)", v0[k], c->nv);
  d->h_vid.insert(d->h_vid.end(), v0, v0 + n);
  const float* r0 = rating + (nusers ? rec_off[0] : 0);
  d->h_rating.insert(d->h_rating.end(), r0, r0 + n);
  d->h_block_off.push_back((int64_t)d->h_run_uid.size());
  return MFB_OK;
}

int mfb_dataset_append_blocks(mfb_ctx* h, int ds, const mfb_blocks* b) {
  MFB_REQUIRE(h && b, "NULL argument");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(!d->finalized, "dataset %d already finalized", ds);
  return append_host(c, d, *blocks_data(b));
}

int mfb_dataset_load_file(mfb_ctx* h, int ds, const char* path) {
  MFB_REQUIRE(h && path, "NULL argument");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(!d->finalized, "dataset %d already finalized", ds);
  int rc = load_blocks_file(path, d);
  if (rc) return rc;
  int32_t badv = 0;
  MFB_REQUIRE(find_out_of_range(d->h_run_uid.data(), (int64_t)d->h_run_uid.size(), c->nu, &badv) < 0,
              "uid %d outside [0,%d) in %s", badv, c->nu, path);
  MFB_REQUIRE(find_out_of_range(d->h_vid.data(), (int64_t)d->h_vid.size(), c->nv, &badv) < 0,
              "vid %d outside [0,%d) in %s", badv, c->nv, path);
  return MFB_OK;
}

int mfb_dataset_finalize(mfb_ctx* h, int ds) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(!d->finalized, "dataset %d already finalized", ds);
  MFB_CUDA(cudaSetDevice(c->device));
  if (d->h_run_off.empty()) d->h_run_off.push_back(0);
  if (d->h_block_off.empty()) d->h_block_off.push_back(0);
  d->nruns = (int64_t)d->h_run_uid.size();
  d->nratings = (int64_t)d->h_vid.size();
  d->nblocks = (int64_t)d->h_block_off.size() - 1;
  int rc;
  if ((rc = to_device(c, d->h_run_uid, &d->d_run_uid))) return rc;
  if ((rc = to_device(c, d->h_run_off, &d->d_run_off))) return rc;
  if ((rc = to_device(c, d->h_vid, &d->d_vid))) return rc;
  if ((rc = to_device(c, d->h_rating, &d->d_rating))) return rc;
  {
    // records of the most rated item: per-thread histograms over the records, merged
    const int64_t hw = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::vector<int32_t>> part((size_t)hw);
    parallel_ranges(d->nratings, [&](int t, int64_t b, int64_t e) {
      std::vector<int32_t>& cnt = part[t];
      cnt.assign(c->nv, 0);
      const int32_t* v = d->h_vid.data();
      for (int64_t i = b; i < e; i++) cnt[v[i]]++;
    });
    int64_t top = 0;
    for (int32_t v = 0; v < c->nv; v++) {
      int64_t sum = 0;
      for (auto& cnt : part)
        if (!cnt.empty()) sum += cnt[v];
      top = std::max(top, sum);
    }
    d->max_item_share = d->nratings ? (double)top / (double)d->nratings : 0.0;
  }
  if (c->arr[MFB_UR]) {  // dpmf enabled: static logical clock + per-row record counts
    std::vector<int32_t> last_u(c->nu, 0), last_v(c->nv, 0), run_uc(d->nruns, 0), vc(d->nratings, 0);
    d->h_ucount.assign(c->nu, 0);
    d->h_vcount.assign(c->nv, 0);
    // users: one step per run (runs are few)
    for (int64_t r = 0; r < d->nruns; r++) {
      const int32_t u = d->h_run_uid[r];
      const int32_t lo = d->h_run_off[r], hi = d->h_run_off[r + 1];
      if (lo == hi) continue;
      run_uc[r] = lo - last_u[u];          // dpmf.h:65: uc = gc - gcountu[uid], gc == record index
      last_u[u] = hi - 1;                  // dpmf.h:66
      d->h_ucount[u] += hi - lo;
    }
    // items: vc[t] = t - (previous record of the same item), dpmf.h:63.  The records are cut into one range per
    // thread; a range resolves every record but the first of each item inside itself, and a serial sweep over the
    // ranges (nv steps each) hands the first ones the last occurrence in the ranges before.
    const int64_t n = d->nratings;
    const int64_t hw = std::max(1u, std::thread::hardware_concurrency());
    const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(hw, n / (1 << 16)));
    std::vector<std::vector<int32_t>> first((size_t)parts), last((size_t)parts), count((size_t)parts);
    const int32_t* v = d->h_vid.data();
    parallel_ranges(n, [&](int t, int64_t b, int64_t e) {
      std::vector<int32_t>&f = first[t], &l = last[t], &k = count[t];
      f.assign(c->nv, -1);
      l.assign(c->nv, -1);
      k.assign(c->nv, 0);
      for (int64_t i = b; i < e; i++) {
        const int32_t it = v[i];
        if (l[it] >= 0) vc[i] = (int32_t)(i - l[it]);
        else f[it] = (int32_t)i;
        l[it] = (int32_t)i;
        k[it]++;
      }
    });
    for (int64_t t = 0; t < parts; t++) {
      if (first[t].empty()) continue;  // (parallel_ranges may use fewer parts than asked for)
      for (int32_t it = 0; it < c->nv; it++) {
        if (first[t][it] >= 0) vc[first[t][it]] = first[t][it] - last_v[it];  // (gcountv starts at 0, model.cc:235/326)
        if (last[t][it] >= 0) last_v[it] = last[t][it];
        d->h_vcount[it] += count[t][it];
      }
    }
    if ((rc = to_device(c, run_uc, &d->d_uc))) return rc;
    if ((rc = to_device(c, vc, &d->d_vc))) return rc;
    if ((rc = to_device(c, last_u, &d->d_last_u))) return rc;
    if ((rc = to_device(c, last_v, &d->d_last_v))) return rc;
    MFB_CUDA(cudaStreamSynchronize(c->stream));  // the vectors above die at the end of this scope
  }
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  // the tiles are resident in HBM; drop the host staging copies (keep the small run tables)
  release(d->h_vid);
  release(d->h_rating);
  d->finalized = true;
  return MFB_OK;
}

int mfb_dataset_free(mfb_ctx* h, int ds) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  cudaStreamSynchronize(c->stream);
  free_ds_device(d);
  *d = Dataset();
  return MFB_OK;
}

int64_t mfb_dataset_num_ratings(mfb_ctx* h, int ds) {
  Dataset* d = h ? get_ds(&h->c, ds) : nullptr;
  return d ? (d->finalized ? d->nratings : (int64_t)d->h_vid.size()) : -1;
}
int64_t mfb_dataset_num_blocks(mfb_ctx* h, int ds) {
  Dataset* d = h ? get_ds(&h->c, ds) : nullptr;
  return d ? (int64_t)d->h_block_off.size() - 1 : -1;
}
int64_t mfb_dataset_num_runs(mfb_ctx* h, int ds) {
  Dataset* d = h ? get_ds(&h->c, ds) : nullptr;
  return d ? (d->finalized ? d->nruns : (int64_t)d->h_run_uid.size()) : -1;
}

// two device staging buffers for packed (3-byte) records on their way to the SoA tiles
static int ensure_stage(Context* c, int64_t capacity) {
  if (c->stage_capacity >= capacity) return MFB_OK;
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->copy_stream) MFB_CUDA(cudaStreamSynchronize(c->copy_stream));
  c->stage_capacity = capacity;
  for (int b = 0; b < 2; b++) {
    cudaFree(c->d_stage_vid[b]);
    cudaFree(c->d_stage_code[b]);
    cudaFree(c->d_stage_vhi[b]);
    MFB_CUDA(cudaMalloc(&c->d_stage_vid[b], c->stage_capacity * sizeof(uint16_t)));
    MFB_CUDA(cudaMalloc(&c->d_stage_code[b], c->stage_capacity));
    MFB_CUDA(cudaMalloc(&c->d_stage_vhi[b], c->stage_capacity));
  }
  if (!c->d_dict) MFB_CUDA(cudaMalloc(&c->d_dict, 256 * sizeof(float)));
  return MFB_OK;
}

// ---- hot path -----------------------------------------------------------------------------------
int mfb_sgd_epoch(mfb_ctx* h, int ds, float eta, float lambda, float gb, int mode) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC,
              "bad mode %d", mode);
  MFB_CUDA(cudaSetDevice(c->device));
  if (int trc = tune_placement(c, &d, 1, gb, mode, c->opt_phi_planes == 1 && c->opt_epoch_launches <= 1 && c->stride == 128)) return trc;
  begin_timing(c);
  int rc = MFB_OK;
  const int64_t parts = std::max(1, c->opt_epoch_launches);  // diagnostic: the epoch as several launches
  c->planes_allowed = true;  // this launch has the item matrix to itself: the stream kernel may transpose it
  for (int64_t k = 0; k < parts && rc == MFB_OK && d->nruns; k++) {
    const int64_t r0 = d->nruns * k / parts, r1 = d->nruns * (k + 1) / parts;
    if (r1 > r0) rc = launch_sgd(c, d, eta, lambda, gb, mode, r0, r1);
  }
  c->planes_allowed = false;
  end_timing(c);
  if (rc == MFB_OK) c->model_age++;
  return rc;
}

// SgdFilter::operator() over Blocks [block_begin, block_end) of the file only (mf.h:76 is called once per
// Block): a slice of an epoch.  The concurrency bounds are those of the whole file.
int mfb_sgd_epoch_blocks(mfb_ctx* h, int ds, int64_t block_begin, int64_t block_end, float eta, float lambda,
                         float gb, int mode) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC, "bad mode %d", mode);
  MFB_REQUIRE(block_begin >= 0 && block_begin <= block_end && block_end <= d->nblocks, "blocks [%lld, %lld) outside 0..%lld",
              (long long)block_begin, (long long)block_end, (long long)d->nblocks);
  MFB_CUDA(cudaSetDevice(c->device));
  begin_timing(c);
  const int64_t r0 = d->h_block_off[block_begin], r1 = d->h_block_off[block_end];
  c->planes_allowed = true;  // (this launch has the item matrix to itself)
  int rc = r1 > r0 ? launch_sgd(c, d, eta, lambda, gb, mode, r0, r1) : MFB_OK;
  c->planes_allowed = false;
  end_timing(c);
  return rc;
}

// One epoch whose inputs start in HOST memory: the rating tiles of `src` are copied H2D chunk
// by chunk on a second stream while the update kernel works on the previous chunk - the
// reference's read -> parse -> update pipeline (main.cc:45-50) with PCIe as the "read" stage.
int mfb_sgd_epoch_from_host(mfb_ctx* h, int ds, const mfb_blocks* src, float eta, float lambda,
                            float gb, int mode, int64_t chunk_ratings) {
  MFB_REQUIRE(h && src, "NULL argument");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  const Dataset* s = blocks_data(src);
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_REQUIRE((int64_t)s->h_run_uid.size() == d->nruns && (int64_t)s->h_vid.size() == d->nratings,
              "host blocks (%zu runs, %zu ratings) do not match dataset %d (%lld, %lld)",
              s->h_run_uid.size(), s->h_vid.size(), ds, (long long)d->nruns, (long long)d->nratings);
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ATOMIC, "streamed epochs are Hogwild/atomic only");
  MFB_CUDA(cudaSetDevice(c->device));
  if (!c->copy_stream) MFB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!d->refresh_pending) {  // (the tiles are resident from finalize: the calibration can use them)
    if (int trc = tune_placement(c, &d, 1, gb, mode, false)) return trc;
  }
  if (chunk_ratings <= 0) chunk_ratings = 3 << 20;
  // Chunks grow geometrically by 7/4 up to 6x the given size: the first kernel starts after a small
  // copy and every later copy is shorter than the kernels it hides behind (measured while kernels
  // run: H2D ~35 GB/s = 11.7 G packed records/s against 6.5 G updates/s).
  // Each launch ends with a tail - the longest user-run still in flight, ~0.9 ms at this shape
  // (tools/exp_launches.py) - so consecutive chunks go to TWO compute streams at half the width
  // each: the tail of one chunk overlaps the body of the next, and the sum of the runs in flight
  // stays within the concurrency bounds.
  // (MFB_CHUNK_CAP / MFB_CHUNK_GROW_PCT: experiment knobs for the two constants)
  const char* env_cap = getenv("MFB_CHUNK_CAP");
  const char* env_grow = getenv("MFB_CHUNK_GROW_PCT");
  const int64_t grow_pct = env_grow ? std::max(100, atoi(env_grow)) : 175;
  const int64_t max_chunk = chunk_ratings * (env_cap ? std::max(1, atoi(env_cap)) : 6);
  begin_timing(c);
  if (!c->stream2) {
    MFB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    MFB_CUDA(cudaEventCreateWithFlags(&c->ev_s2, cudaEventDisableTiming));
  }
  // neither the copy stream nor the second compute stream may run ahead of work queued earlier
  cudaEvent_t start_ev;
  if (c->chunk_events.empty()) {
    cudaEvent_t e;
    MFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->chunk_events.push_back(e);
  }
  start_ev = c->chunk_events[0];
  MFB_CUDA(cudaEventRecord(start_ev, c->stream));
  MFB_CUDA(cudaStreamWaitEvent(c->copy_stream, start_ev, 0));
  MFB_CUDA(cudaStreamWaitEvent(c->stream2, start_ev, 0));
  const bool packed = s->packed && c->opt_packed_h2d;
  if (packed) {
    if (int src = ensure_stage(c, max_chunk + 65536)) return src;
    MFB_CUDA(cudaMemcpyAsync(c->d_dict, s->p_dict, 256 * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  }
  cudaStream_t const main_stream = c->stream;
  // (a file that fits the first chunk is one launch: nothing to overlap, so it keeps the full width)
  const bool two = c->opt_two_streams != 0 && d->nratings > chunk_ratings;
  int rc = MFB_OK;
  int64_t r0 = 0;
  size_t chunk = 0;
  while (r0 < d->nruns && rc == MFB_OK) {
    int64_t r1 = r0;
    const int64_t o0 = s->h_run_off[r0];
    int64_t want = chunk_ratings;
    for (size_t g = 0; g < chunk && want < max_chunk; g++) want = want * grow_pct / 100;
    want = std::min(want, max_chunk);
    {  // last run whose end is within `want` records (run_off is sorted)
      const int32_t* ro = s->h_run_off.data();
      r1 = std::upper_bound(ro + r0 + 1, ro + d->nruns + 1, (int32_t)std::min<int64_t>(o0 + want, INT32_MAX)) - ro - 1;
    }
    if (r1 == r0) r1 = r0 + 1;  // a single run longer than the chunk size
    const int64_t o1 = s->h_run_off[r1];
    const bool staged = packed && o1 - o0 <= c->stage_capacity;
    const int sb = (int)(chunk & 1);
    MFB_CUDA(cudaMemcpyAsync(d->d_run_uid + r0, s->h_run_uid.data() + r0, (r1 - r0) * sizeof(int32_t),
                             cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaMemcpyAsync(d->d_run_off + r0, s->h_run_off.data() + r0, (r1 - r0 + 1) * sizeof(int32_t),
                             cudaMemcpyHostToDevice, c->copy_stream));
    if (staged) {  // 3 bytes per record into a staging buffer, expanded by a kernel on the copy stream
      MFB_CUDA(cudaMemcpyAsync(c->d_stage_vid[sb], s->p_vid + o0, (o1 - o0) * sizeof(uint16_t),
                               cudaMemcpyHostToDevice, c->copy_stream));
      MFB_CUDA(cudaMemcpyAsync(c->d_stage_code[sb], s->p_code + o0, (o1 - o0), cudaMemcpyHostToDevice,
                               c->copy_stream));
      if (s->p_vhi)
        MFB_CUDA(cudaMemcpyAsync(c->d_stage_vhi[sb], s->p_vhi + o0, (o1 - o0), cudaMemcpyHostToDevice, c->copy_stream));
      c->stream = c->copy_stream;  // (stream order frees the staging buffer for the copy after next)
      rc = launch_unpack(c, c->d_stage_vid[sb], s->p_vhi ? c->d_stage_vhi[sb] : nullptr, c->d_stage_code[sb], c->d_dict,
                         d->d_vid + o0, d->d_rating + o0, o1 - o0);
      c->stream = main_stream;
      if (rc) break;
      c->h2d_bytes += (o1 - o0) * (s->p_vhi ? 4 : 3);
    } else {
      MFB_CUDA(cudaMemcpyAsync(d->d_vid + o0, s->h_vid.data() + o0, (o1 - o0) * sizeof(int32_t),
                               cudaMemcpyHostToDevice, c->copy_stream));
      MFB_CUDA(cudaMemcpyAsync(d->d_rating + o0, s->h_rating.data() + o0, (o1 - o0) * sizeof(float),
                               cudaMemcpyHostToDevice, c->copy_stream));
      c->h2d_bytes += (o1 - o0) * 8;
    }
    c->h2d_bytes += (r1 - r0) * 8 + 4;
    ++chunk;
    if (c->chunk_events.size() <= chunk) {
      cudaEvent_t e;
      MFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->chunk_events.push_back(e);
    }
    MFB_CUDA(cudaEventRecord(c->chunk_events[chunk], c->copy_stream));
    // the update kernel of this chunk: streams alternate, each launch gets half the width
    const bool second = two && (chunk & 1) == 0;
    c->stream = second ? c->stream2 : main_stream;
    c->counter_slot = second ? 2 : 0;
    c->width_div = two ? 2 : 1;
    cudaError_t e = cudaStreamWaitEvent(c->stream, c->chunk_events[chunk], 0);
    if (e == cudaSuccess) rc = launch_sgd(c, d, eta, lambda, gb, mode, r0, r1);
    c->stream = main_stream;
    c->counter_slot = 0;
    c->width_div = 1;
    MFB_CUDA(e);
    r0 = r1;
  }
  if (rc) return rc;
  MFB_CUDA(cudaEventRecord(c->ev_s2, c->stream2));
  MFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_s2, 0));
  end_timing(c);
  c->model_age++;
  return MFB_OK;
}

// Re-send the tiles of a finalized dataset from (pinned) host arrays: H2D on the copy stream; the
// next kernel that uses the dataset waits for it.  Lets several datasets (the DSGD cells of one
// epoch) stream in behind the compute of the previous ones.
int mfb_dataset_refresh_from_host(mfb_ctx* h, int ds, const mfb_blocks* src) {
  MFB_REQUIRE(h && src, "NULL argument");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  const Dataset* s = blocks_data(src);
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_REQUIRE((int64_t)s->h_run_uid.size() == d->nruns && (int64_t)s->h_vid.size() == d->nratings,
              "host blocks do not match dataset %d", ds);
  MFB_CUDA(cudaSetDevice(c->device));
  if (!c->copy_stream) MFB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!d->refreshed) MFB_CUDA(cudaEventCreateWithFlags(&d->refreshed, cudaEventDisableTiming));
  // do not overwrite tiles that kernels already queued on the main stream still read
  MFB_CUDA(cudaEventRecord(d->refreshed, c->stream));
  MFB_CUDA(cudaStreamWaitEvent(c->copy_stream, d->refreshed, 0));
  if (d->nruns) {
    MFB_CUDA(cudaMemcpyAsync(d->d_run_uid, s->h_run_uid.data(), d->nruns * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaMemcpyAsync(d->d_run_off, s->h_run_off.data(), (d->nruns + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
  }
  c->h2d_bytes += d->nruns * 8 + 4;
  if (d->nratings && s->packed && c->opt_packed_h2d) {
    // the compact records (u16 item id + u8 rating code, mfb_blocks_pin) into a staging buffer, expanded on the copy
    // stream; stream order frees a staging buffer for the copy after next
    if (int src = ensure_stage(c, d->nratings)) return src;
    const int sb = c->stage_flip;
    c->stage_flip ^= 1;
    MFB_CUDA(cudaMemcpyAsync(c->d_dict, s->p_dict, 256 * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaMemcpyAsync(c->d_stage_vid[sb], s->p_vid, d->nratings * sizeof(uint16_t), cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaMemcpyAsync(c->d_stage_code[sb], s->p_code, d->nratings, cudaMemcpyHostToDevice, c->copy_stream));
    if (s->p_vhi) MFB_CUDA(cudaMemcpyAsync(c->d_stage_vhi[sb], s->p_vhi, d->nratings, cudaMemcpyHostToDevice, c->copy_stream));
    cudaStream_t const main_stream = c->stream;
    c->stream = c->copy_stream;
    const int urc = launch_unpack(c, c->d_stage_vid[sb], s->p_vhi ? c->d_stage_vhi[sb] : nullptr, c->d_stage_code[sb], c->d_dict,
                                  d->d_vid, d->d_rating, d->nratings);
    c->stream = main_stream;
    if (urc) return urc;
    c->h2d_bytes += d->nratings * (s->p_vhi ? 4 : 3);
  } else if (d->nratings) {
    MFB_CUDA(cudaMemcpyAsync(d->d_vid, s->h_vid.data(), d->nratings * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
    MFB_CUDA(cudaMemcpyAsync(d->d_rating, s->h_rating.data(), d->nratings * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    c->h2d_bytes += d->nratings * 8;
  }
  MFB_CUDA(cudaEventRecord(d->refreshed, c->copy_stream));
  d->refresh_pending = true;
  return MFB_OK;
}

int mfb_blocks_pin(mfb_blocks* b) {
  MFB_REQUIRE(b, "NULL argument");
  Dataset* s = blocks_mut(b);
  if (s->pinned) return MFB_OK;
  auto reg = [](const void* p, size_t bytes) {
    return bytes ? cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable) : cudaSuccess;
  };
  MFB_CUDA(reg(s->h_run_uid.data(), s->h_run_uid.size() * sizeof(int32_t)));
  MFB_CUDA(reg(s->h_run_off.data(), s->h_run_off.size() * sizeof(int32_t)));
  MFB_CUDA(reg(s->h_vid.data(), s->h_vid.size() * sizeof(int32_t)));
  MFB_CUDA(reg(s->h_rating.data(), s->h_rating.size() * sizeof(float)));
  s->pinned = true;
  // compact form for streaming: u16 item ids (+ a u8 plane for bits 16..23 when some id needs them) + u8 codes into
  // a table of the distinct rating values: 3 (4) bytes per record instead of 8, lossless
  s->packed = false;
  const size_t n = s->h_vid.size();
  bool fits = n > 0, wide = false;
  for (size_t i = 0; fits && i < n; i++) {
    fits = s->h_vid[i] >= 0 && s->h_vid[i] < (1 << 24);
    wide = wide || s->h_vid[i] >= 65536;
  }
  if (fits) {
    std::vector<uint8_t> code(n);
    int ndict = 0;
    uint32_t last_bits = 0;
    int last_code = -1;
    for (size_t i = 0; fits && i < n; i++) {
      uint32_t bits;
      memcpy(&bits, &s->h_rating[i], 4);
      if (last_code < 0 || bits != last_bits) {
        int k = 0;
        for (; k < ndict; k++) {
          uint32_t db;
          memcpy(&db, &s->p_dict[k], 4);
          if (db == bits) break;
        }
        if (k == ndict) {
          if (ndict == 256) {
            fits = false;
            break;
          }
          s->p_dict[ndict++] = s->h_rating[i];
        }
        last_bits = bits;
        last_code = k;
      }
      code[i] = (uint8_t)last_code;
    }
    if (fits) {
      MFB_CUDA(cudaHostAlloc((void**)&s->p_vid, n * sizeof(uint16_t), cudaHostAllocPortable));
      MFB_CUDA(cudaHostAlloc((void**)&s->p_code, n, cudaHostAllocPortable));
      if (wide) MFB_CUDA(cudaHostAlloc((void**)&s->p_vhi, n, cudaHostAllocPortable));
      for (size_t i = 0; i < n; i++) s->p_vid[i] = (uint16_t)(s->h_vid[i] & 0xffff);
      if (wide)
        for (size_t i = 0; i < n; i++) s->p_vhi[i] = (uint8_t)(s->h_vid[i] >> 16);
      memcpy(s->p_code, code.data(), n);
      s->packed = true;
    }
  }
  return MFB_OK;
}

int mfb_blocks_unpin(mfb_blocks* b) {
  MFB_REQUIRE(b, "NULL argument");
  Dataset* s = blocks_mut(b);
  if (!s->pinned) return MFB_OK;
  if (!s->h_run_uid.empty()) cudaHostUnregister(s->h_run_uid.data());
  if (!s->h_run_off.empty()) cudaHostUnregister(s->h_run_off.data());
  if (!s->h_vid.empty()) cudaHostUnregister(s->h_vid.data());
  if (!s->h_rating.empty()) cudaHostUnregister(s->h_rating.data());
  if (s->packed) {
    cudaFreeHost(s->p_vid);
    cudaFreeHost(s->p_code);
    cudaFreeHost(s->p_vhi);
    s->p_vid = nullptr;
    s->p_code = nullptr;
    s->p_vhi = nullptr;
    s->packed = false;
  }
  s->pinned = false;
  return MFB_OK;
}

int mfb_sse(mfb_ctx* h, int ds, float gb, double* sse, int64_t* n) { return mfb_sse_link(h, ds, gb, 0, sse, n); }

int mfb_sse_link(mfb_ctx* h, int ds, float gb, int link, double* sse, int64_t* n) {
  MFB_REQUIRE(h && sse, "NULL argument");
  MFB_REQUIRE(link == 0 || link == 1, "link must be 0 (identity) or 1 (logistic)");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_CUDA(cudaSetDevice(c->device));
  *sse = 0.0;
  if (n) *n = d->nratings;
  if (d->nruns == 0) return MFB_OK;
  begin_timing(c);
  int rc = launch_sse(c, d, gb, link);
  end_timing(c);
  if (rc) return rc;
  MFB_CUDA(cudaMemcpyAsync(c->h_accum, c->d_accum, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  *sse = c->h_accum[0];
  return MFB_OK;
}

// ---- dpmf ------------------------------------------------------------------------------------
float mfb_dp_bound(float epsilon, int tau, int nv) {  // model.cc:239-242
  if (tau <= 0) tau = nv;
  if (epsilon <= 0.0f) return 1.0f;
  return (float)((double)epsilon * 1.0 / (4.0 * 25.0 * (double)tau));
}

int mfb_dp_weights(mfb_ctx* h, int ds, int32_t* ntrain) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(d->finalized && d->d_vc, "dataset %d was not finalized with dpmf enabled", ds);
  MFB_CUDA(cudaSetDevice(c->device));
  const int32_t n = (int32_t)d->nratings;
  std::vector<float> ur(c->nu), vr(c->nv);
  for (int i = 0; i < c->nu; i++) ur[i] = (float)n / d->h_ucount[i];  // model.cc:294 (inf if unseen)
  for (int i = 0; i < c->nv; i++) vr[i] = (float)n / d->h_vcount[i];  // model.cc:295
  MFB_CUDA(cudaMemcpyAsync(c->arr[MFB_UR], ur.data(), ur.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  MFB_CUDA(cudaMemcpyAsync(c->arr[MFB_VR], vr.data(), vr.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  if (ntrain) *ntrain = n;
  return MFB_OK;
}

static int check_sgld(Context* c, Dataset* d, const mfb_sgld_params* p) {
  MFB_REQUIRE(p, "params are NULL");
  MFB_REQUIRE(c->arr[MFB_UR] && c->arr[MFB_LAMBDA_U], "dpmf arrays not enabled (mfb_enable(ctx, 2))");
  MFB_REQUIRE(d->finalized && d->d_vc, "dataset was not finalized with dpmf enabled");
  MFB_REQUIRE(!p->use_table || c->d_noise_table, "no noise table uploaded");
  return MFB_OK;
}

int mfb_sgld_epoch(mfb_ctx* h, int ds, const mfb_sgld_params* p, float gb, int mode) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  int rc = check_sgld(c, d, p);
  if (rc) return rc;
  MFB_REQUIRE(mode == MFB_MODE_HOGWILD || mode == MFB_MODE_ORDERED, "bad mode %d", mode);
  MFB_REQUIRE(!p->use_table || mode == MFB_MODE_ORDERED, "table noise is an ordered-mode (parity) feature");
  if (p->use_table) {
    int64_t longest = 0;
    for (int64_t r = 0; r < d->nruns; r++) longest = std::max<int64_t>(longest, d->h_run_off[r + 1] - d->h_run_off[r]);
    MFB_REQUIRE(p->table_offset >= 0 && p->table_offset + longest * (c->dim + 1) <= c->noise_table_size,
                "noise table too small for offset %d", p->table_offset);
  }
  MFB_CUDA(cudaSetDevice(c->device));
  begin_timing(c);
  // (dpmf keeps the run bound in every epoch: its parity was established with it)
  const int age = c->model_age;
  c->model_age = 0;
  rc = d->nruns ? launch_sgld(c, d, p, gb, mode) : MFB_OK;
  c->model_age = age + (rc == MFB_OK ? 1 : 0);
  end_timing(c);
  return rc;
}

int mfb_sgld_flush_noise(mfb_ctx* h, int ds, const mfb_sgld_params* p) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  int rc = check_sgld(c, d, p);
  if (rc) return rc;
  MFB_REQUIRE(!p->use_table || p->table_offset + c->dim + 1 <= c->noise_table_size, "noise table too small");
  MFB_CUDA(cudaSetDevice(c->device));
  begin_timing(c);
  rc = launch_flush(c, d, p);
  end_timing(c);
  return rc;
}

int mfb_col_sqnorms(mfb_ctx* h, double* normu, double* normv, double* bu2, double* bv2) {
  MFB_REQUIRE(h && normu && normv && bu2 && bv2, "NULL argument");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  const int w = c->stride + 1;
  if (!c->d_norms) MFB_CUDA(cudaMalloc(&c->d_norms, 2 * w * sizeof(double)));
  begin_timing(c);
  int rc = launch_col_sqnorms(c, c->d_norms);
  end_timing(c);
  if (rc) return rc;
  std::vector<double> hb(2 * w);
  MFB_CUDA(cudaMemcpyAsync(hb.data(), c->d_norms, 2 * w * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->dim; i++) {
    normu[i] = hb[i];
    normv[i] = hb[w + i];
  }
  *bu2 = hb[c->stride];
  *bv2 = hb[w + c->stride];
  return MFB_OK;
}

int mfb_set_noise_table(mfb_ctx* h, const float* host, int64_t n) {
  MFB_REQUIRE(h && host && n > 0, "bad argument");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_noise_table);
  c->d_noise_table = nullptr;
  MFB_CUDA(cudaMalloc(&c->d_noise_table, n * sizeof(float)));
  MFB_CUDA(cudaMemcpy(c->d_noise_table, host, n * sizeof(float), cudaMemcpyHostToDevice));
  c->noise_table_size = n;
  return MFB_OK;
}

// ---- admf ------------------------------------------------------------------------------------
int mfb_admf_set_validation(mfb_ctx* h, int64_t n, const int32_t* u, const int32_t* v, const float* r) {
  MFB_REQUIRE(h && n > 0 && u && v && r, "bad argument");
  Context* c = &h->c;
  for (int64_t i = 0; i < n; i++)
    MFB_REQUIRE(u[i] >= 0 && u[i] < c->nu && v[i] >= 0 && v[i] < c->nv, "validation record %lld out of range", (long long)i);
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->d_val_u); cudaFree(c->d_val_v); cudaFree(c->d_val_r);
  c->d_val_u = c->d_val_v = nullptr;
  c->d_val_r = nullptr;
  MFB_CUDA(cudaMalloc(&c->d_val_u, n * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&c->d_val_v, n * sizeof(int32_t)));
  MFB_CUDA(cudaMalloc(&c->d_val_r, n * sizeof(float)));
  MFB_CUDA(cudaMemcpy(c->d_val_u, u, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  MFB_CUDA(cudaMemcpy(c->d_val_v, v, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  MFB_CUDA(cudaMemcpy(c->d_val_r, r, n * sizeof(float), cudaMemcpyHostToDevice));
  c->nvalid = n;
  return MFB_OK;
}

int mfb_admf_set_draws(mfb_ctx* h, int64_t n, const int32_t* draws) {
  MFB_REQUIRE(h && n >= 0 && (n == 0 || draws), "bad argument");
  Context* c = &h->c;
  MFB_REQUIRE(c->nvalid > 0, "set the validation list first");
  for (int64_t i = 0; i < n; i++) MFB_REQUIRE(draws[i] >= 0 && draws[i] < c->nvalid, "draw %lld out of range", (long long)i);
  MFB_CUDA(cudaSetDevice(c->device));
  if (n > c->ndraws) {
    MFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_draws);
    c->d_draws = nullptr;
    MFB_CUDA(cudaMalloc(&c->d_draws, std::max<int64_t>(n, 1) * sizeof(int32_t)));
  }
  c->ndraws = n;
  if (n) MFB_CUDA(cudaMemcpyAsync(c->d_draws, draws, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  return MFB_OK;
}

static int ensure_lams(Context* c) {
  if (!c->d_lams) {
    MFB_CUDA(cudaMalloc(&c->d_lams, 4 * sizeof(float)));
    MFB_CUDA(cudaMemsetAsync(c->d_lams, 0, 4 * sizeof(float), c->stream));
  }
  return MFB_OK;
}

int mfb_admf_set_lams(mfb_ctx* h, const float lams[4]) {
  MFB_REQUIRE(h && lams, "NULL argument");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  int rc = ensure_lams(c);
  if (rc) return rc;
  MFB_CUDA(cudaMemcpyAsync(c->d_lams, lams, 4 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  return MFB_OK;
}

int mfb_admf_get_lams(mfb_ctx* h, float lams[4]) {
  MFB_REQUIRE(h && lams, "NULL argument");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  int rc = ensure_lams(c);
  if (rc) return rc;
  MFB_CUDA(cudaMemcpyAsync(lams, c->d_lams, 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 4; k++) lams[k] = lams[k] < 0.f ? 0.f : lams[k];  // clamped view (model.h:94)
  return MFB_OK;
}

int mfb_admf_epoch(mfb_ctx* h, int ds, float eta, float eta_reg, int loss, float gb, int mode) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  Dataset* d = get_ds(c, ds);
  if (!d) return MFB_E_ARG;
  MFB_REQUIRE(d->finalized, "dataset %d not finalized", ds);
  MFB_REQUIRE(c->arr[MFB_THETA_OLD], "admf shadows not enabled (mfb_enable(ctx, 1))");
  MFB_REQUIRE(c->nvalid > 0 && c->d_lams, "validation list / regularisers not set");
  MFB_REQUIRE(c->ndraws >= d->nruns, "need one validation draw per user-run (%lld < %lld)",
              (long long)c->ndraws, (long long)d->nruns);
  MFB_REQUIRE(mode == MFB_MODE_ORDERED || mode == MFB_MODE_ATOMIC || mode == MFB_MODE_HOGWILD, "bad mode %d", mode);
  MFB_REQUIRE(loss == 0 || loss == 1, "loss must be 0 (least squares) or 1 (logistic)");
  MFB_CUDA(cudaSetDevice(c->device));
  begin_timing(c);
  // admf keeps the run bound in every epoch: the regularisers are learned from the same stale rows, and with the
  // bound lifted after epoch 1 the ML-1M-shaped run diverges (measured, round 2)
  const int age = c->model_age;
  c->model_age = 0;
  int rc = d->nruns ? launch_admf(c, d, eta, eta_reg, loss, gb, mode) : MFB_OK;
  c->model_age = age + (rc == MFB_OK ? 1 : 0);
  end_timing(c);
  return rc;
}

// Staleness probe: item >= 0 arms it (per-item update counters, zeroed), item < 0 disarms it.
int mfb_probe_arm(mfb_ctx* h, int item) {
  MFB_REQUIRE(h, "ctx is NULL");
  Context* c = &h->c;
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  if (item < 0) {
    cudaFree(c->d_version);
    cudaFree(c->d_probe);
    c->d_version = nullptr;
    c->d_probe = nullptr;
    return MFB_OK;
  }
  MFB_REQUIRE(item < c->nv, "probe item out of range");
  if (!c->d_version) {
    MFB_CUDA(cudaMalloc(&c->d_version, (size_t)c->nv * sizeof(int)));
    MFB_CUDA(cudaMalloc(&c->d_probe, 4 * sizeof(unsigned long long)));
  }
  MFB_CUDA(cudaMemset(c->d_version, 0, (size_t)c->nv * sizeof(int)));
  MFB_CUDA(cudaMemset(c->d_probe, 0, 4 * sizeof(unsigned long long)));
  c->probe_item = item;
  return MFB_OK;
}

// out[0..3] = {stale updates summed over all updates, updates, the same two for the probed item};
// counters restart from zero
int mfb_probe_read(mfb_ctx* h, uint64_t out[4]) {
  MFB_REQUIRE(h && out, "NULL argument");
  Context* c = &h->c;
  MFB_REQUIRE(c->d_probe, "probe not armed");
  MFB_CUDA(cudaSetDevice(c->device));
  MFB_CUDA(cudaStreamSynchronize(c->stream));
  MFB_CUDA(cudaMemcpy(out, c->d_probe, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  MFB_CUDA(cudaMemset(c->d_probe, 0, 4 * sizeof(unsigned long long)));
  return MFB_OK;
}

float mfb_last_kernel_ms(mfb_ctx* h) {
  if (!h || !h->c.timed) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(h->c.ev1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, h->c.ev0, h->c.ev1) != cudaSuccess) return -1.f;
  return ms;
}

int64_t mfb_launch_count(mfb_ctx* h) { return h ? h->c.launches : 0; }
int64_t mfb_h2d_bytes(mfb_ctx* h) { return h ? h->c.h2d_bytes : 0; }

int mfb_placement_report(mfb_ctx* h, int which, float* ms, int n, int* best) {
  MFB_REQUIRE(h && (which == 0 || which == 1), "bad argument");
  const Context* c = &h->c;
  for (int i = 0; i < c->placement_tried[which] && i < n; i++) ms[i] = c->placement_ms[which][i];
  if (best) *best = c->placement_best[which];
  return std::min(n, c->placement_tried[which]);
}

int mfb_last_launch(mfb_ctx* h, int out[4]) {
  MFB_REQUIRE(h && out, "NULL argument");
  out[0] = h->c.use_kernel;
  out[1] = h->c.last_grid;
  out[2] = h->c.last_threads;
  out[3] = h->c.last_ring;
  return MFB_OK;
}

}  // extern "C"
