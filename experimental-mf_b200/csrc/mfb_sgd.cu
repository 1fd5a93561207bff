// K1 sgd_epoch, K5 sse_reduce, K8 philox_normal_fill   - hand-written sm_100a CUDA.
//
// sgd_epoch replaces SgdFilter::operator() (mf.h:76-132).  Per rating, with old values on the
// right-hand side:
//     e      = eta * (r - <theta_u, phi_v> - bu[u] - bv[v] - gb)
//     theta' = lameta*theta + e*phi        phi' = lameta*phi + e*theta
//     bu'    = lameta*bu + e               bv'  = lameta*bv + e         lameta = 1 - eta*lambda
// Mapping: one group of LPR lanes owns a user-run (all records of one user inside one block of the
// file, mf.h:83-88).  theta_u and bu[u] stay in registers for the whole run and are written once;
// phi rows are gathered/scattered as coalesced 128-bit accesses that are served by the L2 when
// the item matrix fits (Netflix shape: 9.1 MB of 126 MB).  Runs are pulled in file order from a
// device-side queue, so execution is Hogwild over item rows exactly like the reference with
// --fly N, only wider.  This path is gather/scatter bound (about 0.5 flop/B): no tensor cores.
#include <algorithm>

#include "mfb_group.cuh"
#include "mfb_internal.h"
#include "mfb_philox.cuh"
#include "mfb_sgd_args.cuh"

namespace mfb {


// One rating, fast arithmetic (fused multiply-adds, butterfly dot).
template <int LPR, int VPL, int MODE>
__device__ __forceinline__ void sgd_update_fast(const SgdArgs& a, Row<VPL>& t, float& bu,
                                                const Row<VPL>& f, float bvv, int v, float r,
                                                int gl, unsigned m) {
  const float d = group_dot<LPR, VPL>(t, f, m);
  const float e = a.eta * (r - d - bu - bvv - a.gb);
  Row<VPL> nf;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const float4 tt = t.v[i], ff = f.v[i];
    if (MODE == MFB_MODE_ATOMIC) {  // increment of phi: (lameta-1)*phi + e*theta
      nf.v[i] = make_float4(fmaf(e, tt.x, a.lm1 * ff.x), fmaf(e, tt.y, a.lm1 * ff.y),
                            fmaf(e, tt.z, a.lm1 * ff.z), fmaf(e, tt.w, a.lm1 * ff.w));
    } else {
      nf.v[i] = make_float4(fmaf(e, tt.x, a.lameta * ff.x), fmaf(e, tt.y, a.lameta * ff.y),
                            fmaf(e, tt.z, a.lameta * ff.z), fmaf(e, tt.w, a.lameta * ff.w));
    }
    t.v[i] = make_float4(fmaf(e, ff.x, a.lameta * tt.x), fmaf(e, ff.y, a.lameta * tt.y),
                         fmaf(e, ff.z, a.lameta * tt.z), fmaf(e, ff.w, a.lameta * tt.w));
  }
  if (MODE == MFB_MODE_ATOMIC) {
    red_add_row<LPR, VPL>(a.phi, v, a.nvec, gl, nf);
    if (gl == 0) atomicAdd(a.bv + v, fmaf(a.lm1, bvv, e));
  } else {
    store_row_f<LPR, VPL>(a.phi, v, a.nvec, gl, nf, a.st_flavour);
    // The item bias is a 4-byte word in a line shared by 32 items: plain stores to it serialise
    // in L2 (measured: they alone doubled the epoch time), so it is updated with a reduction.
    if (gl == 0) {
      if (a.bias_flavour == 0) atomicAdd(a.bv + v, fmaf(a.lm1, bvv, e));
      else if (a.bias_flavour == 2) __stcg(a.bv + v, fmaf(a.lameta, bvv, e));
    }
  }
  bu = fmaf(a.lameta, bu, e);
}

// One rating in the oracle's operation order (mf.h:94-109 as restated in oracle/mf_oracle.c):
// every multiply and add rounded separately, dot accumulated in coordinate order.
template <int LPR, int VPL>
__device__ __forceinline__ void sgd_update_exact(const SgdArgs& a, Row<VPL>& t, float& bu,
                                                 const Row<VPL>& f, float bvv, int v, float r,
                                                 int gl, unsigned m) {
  const float d = group_dot_ordered<LPR, VPL>(t, f, gl, m);
  float e = __fsub_rn(__fsub_rn(__fsub_rn(__fsub_rn(r, d), bu), bvv), a.gb);  // mf.h:99-101
  e = __fmul_rn(a.eta, e);                                                     // mf.h:102
  Row<VPL> nf;
#define MFB_EXACT1(T, F)                                                                   \
  {                                                                                        \
    const float q = __fmul_rn(e, T);                          /* mf.h:103 q = e*theta   */ \
    float th = __fadd_rn(T, __fmul_rn(a.lm1, T));             /* mf.h:104               */ \
    th = __fadd_rn(th, __fmul_rn(e, F));                      /* mf.h:105               */ \
    F = __fadd_rn(q, __fmul_rn(a.lameta, F));                 /* mf.h:106-107           */ \
    T = th;                                                                                \
  }
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    float4 tt = t.v[i], ff = f.v[i];
    MFB_EXACT1(tt.x, ff.x) MFB_EXACT1(tt.y, ff.y) MFB_EXACT1(tt.z, ff.z) MFB_EXACT1(tt.w, ff.w)
    t.v[i] = tt;
    nf.v[i] = ff;
  }
#undef MFB_EXACT1
  store_row<LPR, VPL>(a.phi, v, a.nvec, gl, nf);
  if (gl == 0) __stcg(a.bv + v, __fadd_rn(__fmul_rn(a.lameta, bvv), e));  // mf.h:109
  bu = __fadd_rn(__fmul_rn(a.lameta, bu), e);                             // mf.h:108
}

template <int LPR, int VPL, int MODE, int B>
__global__ void __launch_bounds__(256) sgd_epoch_kernel(const SgdArgs a) {
  constexpr bool ORDERED = (MODE == MFB_MODE_ORDERED);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned m = group_mask<LPR>();
  if (ORDERED && (blockIdx.x != 0 || threadIdx.x >= LPR)) return;  // a single group walks the file
  int next = a.run_begin;
  for (;;) {
    int run;
    if (ORDERED) {
      run = next++;
    } else {
      if (gl == 0) run = a.run_begin + atomicAdd(a.counter, 1);
      run = __shfl_sync(m, run, 0, LPR);
    }
    if (run >= a.nruns) break;
    const int uid = __ldg(a.run_uid + run);
    const int lo = __ldg(a.run_off + run), hi = __ldg(a.run_off + run + 1);
    if (lo == hi) continue;
    Row<VPL> t = load_row<LPR, VPL>(a.theta, uid, a.nvec, gl);
    float bu;
    if (ORDERED) {  // the leader owns the scalar; other lanes must not read memory it wrote
      bu = (gl == 0) ? __ldcg(a.bu + uid) : 0.f;
      bu = __shfl_sync(m, bu, 0, LPR);
    } else {
      bu = __ldcg(a.bu + uid);
    }
    const Row<VPL> t_in = t;
    const float bu_in = bu;
    for (int j0 = lo; j0 < hi; j0 += LPR) {
      const int nb = min(LPR, hi - j0);
      int myvid = 0;
      float myr = 0.f;
      if (gl < nb) {  // one coalesced load of up to LPR records of the run
        myvid = __ldcs(a.vid + j0 + gl);
        myr = __ldcs(a.rating + j0 + gl);
      }
      for (int b0 = 0; b0 < nb; b0 += B) {
        Row<VPL> f[B];
        float bvv[B];
        int v[B];
#pragma unroll
        for (int b = 0; b < B; b++) {  // issue the B row gathers back to back
          v[b] = __shfl_sync(m, myvid, (b0 + b) & (LPR - 1), LPR);
          if (b0 + b < nb) {
            f[b] = ORDERED ? load_row<LPR, VPL>(a.phi, v[b], a.nvec, gl)
                           : load_row_f<LPR, VPL>(a.phi, v[b], a.nvec, gl, a.ld_flavour);
            if (ORDERED) {
              bvv[b] = (gl == 0) ? __ldcg(a.bv + v[b]) : 0.f;
            } else {
              bvv[b] = __ldcg(a.bv + v[b]);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
          const float r = __shfl_sync(m, myr, (b0 + b) & (LPR - 1), LPR);
          if (ORDERED) bvv[b] = __shfl_sync(m, bvv[b], 0, LPR);
          if (b0 + b < nb) {
            if (ORDERED)
              sgd_update_exact<LPR, VPL>(a, t, bu, f[b], bvv[b], v[b], r, gl, m);
            else
              sgd_update_fast<LPR, VPL, MODE>(a, t, bu, f[b], bvv[b], v[b], r, gl, m);
          }
        }
      }
    }
    if (MODE == MFB_MODE_ATOMIC) {
      // the user row receives what this run added to it, as a reduction: a run of the same user in flight in
      // another group at the same time then loses nothing (its row is merely stale, like an item row)
      Row<VPL> dt;
#pragma unroll
      for (int i = 0; i < VPL; i++)
        dt.v[i] = make_float4(t.v[i].x - t_in.v[i].x, t.v[i].y - t_in.v[i].y, t.v[i].z - t_in.v[i].z,
                              t.v[i].w - t_in.v[i].w);
      red_add_row<LPR, VPL>(a.theta, uid, a.nvec, gl, dt);
      if (gl == 0) atomicAdd(a.bu + uid, bu - bu_in);
    } else {
      store_row<LPR, VPL>(a.theta, uid, a.nvec, gl, t);
      if (gl == 0) __stcg(a.bu + uid, bu);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K5: MF::calc_mse (model.cc:41-73).  Same gather, read-only; fp32 error in the reference's
// operation order, squared errors accumulated in fp64 (the reference accumulates per-block fp32
// partials under a mutex in nondeterministic order; fp64 is the order-independent version).
struct SseArgs {
  const float* theta;
  const float* phi;
  const float* bu;
  const float* bv;
  const int32_t* run_uid;
  const int32_t* run_off;
  const int32_t* vid;
  const float* rating;
  double* accum;  // [0] += sum of squared errors
  int nruns, nvec;
  float gb;
  int link;  // 0 identity (MF::calc_mse as written), 1 logistic (util.h:90-95)
};

template <int LPR, int VPL>
__global__ void __launch_bounds__(256) sse_kernel(const SseArgs a) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned m = group_mask<LPR>();
  const int groups_per_cta = blockDim.x / LPR;
  const int g = blockIdx.x * groups_per_cta + threadIdx.x / LPR;
  const int G = gridDim.x * groups_per_cta;
  double acc = 0.0;
  for (int run = g; run < a.nruns; run += G) {
    const int uid = __ldg(a.run_uid + run);
    const int lo = __ldg(a.run_off + run), hi = __ldg(a.run_off + run + 1);
    if (lo == hi) continue;
    const Row<VPL> t = load_row<LPR, VPL>(a.theta, uid, a.nvec, gl);
    const float bu = __ldcg(a.bu + uid);
    for (int j0 = lo; j0 < hi; j0 += LPR) {
      const int nb = min(LPR, hi - j0);
      int myvid = 0;
      float myr = 0.f;
      if (gl < nb) {
        myvid = __ldcs(a.vid + j0 + gl);
        myr = __ldcs(a.rating + j0 + gl);
      }
      for (int b = 0; b < nb; b++) {
        const int v = __shfl_sync(m, myvid, b, LPR);
        const float r = __shfl_sync(m, myr, b, LPR);
        const Row<VPL> f = load_row<LPR, VPL>(a.phi, v, a.nvec, gl);
        const float d = group_dot<LPR, VPL>(t, f, m);
        // model.cc:62-63; link = 1: the prediction goes through the logistic link first, as in the update and in
        // updateReg (util.h:90-95 active(), admf.h:69, model.h:87) - the evaluation that matches --loss 1
        const float bvv = __ldcg(a.bv + v);
        const float err = a.link == 1 ? r - 1.0f / (1.0f + expf(-(d + bu + bvv + a.gb))) : r - d - bu - bvv - a.gb;
        if (gl == 0) acc += (double)(err * err);                 // model.cc:64
      }
    }
  }
  // block reduction: leaders hold partials
  __shared__ double part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += part[w];
    atomicAdd(a.accum, s);
  }
}

// K8: MF::init fill (model.cc:22-33): element = N(0,1)*scale, from Philox keyed by (array, index)
__global__ void fill_normal_kernel(float* p, int64_t rows, int cols, int stride, uint64_t seed,
                                   uint32_t which, float scale) {
  const int cvec = (cols + 3) / 4;
  const int64_t total = rows * cvec;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cvec;
    const int c4 = (int)(i - row * cvec);
    const uint4 ctr = make_uint4((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)c4, 0xF1110000u + which);
    const float4 z = box_muller4(philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))));
    const float zz[4] = {z.x, z.y, z.z, z.w};
    for (int k = 0; k < 4; k++) {
      const int c = c4 * 4 + k;
      if (c < cols) p[row * stride + c] = zz[k] * scale;
    }
  }
}

// expand packed records (the compact host form, mfb_blocks_pin): 3 or 4 B in, 8 B out per record
__global__ void unpack_kernel(const uint16_t* __restrict__ vid16, const uint8_t* __restrict__ vhi,
                              const uint8_t* __restrict__ code, const float* __restrict__ dict,
                              int32_t* __restrict__ vid, float* __restrict__ rating, int64_t n) {
  __shared__ float sdict[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sdict[i] = dict[i];
  __syncthreads();
  // streaming loads and stores (evict-first): 11 bytes per record pass through here every epoch and
  // must not push the item matrix out of the L2
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = (int32_t)__ldcs(vid16 + i);
    if (vhi) v |= (int32_t)__ldcs(vhi + i) << 16;
    __stcs(vid + i, v);
    __stcs(rating + i, sdict[__ldcs(code + i)]);
  }
}

int launch_unpack(Context* c, const uint16_t* vid16, const uint8_t* vhi, const uint8_t* code, const float* dict,
                  int32_t* vid, float* rating, int64_t n) {
  if (n <= 0) return MFB_OK;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 8);
  unpack_kernel<<<grid, 256, 0, c->stream>>>(vid16, vhi, code, dict, vid, rating, n);
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

// ------------------------------------------------------------------------------------------
namespace {

template <int LPR, int VPL>
int launch_sgd_t(Context* c, const Dataset* d, const SgdArgs& a, int mode) {
  if (mode == MFB_MODE_ORDERED) {
    sgd_epoch_kernel<LPR, VPL, MFB_MODE_ORDERED, 1><<<1, 32, 0, c->stream>>>(a);
  } else {
    constexpr int B = VPL == 1 ? 4 : (VPL == 2 ? 2 : 1);
    const void* k = mode == MFB_MODE_ATOMIC ? (const void*)sgd_epoch_kernel<LPR, VPL, MFB_MODE_ATOMIC, B>
                                            : (const void*)sgd_epoch_kernel<LPR, VPL, MFB_MODE_HOGWILD, B>;
    const LaunchShape ls = pick_launch(c, k, LPR, a.nruns - a.run_begin, d->max_item_share, d->nruns,
                                       B, a.eta);
    void* args[] = {(void*)&a};
    MFB_CUDA(cudaLaunchKernel(k, dim3(ls.grid), dim3(ls.threads), args, 0, c->stream));
  }
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

template <int LPR, int VPL>
int launch_sse_t(Context* c, const SseArgs& a) {
  auto k = sse_kernel<LPR, VPL>;
  const LaunchShape ls = pick_launch(c, (const void*)k, LPR, a.nruns, 0.0, 0, 1, 0.f);  // read-only: no bound
  k<<<ls.grid, ls.threads, 0, c->stream>>>(a);
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

}  // namespace

// nvec = stride/4 float4 per row -> (lanes per row, vectors per lane)
#define MFB_DISPATCH_SHAPE(nvec, CALL)                                    \
  do {                                                                    \
    if ((nvec) <= 4) { CALL(4, 1); }                                      \
    else if ((nvec) <= 8) { CALL(8, 1); }                                 \
    else if ((nvec) <= 16) { CALL(16, 1); }                               \
    else if ((nvec) <= 32) { CALL(32, 1); }                               \
    else if ((nvec) <= 64) { CALL(32, 2); }                               \
    else if ((nvec) <= 128) { CALL(32, 4); }                              \
    else if ((nvec) <= 256) { CALL(32, 8); }                              \
    else if ((nvec) <= 512) { CALL(32, 16); }                             \
    else { set_error("dim too large (max 2048)"); return MFB_E_ARG; }     \
  } while (0)

int launch_sgd(Context* c, Dataset* d, float eta, float lambda, float gb, int mode,
               int64_t run_begin, int64_t run_end) {
  SgdArgs a;
  a.theta = c->arr[MFB_THETA];
  a.phi = c->arr[MFB_PHI];
  a.bu = c->arr[MFB_BU];
  a.bv = c->arr[MFB_BV];
  a.run_uid = d->d_run_uid;
  a.run_off = d->d_run_off;
  a.vid = d->d_vid;
  a.rating = d->d_rating;
  a.counter = c->d_counter + c->counter_slot;
  a.big_spans = 0;
  a.span_runs = 32;
  a.run_begin = (int)run_begin;
  a.nruns = (int)run_end;
  a.nvec = c->stride / 4;
  a.eta = eta;
  a.lameta = (float)(1.0 - (double)(eta * lambda));  // mf.h:80
  a.lm1 = (float)((double)a.lameta - 1.0);           // mf.h:104
  a.gb = gb;
  a.ld_flavour = c->opt_memopt & 3;
  a.st_flavour = (c->opt_memopt >> 2) & 3;
  a.bias_flavour = (c->opt_memopt >> 4) & 3;
  a.throttle = c->opt_throttle;
  a.version = c->d_version;
  a.probe_out = c->d_probe;
  a.probe_item = c->probe_item;
  if (d->refresh_pending) {  // tiles are being re-sent from the host (mfb_dataset_refresh_from_host)
    MFB_CUDA(cudaStreamWaitEvent(c->stream, d->refreshed, 0));
    d->refresh_pending = false;
  }
  MFB_CUDA(cudaMemsetAsync(c->d_counter + c->counter_slot, 0, sizeof(int), c->stream));
  // Which kernel.  The streaming sub-warp kernel needs the fewest instructions and L2 transactions
  // per update and wins when the GPU can be filled.  When the bounds on concurrency (mfb_internal.h)
  // leave only a few hundred user-runs in flight - a DSGD cell on one of many GPUs, the first
  // epochs - throughput is (runs in flight) x (updates per second inside one run), and the
  // warp-per-run burst kernel (mfb_sgd_burst.cu) is several times faster per run.
  c->use_kernel = c->opt_kernel;
  if (c->opt_kernel == 0) {
    c->use_kernel = 3;
    const int nvec = a.nvec;
    if (mode != MFB_MODE_ORDERED && nvec <= 32) {
      const int64_t runs = a.nruns - a.run_begin;
      const int64_t cap = (int64_t)c->sm_count * 64;  // more than either kernel holds: bounds only
      const double w_stream = (double)std::min<int64_t>(bounded_groups(c, cap, d->max_item_share, d->nruns, 1.0, eta), runs);
      const double w_burst = (double)std::min<int64_t>(bounded_groups(c, cap, d->max_item_share, d->nruns, 4.0, eta), (runs + 31) / 32);
      const double stream = std::min(w_stream * c->rate_stream, 6.5e9), burst = std::min(w_burst * c->rate_burst, 5.8e9);
      if (burst > stream) c->use_kernel = 4;
    }
  }
  if (mode != MFB_MODE_ORDERED && c->use_kernel == 4) {
    bool handled = false;
    const int rc = launch_sgd_burst(c, d, a, mode, &handled);
    if (rc != MFB_OK || handled) return rc;
    c->use_kernel = 3;
  }
  if (mode != MFB_MODE_ORDERED && c->use_kernel == 3) {
    bool handled = false;
    const int rc = launch_sgd_stream(c, d, a, mode, &handled);
    if (rc != MFB_OK || handled) return rc;
  }
#define CALL(L, V) return launch_sgd_t<L, V>(c, d, a, mode)
  MFB_DISPATCH_SHAPE(a.nvec, CALL);
#undef CALL
  return MFB_OK;
}

int launch_sse(Context* c, Dataset* d, float gb, int link) {
  SseArgs a;
  a.theta = c->arr[MFB_THETA];
  a.phi = c->arr[MFB_PHI];
  a.bu = c->arr[MFB_BU];
  a.bv = c->arr[MFB_BV];
  a.run_uid = d->d_run_uid;
  a.run_off = d->d_run_off;
  a.vid = d->d_vid;
  a.rating = d->d_rating;
  a.accum = c->d_accum;
  a.nruns = (int)d->nruns;
  a.nvec = c->stride / 4;
  a.gb = gb;
  a.link = link;
  MFB_CUDA(cudaMemsetAsync(c->d_accum, 0, sizeof(double), c->stream));
#define CALL(L, V) return launch_sse_t<L, V>(c, a)
  MFB_DISPATCH_SHAPE(a.nvec, CALL);
#undef CALL
  return MFB_OK;
}

int launch_fill_normal(Context* c, uint64_t seed, float scale) {
  const int which[4] = {MFB_THETA, MFB_PHI, MFB_BU, MFB_BV};
  for (int w = 0; w < 4; w++) {
    const int64_t rows = array_rows(c, which[w]);
    const int cols = array_cols(c, which[w]), stride = array_stride(c, which[w]);
    if (rows == 0) continue;
    // the bias vectors are filled as one row-major [rows][1] array
    fill_normal_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->arr[which[w]], rows, cols, stride,
                                                               seed, (uint32_t)which[w], scale);
    MFB_CUDA(cudaGetLastError());
    c->launches++;
  }
  return MFB_OK;
}

}  // namespace mfb
