// K2 sgld_epoch, K6 sgld_flush, K7 col_sqnorm - the dpmf path, sm_100a CUDA.
//
// sgld_epoch replaces SgldFilter::operator() (dpmf.h:41-91).  Per rating:
//   lazy noise   theta += sqrt(temp*eta*uc)*xi_u[0:k]   bu += sqrt(temp*eta*uc)*xi_u[k]
//                phi   += sqrt(temp*eta*vc)*xi_v[0:k]   bv += sqrt(temp*eta*vc)*xi_v[k]
//   residual     e = scal*(r - <theta,phi> - bu - bv - gb),  scal = eta*ntrain*bound*lambda_r
//   update       theta' = theta - eta*ur[u]*bound*(lambda_u .* theta) + e*phi
//                phi'   = phi   - eta*vr[v]*bound*(lambda_v .* phi)   + e*theta   (old noised values)
//                bu' = (1 - eta*lambda_ub*ur[u]*bound)*bu + e,   bv' likewise
// uc / vc are "steps since this row was last touched".  The reference obtains them from a global
// atomic counter under a per-item mutex (dpmf.h:61-66) - a serialisation point.  Here they are a
// STATIC LOGICAL CLOCK: the values that counter takes in file order, computed once at ingest
// (uc is 1 inside a user-run, so only the first record of a run stores one; vc is per record).
// No atomics in the kernel, and every row still receives total noise variance temp*eta*ntrain per
// coordinate per epoch once sgld_flush (DPMF::finish_noise, model.cc:312-332) has run.
// The noise itself is Philox4x32-10 + Box-Muller evaluated in registers (mfb_philox.cuh); the
// reference's 8 GB lookup table is only emulated in the ordered parity mode.
#include <algorithm>

#include "mfb_group.cuh"
#include "mfb_internal.h"
#include "mfb_philox.cuh"

namespace mfb {

struct SgldArgs {
  float* theta;
  float* phi;
  float* bu;
  float* bv;
  const float* ur;
  const float* vr;
  const float* lambda_u;  // [stride], zero padded
  const float* lambda_v;
  const int32_t* run_uid;
  const int32_t* run_off;
  const int32_t* run_uc;  // clock delta of the first record of each run
  const int32_t* vid;
  const int32_t* vc;      // clock delta of the item row, per record
  const float* rating;
  int* counter;
  int nruns, nvec, dim;
  float eta, temp, bound, scal, lambda_ub, lambda_vb, gb;
  uint64_t seed;
  uint32_t round;
  uint32_t rk[20];     // Philox round keys of `seed` (philox_round_keys)
  const float* table;  // ordered parity mode only: the reference's noise_ table
  int table_offset;
};

// dim+1 noise values for (kind,row) at logical time t: factor part as a Row, bias part returned.
// *spare = the unused low bytes of this lane's chunk gl (i = 0), the raw material of the bias value.
template <int LPR, int VPL, bool FAST>
__device__ __forceinline__ Row<VPL> noise_chunks(const SgldArgs& a, int kind, int row, int t, int gl, uint32_t* spare) {
  Row<VPL> z;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    const uint4 x = philox4x32_10_rk(make_uint4((uint32_t)t, (uint32_t)row, (uint32_t)v, (uint32_t)kind + 2u * a.round), a.rk);
    if (i == 0) *spare = spare_bits24(x);
    z.v[i] = FAST ? box_muller4_fast(x) : box_muller4(x);
    // coordinates >= dim are padding: keep them exactly zero (a uniform branch: rows that fill
    // the group's lanes exactly, e.g. k = 128, skip the eight selects)
    if (a.dim < 4 * LPR * VPL) {
      const int c = 4 * v;
      if (c + 0 >= a.dim) z.v[i].x = 0.f;
      if (c + 1 >= a.dim) z.v[i].y = 0.f;
      if (c + 2 >= a.dim) z.v[i].z = 0.f;
      if (c + 3 >= a.dim) z.v[i].w = 0.f;
    }
  }
  return z;
}

template <int LPR, int VPL, bool FAST>
__device__ __forceinline__ Row<VPL> noise_row(const SgldArgs& a, int kind, int row, int t, int gl,
                                              float* bias_noise) {
  static_assert(LPR >= 2, "the bias value needs the chunks of lanes 0 and 1");
  uint32_t spare;
  const Row<VPL> z = noise_chunks<LPR, VPL, FAST>(a, kind, row, t, gl, &spare);
  // the bias value: spare bits of chunk 0 (lane 0) and chunk 1 (lane 1), evaluated by lane 0 and
  // handed to every lane of the group (all of them carry the bias)
  if (bias_noise) {
    const uint32_t s1 = __shfl_sync(group_mask<LPR>(), spare, 1, LPR);
    float b = 0.f;
    if (gl == 0) b = FAST ? box_muller_bias_fast(spare, s1) : box_muller_bias(spare, s1);
    *bias_noise = __shfl_sync(group_mask<LPR>(), b, 0, LPR);
  }
  return z;
}

// Both rows of one record: user noise (kind 0, row uid) and item noise (kind 1, row v) at logical time
// t.  The two bias values come from the spare bits of chunks 0 and 1 of either stream and are
// evaluated in ONE pass - lane 0 the user's, lane 1 the item's - after one exchange between the two
// lanes (lane 0 lacks the user's chunk 1, lane 1 the item's chunk 0).
template <int LPR, int VPL, bool FAST>
__device__ __forceinline__ void noise_pair(const SgldArgs& a, int uid, int v, int t, int gl, unsigned m,
                                           Row<VPL>& xu, Row<VPL>& xv, float* xbu, float* xbv) {
  static_assert(LPR >= 2, "the bias values need the chunks of lanes 0 and 1");
  uint32_t su, sv;
  xu = noise_chunks<LPR, VPL, FAST>(a, 0, uid, t, gl, &su);
  xv = noise_chunks<LPR, VPL, FAST>(a, 1, v, t, gl, &sv);
  const uint32_t other = __shfl_xor_sync(m, gl == 0 ? sv : su, 1, LPR);
  // (every lane evaluates the transform - the same warp instructions as a guarded one, without the
  // divergence bookkeeping; only lanes 0 and 1 hold meaningful bits)
  const uint32_t s0 = gl == 0 ? su : other, s1 = gl == 0 ? other : sv;
  const float b = FAST ? box_muller_bias_fast(s0, s1) : box_muller_bias(s0, s1);
  *xbu = __shfl_sync(m, b, 0, LPR);
  *xbv = __shfl_sync(m, b, 1, LPR);
}

// table source of the ordered parity mode: values [ind, ind+dim] of the reference's noise_ table
template <int LPR, int VPL>
__device__ __forceinline__ Row<VPL> table_row(const SgldArgs& a, int64_t ind, int gl, float* bias_noise) {
  Row<VPL> z;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int c = 4 * (gl + i * LPR);
    z.v[i].x = (c + 0 < a.dim) ? __ldg(a.table + ind + c + 0) : 0.f;
    z.v[i].y = (c + 1 < a.dim) ? __ldg(a.table + ind + c + 1) : 0.f;
    z.v[i].z = (c + 2 < a.dim) ? __ldg(a.table + ind + c + 2) : 0.f;
    z.v[i].w = (c + 3 < a.dim) ? __ldg(a.table + ind + c + 3) : 0.f;
  }
  *bias_noise = __ldg(a.table + ind + a.dim);
  return z;
}

template <int LPR, int VPL>
__device__ __forceinline__ Row<VPL> load_vec(const float* p, int nvec, int gl) {
  Row<VPL> r;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    r.v[i] = (v < nvec) ? __ldg(reinterpret_cast<const float4*>(p) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}

#define MFB_FOR4(EXPR)                    \
  {                                       \
    {                                     \
      auto& T = tt.x; auto& F = ff.x;     \
      const float XU = zu.x, XV = zv.x, LU = lu.x, LV = lv.x; \
      EXPR                                \
    }                                     \
    {                                     \
      auto& T = tt.y; auto& F = ff.y;     \
      const float XU = zu.y, XV = zv.y, LU = lu.y, LV = lv.y; \
      EXPR                                \
    }                                     \
    {                                     \
      auto& T = tt.z; auto& F = ff.z;     \
      const float XU = zu.z, XV = zv.z, LU = lu.z, LV = lv.z; \
      EXPR                                \
    }                                     \
    {                                     \
      auto& T = tt.w; auto& F = ff.w;     \
      const float XU = zu.w, XV = zv.w, LU = lu.w, LV = lv.w; \
      EXPR                                \
    }                                     \
  }

template <int LPR, int VPL, int MODE>
__global__ void __launch_bounds__(256) sgld_epoch_kernel(const SgldArgs a) {
  constexpr bool ORDERED = (MODE == MFB_MODE_ORDERED);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned m = group_mask<LPR>();
  if (ORDERED && (blockIdx.x != 0 || threadIdx.x >= LPR)) return;
  const Row<VPL> lam_u = load_vec<LPR, VPL>(a.lambda_u, a.nvec, gl);
  const Row<VPL> lam_v = load_vec<LPR, VPL>(a.lambda_v, a.nvec, gl);
  int next = 0;
  for (;;) {
    int run;
    if (ORDERED) {
      run = next++;
    } else {
      if (gl == 0) run = atomicAdd(a.counter, 1);
      run = __shfl_sync(m, run, 0, LPR);
    }
    if (run >= a.nruns) break;
    const int uid = __ldg(a.run_uid + run);
    const int lo = __ldg(a.run_off + run), hi = __ldg(a.run_off + run + 1);
    if (lo == hi) continue;
    Row<VPL> t = load_row<LPR, VPL>(a.theta, uid, a.nvec, gl);
    float bu = (gl == 0) ? __ldcg(a.bu + uid) : 0.f;
    bu = __shfl_sync(m, bu, 0, LPR);
    const Row<VPL> t_in = t;
    const float bu_in = bu;
    const float ur = __ldg(a.ur + uid);
    const float au = -a.eta * ur * a.bound;                                       // dpmf.h:78
    const double cbu = 1.0 - (double)(a.eta * a.lambda_ub * ur * a.bound);        // dpmf.h:84
    int uc = __ldg(a.run_uc + run);
    // records are read LPR at a time (one per lane); in the parallel schedule the item row, bias and
    // weight of record j+1 are requested before the noise of record j is evaluated (~350
    // instructions of pure arithmetic), so their latency is hidden behind it
    int myvid = 0, myvc = 0, v_n = 0;
    float myr = 0.f, bv_n = 0.f, vr_n = 0.f;
    Row<VPL> f_n;
    auto fetch = [&](int b) {
      v_n = __shfl_sync(m, myvid, b & (LPR - 1), LPR);
      f_n = load_row<LPR, VPL>(a.phi, v_n, a.nvec, gl);
      bv_n = (gl == 0) ? __ldcg(a.bv + v_n) : 0.f;
      vr_n = __ldg(a.vr + v_n);
    };
    for (int j = lo; j < hi; j++) {
      const int b = (j - lo) & (LPR - 1);
      if (b == 0) {
        const int q = j + gl;
        myvid = q < hi ? __ldcs(a.vid + q) : 0;
        myr = q < hi ? __ldcs(a.rating + q) : 0.f;
        myvc = q < hi ? __ldcs(a.vc + q) : 0;
      }
      if (ORDERED || b == 0) fetch(b);  // the ordered schedule reads every row after the previous update
      const int v = v_n;
      const float r = __shfl_sync(m, myr, b, LPR);
      const int vc = __shfl_sync(m, myvc, b, LPR);
      Row<VPL> f = f_n;
      const Row<VPL> f_in = f;
      float bvv = bv_n;
      const float bv_in = bvv;
      const float vr = vr_n;
      if (!ORDERED && b + 1 < LPR && j + 1 < hi) fetch(b + 1);  // (items of one run are distinct)
      const float av = -a.eta * vr * a.bound;                                     // dpmf.h:81
      const double cbv = 1.0 - (double)(a.eta * a.lambda_vb * vr * a.bound);      // dpmf.h:85
      // dpmf.h:67-70; the parallel schedule takes the square roots with one MUFU each
      const float su = ORDERED ? sqrtf(a.temp * a.eta * uc) : mufu_sqrt(a.temp * a.eta * uc);
      const float sv = ORDERED ? sqrtf(a.temp * a.eta * vc) : mufu_sqrt(a.temp * a.eta * vc);
      float xbu, xbv;
      Row<VPL> xu, xv;
      if (a.table) {  // dpmf.h:53-54,87 with every offset == table_offset
        const int64_t ind = (int64_t)a.table_offset + (int64_t)(j - lo) * (a.dim + 1);
        xu = table_row<LPR, VPL>(a, ind, gl, &xbu);
        xv = xu;
        xbv = xbu;
      } else {
        noise_pair<LPR, VPL, !ORDERED>(a, uid, v, j, gl, m, xu, xv, &xbu, &xbv);
      }
      if (ORDERED) {
        // the oracle's operation order, every product and sum rounded on its own
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          float4 tt = t.v[i], ff = f.v[i];
          const float4 zu = xu.v[i], zv = xv.v[i], lu = lam_u.v[i], lv = lam_v.v[i];
          MFB_FOR4(T = __fadd_rn(T, __fmul_rn(su, XU)); F = __fadd_rn(F, __fmul_rn(sv, XV)); (void)LU; (void)LV;)
          t.v[i] = tt;
          f.v[i] = ff;
        }
        bu = __fadd_rn(bu, __fmul_rn(su, xbu));
        bvv = __fadd_rn(bvv, __fmul_rn(sv, xbv));
        bvv = __shfl_sync(m, bvv, 0, LPR);
        const float d = group_dot_ordered<LPR, VPL>(t, f, gl, m);
        float e = __fsub_rn(__fsub_rn(__fsub_rn(__fsub_rn(r, d), bu), bvv), a.gb);  // dpmf.h:72-74
        e = __fmul_rn(a.scal, e);                                                    // dpmf.h:75
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          float4 tt = t.v[i], ff = f.v[i];
          const float4 zu = xu.v[i], zv = xv.v[i], lu = lam_u.v[i], lv = lam_v.v[i];
          MFB_FOR4(const float q = __fmul_rn(e, T);                       /* dpmf.h:76 */
                   T = __fadd_rn(T, __fmul_rn(au, __fmul_rn(LU, T)));     /* dpmf.h:77-78 */
                   T = __fadd_rn(T, __fmul_rn(e, F));                     /* dpmf.h:79 */
                   F = __fadd_rn(F, __fmul_rn(av, __fmul_rn(LV, F)));     /* dpmf.h:80-81 */
                   F = __fadd_rn(F, q);                                   /* dpmf.h:82 */
                   (void)XU; (void)XV;)
          t.v[i] = tt;
          f.v[i] = ff;
        }
        bu = (float)__dadd_rn(__dmul_rn(cbu, (double)bu), (double)e);      // dpmf.h:84 (unfused)
        bvv = (float)__dadd_rn(__dmul_rn(cbv, (double)bvv), (double)e);    // dpmf.h:85
      } else {
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          float4 tt = t.v[i], ff = f.v[i];
          const float4 zu = xu.v[i], zv = xv.v[i], lu = lam_u.v[i], lv = lam_v.v[i];
          MFB_FOR4(T = fmaf(su, XU, T); F = fmaf(sv, XV, F); (void)LU; (void)LV;)
          t.v[i] = tt;
          f.v[i] = ff;
        }
        bu = fmaf(su, xbu, bu);
        bvv = fmaf(sv, xbv, bvv);
        bvv = __shfl_sync(m, bvv, 0, LPR);
        const float d = group_dot<LPR, VPL>(t, f, m);
        const float e = a.scal * (r - d - bu - bvv - a.gb);
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          float4 tt = t.v[i], ff = f.v[i];
          const float4 zu = xu.v[i], zv = xv.v[i], lu = lam_u.v[i], lv = lam_v.v[i];
          MFB_FOR4(const float q = e * T;
                   T = fmaf(e, F, fmaf(au * LU, T, T));
                   F = fmaf(av * LV, F, F) + q;
                   (void)XU; (void)XV;)
          t.v[i] = tt;
          f.v[i] = ff;
        }
        bu = fmaf((float)cbu, bu, e);   // (the ordered schedule keeps the reference's double arithmetic)
        bvv = fmaf(1.0f - a.eta * a.lambda_vb * vr * a.bound, bvv, e);
      }
      if (ORDERED) {
        store_row<LPR, VPL>(a.phi, v, a.nvec, gl, f);
        if (gl == 0) __stcg(a.bv + v, bvv);
      } else {
        // Parallel schedule: the item row receives its INCREMENT (noise + drift) as a 128-bit fp32
        // reduction, so neither concurrent gradient steps nor concurrent noise injections are lost
        // (a lost noise injection would break the temp*eta*ntrain variance invariant).
        Row<VPL> df;
#pragma unroll
        for (int i = 0; i < VPL; i++)
          df.v[i] = make_float4(f.v[i].x - f_in.v[i].x, f.v[i].y - f_in.v[i].y, f.v[i].z - f_in.v[i].z,
                                f.v[i].w - f_in.v[i].w);
        red_add_row<LPR, VPL>(a.phi, v, a.nvec, gl, df);
        if (gl == 0) atomicAdd(a.bv + v, bvv - bv_in);
      }
      uc = 1;  // consecutive records of a run are consecutive clock ticks
    }
    if (ORDERED) {
      store_row<LPR, VPL>(a.theta, uid, a.nvec, gl, t);
      if (gl == 0) __stcg(a.bu + uid, bu);
    } else {
      // the user row, too, receives its increment (noise + drift) as a reduction: a second run of the same
      // user in flight in another group loses neither its gradient steps nor - the variance invariant - its noise
      Row<VPL> dt;
#pragma unroll
      for (int i = 0; i < VPL; i++)
        dt.v[i] = make_float4(t.v[i].x - t_in.v[i].x, t.v[i].y - t_in.v[i].y, t.v[i].z - t_in.v[i].z,
                              t.v[i].w - t_in.v[i].w);
      red_add_row<LPR, VPL>(a.theta, uid, a.nvec, gl, dt);
      if (gl == 0) atomicAdd(a.bu + uid, bu - bu_in);
    }
  }
}

// K6: DPMF::finish_noise (model.cc:312-332): every row receives the noise it has not yet been
// given: sqrt(temp*eta*(ntrain - last_touch)) * xi.  One group per row, pure streaming.
struct FlushArgs {
  float* mat;
  float* bias;
  const int32_t* last;  // clock of the row's last touch in the epoch (0 if never touched)
  int rows, nvec, dim, kind, ntrain;
  float eta, temp;
  uint64_t seed;
  uint32_t round;
  uint32_t rk[20];
  const float* table;
  int table_offset;
};

template <int LPR, int VPL>
__global__ void __launch_bounds__(256) sgld_flush_kernel(const FlushArgs f) {
  const int gl = threadIdx.x & (LPR - 1);
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int G = gridDim.x * blockDim.x / LPR;
  SgldArgs a;  // only the fields noise_row / table_row read
  a.dim = f.dim;
  a.seed = f.seed;
  a.round = f.round;
#pragma unroll
  for (int i = 0; i < 20; i++) a.rk[i] = f.rk[i];
  a.table = f.table;
  a.table_offset = f.table_offset;
  for (int row = g; row < f.rows; row += G) {
    const int cnt = f.ntrain - __ldg(f.last + row);        // model.cc:318,326 (int arithmetic)
    const float sc = sqrtf(f.temp * f.eta * cnt);          // model.cc:320
    float xb;
    Row<VPL> x = f.table ? table_row<LPR, VPL>(a, f.table_offset, gl, &xb)
                         : noise_row<LPR, VPL, false>(a, f.kind, row, f.ntrain, gl, &xb);
    Row<VPL> r = load_row<LPR, VPL>(f.mat, row, f.nvec, gl);
#pragma unroll
    for (int i = 0; i < VPL; i++) {
      r.v[i].x = __fadd_rn(r.v[i].x, __fmul_rn(sc, x.v[i].x));
      r.v[i].y = __fadd_rn(r.v[i].y, __fmul_rn(sc, x.v[i].y));
      r.v[i].z = __fadd_rn(r.v[i].z, __fmul_rn(sc, x.v[i].z));
      r.v[i].w = __fadd_rn(r.v[i].w, __fmul_rn(sc, x.v[i].w));
    }
    store_row<LPR, VPL>(f.mat, row, f.nvec, gl, r);
    if (gl == 0) f.bias[row] = __fadd_rn(f.bias[row], __fmul_rn(sc, xb));  // model.cc:321
  }
}

// K7: column sums of squares of a [rows][stride] matrix (util.h:156-161) in fp64, plus the sum of
// squares of the bias vector (util.h:111-113).  out[0..stride) columns, out[stride] bias.
__global__ void __launch_bounds__(256) col_sqnorm_kernel(const float* mat, const float* bias, int rows,
                                                         int stride, double* out) {
  // thread x owns column (threadIdx.x % stride) of a row slice; stride <= blockDim handled by loop
  for (int col = threadIdx.x; col < stride; col += blockDim.x) {
    double acc = 0.0;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const float x = __ldcs(mat + (int64_t)row * stride + col);
      acc += (double)(x * x);  // util.h:159 squares in fp32
    }
    atomicAdd(out + col, acc);
  }
  double b = 0.0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x) {
    const float x = __ldcs(bias + row);
    b += (double)(x * x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + stride, b);
}

// ------------------------------------------------------------------------------------------
namespace {

template <int LPR, int VPL>
int launch_sgld_t(Context* c, const Dataset* d, const SgldArgs& a, int mode) {
  if (mode == MFB_MODE_ORDERED) {
    sgld_epoch_kernel<LPR, VPL, MFB_MODE_ORDERED><<<1, 32, 0, c->stream>>>(a);
  } else {
    auto k = sgld_epoch_kernel<LPR, VPL, MFB_MODE_HOGWILD>;
    // one record between gather and write-back per group; the step of a stale update is
    // scal = eta*ntrain*bound*lambda_r (dpmf.h:46), the counterpart of plain SGD's eta
    const LaunchShape ls = pick_launch(c, (const void*)k, LPR, a.nruns, d->max_item_share, d->nruns, 1, a.scal);
    k<<<ls.grid, ls.threads, 0, c->stream>>>(a);
  }
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

template <int LPR, int VPL>
int launch_flush_t(Context* c, const FlushArgs& f) {
  const int threads = 256;
  const int64_t need = ((int64_t)f.rows + threads / LPR - 1) / (threads / LPR);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)c->sm_count * 8));
  sgld_flush_kernel<LPR, VPL><<<grid, threads, 0, c->stream>>>(f);
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

}  // namespace

#define MFB_DISPATCH_SHAPE(nvec, CALL)                                    \
  do {                                                                    \
    if ((nvec) <= 4) { CALL(4, 1); }                                      \
    else if ((nvec) <= 8) { CALL(8, 1); }                                 \
    else if ((nvec) <= 16) { CALL(16, 1); }                               \
    else if ((nvec) <= 32) { CALL(32, 1); }                               \
    else if ((nvec) <= 64) { CALL(32, 2); }                               \
    else if ((nvec) <= 128) { CALL(32, 4); }                              \
    else { set_error("dpmf supports dim <= 512"); return MFB_E_ARG; }     \
  } while (0)

int launch_sgld(Context* c, Dataset* d, const mfb_sgld_params* p, float gb, int mode) {
  SgldArgs a;
  a.theta = c->arr[MFB_THETA];
  a.phi = c->arr[MFB_PHI];
  a.bu = c->arr[MFB_BU];
  a.bv = c->arr[MFB_BV];
  a.ur = c->arr[MFB_UR];
  a.vr = c->arr[MFB_VR];
  a.lambda_u = c->arr[MFB_LAMBDA_U];
  a.lambda_v = c->arr[MFB_LAMBDA_V];
  a.run_uid = d->d_run_uid;
  a.run_off = d->d_run_off;
  a.run_uc = d->d_uc;
  a.vid = d->d_vid;
  a.vc = d->d_vc;
  a.rating = d->d_rating;
  a.counter = c->d_counter;
  a.nruns = (int)d->nruns;
  a.nvec = c->stride / 4;
  a.dim = c->dim;
  a.eta = p->eta;
  a.temp = p->temp;
  a.bound = p->bound;
  a.scal = p->eta * p->ntrain * p->bound * p->lambda_r;  // dpmf.h:46, fp32 left to right
  a.lambda_ub = p->lambda_ub;
  a.lambda_vb = p->lambda_vb;
  a.gb = gb;
  a.seed = p->seed;
  a.round = p->round;
  philox_round_keys(p->seed, a.rk);
  a.table = p->use_table ? c->d_noise_table : nullptr;
  a.table_offset = p->table_offset;
  MFB_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->stream));
#define CALL(L, V) return launch_sgld_t<L, V>(c, d, a, mode)
  MFB_DISPATCH_SHAPE(a.nvec, CALL);
#undef CALL
  return MFB_OK;
}

int launch_flush(Context* c, Dataset* d, const mfb_sgld_params* p) {
  for (int kind = 0; kind < 2; kind++) {
    FlushArgs f;
    f.mat = c->arr[kind == 0 ? MFB_THETA : MFB_PHI];
    f.bias = c->arr[kind == 0 ? MFB_BU : MFB_BV];
    f.last = kind == 0 ? d->d_last_u : d->d_last_v;
    f.rows = kind == 0 ? c->nu : c->nv;
    f.nvec = c->stride / 4;
    f.dim = c->dim;
    f.kind = kind;
    f.ntrain = p->ntrain;
    f.eta = p->eta;
    f.temp = p->temp;
    f.seed = p->seed;
    f.round = p->round;
    philox_round_keys(p->seed, f.rk);
    f.table = p->use_table ? c->d_noise_table : nullptr;
    f.table_offset = p->table_offset;
#define CALL(L, V) { int rc = launch_flush_t<L, V>(c, f); if (rc) return rc; }
    MFB_DISPATCH_SHAPE(f.nvec, CALL);
#undef CALL
  }
  return MFB_OK;
}

int launch_col_sqnorms(Context* c, double* d_out /* [2*(stride+1)] */) {
  MFB_CUDA(cudaMemsetAsync(d_out, 0, 2 * (c->stride + 1) * sizeof(double), c->stream));
  col_sqnorm_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->arr[MFB_THETA], c->arr[MFB_BU], c->nu,
                                                             c->stride, d_out);
  col_sqnorm_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->arr[MFB_PHI], c->arr[MFB_BV], c->nv,
                                                             c->stride, d_out + c->stride + 1);
  MFB_CUDA(cudaGetLastError());
  c->launches += 2;
  return MFB_OK;
}

}  // namespace mfb
