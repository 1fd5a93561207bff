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

// ---- the parallel schedule, sub-warp organisation (round 2) -------------------------------------------------------
// sgld_epoch_kernel above gives a whole warp to one user-run at k = 128: ~205 of its 345 warp instructions per record
// are the generator (2 Philox blocks + 2 Box-Muller quadruples + the bias pair per lane), the other ~140 are the
// update and the bookkeeping of the record - scalar work that one warp instruction does for ONE record.  Here a row
// is owned by LPR = 16 lanes holding VPL = 2 float4 each (k = 64, 32: 8 and 4 lanes), a warp advances 32/LPR user-runs at once and every
// non-generator instruction serves 32/LPR records; the generator work per record is unchanged (the same
// (t, row, chunk, kind) -> normal function, evaluated chunk by chunk and consumed at once, so no noise row is ever
// held in registers).  The loop is FLAT: all groups of a warp execute the same step, the claim / retire of a
// user-run is the only divergent path (one run per group in flight, claimed from the same queue as before).
// lambda_u / lambda_v live in shared memory (the same for every group), the factor row as the run found it in a
// per-lane shared-memory slot (the user row leaves as a reduction of its increment, like the item row).
// The ordered schedule and dim > 128 keep the kernel above.
template <int VPL>
struct FlatItem {  // the item side of one record: row, bias, weight, id
  Row<VPL> f;
  float bv, vr;
  int v;
};

// EXACT: the row is exactly LPR*VPL float4 (k = 32, 64, 128): no per-vector predicates, no padding selects
template <int LPR, int VPL, bool EXACT>
__global__ void __launch_bounds__(128) sgld_flat_kernel(const SgldArgs a) {
  extern __shared__ float4 flat_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int ROW4 = LPR * VPL;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned m = group_mask<LPR>();
  float4* lam = flat_smem;                              // [2][ROW4]: lambda_u, lambda_v (zero padded)
  float4* t0s = flat_smem + 2 * ROW4 + threadIdx.x;     // [VPL][blockDim.x]: this lane's part of the row at run start
  for (int q = threadIdx.x; q < ROW4; q += blockDim.x) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    lam[q] = q < a.nvec ? __ldg(reinterpret_cast<const float4*>(a.lambda_u) + q) : z;
    lam[ROW4 + q] = q < a.nvec ? __ldg(reinterpret_cast<const float4*>(a.lambda_v) + q) : z;
  }
  __syncthreads();
  const float noise_scale = a.temp * a.eta;
  const float4* __restrict__ phi4 = reinterpret_cast<const float4*>(a.phi);

  bool act = false, done = false;
  int uid = 0, j = 0, lo = 0, hi = 0, uc = 1;
  Row<VPL> t;
#pragma unroll
  for (int i = 0; i < VPL; i++) t.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bu = 0.f, bu_in = 0.f, au = 0.f, cbu = 1.f;
  // records [chunk, chunk+LPR) of the run, one per lane, and the LPR after them
  int c_vid = 0, c_vc = 1, n_vid = 0, n_vc = 1;
  float c_r = 0.f, n_r = 0.f;
  // two register sets for the item side: a step works on one and requests the next record's row into the other
  // (the main loop is unrolled twice, so no set is ever copied)
  FlatItem<VPL> A, B;
#pragma unroll
  for (int i = 0; i < VPL; i++) A.f.v[i] = B.f.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  A.bv = B.bv = A.vr = B.vr = 0.f;
  A.v = B.v = 0;

  auto load_chunk = [&](int q0, int* vid, float* r, int* vc) {
    const int q = q0 + gl;
    *vid = q < hi ? __ldcs(a.vid + q) : 0;
    *r = q < hi ? __ldcs(a.rating + q) : 0.f;
    *vc = q < hi ? __ldcs(a.vc + q) : 1;
  };
  auto fetch = [&](FlatItem<VPL>& it, int item) {
    it.v = item;
    const float4* p = phi4 + (int64_t)item * a.nvec + gl;
#pragma unroll
    for (int i = 0; i < VPL; i++)
      if (EXACT || gl + i * LPR < a.nvec) it.f.v[i] = __ldcg(p + i * LPR);
    it.bv = __ldcg(a.bv + item);
    it.vr = __ldg(a.vr + item);
  };
  // claim a user-run (divergent; once per run and group); its first item row goes into `it`
  auto claim = [&](FlatItem<VPL>& it) {
    if (!act && !done) {
      int run = 0;
      if (gl == 0) run = atomicAdd(a.counter, 1);
      run = __shfl_sync(m, run, 0, LPR);
      if (run >= a.nruns) {
        done = true;
      } else {
        lo = __ldg(a.run_off + run);
        hi = __ldg(a.run_off + run + 1);
        if (lo < hi) {
          uid = __ldg(a.run_uid + run);
          t = load_row<LPR, VPL>(a.theta, uid, a.nvec, gl);
          bu = __ldcg(a.bu + uid);
          const float ur = __ldg(a.ur + uid);
          uc = __ldg(a.run_uc + run);
          load_chunk(lo, &c_vid, &c_r, &c_vc);
          load_chunk(lo + LPR, &n_vid, &n_r, &n_vc);
          fetch(it, __shfl_sync(m, c_vid, 0, LPR));
#pragma unroll
          for (int i = 0; i < VPL; i++) t0s[i * blockDim.x] = t.v[i];
          bu_in = bu;
          au = -a.eta * ur * a.bound;                                          // dpmf.h:78
          cbu = (float)(1.0 - (double)(a.eta * a.lambda_ub * ur * a.bound));    // dpmf.h:84
          j = lo;
          act = true;
        }
      }
    }
  };
  // one record per group (converged: every shuffle is a full-warp instruction of width LPR; groups without a run
  // compute on stale registers and write nothing)
  auto step = [&](FlatItem<VPL>& cur, FlatItem<VPL>& nxt) {
    const int b = (j - lo) & (LPR - 1);
    const float r = __shfl_sync(FULL, c_r, b, LPR);
    const int vc = __shfl_sync(FULL, c_vc, b, LPR);
    const int item = cur.v;
    // the next record of the run: its row is requested now and used one step from here
    const bool more = act && j + 1 < hi;
    if (b + 1 == LPR) {  // ... it opens the next chunk of LPR records
      c_vid = n_vid;
      c_r = n_r;
      c_vc = n_vc;
      if (more) load_chunk(j + 1 + LPR, &n_vid, &n_r, &n_vc);
    }
    const int nitem = __shfl_sync(FULL, c_vid, (b + 1) & (LPR - 1), LPR);
    if (more) fetch(nxt, nitem);

    const float su = mufu_sqrt(noise_scale * (float)uc), sv = mufu_sqrt(noise_scale * (float)vc);  // dpmf.h:67-70
    const float av = -a.eta * cur.vr * a.bound;                                                   // dpmf.h:81
    // lazy noise of both rows, chunk by chunk: theta += su * xi_u, phi += sv * xi_v
    uint32_t spare_u = 0, spare_v = 0;
    Row<VPL> nz;  // sv * xi_v: part of the item row's increment
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; i++) {
      const uint32_t chunk = (uint32_t)(gl + i * LPR);
      const uint4 xu = philox4x32_10_rk(make_uint4((uint32_t)j, (uint32_t)uid, chunk, 2u * a.round), a.rk);
      const uint4 xv = philox4x32_10_rk(make_uint4((uint32_t)j, (uint32_t)item, chunk, 1u + 2u * a.round), a.rk);
      if (i == 0) {
        spare_u = spare_bits24(xu);
        spare_v = spare_bits24(xv);
      }
      float4 zu = box_muller4_fast(xu), zv = box_muller4_fast(xv);
      if (!EXACT && a.dim < 4 * ROW4) {  // coordinates >= dim are padding: they stay exactly zero
        const int c = 4 * (int)chunk;
        if (c + 0 >= a.dim) zu.x = zv.x = 0.f;
        if (c + 1 >= a.dim) zu.y = zv.y = 0.f;
        if (c + 2 >= a.dim) zu.z = zv.z = 0.f;
        if (c + 3 >= a.dim) zu.w = zv.w = 0.f;
      }
      float4 tt = t.v[i], f1 = cur.f.v[i];
      tt.x = fmaf(su, zu.x, tt.x); tt.y = fmaf(su, zu.y, tt.y); tt.z = fmaf(su, zu.z, tt.z); tt.w = fmaf(su, zu.w, tt.w);
      nz.v[i] = make_float4(sv * zv.x, sv * zv.y, sv * zv.z, sv * zv.w);
      f1.x += nz.v[i].x; f1.y += nz.v[i].y; f1.z += nz.v[i].z; f1.w += nz.v[i].w;
      d = fmaf(tt.x, f1.x, d); d = fmaf(tt.y, f1.y, d); d = fmaf(tt.z, f1.z, d); d = fmaf(tt.w, f1.w, d);
      t.v[i] = tt;
      cur.f.v[i] = f1;
    }
    // the two bias values from the spare bits of chunks 0 and 1 of either stream: lane 0 evaluates the user's,
    // lane 1 the item's, after one exchange (see noise_pair)
    const uint32_t other = __shfl_xor_sync(FULL, gl == 0 ? spare_v : spare_u, 1, LPR);
    const uint32_t s0 = gl == 0 ? spare_u : other, s1 = gl == 0 ? other : spare_v;
    const float bn = box_muller_bias_fast(s0, s1);
    const float xbu = __shfl_sync(FULL, bn, 0, LPR), xbv = __shfl_sync(FULL, bn, 1, LPR);
    const float bu1 = fmaf(su, xbu, bu);
    const float bv_in = cur.bv;
    const float bv1 = fmaf(sv, xbv, bv_in);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o, LPR);
    const float e = a.scal * (r - d - bu1 - bv1 - a.gb);                        // dpmf.h:72-75
    if (act) {
      // the item row receives its INCREMENT (noise + drift + gradient step) as 128-bit reductions
      float4* dst = reinterpret_cast<float4*>(a.phi) + (int64_t)item * a.nvec + gl;
#pragma unroll
      for (int i = 0; i < VPL; i++) {
        const float4 lu = lam[gl + i * LPR], lv = lam[ROW4 + gl + i * LPR];
        float4 tt = t.v[i];
        const float4 f1 = cur.f.v[i];
        float4 df;
        df.x = fmaf(av * lv.x, f1.x, fmaf(e, tt.x, nz.v[i].x));                 // dpmf.h:76,80-82
        df.y = fmaf(av * lv.y, f1.y, fmaf(e, tt.y, nz.v[i].y));
        df.z = fmaf(av * lv.z, f1.z, fmaf(e, tt.z, nz.v[i].z));
        df.w = fmaf(av * lv.w, f1.w, fmaf(e, tt.w, nz.v[i].w));
        tt.x = fmaf(e, f1.x, fmaf(au * lu.x, tt.x, tt.x));                      // dpmf.h:77-79
        tt.y = fmaf(e, f1.y, fmaf(au * lu.y, tt.y, tt.y));
        tt.z = fmaf(e, f1.z, fmaf(au * lu.z, tt.z, tt.z));
        tt.w = fmaf(e, f1.w, fmaf(au * lu.w, tt.w, tt.w));
        t.v[i] = tt;
        if (EXACT || gl + i * LPR < a.nvec)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i * LPR), "f"(df.x), "f"(df.y),
                       "f"(df.z), "f"(df.w) : "memory");
      }
      if (gl == 0) {
        const float bv_new = fmaf(1.0f - a.eta * a.lambda_vb * cur.vr * a.bound, bv1, e);  // dpmf.h:85
        atomicAdd(a.bv + item, bv_new - bv_in);
      }
      bu = fmaf(cbu, bu1, e);                                                   // dpmf.h:84
      uc = 1;  // consecutive records of a run are consecutive clock ticks
      j++;
      if (j == hi) {  // retire the run: the user row leaves as the reduction of what the run added to it
        float4* dt = reinterpret_cast<float4*>(a.theta) + (int64_t)uid * a.nvec + gl;
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          const float4 t0 = t0s[i * blockDim.x];
          if (EXACT || gl + i * LPR < a.nvec)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt + i * LPR), "f"(t.v[i].x - t0.x),
                         "f"(t.v[i].y - t0.y), "f"(t.v[i].z - t0.z), "f"(t.v[i].w - t0.w) : "memory");
        }
        if (gl == 0) atomicAdd(a.bu + uid, bu - bu_in);
        act = false;
      }
    }
  };

  for (;;) {
    claim(A);
    if (__all_sync(FULL, done)) break;
    __syncwarp();
    step(A, B);
    claim(B);
    if (__all_sync(FULL, done)) break;
    __syncwarp();
    step(B, A);
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

// parallel schedule, sub-warp organisation: FL lanes x FV float4 per row
template <int FL, int FV>
int launch_sgld_flat(Context* c, const Dataset* d, const SgldArgs& a) {
  auto k = a.nvec == FL * FV ? sgld_flat_kernel<FL, FV, true> : sgld_flat_kernel<FL, FV, false>;
  // CTAs of at most 4 warps (155 registers per thread at FV = 4: three such CTAs per SM)
  int per_sm = 0;
  const size_t smem4 = (size_t)(2 * FL * FV + FV * 128) * sizeof(float4);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, smem4);
  per_sm = std::max(per_sm, 1);
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  const int gpw = 32 / FL;  // groups per warp
  int64_t groups = std::min<int64_t>((int64_t)c->sm_count * per_sm * 4 * gpw, std::max(a.nruns, 1));
  // one record between gather and write-back per group; the step of a stale update is
  // scal = eta*ntrain*bound*lambda_r (dpmf.h:46), the counterpart of plain SGD's eta
  groups = bounded_groups(c, groups, d->max_item_share, d->nruns, 1.0, a.scal);
  const int64_t warps = (groups + gpw - 1) / gpw;
  int grid, threads;
  if (warps <= c->sm_count) {
    grid = (int)warps;
    threads = 32;
  } else {
    const int64_t per = (warps + c->sm_count - 1) / c->sm_count;  // warps per SM
    const int ctas = (int)((per + 3) / 4);
    threads = 32 * (int)((per + ctas - 1) / ctas);
    grid = c->sm_count * ctas;
  }
  c->last_grid = grid;
  c->last_threads = threads;
  const size_t smem = (size_t)(2 * FL * FV + FV * threads) * sizeof(float4);
  k<<<grid, threads, smem, c->stream>>>(a);
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

template <int LPR, int VPL>
int launch_sgld_t(Context* c, const Dataset* d, const SgldArgs& a, int mode) {
  if (mode != MFB_MODE_ORDERED && c->opt_sgld_flat) {
    // rows of 4*nvec floats over the fewest lanes that hold them with 4 float4 each
    // (k = 64 as 8 lanes x 2 float4: four runs per warp; 4 x 4 - eight runs - was slower than the warp-per-run
    // kernel's two, 30.3 against 23.6 ms)
    // (k = 128 as 8 x 4 - four runs per warp, 234 warp instructions per record - leaves 3 warps per scheduler under
    // the run bound and stalls on fixed-latency dependencies: 34.8 ms, and 53 / 41 ms in the narrow epochs 1 - 2;
    // 16 x 2: 33.1 ms and 37 / 33 ms.  Option sgld_flat = 2 selects 8 x 4.)
    if (a.nvec > 16 && a.nvec <= 32 && c->opt_sgld_flat == 2) return launch_sgld_flat<8, 4>(c, d, a);
    if (a.nvec > 16 && a.nvec <= 32) return launch_sgld_flat<16, 2>(c, d, a);    // k <= 128
    if (a.nvec > 8 && a.nvec <= 16) {                                           // k <= 64
      // four runs per warp (8 x 2) need the fewest instructions per record (20.5 against 21.0 ms) but leave half as
      // many warps: while the budgets keep the launch narrow (first epochs: 34 / 25 ms against 27 / 21) two runs
      // per warp (16 x 1) hide latency better
      const int64_t allowed = bounded_groups(c, (int64_t)c->sm_count * 64, d->max_item_share, d->nruns, 1.0, a.scal);
      if (c->opt_sgld_flat == 2 || allowed < 6000) return launch_sgld_flat<16, 1>(c, d, a);
      return launch_sgld_flat<8, 2>(c, d, a);
    }
    if (a.nvec > 4 && a.nvec <= 8) return launch_sgld_flat<4, 2>(c, d, a);       // k <= 32
  }
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
