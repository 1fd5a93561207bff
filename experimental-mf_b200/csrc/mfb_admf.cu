// K3 admf_epoch + K4 lambda update - the adaptive-regulariser path, sm_100a CUDA.
//
// Replaces AdRegFilter::operator() (admf.h:52-86) and AdaptRegMF::updateReg/updateUV/updateBias
// (model.h:86-102).  Per rating, with the four regularisers lam_u, lam_v, lam_bu, lam_bv:
//     theta_old[u] <- theta,  phi_old[v] <- phi,  bu_old[u] <- bu,  bv_old[v] <- bv   (snapshots)
//     e      = eta * (r - link(<theta,phi> + bu + bv + gb))
//     theta' = (1 - eta*lam_u)*theta + e*phi      phi' = (1 - eta*lam_v)*phi + e*theta
//     bu'    = (1 - eta*lam_bu)*bu + e            bv'  = (1 - eta*lam_bv)*bv + e
// and once per USER-RUN one validation record (uv, vv, rv) - drawn by the host exactly like the
// reference's rand() % |valid| (admf.h:82) - moves the regularisers:
//     g = rv - link(<theta[uv],phi[vv]> + bu[uv] + bv[vv] + gb)
//     lam_u  = max(0, lam_u  - eta_reg*eta*g*<theta_old[uv], phi[vv]>)
//     lam_v  = max(0, lam_v  - eta_reg*eta*g*<theta[uv], phi_old[vv]>)
//     lam_bu = max(0, lam_bu - eta_reg*eta*g*bu_old[uv]),  lam_bv likewise with bv_old[vv]
// The regularisers are four global scalars read by every rating and written once per user: a
// serial dependence.  ORDERED mode keeps it (one sub-warp, file order, oracle operation order:
// bit-exact lambda trajectory for the identity link).  The parallel mode lets every sub-warp read
// the current values at the start of its run and apply its increment with one 128-bit reduction;
// the clamp at zero is restored with a compare-and-swap in the rare case a value goes negative.
#include <algorithm>

#include "mfb_group.cuh"
#include "mfb_internal.h"

namespace mfb {

struct AdmfArgs {
  float* theta;
  float* phi;
  float* bu;
  float* bv;
  float* theta_old;
  float* phi_old;
  float* bu_old;
  float* bv_old;
  const int32_t* run_uid;
  const int32_t* run_off;
  const int32_t* vid;
  const float* rating;
  const int32_t* val_u;  // validation records (already shuffled by the host, model.cc:413)
  const int32_t* val_v;
  const float* val_r;
  const int32_t* draws;  // one index into the validation list per user-run, file order
  float* lams;           // [4] lam_u, lam_v, lam_bu, lam_bv (16-byte aligned)
  int* counter;
  int nruns, nvec, loss;
  int prefetch;          // parallel schedule: request the item row of record j+1 before working on record j
  float eta, eta_reg, gb;
};

// ---- cp.async (LDGSTS) staging of the item rows requested ahead -----------------------------------
// A register ring does not pipeline (ptxas keeps all LDGs of a loop on one scoreboard slot, so the
// first use of the oldest row waits for the newest request - DESIGN.md 3.1 item 4; measured here: the
// row requested one record ahead by LDG gained 12 % per run).  The parallel schedule therefore lands the
// rows of the next `depth` records in shared memory and retires them in order.
constexpr int ADMF_RING = 4;  // slots per group; depth <= ADMF_RING - 1
__device__ __forceinline__ void admf_cp16(uint32_t dst, const void* src) {  // L2 only (.cg): coherent with the reductions
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void admf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void admf_wait(int depth) {  // all but the `depth` youngest groups have landed
  if (depth <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  else if (depth == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else if (depth == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
  else asm volatile("cp.async.wait_group 3;" ::: "memory");
}
__device__ __forceinline__ float4 admf_lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int admf_lds1i(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float admf_lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ float link_fn(float x, int loss) {  // util.h:90-95
  return loss == 1 ? 1.0f / (1.0f + expf(-x)) : x;
}

template <int LPR, int VPL, int MODE>
__global__ void __launch_bounds__(256) admf_epoch_kernel(const AdmfArgs a) {
  constexpr bool ORDERED = (MODE == MFB_MODE_ORDERED);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned m = group_mask<LPR>();
  if (ORDERED && (blockIdx.x != 0 || threadIdx.x >= LPR)) return;
  // parallel schedule: ADMF_RING slots of (row, 16 bytes around the bias) per group in shared memory
  extern __shared__ __align__(16) unsigned char admf_smem[];
  constexpr int ROW_BYTES = LPR * VPL * 16;
  constexpr int GROUP_BYTES = ADMF_RING * (ROW_BYTES + 16) + LPR * 8;  // ring + the chunk's item ids and ratings
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(admf_smem) + (uint32_t)(threadIdx.x / LPR) * GROUP_BYTES;
  const uint32_t chunk_v = ring0 + ADMF_RING * (ROW_BYTES + 16), chunk_r = chunk_v + LPR * 4;
  const int depth = a.prefetch;  // records requested ahead (0..ADMF_RING-1), uniform
  const float4* const phi4 = reinterpret_cast<const float4*>(a.phi);
  float lam_u = 0.f, lam_v = 0.f, lam_bu = 0.f, lam_bv = 0.f;
  if (ORDERED) {
    lam_u = a.lams[0];
    lam_v = a.lams[1];
    lam_bu = a.lams[2];
    lam_bv = a.lams[3];
  }
  const float ee = __fmul_rn(a.eta_reg, a.eta);  // model.h:94: eta_reg_*eta_ evaluated first
  // parallel schedule: the four regularisers live in ONE 16-byte word that every run would hit with
  // a reduction (1.9 M serialised L2 atomics per epoch at the Netflix shape); a group sums its
  // increments and sends them every MFB_LAM_FLUSH runs (they are of order 1e-6 each)
  constexpr int MFB_LAM_FLUSH = 8;
  float4 pend = make_float4(0.f, 0.f, 0.f, 0.f);
  int npend = 0;
  auto flush_lams = [&]() {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.lams), "f"(pend.x), "f"(pend.y),
                 "f"(pend.z), "f"(pend.w)
                 : "memory");
    // model.h:94 clamps at zero after every update: a float below zero is a negative int, so a
    // signed integer max with 0 (= +0.0f) restores the clamp without a compare-and-swap loop
#pragma unroll
    for (int k = 0; k < 4; k++) atomicMax(reinterpret_cast<int*>(a.lams + k), 0);
    pend = make_float4(0.f, 0.f, 0.f, 0.f);
    npend = 0;
  };
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
    if (!ORDERED) {  // the regularisers as they are now (clamped view)
      const float4 l4 = __ldcg(reinterpret_cast<const float4*>(a.lams));
      lam_u = fmaxf(l4.x, 0.f);
      lam_v = fmaxf(l4.y, 0.f);
      lam_bu = fmaxf(l4.z, 0.f);
      lam_bv = fmaxf(l4.w, 0.f);
    }
    if (lo < hi) {
      Row<VPL> t = load_row<LPR, VPL>(a.theta, uid, a.nvec, gl);
      Row<VPL> t_prev = t;
      float bu = (gl == 0) ? __ldcg(a.bu + uid) : 0.f;
      bu = __shfl_sync(m, bu, 0, LPR);
      float bu_prev = bu;
      const Row<VPL> t_in = t;
      const float bu_in = bu;
      const float cu = __fmul_rn(-a.eta, lam_u);                       // admf.h:73
      const float cv = __fsub_rn(1.0f, __fmul_rn(a.eta, lam_v));       // admf.h:75
      const float cbu = __fsub_rn(1.0f, __fmul_rn(a.eta, lam_bu));     // admf.h:79
      const float cbv = __fsub_rn(1.0f, __fmul_rn(a.eta, lam_bv));     // admf.h:80
      // records are read LPR at a time (one per lane); in the parallel schedule the item row and
      // bias of record j+1 are requested before record j is worked on (items of a run are distinct)
      int myvid = 0, v_n = 0;
      float myr = 0.f, bv_n = 0.f;
      Row<VPL> f_n;
      auto fetch = [&](int b) {
        v_n = __shfl_sync(m, myvid, b & (LPR - 1), LPR);
        f_n = load_row<LPR, VPL>(a.phi, v_n, a.nvec, gl);
        bv_n = (gl == 0) ? __ldcg(a.bv + v_n) : 0.f;
      };
      // parallel schedule: request the row of record (chunk position) q into slot q % ADMF_RING, or
      // commit an empty group when the chunk / the run has no such record - the number of groups in
      // flight is then always depth + 1 and the wait below takes a constant
      auto request = [&](int j0, int q) {
        if (q < LPR && j0 + q < hi) {
          const int vq = admf_lds1i(chunk_v + q * 4);
          const uint32_t slot = ring0 + (uint32_t)(q & (ADMF_RING - 1)) * (ROW_BYTES + 16);
          const float4* src = phi4 + (int64_t)vq * a.nvec;
#pragma unroll
          for (int i = 0; i < VPL; i++) {
            const int w = gl + i * LPR;
            if (w < a.nvec) admf_cp16(slot + w * 16, src + w);
          }
          if (gl == 0) admf_cp16(slot + ROW_BYTES, a.bv + (vq & ~3));
        }
        admf_commit();
      };
      for (int j = lo; j < hi; j++) {
        const int b = (j - lo) & (LPR - 1);
        if (b == 0) {
          const int q = j + gl;
          myvid = q < hi ? __ldcs(a.vid + q) : 0;
          myr = q < hi ? __ldcs(a.rating + q) : 0.f;
          if (!ORDERED) {
            // the chunk's item ids and ratings go to shared memory: reading record b's from there is one
            // LDS, a sub-warp shuffle in this (divergent) loop is a ten-instruction protocol
            __syncwarp(m);  // the previous chunk has been read by every lane of the group
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(chunk_v + gl * 4), "r"(myvid) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(chunk_r + gl * 4), "f"(myr) : "memory");
            __syncwarp(m);
            for (int q0 = 0; q0 < depth; q0++) request(j, q0);
          }
        }
        int v;
        Row<VPL> f;
        float bvv;
        if (ORDERED) {
          fetch(b);  // the ordered schedule reads every row after the previous update
          v = v_n;
          f = f_n;
          bvv = __shfl_sync(m, bv_n, 0, LPR);
        } else {
          request(j - b, b + depth);
          admf_wait(depth);
          v = admf_lds1i(chunk_v + b * 4);
          const uint32_t slot = ring0 + (uint32_t)(b & (ADMF_RING - 1)) * (ROW_BYTES + 16);
#pragma unroll
          for (int i = 0; i < VPL; i++) {
            const int w = gl + i * LPR;
            f.v[i] = (w < a.nvec) ? admf_lds4(slot + w * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          bvv = (gl == 0) ? admf_lds1(slot + ROW_BYTES + (v & 3) * 4) : 0.f;
          bvv = __shfl_sync(m, bvv, 0, LPR);
        }
        const float r = ORDERED ? __shfl_sync(m, myr, b, LPR) : admf_lds1(chunk_r + b * 4);
        t_prev = t;                                                    // admf.h:67
        bu_prev = bu;                                                  // admf.h:77
        store_row_f<LPR, VPL>(a.phi_old, v, a.nvec, gl, f, ORDERED ? 0 : 1);  // admf.h:68
        if (ORDERED) {
          const float d = group_dot_ordered<LPR, VPL>(t, f, gl, m);
          const float pred = link_fn(__fadd_rn(__fadd_rn(__fadd_rn(d, bu), bvv), a.gb), a.loss);  // admf.h:69
          const float e = __fmul_rn(a.eta, __fsub_rn(r, pred));        // admf.h:70-71
#pragma unroll
          for (int i = 0; i < VPL; i++) {
            float* tt = reinterpret_cast<float*>(&t.v[i]);
            float* ff = reinterpret_cast<float*>(&f.v[i]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float q = __fmul_rn(e, tt[k]);                                   // admf.h:72
              float th = __fadd_rn(tt[k], __fmul_rn(cu, tt[k]));                     // admf.h:73
              th = __fadd_rn(th, __fmul_rn(e, ff[k]));                               // admf.h:74
              ff[k] = __fadd_rn(q, __fmul_rn(cv, ff[k]));                            // admf.h:75-76
              tt[k] = th;
            }
          }
          store_row<LPR, VPL>(a.phi, v, a.nvec, gl, f);
          if (gl == 0) {
            __stcg(a.bv_old + v, bvv);                                               // admf.h:78
            __stcg(a.bv + v, __fadd_rn(__fmul_rn(cbv, bvv), e));                     // admf.h:80
          }
          bu = __fadd_rn(__fmul_rn(cbu, bu), e);                                     // admf.h:79
        } else {
          const float d = group_dot<LPR, VPL>(t, f, m);
          const float e = a.eta * (r - link_fn(d + bu + bvv + a.gb, a.loss));
          Row<VPL> df;
#pragma unroll
          for (int i = 0; i < VPL; i++) {
            const float4 tt = t.v[i], ff = f.v[i];
            // increment of phi: (cv - 1)*phi + e*theta, applied as a reduction
            df.v[i] = make_float4(fmaf(e, tt.x, (cv - 1.0f) * ff.x), fmaf(e, tt.y, (cv - 1.0f) * ff.y),
                                  fmaf(e, tt.z, (cv - 1.0f) * ff.z), fmaf(e, tt.w, (cv - 1.0f) * ff.w));
            t.v[i] = make_float4(fmaf(e, ff.x, fmaf(cu, tt.x, tt.x)), fmaf(e, ff.y, fmaf(cu, tt.y, tt.y)),
                                 fmaf(e, ff.z, fmaf(cu, tt.z, tt.z)), fmaf(e, ff.w, fmaf(cu, tt.w, tt.w)));
          }
          red_add_row<LPR, VPL>(a.phi, v, a.nvec, gl, df);
          if (gl == 0) {
            // the shadow value: a plain 4-byte store (an atomicExch - "performed by the L2 atomic unit like
            // the reduction" - returns a value nobody reads and its round trip sat on the scoreboard:
            // 51.9 -> 42.1 ms in epoch 1, 20.8 -> 17.5 ms at full width, same trajectory)
            __stcg(a.bv_old + v, bvv);
            atomicAdd(a.bv + v, fmaf(cbv - 1.0f, bvv, e));
          }
          bu = fmaf(cbu, bu, e);
        }
      }
      if (MODE == MFB_MODE_ATOMIC) {
        // the user row receives what this run added to it as a reduction (a second run of the same user in
        // flight in another group loses nothing); the shadow copy stays a plain store, as for items
        Row<VPL> dt;
#pragma unroll
        for (int i = 0; i < VPL; i++)
          dt.v[i] = make_float4(t.v[i].x - t_in.v[i].x, t.v[i].y - t_in.v[i].y, t.v[i].z - t_in.v[i].z,
                                t.v[i].w - t_in.v[i].w);
        red_add_row<LPR, VPL>(a.theta, uid, a.nvec, gl, dt);
      } else {
        store_row<LPR, VPL>(a.theta, uid, a.nvec, gl, t);
      }
      store_row<LPR, VPL>(a.theta_old, uid, a.nvec, gl, t_prev);
      if (gl == 0) {
        if (MODE == MFB_MODE_ATOMIC) atomicAdd(a.bu + uid, bu - bu_in);
        else __stcg(a.bu + uid, bu);
        __stcg(a.bu_old + uid, bu_prev);
      }
    }
    // ---- admf.h:82-83: one validation record per user (also for users without records) ----
    const int ii = __ldg(a.draws + run);
    const int uv = __ldg(a.val_u + ii), vv = __ldg(a.val_v + ii);
    const float rv = __ldg(a.val_r + ii);
    if (ORDERED) __syncwarp(m);  // rows just stored by other lanes of this group are read below
    const Row<VPL> tv = load_row<LPR, VPL>(a.theta, uv, a.nvec, gl);
    const Row<VPL> fv = load_row<LPR, VPL>(a.phi, vv, a.nvec, gl);
    const Row<VPL> tvo = load_row<LPR, VPL>(a.theta_old, uv, a.nvec, gl);
    const Row<VPL> fvo = load_row<LPR, VPL>(a.phi_old, vv, a.nvec, gl);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (gl == 0) {
      s0 = __ldcg(a.bu + uv);
      s1 = __ldcg(a.bv + vv);
      s2 = __ldcg(a.bu_old + uv);
      s3 = __ldcg(a.bv_old + vv);
    }
    s0 = __shfl_sync(m, s0, 0, LPR);
    s1 = __shfl_sync(m, s1, 0, LPR);
    s2 = __shfl_sync(m, s2, 0, LPR);
    s3 = __shfl_sync(m, s3, 0, LPR);
    float d0, d1, d2;
    if (ORDERED) {
      d0 = group_dot_ordered<LPR, VPL>(tv, fv, gl, m);
      d1 = group_dot_ordered<LPR, VPL>(tvo, fv, gl, m);
      d2 = group_dot_ordered<LPR, VPL>(tv, fvo, gl, m);
    } else {
      d0 = group_dot<LPR, VPL>(tv, fv, m);
      d1 = group_dot<LPR, VPL>(tvo, fv, m);
      d2 = group_dot<LPR, VPL>(tv, fvo, m);
    }
    const float pred = link_fn(__fadd_rn(__fadd_rn(__fadd_rn(d0, s0), s1), a.gb), a.loss);  // model.h:87
    const float g = __fsub_rn(rv, pred);                                                     // model.h:88
    const float eg = __fmul_rn(ee, g);
    if (ORDERED) {
      lam_u = fmaxf(0.0f, __fsub_rn(lam_u, __fmul_rn(eg, d1)));    // model.h:93-94
      lam_v = fmaxf(0.0f, __fsub_rn(lam_v, __fmul_rn(eg, d2)));    // model.h:95-96
      lam_bu = fmaxf(0.0f, __fsub_rn(lam_bu, __fmul_rn(eg, s2)));  // model.h:100
      lam_bv = fmaxf(0.0f, __fsub_rn(lam_bv, __fmul_rn(eg, s3)));  // model.h:101
    } else if (gl == 0) {
      pend.x -= eg * d1;
      pend.y -= eg * d2;
      pend.z -= eg * s2;
      pend.w -= eg * s3;
      if (++npend == MFB_LAM_FLUSH) flush_lams();
    }
  }
  if (!ORDERED && gl == 0 && npend) flush_lams();
  if (ORDERED && gl == 0) {
    a.lams[0] = lam_u;
    a.lams[1] = lam_v;
    a.lams[2] = lam_bu;
    a.lams[3] = lam_bv;
  }
}

namespace {

template <int LPR, int VPL>
int launch_admf_t(Context* c, const Dataset* d, const AdmfArgs& a, int mode) {
  if (mode == MFB_MODE_ORDERED) {
    admf_epoch_kernel<LPR, VPL, MFB_MODE_ORDERED><<<1, 32, 0, c->stream>>>(a);
  } else {
    auto k = admf_epoch_kernel<LPR, VPL, MFB_MODE_ATOMIC>;
    // The regularisers are learned from the same stale rows, which makes this path less tolerant than
    // plain SGD; the step of a stale update is eta.  A run holds depth + 1 item rows between gather and
    // reduction (the current one and the `depth` requested ahead into shared memory); it counts for
    // depth + 2 in the hot-row budget (option admf_weight overrides).  Measured at the Netflix shape,
    // k = 64, with one row ahead (tools/exp_admf_width.py): weight 6 -> 3 halves the first epochs
    // (102.7/54.0/37.8/29.7 -> 54.0/29.7/21.6/21.6 ms) and moves the test RMSE by <= 4e-4 and lam_bu by
    // 0.5 %; weight 2: <= 6e-4 and 1.3 %.
    AdmfArgs aa = a;
    aa.prefetch = std::max(0, std::min(c->opt_admf_prefetch, ADMF_RING - 1));
    const int weight = c->opt_admf_weight > 0 ? c->opt_admf_weight : aa.prefetch + 2;
    const LaunchShape ls = pick_launch(c, (const void*)k, LPR, a.nruns, d->max_item_share, d->nruns, weight, a.eta);
    const size_t smem = (size_t)(ls.threads / LPR) * (ADMF_RING * (LPR * VPL * 16 + 16) + LPR * 8);
    k<<<ls.grid, ls.threads, smem, c->stream>>>(aa);
  }
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

}  // namespace

int launch_admf(Context* c, Dataset* d, float eta, float eta_reg, int loss, float gb, int mode) {
  AdmfArgs a;
  a.theta = c->arr[MFB_THETA];
  a.phi = c->arr[MFB_PHI];
  a.bu = c->arr[MFB_BU];
  a.bv = c->arr[MFB_BV];
  a.theta_old = c->arr[MFB_THETA_OLD];
  a.phi_old = c->arr[MFB_PHI_OLD];
  a.bu_old = c->arr[MFB_BU_OLD];
  a.bv_old = c->arr[MFB_BV_OLD];
  a.run_uid = d->d_run_uid;
  a.run_off = d->d_run_off;
  a.vid = d->d_vid;
  a.rating = d->d_rating;
  a.val_u = c->d_val_u;
  a.val_v = c->d_val_v;
  a.val_r = c->d_val_r;
  a.draws = c->d_draws;
  a.lams = c->d_lams;
  a.counter = c->d_counter;
  a.nruns = (int)d->nruns;
  a.nvec = c->stride / 4;
  a.loss = loss;
  a.prefetch = c->opt_admf_prefetch;
  a.eta = eta;
  a.eta_reg = eta_reg;
  a.gb = gb;
  MFB_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(int), c->stream));
  const int nvec = a.nvec;
  if (nvec <= 4) return launch_admf_t<4, 1>(c, d, a, mode);
  if (nvec <= 8) return launch_admf_t<8, 1>(c, d, a, mode);
  if (nvec <= 16) return launch_admf_t<16, 1>(c, d, a, mode);
  if (nvec <= 32) return launch_admf_t<32, 1>(c, d, a, mode);
  if (nvec <= 64) return launch_admf_t<32, 2>(c, d, a, mode);
  set_error("admf supports dim <= 256");
  return MFB_E_ARG;
}

}  // namespace mfb
