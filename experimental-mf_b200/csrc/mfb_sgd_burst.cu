// sgd_burst_kernel - plain-SGD epoch (SgdFilter::operator(), mf.h:76-132) for the regime in which
// only a few hundred user-runs may be in flight per GPU (a DSGD cell on one of many GPUs, the first
// epochs of a file with hot items; bounds: mfb_internal.h).  Throughput is then
//     (runs in flight) x (updates per second inside ONE run),
// and a run is a sequential chain through theta_u.  This kernel shortens that chain:
//
//  * one warp per run, one float4 of the row per lane (k <= 128);
//  * B consecutive records of the run are advanced together, with the result of updating them one
//    after the other in exact arithmetic:
//        theta^b = lameta*theta^(b-1) + e_b*phi_b  =>
//        <theta^(b-1), phi_b> = lameta^b <theta^0, phi_b> + sum_{a<b} lameta^(b-1-a) e_a <phi_a, phi_b>
//    so the B + B(B-1)/2 inner products <theta^0,phi_b>, <phi_a,phi_b> are reduced TOGETHER (5 shuffle
//    rounds for all of them), e_1..e_B follow from a scalar recurrence (4 dependent operations per
//    record), and the row updates are straight vector FMAs + reductions.  (Records of one run are
//    distinct items - getdata.cc groups a user's ratings - so no phi row appears twice in a batch.)
//  * everything the next batch needs is requested one batch ahead - its item rows and biases, across
//    run boundaries; the factor row of the next user one run ahead; the record ids/ratings 32 records
//    ahead - by cp.async (LDGSTS) into shared memory, retired in order by cp.async.wait_group.  With
//    plain loads into registers nothing overlaps: ptxas tracks every LDG (and the shuffles) of the
//    loop on one scoreboard slot, so the first use of the CURRENT batch waits for the requests of
//    the NEXT one (measured: 2k cycles per batch of four).
//  * a warp walks a span of 32 consecutive runs: lane l keeps the user id and end of run l.
#include <algorithm>
#include <type_traits>

#include "mfb_internal.h"
#include "mfb_sgd_args.cuh"

namespace mfb {

namespace {

__device__ __forceinline__ float dot4(const float4& x, const float4& y) {
  const float2 p = __ffma2_rn(make_float2(x.x, x.y), make_float2(y.x, y.y),
                              __fmul2_rn(make_float2(x.z, x.w), make_float2(y.z, y.w)));
  return p.x + p.y;
}
// a*x + b*y on four lanes
__device__ __forceinline__ float4 axpby4(float a, const float4& x, float b, const float4& y) {
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  const float2 lo = __ffma2_rn(a2, make_float2(x.x, x.y), __fmul2_rn(b2, make_float2(y.x, y.y)));
  const float2 hi = __ffma2_rn(a2, make_float2(x.z, x.w), __fmul2_rn(b2, make_float2(y.z, y.w)));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {  // L2 only (.cg)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {  // immutable data (.ca)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds1i(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void burst_red4(float4* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void burst_red1(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace

// shared memory of one warp: D+1 slots of (B item rows + B bias quads), three chunks of 32 record
// ids/ratings, two prefetched factor rows + bias quads (the next two users)
template <int B, int D>
struct BurstSmem {
  static constexpr int S = D + 1;
  static constexpr int ROWS = S * B * 512;   // [slot][b][lane] float4
  static constexpr int BIAS = S * B * 16;    // [slot][b] the 16 aligned bytes around bv[v]
  static constexpr int CHUNK = 3 * 256;      // [buf][vid 32 x int | rating 32 x float]
  static constexpr int PFT = 2 * (512 + 16); // [k]: a user's row + the 16 aligned bytes around bu[u]
  static constexpr int VER = S * B * 4;      // staleness probe: item version at request time
  static constexpr int WARP_BYTES = ROWS + BIAS + CHUNK + PFT + VER;
};

// EXACT: rows of exactly 32 float4 (k = 128): no lane predicates.
// D: batches requested ahead of the one being computed (1 or 2).
template <int B, int D, int MODE, bool EXACT>
__global__ void __launch_bounds__(128) sgd_burst_kernel(const SgdArgs a, const int nspans) {
  using SM = BurstSmem<B, D>;
  constexpr int S = SM::S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool lane_ok = EXACT || lane < a.nvec;
  const bool lane0 = lane == 0;
  const float eta = a.eta, lameta = a.lameta, lm1 = a.lm1, gb = a.gb;
  const int nvec = EXACT ? 32 : a.nvec;
  const float4* __restrict__ phi4 = reinterpret_cast<const float4*>(a.phi);
  float4* theta4 = reinterpret_cast<float4*>(a.theta);
  const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + (threadIdx.x >> 5) * SM::WARP_BYTES;
  const uint32_t rows_me = wbase + lane * 16;           // + (slot*B + b)*512
  const uint32_t bias0 = wbase + SM::ROWS;              // + (slot*B + b)*16
  const uint32_t chunk0 = bias0 + SM::BIAS;             // + buf*256 (+128: ratings)
  const uint32_t pft0 = chunk0 + SM::CHUNK;             // + k*528: row (lane*16), then the bias quad at +512
  int* const ver = reinterpret_cast<int*>(smem_raw + (threadIdx.x >> 5) * SM::WARP_BYTES + SM::ROWS + SM::BIAS +
                                          SM::CHUNK + SM::PFT);  // [slot*B + b]
  unsigned long long pr_sum = 0, pr_n = 0, pr_hsum = 0, pr_hn = 0;
  {  // lanes beyond the row length never copy: their vectors must read as zeros
    float4* w = reinterpret_cast<float4*>(smem_raw + (threadIdx.x >> 5) * SM::WARP_BYTES);
    for (int q = lane; q < SM::WARP_BYTES / 16; q += 32) w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }
  int iter = 0;  // iterations of the batch loop = cp.async groups committed there

  for (;;) {
    // ---- claim a span of 32 consecutive user-runs ---------------------------------------------
    int sp = 0;
    if (lane0) sp = atomicAdd(a.counter, 1);
    sp = __shfl_sync(FULL, sp, 0);
    if (sp >= nspans) break;
    int run0 = a.run_begin + sp * a.span_runs, span_n = a.span_runs;
    if (sp >= a.big_spans) {  // the tail of the launch: single runs
      run0 = a.run_begin + a.big_spans * a.span_runs + (sp - a.big_spans);
      span_n = 1;
    }
    int s_uid = 0, s_end = 0;  // per lane: user and record-end of run `lane` of the span
    if (lane < span_n) {
      s_uid = __ldg(a.run_uid + run0 + lane);
      s_end = __ldg(a.run_off + run0 + lane + 1);
    }
    const int span_lo = __ldg(a.run_off + run0);
    const int span_hi = __shfl_sync(FULL, s_end, span_n - 1);
    if (span_lo >= span_hi) continue;

    // next run with records after run i, whose predecessor ends at `prev_end` (runs without records
    // share their predecessor's end); index span_n = none
    auto next_run = [&](int i, int prev_end, int* end, int* user) {
      int k = i + 1, e = 0;
      for (;;) {
        if (k >= span_n) break;
        e = __shfl_sync(FULL, s_end, k & 31);
        if (e != prev_end) break;
        k++;
      }
      *end = e;
      *user = k < span_n ? __shfl_sync(FULL, s_uid, k & 31) : -1;
      return k;
    };
    // the factor row + bias of run-queue entry k go to prefetch slot `slot`
    auto prefetch_user = [&](int slot, int user) {
      if (lane_ok) cp_async16(pft0 + slot * 528 + lane * 16, theta4 + (int64_t)user * nvec + lane);
      if (lane0) cp_async16(pft0 + slot * 528 + 512, a.bu + (user & ~3));
    };
    // records [32c, 32c+32) of the span live in chunk buffer c % 3
    auto chunk_fetch = [&](int buf, int q0) {
      const int q = q0 + lane;
      if (q < span_hi) {
        cp_async4(chunk0 + buf * 256 + lane * 4, a.vid + q);
        cp_async4(chunk0 + buf * 256 + 128 + lane * 4, a.rating + q);
      }
    };

    // ---- compute side: current run, the two runs after it (their rows are prefetched) -------------
    int ri = -1, cur_end = span_lo, uid = -1, uid_before = -1;  // uid_before: user of the run before ri
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    float bu = 0.f;
    float4 t0 = t;  // the row and bias as the current run found them
    float bu0 = 0.f;
    // write-back of a finished run.  ATOMIC: the user row receives what the run added to it as a reduction, so
    // a second run of the same user in flight elsewhere (another split of the file, another DSGD piece, the
    // other compute stream of a chunked epoch) loses nothing - its row is merely stale, like an item row
    auto retire = [&]() {
      if (MODE == MFB_MODE_ATOMIC) {
        if (lane_ok)
          burst_red4(theta4 + (int64_t)uid * nvec + lane, make_float4(t.x - t0.x, t.y - t0.y, t.z - t0.z, t.w - t0.w));
        if (lane0) burst_red1(a.bu + uid, bu - bu0);
      } else {
        if (lane_ok) __stcg(theta4 + (int64_t)uid * nvec + lane, t);
        if (lane0) __stcg(a.bu + uid, bu);
      }
    };
    int up_i[2], up_end[2], up_uid[2], up_iter[2];
    up_i[0] = next_run(-1, span_lo, &up_end[0], &up_uid[0]);
    up_i[1] = next_run(up_i[0], up_end[0], &up_end[1], &up_uid[1]);
    int nswitch = 0;  // run switches so far: the run entered at switch k has its row in slot k & 1
    prefetch_user(0, up_uid[0]);
    if (up_i[1] < span_n) prefetch_user(1, up_uid[1]);
    up_iter[0] = up_iter[1] = iter - 8;  // both land before the loop starts (wait<0> below)
    // ---- request side: cursor, run under the cursor, chunk under the cursor -----------------------
    int jr = span_lo, rq_i = up_i[0], rq_end = up_end[0];
    bool rq_start = true;  // the cursor stands at the first record of run rq_i
    int rq_cbase = span_lo, rq_cbuf = 0;
    chunk_fetch(0, span_lo);
    chunk_fetch(1, span_lo + 32);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    // request the next batch (rows + bias quads) into ring slot `slot`; returns its descriptor.
    // Branch-free inside: a batch never crosses a run or a 32-record chunk; ring entries b >= n are
    // filled with the batch's last row again (never used).
    auto request = [&](int slot, int* qj, int* qnew, int* qbuf) {
      if (jr >= span_hi) return 0;
      if (jr == rq_cbase + 32) {  // entering the next chunk: fetch the one after it over the oldest
        rq_cbase += 32;
        rq_cbuf = rq_cbuf == 2 ? 0 : rq_cbuf + 1;
        chunk_fetch(rq_cbuf == 2 ? 0 : rq_cbuf + 1, rq_cbase + 32);
      }
      const int n = min(B, min(rq_end, rq_cbase + 32) - jr);
      const uint32_t cb = chunk0 + rq_cbuf * 256 + (jr - rq_cbase) * 4;
#pragma unroll
      for (int b = 0; b < B; b++) {
        const int vv = lds1i(cb + min(b, n - 1) * 4);
        if (lane_ok) cp_async16(rows_me + (slot * B + b) * 512, phi4 + (int64_t)vv * nvec + lane);
        if (lane0) cp_async16(bias0 + (slot * B + b) * 16, a.bv + (vv & ~3));
        if (a.version && lane0) ver[slot * B + b] = __ldcg(a.version + vv);
      }
      *qj = jr;
      *qnew = rq_start;
      *qbuf = rq_cbuf;
      jr += n;
      rq_start = false;
      if (jr == rq_end && jr < span_hi) {
        int dummy;
        rq_i = next_run(rq_i, rq_end, &rq_end, &dummy);
        rq_start = true;
      }
      return n;
    };

    // ---- prime: D batches requested, one group each ------------------------------------------------
    int qj[D], qnb[D], qnew[D], qbuf[D];
    int sr = 0, sc = 0;  // ring slots of the next request / the next batch to compute
#pragma unroll
    for (int d = 0; d < D; d++) {
      qj[d] = qnew[d] = qbuf[d] = 0;
      qnb[d] = request(sr, &qj[d], &qnew[d], &qbuf[d]);
      sr = sr == S - 1 ? 0 : sr + 1;
      cp_async_commit();
    }

    for (;;) {
      const int j = qj[0], nb = qnb[0], cbuf = qbuf[0];
      if (nb == 0) break;
      // ---- run switch: the batch about to be computed opens run up_i[0] ---------------------------
      if (qnew[0]) {
        if (ri >= 0) {
          retire();
        }
        const int nuid = up_uid[0];
        if (nuid == uid) {
          // the same user again: theta/bu continue in registers
        } else if (nuid == uid_before) {
          // its row was prefetched before the run in between (same user) wrote it back: reload
          t = lane_ok ? __ldcg(theta4 + (int64_t)nuid * nvec + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
          bu = __ldcg(a.bu + nuid);
        } else {
          if (iter - up_iter[0] < D + 1) {  // requested too recently for wait<D> to cover it (short runs)
            cp_async_wait<0>();
            __syncwarp();
          }
          const uint32_t at = pft0 + (nswitch & 1) * 528;
          t = lds4(at + lane * 16);
          bu = lds1(at + 512 + (nuid & 3) * 4);
        }
        uid_before = uid;
        uid = nuid;
        t0 = t;  // what the run starts from (a continued user: what was just written back)
        bu0 = bu;
        ri = up_i[0];
        cur_end = up_end[0];
        up_i[0] = up_i[1];
        up_end[0] = up_end[1];
        up_uid[0] = up_uid[1];
        up_iter[0] = up_iter[1];
        up_i[1] = next_run(up_i[0], up_end[0], &up_end[1], &up_uid[1]);
        if (up_i[1] < span_n) prefetch_user(nswitch & 1, up_uid[1]);  // the slot just read
        up_iter[1] = iter;
        nswitch++;
      }
      // ---- request the batch D ahead while this one is computed -----------------------------------
#pragma unroll
      for (int d = 0; d + 1 < D; d++) {
        qj[d] = qj[d + 1];
        qnb[d] = qnb[d + 1];
        qnew[d] = qnew[d + 1];
        qbuf[d] = qbuf[d + 1];
      }
      qnb[D - 1] = request(sr, &qj[D - 1], &qnew[D - 1], &qbuf[D - 1]);
      sr = sr == S - 1 ? 0 : sr + 1;
      cp_async_commit();
      iter++;
      cp_async_wait<D>();  // everything but the D newest requests has landed
      __syncwarp();

      // ---- this batch: all inner products at once, residuals by recurrence, row updates ---------------
      auto compute = [&](auto full_tag) {
        constexpr bool FULLB = decltype(full_tag)::value;  // all B records present: no per-record tests
        const uint32_t cb = chunk0 + cbuf * 256 + ((j - span_lo) & 31) * 4;
        float4 f[B];
        float bvv[B], r[B];
        int v[B];
#pragma unroll
        for (int b = 0; b < B; b++) {
          const uint32_t at = cb + (FULLB ? b : min(b, nb - 1)) * 4;
          v[b] = lds1i(at);
          r[b] = lds1(at + 128);
          f[b] = lds4(rows_me + (sc * B + b) * 512);
          bvv[b] = lds1(bias0 + (sc * B + b) * 16 + (v[b] & 3) * 4);
        }
        float D_[B], G[B][B];
#pragma unroll
        for (int b = 0; b < B; b++) {
          D_[b] = dot4(t, f[b]);
#pragma unroll
          for (int c = b + 1; c < B; c++) G[b][c] = dot4(f[b], f[c]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int b = 0; b < B; b++) {
            D_[b] += __shfl_xor_sync(FULL, D_[b], o);
#pragma unroll
            for (int c = b + 1; c < B; c++) G[b][c] += __shfl_xor_sync(FULL, G[b][c], o);
          }
        }
        float coef[B];
        float tpow = 1.0f;
        float my_bias = 0.f;  // lane b carries the bias increment of record b
        int my_v = 0;
#pragma unroll
        for (int b = 0; b < B; b++) {
          if (FULLB || b < nb) {
            float d = tpow * D_[b];
#pragma unroll
            for (int c = 0; c < b; c++) d = fmaf(coef[c], G[c][b], d);
            const float e = eta * (((r[b] - bvv[b] - gb) - d) - bu);
            float4* dst = reinterpret_cast<float4*>(a.phi) + (int64_t)v[b] * nvec + lane;
            if (MODE == MFB_MODE_ATOMIC) {  // increment of phi_b, from theta BEFORE this record
              const float4 nf = axpby4(e, t, lm1, f[b]);
              if (lane_ok) burst_red4(dst, nf);
            } else {
              const float4 nf = axpby4(e, t, lameta, f[b]);
              if (lane_ok) __stcg(dst, nf);
            }
            if (lane == b) {
              my_bias = fmaf(lm1, bvv[b], e);
              my_v = v[b];
            }
            t = axpby4(e, f[b], lameta, t);
            bu = fmaf(lameta, bu, e);
#pragma unroll
            for (int c = 0; c < b; c++) coef[c] *= lameta;
            coef[b] = e;
            tpow *= lameta;
          }
        }
        if (lane < (FULLB ? B : nb)) burst_red1(a.bv + my_v, my_bias);
        if (a.version && lane0) {  // staleness probe: updates of the item performed since its row was requested
#pragma unroll
          for (int b = 0; b < B; b++)
            if (FULLB || b < nb) {
              const int seen = atomicAdd(a.version + v[b], 1) - ver[sc * B + b];
              pr_sum += (unsigned)seen;
              pr_n++;
              if (v[b] == a.probe_item) {
                pr_hsum += (unsigned)seen;
                pr_hn++;
              }
            }
        }
      };
      if (nb == B) compute(std::true_type{});
      else compute(std::false_type{});
      sc = sc == S - 1 ? 0 : sc + 1;
    }
    // last run of the span
    retire();
    cp_async_wait<0>();  // nothing of this span may land in the buffers of the next one
    __syncwarp();
  }
  if (a.version && lane0) {
    atomicAdd(a.probe_out + 0, pr_sum);
    atomicAdd(a.probe_out + 1, pr_n);
    atomicAdd(a.probe_out + 2, pr_hsum);
    atomicAdd(a.probe_out + 3, pr_hn);
  }
}

// ------------------------------------------------------------------------------------------
namespace {

template <int B, int D>
int launch_burst_t(Context* c, const Dataset* d, const SgdArgs& a, int mode) {
  const bool exact = a.nvec == 32;
  const void* k = mode == MFB_MODE_ATOMIC
                      ? (exact ? (const void*)sgd_burst_kernel<B, D, MFB_MODE_ATOMIC, true>
                               : (const void*)sgd_burst_kernel<B, D, MFB_MODE_ATOMIC, false>)
                      : (exact ? (const void*)sgd_burst_kernel<B, D, MFB_MODE_HOGWILD, true>
                               : (const void*)sgd_burst_kernel<B, D, MFB_MODE_HOGWILD, false>);
  const int nruns = a.nruns - a.run_begin;
  int per_sm = 0;
  constexpr int WARP_BYTES = BurstSmem<B, D>::WARP_BYTES;
  if (4 * WARP_BYTES > 48 * 1024)
    MFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * WARP_BYTES));
  MFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, 4 * WARP_BYTES));
  per_sm = std::max(per_sm, 1);
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  // runs per claim: 32, fewer when the file is small - the claims in flight then cover a shorter
  // stretch of the file and the order of updates stays closer to the file order
  const int span_runs = c->opt_span_runs > 0 ? c->opt_span_runs : (d->nruns >= 400000 ? 32 : (d->nruns >= 100000 ? 16 : 8));
  int64_t warps = std::min<int64_t>((int64_t)c->sm_count * per_sm * 4, std::max((nruns + span_runs - 1) / span_runs, 1));
  // a run holds the batch being computed and the D requested ones between gather and reduction
  // (The hot-row budget, row_concurrency = 16, was re-derived in round 2 for the STREAM kernel, whose wide launches
  // are what went unstable.  This kernel was validated at what is now "32": round 1's 1-8 GPU runs and round 2's
  // 8-rank schedule against the reference's trajectory.  Its run counts for half its rows so that it keeps that width.)
  warps = bounded_groups(c, warps, d->max_item_share, d->nruns, 0.5 * (double)(D + 1) * B, a.eta);
  SgdArgs aa = a;
  aa.span_runs = span_runs;
  aa.big_spans = (int)std::max<int64_t>(0, (nruns - c->opt_tail_runs * warps) / span_runs);  // single runs at the end
  const int nspans = aa.big_spans + (nruns - aa.big_spans * span_runs);
  int grid, threads;
  if (warps <= c->sm_count) {
    grid = (int)warps;
    threads = 32;
  } else {
    const int64_t per = (warps + c->sm_count - 1) / c->sm_count;
    const int ctas = (int)((per + 3) / 4);
    threads = 32 * (int)((per + ctas - 1) / ctas);
    grid = c->sm_count * ctas;
  }
  c->last_grid = grid;
  c->last_threads = threads;
  c->last_ring = B * 10 + D;
  void* args[] = {(void*)&aa, (void*)&nspans};
  MFB_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), args, (size_t)(threads / 32) * WARP_BYTES, c->stream));
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

}  // namespace

int launch_sgd_burst(Context* c, const Dataset* d, const SgdArgs& a, int mode, bool* handled) {
  // rows of up to 128 floats: one float4 per lane; shorter rows leave lanes idle, which costs nothing
  // where this kernel is used (few runs in flight, latency of a single run's chain is what counts)
  *handled = a.nvec <= 32;
  if (!*handled) return MFB_OK;
  // depth: batches requested ahead.  One is enough wherever it was measured (tools/exp_wbound.py,
  // tools/exp_e2e2.py): two ahead gains 5% per run at 210 runs in flight and loses 10% when the
  // machine is full, and it holds one more batch of rows in flight per run.
  const int depth = c->opt_depth == 2 ? 2 : 1;
  if (c->opt_batch == 8) return depth == 2 ? launch_burst_t<8, 2>(c, d, a, mode) : launch_burst_t<8, 1>(c, d, a, mode);
  return depth == 2 ? launch_burst_t<4, 2>(c, d, a, mode) : launch_burst_t<4, 1>(c, d, a, mode);
}

}  // namespace mfb
