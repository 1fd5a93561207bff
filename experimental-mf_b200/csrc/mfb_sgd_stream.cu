// sgd_stream_kernel - the production schedule of the plain-SGD epoch (SgdFilter::operator(),
// mf.h:76-132) for B200.
//
// What bounds this path (profiles/): the item matrix is L2-resident, so an update is
// one 16*k-byte row gather from L2, ~6k flops and one 16*k-byte row reduction back into L2; the
// number of updates that may be *in flight* (row read, increment not yet issued) is capped by the
// staleness budget of the hottest item row.  Throughput = in-flight updates / time in flight, so
// the kernel is organised to keep that time at one L2 round trip plus one dependent chain:
//
//  * a row is owned by a sub-warp of LPR lanes holding VPL float4 each (k=128: 8 lanes x 4), so a
//    warp advances 32/LPR user-runs at once and every warp instruction does 32/LPR updates' worth
//    of work: 3 shuffle rounds per dot product instead of 5, scalar work amortised, packed
//    FFMA2/FMUL2 (fp32x2) for the row arithmetic;
//  * each sub-warp streams through a span of LPR consecutive user-runs (one contiguous range of
//    the record tiles) behind a shared-memory ring of R item rows filled by cp.async (LDGSTS,
//    L1-bypassing): row j+R is in flight while row j is used, across run boundaries, and
//    cp.async.wait_group retires the rows in order.  (A register ring does not pipeline: ptxas
//    puts every LDG of the loop on one scoreboard slot, so waiting for the oldest row waits for
//    the newest - measured 1.8k cycles per step.)  The record ids/ratings and the factor row of
//    the next user-run travel through the same asynchronous queue;
//  * the warp stays converged: all sub-warps execute the same step; run/span boundaries are the
//    only divergent (rare) paths.
// Item rows receive their increment (lameta-1)*phi + e*theta by red.global.add.v4.f32 (ATOMIC) or
// are overwritten with plain stores (HOGWILD, the reference's literal --fly N semantics).
#include <algorithm>

#include "mfb_internal.h"
#include "mfb_sgd_args.cuh"

namespace mfb {

namespace {

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat4(const float2& a, const float2& b) { return make_float4(a.x, a.y, b.x, b.y); }

__device__ __forceinline__ void red_add4(float4* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add4p(float4* p, const float4& v, bool on) {
  if (on) red_add4(p, v);
}
// the same with the old value returned: its arrival tells the issuer that the L2 has performed it
__device__ __forceinline__ float atom_add1(float* p, float v) {
  float old;
  asm volatile("atom.global.add.f32 %0, [%1], %2;" : "=f"(old) : "l"(p), "f"(v) : "memory");
  return old;
}

template <int LPR>
__device__ __forceinline__ unsigned sub_mask(int lane) {
  if (LPR == 32) return 0xffffffffu;
  return ((1u << (LPR & 31)) - 1u) << (lane & ~(LPR - 1));
}

// ---- cp.async helpers (LDGSTS): global -> shared without passing through registers ------------
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

// shared memory of one warp: R ring slots of (32/LPR rows + 32/LPR bias quads), two record chunks,
// one prefetched factor row per sub-warp
template <int LPR, int VPL, int R>
struct StreamSmem {
  static constexpr int SUBS = 32 / LPR;
  static constexpr int ROW_F4 = LPR * VPL;              // float4 per row slot (padded row)
  static constexpr int RING_BYTES = R * 32 * VPL * 16;  // R * SUBS * ROW_F4 float4
  static constexpr int BIAS_BYTES = R * SUBS * 16;
  static constexpr int CHUNK_BYTES = 2 * 32 * 8;        // [2][vid 32 x int | rating 32 x float]
  static constexpr int PFT_BYTES = 32 * VPL * 16;
  static constexpr int T0_BYTES = 32 * VPL * 16;        // the factor row as the run found it (ATOMIC: delta write-back)
  static constexpr int WARP_BYTES = RING_BYTES + BIAS_BYTES + CHUNK_BYTES + PFT_BYTES + T0_BYTES;
};

}  // namespace

// EXACT: the row is exactly LPR*VPL float4 (k = 32, 64, 128, 256, 512): no per-vector predicates.
template <int LPR, int VPL, int MODE, int R, bool EXACT>
__global__ void __launch_bounds__(128) sgd_stream_kernel(const SgdArgs a, const int nspans) {
  using SM = StreamSmem<LPR, VPL, R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const int sub = lane / LPR;
  const unsigned m = sub_mask<LPR>(lane);
  constexpr unsigned FULL = 0xffffffffu;
  const float4* __restrict__ phi4 = reinterpret_cast<const float4*>(a.phi);
  const int64_t lane_off = a.phi_pair4 ? (int64_t)(gl >> 1) * a.phi_pair4 + (gl & 1) : (int64_t)gl;
  float4* theta4 = reinterpret_cast<float4*>(a.theta);
  bool ok[VPL];  // this lane's vector i exists (rows whose length is not a multiple of LPR*4 floats)
#pragma unroll
  for (int i = 0; i < VPL; i++) ok[i] = EXACT || gl + i * LPR < a.nvec;
  const float2 lameta2 = make_float2(a.lameta, a.lameta);
  const float2 lm12 = make_float2(a.lm1, a.lm1);

  // ---- this warp's shared memory -----------------------------------------------------------------
  const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + (threadIdx.x >> 5) * SM::WARP_BYTES;
  const uint32_t ring_me = wbase + (sub * SM::ROW_F4 + gl) * 16;  // + s*32*VPL*16 + i*LPR*16
  const uint32_t bias_sub = wbase + SM::RING_BYTES + sub * 16;    // + s*SUBS*16
  const uint32_t chunk0 = wbase + SM::RING_BYTES + SM::BIAS_BYTES;  // + buf*256 (+128 for ratings)
  const uint32_t pft_me = chunk0 + SM::CHUNK_BYTES + (sub * SM::ROW_F4 + gl) * 16;
  const uint32_t t0_me = pft_me + SM::PFT_BYTES;  // written and read by this lane only: no barrier needed
  {  // zero everything once: vectors a lane never copies (ok[i] false) must read as zeros
    float4* w = reinterpret_cast<float4*>(smem_raw + (threadIdx.x >> 5) * SM::WARP_BYTES);
    for (int q = lane; q < SM::WARP_BYTES / 16; q += 32) w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }

  // ---- per sub-warp state (uniform across the LPR lanes unless noted) ------------------------
  bool act = false, done = false;
  int s_uid = 0, s_end = 0;          // per lane: uid and record-end of run gl of the span
  int span_n = 0, span_hi = 0;       // runs in the span, one past its last record
  int ri = 0, cur_end = 0, uid = 0;  // current run (index in span), one past its last record
  int pf_ri = -1;                    // run whose theta sits in the prefetch slot (-1: none)
  int pf_step = 0;                   // step at which that prefetch was issued
  bool pf_same = false;              // ... it is the same user as the current run: nothing fetched
  int jc = 0, jp = 0;                // next record to update / next record to gather
  int cbase = 0, cbuf = 0;           // records [cbase, cbase+LPR) are in chunk buffer cbuf
  int step = 0;                      // steps executed (= cp.async groups committed in the main loop)
  float4 t[VPL];
  float bu = 0.f, pf_bu = 0.f, bu0 = 0.f;
  float fr[R];
  int fv[R];
  int fver[R];                       // probe: version of the item when its row was gathered
  unsigned long long pr_sum = 0, pr_n = 0, pr_hsum = 0, pr_hn = 0;
#pragma unroll
  for (int s = 0; s < R; s++) fver[s] = 0;
  constexpr int U = R > 2 ? R : 2;   // steps per trip of the main loop; ring slot of step u is u % R
  float ack[2] = {0.f, 0.f};         // values returned by the bias atomics of the last two updates
  unsigned sink = 0;
#pragma unroll
  for (int i = 0; i < VPL; i++) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < R; s++) {
    fr[s] = 0.f;
    fv[s] = 0;
  }

  // copy records [q0, q0+LPR) of the span (one per lane) into chunk buffer b
  auto chunk_fetch = [&](int b, int q0) {
    const int q = q0 + gl;
    if (q < span_hi) {
      cp_async4(chunk0 + b * 256 + lane * 4, a.vid + q);
      cp_async4(chunk0 + b * 256 + 128 + lane * 4, a.rating + q);
    }
  };
  // start the gather of record jp into ring slot s
  auto gather = [&](int s) {
    if (act && jp < span_hi) {
      int idx = jp - cbase;
      if (idx == LPR) {  // chunk used up: the other buffer holds the next LPR records; refill this one
        chunk_fetch(cbuf, cbase + 2 * LPR);
        cbuf ^= 1;
        cbase += LPR;
        idx = 0;
      }
      const int v = lds1i(chunk0 + cbuf * 256 + (sub * LPR + idx) * 4);
      fr[s] = lds1(chunk0 + cbuf * 256 + 128 + (sub * LPR + idx) * 4);
      fv[s] = v;
      const float4* p = phi4 + (int64_t)v * a.phi_row4 + lane_off;
#pragma unroll
      for (int i = 0; i < VPL; i++)
        if (ok[i]) cp_async16(ring_me + (s * 32 * VPL + i * LPR) * 16, p + i * a.phi_line4);
      // the 16 aligned bytes around bv[v] (4-byte cp.async would go through L1, which is not coherent)
      if (gl == 0) cp_async16(bias_sub + s * SM::SUBS * 16, a.bv + (v & ~3));
      if (a.version) fver[s] = __ldcg(a.version + v);
      jp++;
    }
  };

  for (;;) {
    // ---- span acquisition (divergent, once per LPR user-runs) --------------------------------
    if (!act && !done) {
      int sp = 0;
      if (gl == 0) sp = atomicAdd(a.counter, 1);
      sp = __shfl_sync(m, sp, 0, LPR);
      if (sp >= nspans) {
        done = true;
      } else {
        int run0 = a.run_begin + sp * LPR;
        span_n = LPR;
        if (sp >= a.big_spans) {  // the tail of the launch: single runs
          run0 = a.run_begin + a.big_spans * LPR + (sp - a.big_spans);
          span_n = 1;
        }
        s_uid = 0;
        s_end = 0;
        if (gl < span_n) {
          s_uid = __ldg(a.run_uid + run0 + gl);
          s_end = __ldg(a.run_off + run0 + gl + 1);
        }
        const int span_lo = __ldg(a.run_off + run0);
        span_hi = __shfl_sync(m, s_end, span_n - 1, LPR);
        if (span_lo < span_hi) {
          jc = jp = cbase = span_lo;
          cur_end = span_lo;  // forces the run switch before the first update
          ri = -1;
          pf_ri = -1;
          cbuf = 0;
          act = true;
          chunk_fetch(0, span_lo);
          chunk_fetch(1, span_lo + LPR);
          cp_async_commit();
          cp_async_wait<0>();
          __syncwarp(m);
          // prime the ring: one group per slot, so that "all but the newest R-1 groups" always
          // covers the row about to be used
#pragma unroll
          for (int s = 0; s < R; s++) {
            gather(s);
            cp_async_commit();
          }
        }
      }
    }
    if (__all_sync(FULL, done)) break;

#pragma unroll
    for (int u = 0; u < U; u++) {
      const int s = u % R;
      cp_async_wait<R - 1>();  // the row gathered R steps ago (and everything older) has landed
      __syncwarp();            // ... for every lane: bias quads and record chunks are read across lanes
      // ---- run switch / span end (divergent, once per user-run) -------------------------------
      if (act && jc == cur_end) {
        if (ri >= 0) {  // retire the finished run
          if (MODE == MFB_MODE_ATOMIC) {
            // the user row receives what this run added to it, as a reduction: a second run of the same
            // user that is in flight elsewhere (another split of the file, the other compute stream of a
            // chunked epoch, a neighbouring span) then loses nothing - its row is merely stale, like an
            // item row (the reference's in-place Hogwild loses single stores at most, mf.h:103-107)
#pragma unroll
            for (int i = 0; i < VPL; i++) {
              const float4 t0 = lds4(t0_me + i * LPR * 16);
              red_add4p(theta4 + (int64_t)uid * a.nvec + gl + i * LPR,
                        make_float4(t[i].x - t0.x, t[i].y - t0.y, t[i].z - t0.z, t[i].w - t0.w), ok[i]);
            }
            if (gl == 0) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a.bu + uid), "f"(bu - bu0) : "memory");
          } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
              if (ok[i]) __stcg(theta4 + (int64_t)uid * a.nvec + gl + i * LPR, t[i]);
            if (gl == 0) __stcg(a.bu + uid, bu);
          }
        }
        if (jc == span_hi) {
          act = false;  // a new span is claimed at the top of the outer loop
        } else {
          int nri = ri + 1;
          int e = __shfl_sync(m, s_end, nri & (LPR - 1), LPR);
          while (e == jc) {  // skip runs without records (jc < span_hi: a non-empty one follows)
            nri++;
            e = __shfl_sync(m, s_end, nri & (LPR - 1), LPR);
          }
          const bool reuse = (nri == pf_ri) && ri >= 0;
          ri = nri;
          cur_end = e;
          const int nuid = __shfl_sync(m, s_uid, ri & (LPR - 1), LPR);
          if (reuse && pf_same) {
            // same user again: theta/bu continue in registers
          } else if (reuse) {
            if (step - pf_step < R) cp_async_wait<0>();  // prefetched less than R steps ago (short run)
#pragma unroll
            for (int i = 0; i < VPL; i++) t[i] = lds4(pft_me + i * LPR * 16);
            bu = pf_bu;
          } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
              t[i] = ok[i] ? __ldcg(theta4 + (int64_t)nuid * a.nvec + gl + i * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
            bu = __ldcg(a.bu + nuid);
          }
          uid = nuid;
          if (MODE == MFB_MODE_ATOMIC) {  // what the run starts from (a continued user: what was just sent)
#pragma unroll
            for (int i = 0; i < VPL; i++)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(t0_me + i * LPR * 16), "f"(t[i].x), "f"(t[i].y),
                           "f"(t[i].z), "f"(t[i].w) : "memory");
            bu0 = bu;
          }
          // fetch the next run's factor row now; it is needed one whole run from here
          pf_ri = ri + 1;
          if (pf_ri < span_n) {
            const int puid = __shfl_sync(m, s_uid, pf_ri & (LPR - 1), LPR);
            pf_same = (puid == uid);
            if (!pf_same) {
#pragma unroll
              for (int i = 0; i < VPL; i++)
                if (ok[i]) cp_async16(pft_me + i * LPR * 16, theta4 + (int64_t)puid * a.nvec + gl + i * LPR);
              pf_bu = __ldcg(a.bu + puid);
              pf_step = step;
            }
          } else {
            pf_ri = -1;
          }
        }
      }

      // ---- one update per sub-warp (converged) -------------------------------------------------
      float4 f[VPL];
#pragma unroll
      for (int i = 0; i < VPL; i++) f[i] = lds4(ring_me + (s * 32 * VPL + i * LPR) * 16);
      const float bvv = lds1(bias_sub + s * SM::SUBS * 16 + (fv[s] & 3) * 4);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < VPL; i++) {
        acc0 = __ffma2_rn(lo2(t[i]), lo2(f[i]), acc0);
        acc1 = __ffma2_rn(hi2(t[i]), hi2(f[i]), acc1);
      }
      float d = (acc0.x + acc0.y) + (acc1.x + acc1.y);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o, LPR);
      const float e = a.eta * (fr[s] - d - bu - bvv - a.gb);
      const float2 e2 = make_float2(e, e);
      if (act) {
        // Closed loop on the memory system: reductions are fire-and-forget, and when the L2 atomic
        // units saturate they queue up - rows stay stale for longer than the ring accounts for
        // (measured: divergence at a budget that is stable below saturation).  The bias update of a
        // record is an atomic WITH return; waiting here for the one issued two steps ago keeps at
        // most two records' reductions of this sub-warp queued behind the L2, and costs nothing
        // while the L2 keeps up (the round trip overlaps a whole step).
        if (a.throttle) sink ^= __float_as_uint(ack[u & 1]);
        // the reductions come first: they end the window in which this row is stale elsewhere
        float4* dst = reinterpret_cast<float4*>(a.phi) + (int64_t)fv[s] * a.phi_row4 + lane_off;
        const float2 keep = (MODE == MFB_MODE_ATOMIC) ? lm12 : lameta2;  // increment vs new value
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          const float4 nf = cat4(__ffma2_rn(e2, lo2(t[i]), __fmul2_rn(keep, lo2(f[i]))),
                                 __ffma2_rn(e2, hi2(t[i]), __fmul2_rn(keep, hi2(f[i]))));
          if (MODE == MFB_MODE_ATOMIC) {
            if (EXACT) red_add4(dst + i * a.phi_line4, nf);
            else red_add4p(dst + i * a.phi_line4, nf, ok[i]);
          } else if (ok[i]) {
            __stcg(dst + i * a.phi_line4, nf);
          }
        }
        if (gl == 0) {
          if (a.throttle) ack[u & 1] = atom_add1(a.bv + fv[s], fmaf(a.lm1, bvv, e));
          else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a.bv + fv[s]), "f"(fmaf(a.lm1, bvv, e)) : "memory");
        }
        if (a.version && gl == 0) {
          const int seen = atomicAdd(a.version + fv[s], 1) - fver[s];  // updates I did not see
          pr_sum += (unsigned)seen;
          pr_n++;
          if (fv[s] == a.probe_item) {
            pr_hsum += (unsigned)seen;
            pr_hn++;
          }
        }
#pragma unroll
        for (int i = 0; i < VPL; i++)
          t[i] = cat4(__ffma2_rn(e2, lo2(f[i]), __fmul2_rn(lameta2, lo2(t[i]))),
                      __ffma2_rn(e2, hi2(f[i]), __fmul2_rn(lameta2, hi2(t[i]))));
        bu = fmaf(a.lameta, bu, e);
        jc++;
      }
      gather(s);  // refill the slot with record jc-1+R
      cp_async_commit();
      step++;
    }
  }
  if (sink == 0x7fc0dead && a.nruns < 0) a.counter[1] = 1;  // never true: keeps the waits on `ack` alive
  if (a.version && gl == 0) {
    atomicAdd(a.probe_out + 0, pr_sum);
    atomicAdd(a.probe_out + 1, pr_n);
    atomicAdd(a.probe_out + 2, pr_hsum);
    atomicAdd(a.probe_out + 3, pr_hn);
  }
}

// Plane layout of the item matrix: plane i holds float4 [i*L, i*L+L) of every item, L
// float4 (128 bytes for L = 8) per item - the pieces of one row lie nv*L*16 bytes apart and hash to
// different L2 slices, and a 256-byte hash chunk holds pieces of two items, so a row's reductions are
// spread over twice as many slices in pieces half as heavy.  The matrix is transposed into a scratch
// copy before the kernel and back after it (9 MB each way: microseconds).
__global__ void phi_rows_to_planes_kernel(const float4* __restrict__ rows, float4* __restrict__ planes, int nv,
                                          int nvec, int L) {
  const int64_t n = (int64_t)nv * nvec;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / nvec;
    const int q = (int)(t - v * nvec), i = q / L, gl = q - i * L;
    planes[v * L + (int64_t)i * nv * L + gl] = rows[t];
  }
}
__global__ void phi_planes_to_rows_kernel(const float4* __restrict__ planes, float4* __restrict__ rows, int nv,
                                          int nvec, int L) {
  const int64_t n = (int64_t)nv * nvec;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = t / nvec;
    const int q = (int)(t - v * nvec), i = q / L, gl = q - i * L;
    rows[t] = planes[v * L + (int64_t)i * nv * L + gl];
  }
}

// ------------------------------------------------------------------------------------------
namespace {

// Launch shape.  A sub-warp with a ring of R rows holds about R item rows between gather and
// reduction (R = 1: the row is in flight for the L2 round trip plus the chain up to the reduction,
// about 0.6 of a step).  The hottest item (share p of the records) is hit by inflight*p stale updates
// at once, each applied with step eta: the product eta*inflight*p is what must stay bounded
// (measured: divergence between 1.3 and 1.8 at eta = 0.02).  row_concurrency is that
// bound expressed as a count at the reference's default eta = 0.02 (main.cc:97).
// (Full Netflix shape, epoch 1, round 2: 6,720 sub-warps with R = 1 - 24 stale updates of the hottest row by the
// probe's 0.75 rows per sub-warp, eta*c = 0.47 - is stable; 11,348 - eta*c = 0.8 - ends in NaN.  A sub-warp with
// R = 1 therefore counts for one whole row: row_concurrency = 16 means eta*c <= 0.24 measured; mfb_internal.h.)
template <int R>
constexpr double ring_weight() { return (double)R; }

template <int LPR, int VPL, int R>
int64_t stream_capacity(Context* c, const void* k) {
  int per_sm = 0;
  if (4 * StreamSmem<LPR, VPL, R>::WARP_BYTES > 48 * 1024)  // k = 128 with the deepest ring: 51 KB per CTA of 4 warps
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * StreamSmem<LPR, VPL, R>::WARP_BYTES);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 128, 4 * StreamSmem<LPR, VPL, R>::WARP_BYTES);
  per_sm = std::max(per_sm, 1);
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  return (int64_t)c->sm_count * per_sm * 4 * (32 / LPR);  // sub-warps the hardware holds
}

template <int LPR, int VPL, int R>
const void* stream_kernel(int mode, bool exact) {
  return mode == MFB_MODE_ATOMIC
             ? (exact ? (const void*)sgd_stream_kernel<LPR, VPL, MFB_MODE_ATOMIC, R, true>
                      : (const void*)sgd_stream_kernel<LPR, VPL, MFB_MODE_ATOMIC, R, false>)
             : (exact ? (const void*)sgd_stream_kernel<LPR, VPL, MFB_MODE_HOGWILD, R, true>
                      : (const void*)sgd_stream_kernel<LPR, VPL, MFB_MODE_HOGWILD, R, false>);
}

template <int LPR, int VPL, int R>
int launch_stream_t(Context* c, const Dataset* d, const SgdArgs& a, int mode) {
  const void* k = stream_kernel<LPR, VPL, R>(mode, a.nvec == LPR * VPL);
  const int nruns = a.nruns - a.run_begin;
  const int subs_per_warp = 32 / LPR;
  constexpr int WARP_BYTES = StreamSmem<LPR, VPL, R>::WARP_BYTES;
  if (4 * WARP_BYTES > 48 * 1024) MFB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * WARP_BYTES));
  int64_t subs = std::min<int64_t>(stream_capacity<LPR, VPL, R>(c, k), std::max((nruns + LPR - 1) / LPR, 1));
  subs = bounded_groups(c, subs, d->max_item_share, d->nruns, ring_weight<R>(), a.eta);
  SgdArgs aa = a;
  aa.big_spans = (int)std::max<int64_t>(0, (nruns - c->opt_tail_runs * subs) / LPR);  // single runs at the end
  aa.phi_row4 = a.nvec;
  aa.phi_line4 = LPR;
  aa.phi_pair4 = 0;
  if (c->opt_phi_planes == 2 && a.nvec == LPR * VPL) {  // EXPERIMENT: sector layout, addressing only
    aa.phi_row4 = 2;
    aa.phi_pair4 = (int64_t)c->nv * 2;
    aa.phi_line4 = (int64_t)(LPR / 2) * aa.phi_pair4;
  }
  // production: 128-byte planes, for whole-epoch launches of rows that fill the lanes exactly
  const bool planes = c->opt_phi_planes == 1 && c->planes_allowed && LPR == 8 && a.nvec == LPR * VPL;
  const int tgrid = c->sm_count * 4;
  if (planes) {
    if (!c->d_phi_planes) MFB_CUDA(cudaMalloc(&c->d_phi_planes, (size_t)c->nv * c->stride * sizeof(float)));
    phi_rows_to_planes_kernel<<<tgrid, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(a.phi),
                                                           reinterpret_cast<float4*>(c->d_phi_planes), c->nv, a.nvec, LPR);
    aa.phi = c->d_phi_planes;
    aa.phi_row4 = LPR;
    aa.phi_line4 = (int64_t)c->nv * LPR;
    c->launches++;
  }
  const int nspans = aa.big_spans + (nruns - aa.big_spans * LPR);
  // spread the warps over all SMs before stacking them: 1..4 warps per CTA
  const int64_t warps = (subs + subs_per_warp - 1) / subs_per_warp;
  int grid, threads;
  if (warps <= c->sm_count) {
    grid = (int)warps;
    threads = 32;
  } else {
    const int64_t per = (warps + c->sm_count - 1) / c->sm_count;   // warps per SM
    const int ctas = (int)((per + 3) / 4);                          // CTAs per SM
    threads = 32 * (int)((per + ctas - 1) / ctas);
    grid = c->sm_count * ctas;
  }
  c->last_grid = grid;
  c->last_threads = threads;
  c->last_ring = R;
  void* args[] = {(void*)&aa, (void*)&nspans};
  MFB_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(threads), args, (size_t)(threads / 32) * WARP_BYTES, c->stream));
  if (planes) {
    phi_planes_to_rows_kernel<<<tgrid, 256, 0, c->stream>>>(reinterpret_cast<const float4*>(c->d_phi_planes),
                                                           reinterpret_cast<float4*>(a.phi), c->nv, a.nvec, LPR);
    c->launches++;
  }
  MFB_CUDA(cudaGetLastError());
  c->launches++;
  return MFB_OK;
}

template <int LPR, int VPL>
int launch_stream_r(Context* c, const Dataset* d, const SgdArgs& a, int mode) {
  int ring = c->opt_ring;
  if (ring == 0) {
    // The bound on user-runs in flight does not depend on the ring depth, the budget of the hottest
    // row does (R rows in flight per sub-warp).  Take the deepest ring (4, 2) that the row budget
    // does not make narrower than the run bound / the hardware allows; otherwise R = 1, which gets
    // the most updates per second out of a given row budget.
    const int64_t spans = std::max((a.nruns - a.run_begin + LPR - 1) / LPR, 1);
    const void* k1 = stream_kernel<LPR, VPL, 1>(mode, a.nvec == LPR * VPL);
    const int64_t cap = std::min<int64_t>(stream_capacity<LPR, VPL, 1>(c, k1), spans);
    const int64_t free_of_row = bounded_groups(c, cap, 0.0, d->nruns, 1.0, a.eta);  // run bound only
    ring = 1;
    if (bounded_groups(c, cap, d->max_item_share, d->nruns, 2.0, a.eta) >= free_of_row) ring = 2;
    if (bounded_groups(c, cap, d->max_item_share, d->nruns, 4.0, a.eta) >= free_of_row) ring = 4;
  }
  switch (ring) {
    case 1: return launch_stream_t<LPR, VPL, 1>(c, d, a, mode);
    case 2: return launch_stream_t<LPR, VPL, 2>(c, d, a, mode);
    case 3: return launch_stream_t<LPR, VPL, 3>(c, d, a, mode);
    default: return launch_stream_t<LPR, VPL, 4>(c, d, a, mode);
  }
}

}  // namespace

int launch_sgd_stream(Context* c, const Dataset* d, const SgdArgs& a, int mode, bool* handled) {
  *handled = true;
  const int nvec = a.nvec;  // float4 per row
  if (nvec <= 8) return launch_stream_r<4, 2>(c, d, a, mode);      // k <= 32
  if (nvec <= 16) return launch_stream_r<4, 4>(c, d, a, mode);     // k <= 64
  if (nvec <= 32) return launch_stream_r<8, 4>(c, d, a, mode);     // k <= 128
  if (nvec <= 64) return launch_stream_r<16, 4>(c, d, a, mode);    // k <= 256
  if (nvec <= 128) return launch_stream_r<32, 4>(c, d, a, mode);   // k <= 512
  *handled = false;
  return MFB_OK;
}

}  // namespace mfb
