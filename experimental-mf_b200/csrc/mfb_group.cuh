// Sub-warp ("group") primitives shared by the epoch kernels.
//
// A factor row of `stride` floats (stride = padding(dim), a multiple of 16) is nvec = stride/4
// float4 vectors.  A group of LPR lanes (4, 8, 16 or 32) owns one user-run at a time; lane gl
// holds vectors gl, gl+LPR, ... (VPL of them), so one row access is VPL fully coalesced 128-bit
// loads per lane: 32 lanes x 16 B = 512 contiguous bytes per request at dim 128.
#ifndef MFB_GROUP_CUH
#define MFB_GROUP_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace mfb {

template <int LPR>
__device__ __forceinline__ unsigned group_mask() {
  if (LPR == 32) return 0xffffffffu;
  const int lane = threadIdx.x & 31;
  return ((1u << (LPR & 31)) - 1u) << (lane & ~(LPR - 1));
}

// butterfly sum over the group; every lane ends with the total
template <int LPR>
__device__ __forceinline__ float group_sum(float v, unsigned m) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, LPR);
  return v;
}

template <int VPL>
struct Row {
  float4 v[VPL];
};

// L2-only loads/stores for factor rows: other SMs write these rows concurrently, and L1 is not
// coherent across SMs, so rows are never cached in L1 (ld.global.cg / st.global.cg).
template <int LPR, int VPL>
__device__ __forceinline__ Row<VPL> load_row(const float* base, int64_t row, int nvec, int gl) {
  Row<VPL> r;
  const float4* p = reinterpret_cast<const float4*>(base) + row * nvec;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    r.v[i] = (v < nvec) ? __ldcg(p + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}

template <int LPR, int VPL>
__device__ __forceinline__ void store_row(float* base, int64_t row, int nvec, int gl,
                                          const Row<VPL>& r) {
  float4* p = reinterpret_cast<float4*>(base) + row * nvec;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    if (v < nvec) __stcg(p + v, r.v[i]);
  }
}

// Row access flavours (chosen per kernel launch; see DESIGN.md "memory operations"):
//   0  ld/st.global.cg      -> SASS LDG/STG.E.128.STRONG.GPU (performed at the home L2 slice)
//   1  plain ld/st.global   -> weak; loads may allocate and hit in L1 (stale reads possible)
//   2  ld.global.L1::no_allocate / plain st -> weak but never resident in L1
__device__ __forceinline__ float4 ld_na(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
template <int LPR, int VPL>
__device__ __forceinline__ Row<VPL> load_row_f(const float* base, int64_t row, int nvec, int gl, int flavour) {
  Row<VPL> r;
  const float4* p = reinterpret_cast<const float4*>(base) + row * nvec;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    if (v < nvec) {
      r.v[i] = flavour == 0 ? __ldcg(p + v) : (flavour == 1 ? p[v] : ld_na(p + v));
    } else {
      r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  return r;
}
template <int LPR, int VPL>
__device__ __forceinline__ void store_row_f(float* base, int64_t row, int nvec, int gl,
                                            const Row<VPL>& r, int flavour) {
  float4* p = reinterpret_cast<float4*>(base) + row * nvec;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    if (v < nvec) {
      if (flavour == 0) __stcg(p + v, r.v[i]);
      else p[v] = r.v[i];
    }
  }
}

// row += r with one 128-bit fp32 reduction per vector (sm_90+: red.global.add.v4.f32)
template <int LPR, int VPL>
__device__ __forceinline__ void red_add_row(float* base, int64_t row, int nvec, int gl,
                                            const Row<VPL>& r) {
  float4* p = reinterpret_cast<float4*>(base) + row * nvec;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int v = gl + i * LPR;
    if (v < nvec)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + v), "f"(r.v[i].x),
                   "f"(r.v[i].y), "f"(r.v[i].z), "f"(r.v[i].w)
                   : "memory");
  }
}

// fused partial dot + butterfly reduce (fast path)
template <int LPR, int VPL>
__device__ __forceinline__ float group_dot(const Row<VPL>& a, const Row<VPL>& b, unsigned m) {
  float d = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    d = fmaf(a.v[i].x, b.v[i].x, d);
    d = fmaf(a.v[i].y, b.v[i].y, d);
    d = fmaf(a.v[i].z, b.v[i].z, d);
    d = fmaf(a.v[i].w, b.v[i].w, d);
  }
  return group_sum<LPR>(d, m);
}

// The oracle's sdot (oracle/shim/mkl.h, oracle/mf_oracle.c o_sdot): acc = 0; acc += x[i]*y[i] for
// i ascending, product and sum rounded separately.  The running sum is handed from lane to lane
// in coordinate order, so the result is bit-identical to the CPU loop.  LPR*VPL shuffles: only
// for the ordered (parity) mode.  Padding coordinates hold zeros and add +0.
template <int LPR, int VPL>
__device__ __forceinline__ float group_dot_ordered(const Row<VPL>& a, const Row<VPL>& b, int gl,
                                                   unsigned m) {
  float carry = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const float p0 = __fmul_rn(a.v[i].x, b.v[i].x), p1 = __fmul_rn(a.v[i].y, b.v[i].y);
    const float p2 = __fmul_rn(a.v[i].z, b.v[i].z), p3 = __fmul_rn(a.v[i].w, b.v[i].w);
    for (int l = 0; l < LPR; l++) {
      const float in = __shfl_sync(m, carry, (l + LPR - 1) & (LPR - 1), LPR);
      if (gl == l) {
        float s = (i == 0 && l == 0) ? 0.f : in;
        s = __fadd_rn(s, p0);
        s = __fadd_rn(s, p1);
        s = __fadd_rn(s, p2);
        s = __fadd_rn(s, p3);
        carry = s;
      }
    }
  }
  return __shfl_sync(m, carry, LPR - 1, LPR);
}

}  // namespace mfb
#endif
