// Philox4x32-10 counter-based generator + Box-Muller, device side.
//
// The reference draws its SGLD noise from an 8 GB table of pre-generated N(0,1) floats at random
// offsets (model.cc:229-231, dpmf.h:53-54,67-70).  On the GPU the table disappears: every noise
// value is a pure function of (seed, round, row kind, row, logical time, coordinate), so any
// thread can produce the values it needs with no state and no memory traffic.
//
// Stream layout (also restated on the CPU in oracle/mf_oracle.c, mfo_philox_normal4):
//   counter = ( t , row , chunk , kind + 2*round ),  key = ( seed_lo , seed_hi )
//   t     : logical clock of the rating in the epoch (dpmf.h:62 `gc`); ntrain for finish_noise
//   row   : user or item index;  kind : 0 = user row, 1 = item row
//   chunk : coordinate/4 for the factor row
// The 4 outputs become 4 normals by two Box-Muller transforms on the TOP 24 bits of each word.
// The bias term of the same (kind,row,t) is made from the bits those transforms leave unused: the
// low bytes of words x,y,z of chunk 0 form the 24-bit u1 and those of chunk 1 the 24-bit u2
// (spare_bits24 / box_muller_bias), so the bias costs no Philox evaluation of its own - it was a
// third of the generator work of a record (chunks 0 and 1 are evaluated by lanes 0 and 1 anyway).
#ifndef MFB_PHILOX_CUH
#define MFB_PHILOX_CUH

#include <stdint.h>

namespace mfb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    // one IMAD.WIDE.U32 per product (hi and lo halves together): 20 multiplies per call, not 40
    uint32_t hi0, lo0, hi1, lo1;
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo0), "=r"(hi0) : "r"(c.x), "r"(0xD2511F53u));
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo1), "=r"(hi1) : "r"(c.z), "r"(0xCD9E8D57u));
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// The ten round keys of one seed, computed once on the host and passed with the kernel arguments: the
// kernel then reads them as constant-bank operands of the XORs instead of re-deriving them (18
// uniform-datapath additions per record in the dpmf kernel, each an issue slot).
static inline void philox_round_keys(uint64_t seed, uint32_t rk[20]) {
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    rk[2 * r] = kx;
    rk[2 * r + 1] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
}
__device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const uint32_t (&rk)[20]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, lo0, hi1, lo1;
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo0), "=r"(hi0) : "r"(c.x), "r"(0xD2511F53u));
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0, %1}, p; }" : "=r"(lo1), "=r"(hi1) : "r"(c.z), "r"(0xCD9E8D57u));
    c = make_uint4(hi1 ^ c.y ^ rk[2 * r], lo1, hi0 ^ c.w ^ rk[2 * r + 1], lo0);
  }
  return c;
}

// the 24 bits of a Philox block that box_muller4 does not use: low bytes of words x, y, z
__device__ __forceinline__ uint32_t spare_bits24(uint4 x) {
  return (x.x & 0xFFu) | ((x.y & 0xFFu) << 8) | ((x.z & 0xFFu) << 16);
}

// u1 = ((x>>8)+1)/2^24 in (0,1], u2 = (x>>8)/2^24 in [0,1); z = sqrt(-2 ln u1) * (cos, sin)(2 pi u2)
__device__ __forceinline__ float4 box_muller4(uint4 x) {
  const float s = 1.0f / 16777216.0f;
  const float u1 = ((float)(x.x >> 8) + 1.0f) * s, u2 = (float)(x.y >> 8) * s;
  const float u3 = ((float)(x.z >> 8) + 1.0f) * s, u4 = (float)(x.w >> 8) * s;
  const float r0 = sqrtf(-2.0f * logf(u1)), r1 = sqrtf(-2.0f * logf(u3));
  float s0, c0, s1, c1;
  sincosf(6.28318530717958647692f * u2, &s0, &c0);
  sincosf(6.28318530717958647692f * u4, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// bias normal from the spare bits of chunks 0 and 1: u1 = (s0+1)/2^24, u2 = s1/2^24, z = r cos
__device__ __forceinline__ float box_muller_bias(uint32_t s0, uint32_t s1) {
  const float s = 1.0f / 16777216.0f;
  const float u1 = ((float)s0 + 1.0f) * s, u2 = (float)s1 * s;
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

// same transforms with the hardware approximations (MUFU lg2 / sqrt / sin / cos, one instruction each):
// used by the Hogwild kernels, where noise parity is statistical.  |error| ~ 1e-6, far below the
// noise itself.  sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u).
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// r^2 = -2 ln(m 2^-24) = -2 ln2 (lg2 m - 24): the scaling of u1 folds into one FFMA after the MUFU;
// the angle takes the whole 32-bit word (rounded to 24 bits by the conversion, at most half a unit
// of the truncated value the exact transform uses).
__device__ __forceinline__ float4 box_muller4_fast(uint4 x) {
  const float c2 = -1.38629436111989061883f, c0 = 24.0f * 1.38629436111989061883f;
  const float w = 6.28318530717958647692f / 4294967296.0f;
  const float r0 = mufu_sqrt(fmaf(mufu_lg2((float)((x.x >> 8) + 1u)), c2, c0));
  const float r1 = mufu_sqrt(fmaf(mufu_lg2((float)((x.z >> 8) + 1u)), c2, c0));
  const float a0 = (float)x.y * w, a1 = (float)x.w * w;
  return make_float4(r0 * mufu_cos(a0), r0 * mufu_sin(a0), r1 * mufu_cos(a1), r1 * mufu_sin(a1));
}
__device__ __forceinline__ float box_muller_bias_fast(uint32_t s0, uint32_t s1) {
  const float c2 = -1.38629436111989061883f, c0 = 24.0f * 1.38629436111989061883f;
  return mufu_sqrt(fmaf(mufu_lg2((float)(s0 + 1u)), c2, c0)) * mufu_cos((float)s1 * (6.28318530717958647692f / 16777216.0f));
}

__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t round, int kind,
                                                 int32_t row, int32_t t, uint32_t chunk) {
  const uint4 ctr = make_uint4((uint32_t)t, (uint32_t)row, chunk, (uint32_t)kind + 2u * round);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return box_muller4(philox4x32_10(ctr, key));
}

}  // namespace mfb
#endif
