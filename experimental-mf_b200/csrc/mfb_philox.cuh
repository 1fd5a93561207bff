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
//   chunk : coordinate/4 for the factor row, MFB_BIAS_CHUNK for the bias term (value [0] used)
// The 4 outputs become 4 normals by two Box-Muller transforms on 24-bit uniforms.
#ifndef MFB_PHILOX_CUH
#define MFB_PHILOX_CUH

#include <stdint.h>

#define MFB_BIAS_CHUNK 0x7FFFFFFFu

namespace mfb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
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

// same transform with the hardware approximations (MUFU lg2 / sin / cos): used by the Hogwild
// kernels, where noise parity is statistical.  |error| ~ 1e-6, far below the noise itself.
__device__ __forceinline__ float4 box_muller4_fast(uint4 x) {
  const float s = 1.0f / 16777216.0f;
  const float u1 = ((float)(x.x >> 8) + 1.0f) * s, u2 = (float)(x.y >> 8) * s;
  const float u3 = ((float)(x.z >> 8) + 1.0f) * s, u4 = (float)(x.w >> 8) * s;
  const float r0 = sqrtf(-2.0f * __logf(u1)), r1 = sqrtf(-2.0f * __logf(u3));
  float s0, c0, s1, c1;
  __sincosf(6.28318530717958647692f * u2, &s0, &c0);
  __sincosf(6.28318530717958647692f * u4, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t round, int kind,
                                                 int32_t row, int32_t t, uint32_t chunk) {
  const uint4 ctr = make_uint4((uint32_t)t, (uint32_t)row, chunk, (uint32_t)kind + 2u * round);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return box_muller4(philox4x32_10(ctr, key));
}

}  // namespace mfb
#endif
