// philox.cuh -- counter-based Philox4x32-10 + Box-Muller, shared by the fused sampler stages and
// dsk_philox_normal so that both draw the SAME N(0,1) value for a given (seed, stream, element).
#pragma once
#include "common.cuh"

namespace dsk {

// ---- Philox4x32-10 counter-based generator + Box-Muller -----------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t stream_id, uint64_t quad, float out[4]) {
  uint4 ctr = make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), stream_id, 0x5EEDu);
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  float u0 = ((float)r.x + 0.5f) * k, u1 = (float)r.y * k;
  float u2 = ((float)r.z + 0.5f) * k, u3 = (float)r.w * k;
  u0 = fminf(u0, 0.99999994f);
  u2 = fminf(u2, 0.99999994f);
  float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  out[0] = m0 * c0;
  out[1] = m0 * s0;
  out[2] = m1 * c1;
  out[3] = m1 * s1;
}

}  // namespace dsk
