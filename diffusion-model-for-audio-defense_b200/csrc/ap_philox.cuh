// Counter-based Philox4x32-10 + Box-Muller normals, shared by the update kernels (ap_update.cu) and the
// black-box query kernels (ap_query.cu).  Element e of a noise tensor is lane e % 4 of block offset + e / 4, so a
// consumer can REGENERATE the noise another kernel added instead of storing it in HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ap {

// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __device__ static uint4 run(uint64_t counter, uint64_t key) {
    uint4 c = make_uint4(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 0u, 0u);
    uint32_t k0 = static_cast<uint32_t>(key), k1 = static_cast<uint32_t>(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
      const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
      c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
      k0 += W0;
      k1 += W1;
    }
    return c;
  }
};

// 4 standard normals from one Philox block (Box-Muller on two uniform pairs).  The radius and the angle use the
// transcendental-unit approximations directly -- lg2.approx, sqrt.approx, sin/cos.approx: 8 MUFU operations per 4 normals,
// absolute error < 1e-6 -- instead of libm's logf / sincospif (~150 FMA-pipe instructions per block, which made every
// noise-drawing update kernel compute-bound at 0.15-0.4 of HBM bandwidth).  The angle is taken in (-pi, pi), where
// sin.approx / cos.approx are accurate to 2^-21.
__device__ __forceinline__ void normal4(uint64_t counter, uint64_t seed, float (&z)[4]) {
  const uint4 r = Philox::run(counter, seed);
  const float u0 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1), 24 bits
  const float a0 = (static_cast<float>(r.y >> 8) + 0.5f) * (6.283185307179586f / 16777216.0f) - 3.14159265358979f;
  const float u2 = (static_cast<float>(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float a1 = (static_cast<float>(r.w >> 8) + 0.5f) * (6.283185307179586f / 16777216.0f) - 3.14159265358979f;
  float l0, l1, r0, r1, s0, c0, s1, c1;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(u2));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(l0 * -1.3862943611198906f));   // sqrt(-2 ln u) = sqrt(-2 ln2 lg2 u)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(l1 * -1.3862943611198906f));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(a0));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(a0));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(a1));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(a1));
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

}  // namespace ap
