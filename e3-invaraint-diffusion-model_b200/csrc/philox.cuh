// philox.cuh -- Philox4x32-10 (Salmon et al., SC'11), the counter-based generator behind every stochastic kernel of the library:
// the reverse-step race noise, the Gaussian step, the q-sample and the training dropout masks.  Counter-based => a value is a
// pure function of (key, counter): identical for any launch shape / sharding, and a dropout mask never has to be stored (the
// backward regenerates it).
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace seqdiff {

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
  const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
  const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = static_cast<uint32_t>(p1);
  const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = static_cast<uint32_t>(p0);
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// ---- training dropout ---------------------------------------------------------------------------------------------
// One mask per (site, element): `site` numbers the dropout modules of the network in forward order, `step` the optimizer step.
// keep-scale of element e = (word(e) >= p * 2^32) / (1 - p); word(e) = word e % 4 of the Philox call with counter
// (e / 4 low, e / 4 high, site, step) and key = seed.  p == 0 => all ones, no generator call.
struct DropSpec {
  float p;
  uint32_t site;
  uint32_t step;
  uint64_t seed;
};
inline DropSpec no_drop() { return DropSpec{0.f, 0u, 0u, 0ull}; }

__device__ __forceinline__ void drop_scales8(const DropSpec& d, size_t e0 /* multiple of 8 */, float (&keep)[8]) {
  if (d.p <= 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = 1.0f;
    return;
  }
  const uint32_t thr = static_cast<uint32_t>(static_cast<double>(d.p) * 4294967296.0);
  const float sc = 1.0f / (1.0f - d.p);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint64_t q = (e0 >> 2) + h;
    uint32_t c[4] = {static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), d.site, d.step};
    philox4x32_10(c, static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; ++j) keep[4 * h + j] = c[j] >= thr ? sc : 0.f;
  }
}
// scalar variant for kernels that walk single elements (attention probabilities)
__device__ __forceinline__ float drop_scale1(const DropSpec& d, size_t e) {
  if (d.p <= 0.f) return 1.0f;
  const uint32_t thr = static_cast<uint32_t>(static_cast<double>(d.p) * 4294967296.0);
  const uint64_t q = e >> 2;
  uint32_t c[4] = {static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), d.site, d.step};
  philox4x32_10(c, static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
  return c[e & 3] >= thr ? 1.0f / (1.0f - d.p) : 0.f;
}

}  // namespace seqdiff
