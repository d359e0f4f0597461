// skew.cuh -- register barrel shift used for the relative_key skew (attention_pipe.cu forward, attention_bwd_pipe.cu backward):
// a lane-dependent shift of a 64-entry register window, applied as 5 stages of opaque selects on the ALUs.
#pragma once
#include <cstdint>

namespace seqdiff {

// dst = p ? a : b as an opaque SELP (written as `p ? x[k + sh] : x[k]` the compiler turns the barrel shift back into a
// dynamically indexed array in local memory)
__device__ __forceinline__ uint32_t selp_u32(uint32_t a, uint32_t b, uint32_t p) {
  uint32_t d;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.b32 %0, %1, %2, q;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(p));
  return d;
}
// one stage of the barrel shift  X[k] = on ? X[k + SH] : X[k]  over the 32 + SH - 1 entries later stages still need
// (X = x0 ++ x1; every index is a compile-time constant after unrolling)
template <int SH>
__device__ __forceinline__ void shift_stage(uint32_t (&x0)[32], uint32_t (&x1)[32], uint32_t on) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    constexpr int dummy = 0;
    (void)dummy;
    const uint32_t src = (k + SH < 32) ? x0[(k + SH) & 31] : x1[(k + SH - 32) & 31];
    x0[k] = selp_u32(src, x0[k], on);
  }
#pragma unroll
  for (int k = 0; k < SH - 1; ++k) x1[k] = selp_u32(x1[(k + SH) & 31], x1[k], on);
}

}  // namespace seqdiff
