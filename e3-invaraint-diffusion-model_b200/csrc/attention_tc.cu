// attention_tc.cu -- tcgen05 version of the attention core (same math as attention.cu):
//
//   S[l,r] = ( q_l . k_r + q_l . E[l - r + P - 1] ) / 8 + (1 - mask[r]) * -10000 ;  out_l = softmax_r(S[l,:]) @ V
//
// One CTA = 128 query rows of one (graph, head); 128 threads, thread i owns query row i = TMEM lane i.
// Per 128-key block:
//   * TMA brings Q (once), K, the 256-row window of E this (query block, key block) pair can touch, and V into 128B-swizzled
//     smem tiles; one thread issues  S = Q K^T (128x128)  and  QE = Q Ewin^T (128x256)  as tcgen05.mma into TMEM.
//   * The relative-key "skew" S[i,r] += QE[i, i - r + 127] needs no cross-thread traffic in this layout: thread i reads ITS
//     row of QE from TMEM and scatters it into a private smem row at index i + 127 - j, then reads it back aligned with S.
//   * scale / mask / online softmax run thread-locally on the 128 scores of the row (registers); P is written to smem as the
//     K-major A operand of  O (+)= P V  with V as an MN-major B operand (V is [key][d] in memory: d contiguous).
//   * O lives in TMEM across key blocks; it is rescaled in place (tcgen05.ld / st) when the running row maximum moves.
// K/E for block kb+1 are re-fetched as soon as the S/QE MMAs of block kb have completed (under the softmax of block kb); V is
// double-buffered.  The legacy mma.sync kernel (attention.cu) saturates the HMMA pipe at ~160 TFLOP/s on this chip.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace seqdiff {

constexpr int kTQ = 128;       // query rows per CTA (= UMMA M = TMEM lanes)
constexpr int kTK = 128;       // keys per block
constexpr int kTcThreads = 256;  // two threads per query row: warps 0-3 own keys 0..63 of the block, warps 4-7 keys 64..127
constexpr int kPrivPitch = 65;   // floats per private skew row (64 used; odd pitch: conflict-free aligned reads)

template <bool REL> struct TcSmem {
  static constexpr int kQ = 0;                                    // [128][64] 16-bit, SW128
  static constexpr int kK = kQ + 16384;                           // [128][64]
  static constexpr int kV = kK + 16384;                           // 2 x [128][64]
  static constexpr int kE = kV + 2 * 16384;                       // [256][64]
  static constexpr int kP = kE + (REL ? 32768 : 0);               // 2 x [128][64] (keys 0..63 | 64..127)
  static constexpr int kPriv = kP + 32768;                        // [256 threads][65] fp32
  static constexpr int kMask = kPriv + (REL ? kTcThreads * kPrivPitch * 4 : 0);  // 2 x [128] fp32
  static constexpr int kXch = kMask + 2 * kTK * 4;                // [2][128] fp32 row max / row sum exchange
  static constexpr int kBar = kXch + 2 * kTQ * 4;                 // mbarriers + tmem slot
  static constexpr int kBytes = kBar + 128 + 1024;                // + alignment slack
  static constexpr int kTmemCols = REL ? 512 : 256;               // S 128 | QE 256 | O 64  (REL) ; S 128 | O 64
  static constexpr int kColS = 0, kColQE = 128, kColO = REL ? 384 : 128;
};

template <typename T, bool REL>
__global__ void __launch_bounds__(kTcThreads, REL ? 1 : 2) attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                                                                     const float* __restrict__ key_mask, T* __restrict__ out, int heads, int Lq,
                                                                     int Lk, int P, uint32_t fmt) {
  using SM = TcSmem<REL>;
  constexpr float kScale2 = 0.125f * 1.44269504088896f;
  extern __shared__ uint8_t smem_raw[];
  // 1024B-align by OFFSET (not through an integer cast): keeps the pointer provably shared, so accesses compile to LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
  uint64_t* bar_q = bars;       // Q landed
  uint64_t* bar_k = bars + 1;   // K (+E) of the current block landed
  uint64_t* bar_v = bars + 2;   // [2] V buffers
  uint64_t* bar_s = bars + 4;   // S (+QE) MMAs complete
  uint64_t* bar_o = bars + 5;   // PV MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* sMask = reinterpret_cast<float*>(smem + SM::kMask);
  float* xch = reinterpret_cast<float*>(smem + SM::kXch);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3;        // TMEM lane quarter of this warp
  const int hf = warp >> 2;       // which 64-key half of the block this thread owns
  const int row = wq * 32 + lane; // query row inside the CTA tile = TMEM lane
  const int q0 = blockIdx.x * kTQ, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (Lk + kTK - 1) / kTK;
  const uint32_t idesc_s = umma_idesc_16(kTQ, 128, fmt, fmt);
  const uint32_t idesc_e = umma_idesc_16(kTQ, 256, fmt, fmt);
  const uint32_t idesc_o = umma_idesc_16(kTQ, 64, fmt, fmt) | (1u << 16);  // B (= V) is MN-major

  auto load_mask = [&](int kb) {
    if (tid < kTK) {
      const int r = kb * kTK + tid;
      sMask[(kb & 1) * kTK + tid] = (r < Lk) ? (1.0f - key_mask[static_cast<size_t>(b) * Lk + r]) * (-10000.0f * 1.44269504088896f) : -INFINITY;
    }
  };
  auto issue_ke = [&](int kb) {  // thread 0 only
    mbar_expect_tx(bar_k, REL ? 16384 + 32768 : 16384);
    tma_load_2d(smem + SM::kK, &tmK, bar_k, h * 64, b * Lk + kb * kTK);
    if (REL) tma_load_2d(smem + SM::kE, &tmE, bar_k, 0, q0 - kb * kTK + P - 1 - 127);
  };
  auto issue_v = [&](int kb) {  // thread 0 only
    mbar_expect_tx(&bar_v[kb & 1], 16384);
    tma_load_2d(smem + SM::kV + (kb & 1) * 16384, &tmV, &bar_v[kb & 1], h * 64, b * Lk + kb * kTK);
  };

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (REL) tma_prefetch_desc(&tmE);
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, SM::kTmemCols);
    tmem_relinquish();
  }
  pdl_trigger();
  pdl_wait();  // key_mask / q / k / v come from the predecessor kernels
  load_mask(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);

  if (tid == 0) {
    mbar_expect_tx(bar_q, 16384);
    tma_load_2d(smem + SM::kQ, &tmQ, bar_q, h * 64, b * Lq + q0);
    issue_ke(0);
    issue_v(0);
    if (nkb > 1) issue_v(1);
  }

  float m_run = -INFINITY, l_run = 0.f;  // l_run: partial row sum over this thread's key half (same rescaling on both halves)
  for (int kb = 0; kb < nkb; ++kb) {
    const uint32_t ph = kb & 1;
    // ---- S = Q K^T, QE = Q Ewin^T ----
    if (tid == 0) {
      if (kb == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, ph);
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + SM::kQ), ka = smem_u32(smem + SM::kK), ea = smem_u32(smem + SM::kE);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base + SM::kColS, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ka + ks * 32), idesc_s, ks ? 1u : 0u);
      if (REL) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem_base + SM::kColQE, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ea + ks * 32), idesc_e, ks ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    if (kb + 1 < nkb) load_mask(kb + 1);  // other half of sMask; published by this iteration's barriers
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (tid == 0 && kb + 1 < nkb) issue_ke(kb + 1);  // K / E tiles are free again: refill under the softmax

    // ---- scores of (row, key half): S + skewed QE, scaled, masked (log2 domain) ----
    float t[2][32];
    float* prow = reinterpret_cast<float*>(smem + SM::kPriv) + tid * kPrivPitch;
    if (REL) {
      // keys r in [64 hf, 64 hf + 64) pair with window columns j = row + 127 - r; over the warp's 32 rows that is the
      // 3-chunk band starting at chunk  wq + 2 - 2 hf  (warp-uniform), instead of the whole 8-chunk window
      const int cc0 = wq + 2 - 2 * hf;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + SM::kColQE + (cc0 + u) * 32, r);
        tmem_ld_wait();
        const int base = row + 127 - 64 * hf - (cc0 + u) * 32;  // local key index of column jj is  base - jj
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int rr = base - jj;
          if (rr >= 0 && rr < 64) prow[rr] = __uint_as_float(r[jj]);
        }
      }
    }
    const float* mk = sMask + (kb & 1) * kTK + 64 * hf;
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + SM::kColS + (2 * hf + c) * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float sv = __uint_as_float(r[j]);
        if (REL) sv += prow[c * 32 + j];
        sv = fmaf(sv, kScale2, mk[c * 32 + j]);
        t[c][j] = sv;
        mx = fmaxf(mx, sv);
      }
    }
    xch[hf * kTQ + row] = mx;
    tc_fence_before();
    __syncthreads();  // (also: every thread is done reading S / QE of this block from TMEM)
    const float m_new = fmaxf(m_run, fmaxf(mx, xch[(hf ^ 1) * kTQ + row]));
    const float corr = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_new);
    m_run = m_new;

    // ---- O *= corr (in TMEM, this thread's 32 of the 64 columns) once the previous PV has completed ----
    if (kb > 0) {
      mbar_wait(bar_o, (kb - 1) & 1);
      tc_fence_after();
      if (tid == 0 && kb + 1 < nkb) issue_v(kb + 1);  // the V buffer of block kb-1 is free
      uint32_t r[32];
      tmem_ld_32x32(t_lane + SM::kColO + hf * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * corr);
      tmem_st_32x32(t_lane + SM::kColO + hf * 32, r);
      tmem_st_wait();
    }

    // ---- P = exp2(t - m) -> smem (A operand, K-major): key half hf is SW128 tile hf ----
    float rs = 0.f;
    uint8_t* prow16 = smem + SM::kP + hf * 16384 + row * 128;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        float p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          p[j] = ex2_approx(t[c][qd * 8 + j] - m_new);
          rs += p[j];
        }
        const int ch = c * 4 + qd;
        *reinterpret_cast<uint4*>(prow16 + ((ch ^ (row & 7)) << 4)) =
            make_uint4(pack2<T>(p[0], p[1]), pack2<T>(p[2], p[3]), pack2<T>(p[4], p[5]), pack2<T>(p[6], p[7]));
      }
    }
    l_run = l_run * corr + rs;
    fence_proxy_async_smem();  // generic-proxy smem writes (P) -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    // ---- O (+)= P V ----
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&bar_v[kb & 1], (kb >> 1) & 1);
      tc_fence_after();
      const uint32_t pa = smem_u32(smem + SM::kP), va = smem_u32(smem + SM::kV + (kb & 1) * 16384);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem_base + SM::kColO, umma_desc_kmajor_sw128(pa + (ks >> 2) * 16384 + (ks & 3) * 32),
                  umma_desc_kmajor_sw128(va + ks * 2048), idesc_o, (kb | ks) ? 1u : 0u);
      umma_commit(bar_o);
    }
  }

  // ---- O / l -> global: each half-thread stores its 32 of the row's 64 output columns ----
  xch[hf * kTQ + row] = l_run;
  __syncthreads();
  const float inv = 1.0f / (l_run + xch[(hf ^ 1) * kTQ + row]);
  mbar_wait(bar_o, (nkb - 1) & 1);
  tc_fence_after();
  const int grow = q0 + row;
  {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + SM::kColO + hf * 32, r);
    tmem_ld_wait();
    if (grow < Lq) {
      T* orow = out + (static_cast<size_t>(b) * Lq + grow) * (heads * 64) + h * 64 + hf * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        float o8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o8[u] = __uint_as_float(r[j + u]) * inv;
        store8<T>(orow + j, o8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SM::kTmemCols);
  }
}

template <typename T> struct FmtOf;
template <> struct FmtOf<f16> { static constexpr int v = 0; };
template <> struct FmtOf<bf16> { static constexpr int v = 1; };

template <typename T, bool REL>
static int launch_tc_attn(int B, int heads, int Lq, int Lk, const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, const T* E, int P,
                          const float* mask, T* out, cudaStream_t s) {
  using SM = TcSmem<REL>;
  auto kfn = attention_tc_kernel<T, REL>;
  static bool configured = false;
  if (!configured) {
    SD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    configured = true;
  }
  constexpr int fmt = FmtOf<T>::v;
  CUtensorMap tq, tk, tv, te;
  SD_TRY(make_tmap(q, fmt, B * Lq, ldq, 128, &tq));
  SD_TRY(make_tmap(k, fmt, B * Lk, ldk, 128, &tk));
  SD_TRY(make_tmap(v, fmt, B * Lk, ldv, 128, &tv));
  if (REL) SD_TRY(make_tmap(E, fmt, 2 * P - 1, 64, 256, &te));
  else te = tq;
  dim3 grid(ceil_div(Lq, kTQ), heads, B);
  SD_CUDA(launch_k(kfn, dim3(grid), dim3(kTcThreads), SM::kBytes, s, tq, tk, tv, te, mask, out, heads, Lq, Lk, P, static_cast<uint32_t>(fmt)));
  SD_LAUNCHED(REL ? "attention_tc_rel" : "attention_tc_norel", s);
  return SEQDIFF_OK;
}

template <typename T>
int attention_tc(int B, int heads, int Lq, int Lk, const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, const T* dist_emb, int P,
                 const float* key_mask, T* out, cudaStream_t s) {
  SD_CHECK(B > 0 && heads > 0 && Lq > 0 && Lk > 0, "empty attention");
  SD_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "row strides must be multiples of 8 elements");
  SD_CHECK((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) % 16 == 0, "q/k/v must be 16B aligned");
  SD_CHECK(!dist_emb || (Lq <= P && Lk <= P), "sequence longer than max_position_embeddings");
  if (dist_emb) return launch_tc_attn<T, true>(B, heads, Lq, Lk, q, ldq, k, ldk, v, ldv, dist_emb, P, key_mask, out, s);
  return launch_tc_attn<T, false>(B, heads, Lq, Lk, q, ldq, k, ldk, v, ldv, dist_emb, P, key_mask, out, s);
}
template int attention_tc<bf16>(int, int, int, int, const bf16*, int, const bf16*, int, const bf16*, int, const bf16*, int, const float*, bf16*, cudaStream_t);
template int attention_tc<f16>(int, int, int, int, const f16*, int, const f16*, int, const f16*, int, const f16*, int, const float*, f16*, cudaStream_t);

}  // namespace seqdiff
