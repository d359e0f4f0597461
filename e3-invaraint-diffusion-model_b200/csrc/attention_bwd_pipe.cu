// attention_bwd_pipe.cu -- backward of the attention core on tcgen05 (training step, BASELINE configs[3]); same math and interface as
// attention_train_tc.cu (nvcuda::wmma, 20 % issue utilisation at 8 warps per SM: 31 % of the training step) and attention_train.cu (fp32 SIMT,
// the parity reference).  Replaces what autograd derives for HF BertSelfAttention incl. the attention-
// probability dropout; reference call sites sequence_model/model.py:61 and 226-231.
//
// One work item = one (graph, head) with Lq, Lk <= 128: every product of the backward is a single 128-row UMMA tile.
//     phase 1 (tensor core)   S  = Q K^T                     dP~ = dO V^T                                  -> TMEM
//     phase 2 (CUDA cores)    P  = softmax(S / 8 + mask)     keep = Philox dropout mask (regenerated)
//                             D  = rowsum(P * keep * dP~)    dS = P * (keep * dP~ - D) / 8
//                             drop(P) = P * keep and dS -> shared memory, 16-bit, [query][key] rows of 128 B (SW128)
//     phase 3 (tensor core)   dQ = dS K                      dV = drop(P)^T dO          dK = dS^T Q        -> TMEM
//     phase 4 (CUDA cores)    TMEM -> 16-bit -> global (rows past the graph's lengths are not written)
// The [query][key] tiles written in phase 2 are read twice in phase 3: K-major as the A operand of dS K, and MN-major as the A
// operand of the two transposed products (the same bytes: an MN-major SW128 operand is exactly a [k][mn] tile with 128 B rows).
// Q, K, V, dO arrive by TMA and serve as K-major operands in phase 1 and as MN-major B operands in phase 3 -- nothing is
// transposed and nothing of size L x L touches HBM.
// Roles as in attention_pipe.cu: warp 0 TMA producer (operands of item i+1 are in flight while item i is processed), warp 1
// MMA issuer (phase 1 of item i+1 is issued right behind phase 3 of item i, i.e. under item i's epilogue), warps 2..9 two
// threads per row (64 keys each) for phases 2 and 4.
// relative_key (self-attention with the distance embedding) stays on the wmma kernel for now: REL needs the skewed copy of dS
// ([128 x 256], 64 KB) next to these tiles -- see DESIGN.md section 9.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace seqdiff {

namespace {
constexpr int kBT = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 elementwise
constexpr float kL2e = 1.44269504088896f;

struct BwdSmem {
  static constexpr int kIn = 0;                    // 2 stages x (Q | K | V | dO), each [128][64] 16-bit SW128 = 16 KB
  static constexpr int kP = kIn + 2 * 4 * 16384;   // drop(P): 2 tiles [128 q][64 keys]
  static constexpr int kdS = kP + 32768;           // dS: 2 tiles
  static constexpr int kXch = kdS + 32768;         // [3 exchanges][2 halves][128] fp32
  static constexpr int kMask = kXch + 3 * 2 * 128 * 4;  // [128] additive mask (log2 domain)
  static constexpr int kBar = kMask + 128 * 4;
  static constexpr int kBytes = kBar + 256 + 1024;
};
constexpr int kColS = 0, kColdP = 128, kColdQ = 256, kColdV = 320, kColdK = 384;

__device__ __forceinline__ void ew_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace

template <typename T>
__global__ void __launch_bounds__(kBT, 1)
attention_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmdO, const float* __restrict__ key_mask, int heads, int Lq, int Lk, uint32_t fmt,
                          int n_items, const DropSpec dr, T* __restrict__ dq, int lddq, T* __restrict__ dk, int lddk, T* __restrict__ dv, int lddv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* in_full = bars;       // [2]
  uint64_t* in_empty = bars + 2;  // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* s_empty = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + BwdSmem::kXch);
  float* sMask = reinterpret_cast<float*>(smem + BwdSmem::kMask);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = warp_id_uniform();
  const int my_items = (static_cast<int>(blockIdx.x) < n_items) ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);
    mbar_init(p_ready, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------ TMA producer ------------------------------------------
    for (int i = 0; i < my_items; ++i) {
      const int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int h = item % heads, b = item / heads;
      const int st = i & 1;
      mbar_wait(&in_empty[st], ((i >> 1) & 1) ^ 1);
      mbar_expect_tx_e(&in_full[st], 4 * 16384);
      uint8_t* base = smem + BwdSmem::kIn + st * 65536;
      tma_load_2d_e(base, &tmQ, &in_full[st], h * 64, b * Lq);
      tma_load_2d_e(base + 16384, &tmK, &in_full[st], h * 64, b * Lk);
      tma_load_2d_e(base + 32768, &tmV, &in_full[st], h * 64, b * Lk);
      tma_load_2d_e(base + 49152, &tmdO, &in_full[st], h * 64, b * Lq);
    }
  } else if (warp == 1) {
    // ------------------------------------------ MMA issuer --------------------------------------------
    const uint32_t idesc_s = umma_idesc_16(128, 128, fmt, fmt);                                    // A, B K-major
    const uint32_t idesc_q = umma_idesc_16(128, 64, fmt, fmt) | kUmmaBMnMajor;                     // dQ = dS K: A K-major, B = K MN-major
    const uint32_t idesc_t = umma_idesc_16(128, 64, fmt, fmt) | kUmmaAMnMajor | kUmmaBMnMajor;     // dV, dK: A^T products
    auto phase1 = [&](int i) {
      const int st = i & 1;
      mbar_wait(&in_full[st], (i >> 1) & 1);
      mbar_wait(s_empty, (i & 1) ^ 1);  // the elementwise threads have drained S / dP of item i-1
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + BwdSmem::kIn + st * 65536), ka = qa + 16384, va = qa + 32768, oa = qa + 49152;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_e(tmem_base + kColS, umma_desc_kmajor_sw128(qa + k * 32), umma_desc_kmajor_sw128(ka + k * 32), idesc_s, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_e(tmem_base + kColdP, umma_desc_kmajor_sw128(oa + k * 32), umma_desc_kmajor_sw128(va + k * 32), idesc_s, k ? 1u : 0u);
      umma_commit_e(s_full);
    };
    if (my_items > 0) phase1(0);
    for (int i = 0; i < my_items; ++i) {
      const int st = i & 1;
      mbar_wait(p_ready, i & 1);
      mbar_wait(o_empty, (i & 1) ^ 1);  // the epilogue of item i-1 has read its outputs
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + BwdSmem::kIn + st * 65536), ka = qa + 16384, oa = qa + 49152;
      const uint32_t pa = smem_u32(smem + BwdSmem::kP), sa = smem_u32(smem + BwdSmem::kdS);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dQ[q, d] = sum_key dS[q, key] K[key, d]: k-step = 16 keys
        umma_bf16_e(tmem_base + kColdQ, umma_desc_kmajor_sw128(sa + (k >> 2) * 16384 + (k & 3) * 32), umma_desc_mnmajor_sw128(ka + k * 2048, 16384), idesc_q,
                    k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dV[key, d] = sum_q drop(P)[q, key] dO[q, d]: k-step = 16 queries; A = P^T: MN blocks = the two key tiles
        umma_bf16_e(tmem_base + kColdV, umma_desc_mnmajor_sw128(pa + k * 2048, 16384), umma_desc_mnmajor_sw128(oa + k * 2048, 16384), idesc_t, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dK[key, d] = sum_q dS[q, key] Q[q, d]
        umma_bf16_e(tmem_base + kColdK, umma_desc_mnmajor_sw128(sa + k * 2048, 16384), umma_desc_mnmajor_sw128(qa + k * 2048, 16384), idesc_t, k ? 1u : 0u);
      umma_commit_e(o_full);
      umma_commit_e(&in_empty[st]);
      if (i + 1 < my_items) phase1(i + 1);
    }
  } else {
    // ------------------------------------------ elementwise threads -----------------------------------
    const int st_ = tid - 64;         // 0..255
    const int wq = warp & 3;          // TMEM lane quarter of this warp
    const int hf = (warp - 2) >> 2;   // key half (phase 2) / column half of the 64 output columns (phase 4)
    const int row = wq * 32 + lane;   // TMEM lane: query row in phases 1-2, key row for dV / dK in phase 4
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t thr = dr.p > 0.f ? static_cast<uint32_t>(static_cast<double>(dr.p) * 4294967296.0) : 0u;
    const float sc = dr.p > 0.f ? 1.0f / (1.0f - dr.p) : 1.0f;
    constexpr float kScale2 = 0.125f * kL2e;
    for (int i = 0; i < my_items; ++i) {
      const int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int h = item % heads, b = item / heads;
      if (st_ < 128) sMask[st_] = st_ < Lk ? (1.0f - __ldg(key_mask + static_cast<size_t>(b) * Lk + st_)) * (-10000.0f * kL2e) : -INFINITY;
      ew_bar();
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      const bool live = row < Lq;
      // ---- P of (row, key half) ----
      float p[64];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColS + hf * 64 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sv = fmaf(__uint_as_float(r[j]), kScale2, sMask[hf * 64 + c * 32 + j]);
          p[c * 32 + j] = sv;
          mx = fmaxf(mx, sv);
        }
      }
      xch[(0 * 2 + hf) * 128 + row] = mx;
      ew_bar();
      mx = fmaxf(mx, xch[(0 * 2 + (hf ^ 1)) * 128 + row]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        p[j] = ex2_approx(p[j] - mx);  // keys >= Lk carry -inf -> 0 (key 0 is always in range: mx is finite)
        sum += p[j];
      }
      xch[(1 * 2 + hf) * 128 + row] = sum;
      ew_bar();
      sum += xch[(1 * 2 + (hf ^ 1)) * 128 + row];
      const float inv = live ? 1.0f / sum : 0.f;  // rows past the graph's length contribute nothing to dK / dV
      // ---- dropout mask of this thread's 64 keys: one bit per key (regenerated from the forward's Philox stream) ----
      uint64_t keepbits = ~0ull;
      if (dr.p > 0.f) {
        keepbits = 0ull;
        const size_t e_row = ((static_cast<size_t>(b) * heads + h) * Lq + row) * Lk + 64 * hf;
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const uint64_t qd = (e_row + 4 * j4) >> 2;
          uint32_t w4[4] = {static_cast<uint32_t>(qd), static_cast<uint32_t>(qd >> 32), dr.site, dr.step};
          philox4x32_10(w4, static_cast<uint32_t>(dr.seed), static_cast<uint32_t>(dr.seed >> 32));
#pragma unroll
          for (int u = 0; u < 4; ++u) keepbits |= static_cast<uint64_t>(w4[u] >= thr ? 1u : 0u) << (4 * j4 + u);
        }
      }
      // ---- D = sum_key P keep dP~ ----
      float dsum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColdP + hf * 64 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          p[c * 32 + j] *= inv;
          const float kp = ((keepbits >> (c * 32 + j)) & 1ull) ? sc : 0.f;
          dsum = fmaf(p[c * 32 + j] * kp, __uint_as_float(r[j]), dsum);
        }
      }
      xch[(2 * 2 + hf) * 128 + row] = dsum;
      ew_bar();
      dsum += xch[(2 * 2 + (hf ^ 1)) * 128 + row];
      // ---- dS, drop(P) -> shared memory (tile hf, row `row`: 128 B = 8 chunks of 16 B, XOR-swizzled by row & 7) ----
      uint8_t* prow = smem + BwdSmem::kP + hf * 16384 + row * 128;
      uint8_t* srow = smem + BwdSmem::kdS + hf * 16384 + row * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColdP + hf * 64 + c * 32, r);
        tmem_ld_wait();
        uint32_t pp[16], ds[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float k0 = ((keepbits >> (c * 32 + j)) & 1ull) ? sc : 0.f, k1 = ((keepbits >> (c * 32 + j + 1)) & 1ull) ? sc : 0.f;
          const float p0 = p[c * 32 + j], p1 = p[c * 32 + j + 1];
          pp[j >> 1] = pack2<T>(p0 * k0, p1 * k1);
          ds[j >> 1] = pack2<T>(p0 * (__uint_as_float(r[j]) * k0 - dsum) * 0.125f, p1 * (__uint_as_float(r[j + 1]) * k1 - dsum) * 0.125f);
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int chunk = ((c * 4 + qd) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + chunk) = make_uint4(pp[4 * qd], pp[4 * qd + 1], pp[4 * qd + 2], pp[4 * qd + 3]);
          *reinterpret_cast<uint4*>(srow + chunk) = make_uint4(ds[4 * qd], ds[4 * qd + 1], ds[4 * qd + 2], ds[4 * qd + 3]);
        }
      }
      // S / dP are in registers / shared memory now: phase 1 of the next item may overwrite them
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes (P, dS) -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_ready);
      }
      // ---- phase 4: outputs ----
      mbar_wait(o_full, i & 1);
      tc_fence_after();
      auto store_rows = [&](int col0, T* dst_base, int ld, int n_rows, int L) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + col0 + hf * 32, r);
        tmem_ld_wait();
        if (row < n_rows) {
          T* dst = dst_base + (static_cast<size_t>(b) * L + row) * ld + h * 64 + hf * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(dst + 8 * j) =
                make_uint4(pack2<T>(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])), pack2<T>(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])),
                           pack2<T>(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), pack2<T>(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
        }
      };
      store_rows(kColdQ, dq, lddq, Lq, Lq);
      store_rows(kColdV, dv, lddv, Lk, Lk);
      store_rows(kColdK, dk, lddk, Lk, Lk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename T> struct BwdFmt;
template <> struct BwdFmt<f16> { static constexpr int v = 0; };
template <> struct BwdFmt<bf16> { static constexpr int v = 1; };

// usable(...) = the shapes this kernel covers; everything else stays on attention_bwd_tc (wmma)
bool attention_bwd_pipe_usable(int Lq, int Lk, const void* dist_emb, float p_drop) {
  static const bool off = [] { const char* e = getenv("SEQDIFF_TRAIN_ATTN"); return e && (std::string(e) == "wmma" || std::string(e) == "simt"); }();
  return !off && dist_emb == nullptr && Lq <= 128 && Lk <= 128 && Lq >= 1 && Lk >= 1 && (p_drop <= 0.f || Lk % 4 == 0);
}

template <typename T>
int attention_bwd_pipe(int B, int heads, int Lq, int Lk, const T* q, int ldq, const T* k, int ldk, const T* v, int ldv, const float* key_mask, DropSpec dr,
                       const T* dout, T* dq, int lddq, T* dk, int lddk, T* dv, int lddv, cudaStream_t s) {
  SD_CHECK(B > 0 && heads > 0 && Lq >= 1 && Lk >= 1 && Lq <= 128 && Lk <= 128, "attention_bwd_pipe: one 128-row tile per (graph, head)");
  SD_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "row strides must be multiples of 8 elements");
  SD_CHECK(dr.p <= 0.f || Lk % 4 == 0, "dropout: Lk must be a multiple of 4");
  SD_CHECK(dout && dq && dk && dv && key_mask, "attention_bwd_pipe: null argument");
  auto kfn = attention_bwd_pipe_kernel<T>;
  static bool configured = false;
  if (!configured) {
    SD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kBytes));
    configured = true;
  }
  constexpr int fmt = BwdFmt<T>::v;
  CUtensorMap tq, tk, tv, to;
  SD_TRY(make_tmap(q, fmt, B * Lq, ldq, 128, &tq));
  SD_TRY(make_tmap(k, fmt, B * Lk, ldk, 128, &tk));
  SD_TRY(make_tmap(v, fmt, B * Lk, ldv, 128, &tv));
  SD_TRY(make_tmap(dout, fmt, B * Lq, heads * 64, 128, &to));
  const int n_items = B * heads;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  SD_CUDA(launch_k(kfn, dim3(grid), dim3(kBT), BwdSmem::kBytes, s, tq, tk, tv, to, key_mask, heads, Lq, Lk, static_cast<uint32_t>(fmt), n_items, dr, dq, lddq, dk,
                   lddk, dv, lddv));
  SD_LAUNCHED("attention_bwd_pipe", s);
  return SEQDIFF_OK;
}
template int attention_bwd_pipe<bf16>(int, int, int, int, const bf16*, int, const bf16*, int, const bf16*, int, const float*, DropSpec, const bf16*, bf16*, int,
                                      bf16*, int, bf16*, int, cudaStream_t);
template int attention_bwd_pipe<f16>(int, int, int, int, const f16*, int, const f16*, int, const f16*, int, const float*, DropSpec, const f16*, f16*, int, f16*,
                                     int, f16*, int, cudaStream_t);

}  // namespace seqdiff
