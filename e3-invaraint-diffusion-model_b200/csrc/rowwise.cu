// rowwise.cu -- the HBM-bound per-token kernels of the denoiser: timestep features, BertEmbeddings
// (Linear + LayerNorm), post-LN, the SELayer adaLN residual update, and the predictor tail.
// One warp owns one token row (H = 256*VPL features, 8 contiguous elements per lane per vector => 16 B
// (bf16) / 32 B (f32) per lane, fully coalesced); LayerNorm statistics are always fp32, two-pass
// (mean, then centred variance) like ATen's CPU kernel.
#include "common.cuh"
#include "kernels.h"

namespace seqdiff {

constexpr int kRowThreads = 256;  // 8 warps = 8 rows per CTA

template <typename T, int VPL>
__device__ __forceinline__ void load_row(const T* __restrict__ row, int lane, float (&v)[VPL][8]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) load8<T>(row + (i * 32 + lane) * 8, v[i]);
}
template <typename T, int VPL>
__device__ __forceinline__ void store_row(T* __restrict__ row, int lane, const float (&v)[VPL][8]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) store8<T>(row + (i * 32 + lane) * 8, v[i]);
}
template <int VPL>
__device__ __forceinline__ void row_stats(const float (&v)[VPL][8], int H, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  mean = warp_sum(s) / static_cast<float>(H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      q = fmaf(d, d, q);
    }
  const float var = warp_sum(q) / static_cast<float>(H);
  rstd = 1.0f / sqrtf(var + eps);
}

// ---------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ timestep, const int* __restrict__ step_ptr,
                                      const float* __restrict__ W, int B, int H, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  const int half = H / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i % half;
  const float t = step_ptr ? static_cast<float>(*step_ptr) : timestep[b];
  // x[:, None] * W[None, :] * 2 * torch.pi  -> three separately rounded fp32 multiplies (model.py:95)
  const float x = __fmul_rn(__fmul_rn(__fmul_rn(t, W[j]), 2.0f), 3.14159274101257324f);
  out[static_cast<size_t>(b) * H + j] = sinf(x);
  out[static_cast<size_t>(b) * H + half + j] = cosf(x);
}

int timestep_embed(const float* timestep, const int* step_ptr, const float* W, int B, int H, float* out, cudaStream_t s) {
  const int n = B * (H / 2);
  SD_CUDA(launch_k(timestep_embed_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, s, timestep, step_ptr, W, B, H, out));
  SD_LAUNCHED("timestep_embed", s);
  return SEQDIFF_OK;
}

// ---------------------------------------------------------------------------------------------------
// BertEmbeddings: 4 tokens per warp pass so each weight vector load is reused 4x.
template <typename T, int VPL>
__global__ void __launch_bounds__(kRowThreads) embed_ln_kernel(const float* __restrict__ x, int M, int fin,
                                                               const float* __restrict__ Wt, const float* __restrict__ bias,
                                                               const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                               float eps, const float* __restrict__ te, int L, int H,
                                                               float* __restrict__ out32, T* __restrict__ outT) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  constexpr int TOK = 4;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kRowThreads + threadIdx.x) >> 5;
  const int tok0 = warp * TOK;
  if (tok0 >= M) return;
  float acc[TOK][VPL][8];
#pragma unroll
  for (int t = 0; t < TOK; ++t)
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][i][j] = 0.f;
  for (int k = 0; k < fin; ++k) {
    float xv[TOK];
#pragma unroll
    for (int t = 0; t < TOK; ++t) xv[t] = (tok0 + t < M) ? __ldg(x + static_cast<size_t>(tok0 + t) * fin + k) : 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float w[8];
      load8<float>(Wt + static_cast<size_t>(k) * H + (i * 32 + lane) * 8, w);
#pragma unroll
      for (int t = 0; t < TOK; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][i][j] = fmaf(xv[t], w[j], acc[t][i][j]);
    }
  }
#pragma unroll
  for (int t = 0; t < TOK; ++t) {
    if (tok0 + t >= M) break;
    float v[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float b8[8];
      load8<float>(bias + (i * 32 + lane) * 8, b8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = acc[t][i][j] + b8[j];
    }
    float mean, rstd;
    row_stats<VPL>(v, H, eps, mean, rstd);
    const float* terow = te ? te + static_cast<size_t>((tok0 + t) / L) * H : nullptr;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float g8[8], b8[8];
      load8<float>(lnw + (i * 32 + lane) * 8, g8);
      load8<float>(lnb + (i * 32 + lane) * 8, b8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * g8[j] + b8[j];
      if (terow) {
        float t8[8];
        load8<float>(terow + (i * 32 + lane) * 8, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += t8[j];
      }
    }
    if (out32) store_row<float, VPL>(out32 + static_cast<size_t>(tok0 + t) * H, lane, v);
    if (outT) store_row<T, VPL>(outT + static_cast<size_t>(tok0 + t) * H, lane, v);
  }
}

#define SD_VPL_DISPATCH(H_, ...)                                        \
  switch ((H_) / 256) {                                                  \
    case 1: { constexpr int VPL = 1; __VA_ARGS__; } break;                      \
    case 2: { constexpr int VPL = 2; __VA_ARGS__; } break;                      \
    case 3: { constexpr int VPL = 3; __VA_ARGS__; } break;                      \
    case 4: { constexpr int VPL = 4; __VA_ARGS__; } break;                      \
    default: set_error("hidden_size must be 256, 512, 768 or 1024"); return SEQDIFF_ERR_INVALID; \
  }

template <typename T>
int embed_ln(const float* x, int M, int fin, const float* Wt, const float* b, const float* lnw, const float* lnb, float eps,
             const float* te, int L, int H, float* out32, T* outT, cudaStream_t s) {
  SD_CHECK(H % 256 == 0, "hidden_size must be a multiple of 256");
  const int warps = ceil_div(M, 4);
  const int grid = ceil_div(warps * 32, kRowThreads);
  SD_VPL_DISPATCH(H, SD_CUDA(launch_k(embed_ln_kernel<T, VPL>, dim3(grid), dim3(kRowThreads), 0, s, x, M, fin, Wt, b, lnw, lnb, eps, te, L, H, out32, outT)));
  SD_LAUNCHED("embed_ln", s);
  return SEQDIFF_OK;
}
#define SD_INST_EMBED(T) \
  template int embed_ln<T>(const float*, int, int, const float*, const float*, const float*, const float*, float, const float*, int, int, float*, T*, cudaStream_t)
SD_INST_EMBED(float);
SD_INST_EMBED(bf16);
SD_INST_EMBED(f16);

// ---------------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(kRowThreads) layernorm_kernel(const float* __restrict__ in, int M, int H, const float* __restrict__ w,
                                                                const float* __restrict__ b, float eps, float* __restrict__ out32,
                                                                T* __restrict__ outT, float2* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * kRowThreads + threadIdx.x) >> 5;
  if (row >= M) return;
  float v[VPL][8];
  load_row<float, VPL>(in + static_cast<size_t>(row) * H, lane, v);
  float mean, rstd;
  row_stats<VPL>(v, H, eps, mean, rstd);
  if (stats && lane == 0) stats[row] = make_float2(mean, rstd);  // lets a consumer re-derive LN(in) without the fp32 copy
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float g8[8], b8[8];
    load8<float>(w + (i * 32 + lane) * 8, g8);
    load8<float>(b + (i * 32 + lane) * 8, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * g8[j] + b8[j];
  }
  if (out32) store_row<float, VPL>(out32 + static_cast<size_t>(row) * H, lane, v);
  if (outT) store_row<T, VPL>(outT + static_cast<size_t>(row) * H, lane, v);
}

template <typename T>
int layernorm(const float* in, int M, int H, const float* w, const float* b, float eps, float* out32, T* outT, float2* stats, cudaStream_t s) {
  const int grid = ceil_div(M * 32, kRowThreads);
  SD_VPL_DISPATCH(H, SD_CUDA(launch_k(layernorm_kernel<T, VPL>, dim3(grid), dim3(kRowThreads), 0, s, in, M, H, w, b, eps, out32, outT, stats)));
  SD_LAUNCHED("layernorm", s);
  return SEQDIFF_OK;
}
#define SD_INST_LN(T) template int layernorm<T>(const float*, int, int, const float*, const float*, float, float*, T*, float2*, cudaStream_t)
SD_INST_LN(float);
SD_INST_LN(bf16);
SD_INST_LN(f16);

// ---------------------------------------------------------------------------------------------------
template <typename T, int VPL, bool AFFINE_FIRST>
__global__ void __launch_bounds__(kRowThreads) ln_modulate_kernel(const float* __restrict__ in, int M, int H, const float* __restrict__ lnw,
                                                                  const float* __restrict__ lnb, float eps1, const float* __restrict__ x,
                                                                  const T* __restrict__ mod, int mod_div, int chunk0,
                                                                  float* __restrict__ out32, T* __restrict__ outT) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * kRowThreads + threadIdx.x) >> 5;
  if (row >= M) return;
  float v[VPL][8];
  load_row<float, VPL>(in + static_cast<size_t>(row) * H, lane, v);
  float mean, rstd;
  if (AFFINE_FIRST) {  // BertSelfOutput.LayerNorm (eps 1e-12, affine) -- output of self.attn(x, mask)[0]
    row_stats<VPL>(v, H, eps1, mean, rstd);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float g8[8], b8[8];
      load8<float>(lnw + (i * 32 + lane) * 8, g8);
      load8<float>(lnb + (i * 32 + lane) * 8, b8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * g8[j] + b8[j];
    }
  }
  row_stats<VPL>(v, H, 1e-5f, mean, rstd);  // SELayer.norm1/norm2: elementwise_affine=False, default eps
  const T* mrow = mod + static_cast<size_t>(row / mod_div) * (6 * H);
  const float* xrow = x + static_cast<size_t>(row) * H;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = (i * 32 + lane) * 8;
    float sh[8], sc[8], gt[8], xr[8];
    load8<T>(mrow + (chunk0 + 0) * H + e, sh);
    load8<T>(mrow + (chunk0 + 1) * H + e, sc);
    load8<T>(mrow + (chunk0 + 2) * H + e, gt);
    load8<float>(xrow + e, xr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = (v[i][j] - mean) * rstd;
      v[i][j] = xr[j] + gt[j] * (n * (1.0f + sc[j]) + sh[j]);
    }
  }
  if (out32) store_row<float, VPL>(out32 + static_cast<size_t>(row) * H, lane, v);
  if (outT) store_row<T, VPL>(outT + static_cast<size_t>(row) * H, lane, v);
}

template <typename T>
int ln_modulate(const float* in, int M, int H, bool affine_first, const float* lnw, const float* lnb, float eps1, const float* x,
                const T* mod, int mod_div, int chunk0, float* out32, T* outT, cudaStream_t s) {
  const int grid = ceil_div(M * 32, kRowThreads);
  if (affine_first) {
    SD_VPL_DISPATCH(H, SD_CUDA(launch_k(ln_modulate_kernel<T, VPL, true>, dim3(grid), dim3(kRowThreads), 0, s, in, M, H, lnw, lnb, eps1, x, mod, mod_div, chunk0, out32, outT)));
  } else {
    SD_VPL_DISPATCH(H, SD_CUDA(launch_k(ln_modulate_kernel<T, VPL, false>, dim3(grid), dim3(kRowThreads), 0, s, in, M, H, lnw, lnb, eps1, x, mod, mod_div, chunk0, out32, outT)));
  }
  SD_LAUNCHED("ln_modulate", s);
  return SEQDIFF_OK;
}
#define SD_INST_LNMOD(T) \
  template int ln_modulate<T>(const float*, int, int, bool, const float*, const float*, float, const float*, const T*, int, int, float*, T*, cudaStream_t)
SD_INST_LNMOD(float);
SD_INST_LNMOD(bf16);
SD_INST_LNMOD(f16);

// ---------------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(kRowThreads) predictor_tail_kernel(const T* __restrict__ y, int M, int H, const float* __restrict__ lnw,
                                                                     const float* __restrict__ lnb, float eps,
                                                                     const float* __restrict__ W2, const float* __restrict__ b2, int F,
                                                                     float* __restrict__ logits) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * kRowThreads + threadIdx.x) >> 5;
  if (row >= M) return;
  float v[VPL][8];
  load_row<T, VPL>(y + static_cast<size_t>(row) * H, lane, v);
  float mean, rstd;
  row_stats<VPL>(v, H, eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float g8[8], b8[8];
    load8<float>(lnw + (i * 32 + lane) * 8, g8);
    load8<float>(lnb + (i * 32 + lane) * 8, b8);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd * g8[j] + b8[j];
  }
  float mine = 0.f;  // lane f keeps logit f
  for (int f = 0; f < F; ++f) {
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float w[8];
      load8<float>(W2 + static_cast<size_t>(f) * H + (i * 32 + lane) * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) p = fmaf(v[i][j], w[j], p);
    }
    p = warp_sum(p);
    if (lane == (f & 31)) {
      mine = p + b2[f];
      if (f >= 32) logits[static_cast<size_t>(row) * F + f] = mine;  // F > 32 never happens for this model
    }
  }
  if (lane < F && lane < 32) logits[static_cast<size_t>(row) * F + lane] = mine;
}

template <typename T>
int predictor_tail(const T* y, int M, int H, const float* lnw, const float* lnb, float eps, const float* W2, const float* b2, int F,
                   float* logits, cudaStream_t s) {
  SD_CHECK(F <= 32, "feature_size > 32 not supported");
  const int grid = ceil_div(M * 32, kRowThreads);
  SD_VPL_DISPATCH(H, SD_CUDA(launch_k(predictor_tail_kernel<T, VPL>, dim3(grid), dim3(kRowThreads), 0, s, y, M, H, lnw, lnb, eps, W2, b2, F, logits)));
  SD_LAUNCHED("predictor_tail", s);
  return SEQDIFF_OK;
}
template int predictor_tail<float>(const float*, int, int, const float*, const float*, float, const float*, const float*, int, float*, cudaStream_t);
template int predictor_tail<bf16>(const bf16*, int, int, const float*, const float*, float, const float*, const float*, int, float*, cudaStream_t);
template int predictor_tail<f16>(const f16*, int, int, const float*, const float*, float, const float*, const float*, int, float*, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void f32_to_16_kernel(const float* __restrict__ in, size_t n, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = from_f32<T>(in[i]);
}
template <typename T>
int f32_to_16(const float* in, size_t n, T* out, cudaStream_t s) {
  if (n == 0) return SEQDIFF_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  SD_CUDA(launch_k(f32_to_16_kernel<T>, dim3(static_cast<int>(blocks)), dim3(256), 0, s, in, n, out));
  SD_LAUNCHED("f32_to_16", s);
  return SEQDIFF_OK;
}
template int f32_to_16<bf16>(const float*, size_t, bf16*, cudaStream_t);
template int f32_to_16<f16>(const float*, size_t, f16*, cudaStream_t);

__global__ void transpose_f32_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  out[static_cast<size_t>(c) * rows + r] = in[i];
}
int transpose_f32(const float* in, int rows, int cols, float* out, cudaStream_t s) {
  SD_CUDA(launch_k(transpose_f32_kernel, dim3(ceil_div(rows * cols, 256)), dim3(256), 0, s, in, rows, cols, out));
  SD_LAUNCHED("transpose_f32", s);
  return SEQDIFF_OK;
}

__global__ void step_advance_kernel(int* p) {
  pdl_trigger();
  pdl_wait();
  *p -= 1;
}
int step_advance(int* step_ptr, cudaStream_t s) {
  SD_CUDA(launch_k(step_advance_kernel, dim3(1), dim3(1), 0, s, step_ptr));
  SD_LAUNCHED("step_advance", s);
  return SEQDIFF_OK;
}

}  // namespace seqdiff
