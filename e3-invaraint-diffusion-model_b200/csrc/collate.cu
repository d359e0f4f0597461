// collate.cu -- LigandBindingSiteDataset.__getitem__ (sequence_model/dataset.py:97-129) for a whole batch of
// complexes on the GPU: pocket-mask dilation by exactly +-ext (torch.roll semantics incl. wrap-around, only index 0 /
// n-1 protected: quirk Q9), boolean-mask compaction of ligand / extended-pocket rows in residue order, zero padding to
// max_len, prefix-ones attention masks.  Complexes are stored ragged (CSR-style node offsets).  One CTA per complex;
// order-preserving compaction = block-wide exclusive scan of the mask (ballot + popc per warp, warp totals in smem).
// Integer / copy work only: bit-exact against the reference.  Traffic: 2*(8+20+1)*4*L bytes written per complex.
#include "common.cuh"
#include "kernels.h"

namespace seqdiff {

constexpr int kColThreads = 256;

__device__ __forceinline__ int block_excl_scan(bool flag, int* warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kColThreads / 32; ++w) {
    const int t = warp_tot[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + within;
}

__device__ __forceinline__ void copy_row(const float* __restrict__ ang, const float* __restrict__ aa, size_t src, float* __restrict__ o_ang,
                                         float* __restrict__ o_seq, size_t dst) {
  const float4* a = reinterpret_cast<const float4*>(ang + src * 8);
  float4* oa = reinterpret_cast<float4*>(o_ang + dst * 8);
  oa[0] = a[0]; oa[1] = a[1];
  const float4* s = reinterpret_cast<const float4*>(aa + src * 20);
  float4* os = reinterpret_cast<float4*>(o_seq + dst * 20);
#pragma unroll
  for (int i = 0; i < 5; ++i) os[i] = s[i];
}

__global__ void __launch_bounds__(kColThreads) collate_kernel(const int* __restrict__ offsets, const uint8_t* __restrict__ lig_mask,
                                                              const uint8_t* __restrict__ poc_mask, const float* __restrict__ ang,
                                                              const float* __restrict__ aa, int ext, int L, float* __restrict__ lig_ang,
                                                              float* __restrict__ lig_seq, float* __restrict__ lig_attn,
                                                              float* __restrict__ rec_ang, float* __restrict__ rec_seq,
                                                              float* __restrict__ rec_attn, int* __restrict__ lengths) {
  pdl_trigger();
  pdl_wait();  // predecessor grid complete + flushed before any dependent global access
  __shared__ int warp_tot[kColThreads / 32];
  const int g = blockIdx.x;
  const int n0 = offsets[g], n = offsets[g + 1] - n0;
  const size_t out0 = static_cast<size_t>(g) * L;
  int nl = 0, nr = 0;
  for (int base = 0; base < n; base += kColThreads) {
    const int i = base + threadIdx.x;
    bool fl = false, fp = false;
    if (i < n) {
      fl = lig_mask[n0 + i] != 0;
      // roll(pocket, +ext)[i] = pocket[(i - ext) mod n], cleared at i == 0; roll(pocket, -ext)[i] = pocket[(i + ext) mod n], cleared at i == n-1
      const int e = ext % n;
      const int il = ((i - e) % n + n) % n, ir = (i + e) % n;
      fp = poc_mask[n0 + i] != 0 || (i != 0 && poc_mask[n0 + il] != 0) || (i != n - 1 && poc_mask[n0 + ir] != 0);
    }
    int tl, tp;
    const int pl = block_excl_scan(fl, warp_tot, tl);
    const int pp = block_excl_scan(fp, warp_tot, tp);
    if (fl && nl + pl < L) copy_row(ang, aa, n0 + i, lig_ang, lig_seq, out0 + nl + pl);
    if (fp && nr + pp < L) copy_row(ang, aa, n0 + i, rec_ang, rec_seq, out0 + nr + pp);
    nl += tl;
    nr += tp;
  }
  // zero padding (dataset.py:41-49) and prefix-ones masks (dataset.py:110-114)
  for (int r = threadIdx.x; r < L; r += kColThreads) {
    lig_attn[out0 + r] = r < nl ? 1.0f : 0.0f;
    rec_attn[out0 + r] = r < nr ? 1.0f : 0.0f;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= nl) {
      float4* oa = reinterpret_cast<float4*>(lig_ang + (out0 + r) * 8);
      float4* os = reinterpret_cast<float4*>(lig_seq + (out0 + r) * 20);
      oa[0] = z; oa[1] = z;
      for (int k = 0; k < 5; ++k) os[k] = z;
    }
    if (r >= nr) {
      float4* oa = reinterpret_cast<float4*>(rec_ang + (out0 + r) * 8);
      float4* os = reinterpret_cast<float4*>(rec_seq + (out0 + r) * 20);
      oa[0] = z; oa[1] = z;
      for (int k = 0; k < 5; ++k) os[k] = z;
    }
  }
  if (threadIdx.x == 0) {
    lengths[2 * g] = nl;      // > L means the reference would raise RuntimeError("Length exceed") (dataset.py:42-43)
    lengths[2 * g + 1] = nr;
  }
}

int collate(int G, const int* offsets, const uint8_t* lig_mask, const uint8_t* poc_mask, const float* ang, const float* aa, int ext, int L,
            float* lig_ang, float* lig_seq, float* lig_attn, float* rec_ang, float* rec_seq, float* rec_attn, int* lengths, cudaStream_t s) {
  SD_CHECK(G > 0 && L > 0, "empty collation");
  SD_CUDA(launch_k(collate_kernel, dim3(G), dim3(kColThreads), 0, s, offsets, lig_mask, poc_mask, ang, aa, ext, L, lig_ang, lig_seq, lig_attn, rec_ang, rec_seq, rec_attn,
                                          lengths));
  SD_LAUNCHED("collate", s);
  return SEQDIFF_OK;
}

}  // namespace seqdiff
