"""Per-shape timing of the denoiser's GEMM launches at the cfg-2 batch (B=64, L=128) through the C ABI."""
import ctypes, json, math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seqdiff_b200 as sd
lib = sd.lib(); dev = "cuda:0"
H, I = 768, 1024
Mt, Ml, B = 16384, 8192, 64
# (name, count per forward, M, N, K, epi, kind)   kind: 16 -> 16-bit out, r -> fp32 resid+out
SHAPES = [("se_lig.ada0", 1, Mt, H, H, 2, "16"), ("se_lig.ada2", 1, Mt, 6 * H, H, 0, "16"), ("se_lig.qkv", 1, Mt, 3 * H, H, 0, "16"),
          ("se_lig.attn_out", 1, Mt, H, H, 0, "r"), ("se_lig.mlp0", 1, Mt, 4 * H, H, 1, "16"), ("se_lig.mlp3", 1, Mt, H, 4 * H, 0, "r"),
          ("cross_kv_all", 1, Ml, 12 * H, H, 0, "16"), ("dec.qkv", 7, Ml, 3 * H, H, 0, "16"), ("dec.out(+resid)", 13, Ml, H, H, 0, "r"),
          ("dec.cq/p1", 7, Ml, H, H, 0, "16"), ("dec.ffn_up", 6, Ml, I, H, 1, "16"), ("dec.ffn_down", 6, Ml, H, I, 0, "r"),
          ("se_dec.mlp0", 1, Ml, 4 * H, H, 1, "16"), ("se_dec.mlp3", 1, Ml, H, 4 * H, 0, "r"), ("se_dec.ada2", 1, B, 6 * H, H, 0, "16")]
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
tot = 0.0
for name, cnt, M, N, K, epi, kind in SHAPES:
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev); resid = torch.randn(M, N, device=dev) if kind == "r" else None
    C = torch.empty(M, N, device=dev, dtype=torch.float32 if kind == "r" else torch.bfloat16)
    best = {}
    for bn in (128, 192, 256, 1128, 1192, 1256):   # 1xxx = CTA-pair (cta_group::2) kernel with tile width xxx
        cg2, bn = bn // 1000, bn % 1000
        if N % bn: continue
        def call():
            rc = lib.seqdiff_op_gemm(1 | (bn << 8) | (cg2 << 20), M, N, K, p(A), p(W), p(bias), p(resid), epi, p(C), stream)
            assert rc == 0, lib.seqdiff_last_error()
        for _ in range(3): call()
        torch.cuda.synchronize()
        # device time per launch from a captured graph of 20 launches (eager back-to-back launches from Python are
        # CPU-bound below ~15 us per kernel and hide everything the kernel does better than that)
        n = 20
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st2 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(n):
                rc = lib.seqdiff_op_gemm(1 | (bn << 8) | (cg2 << 20), M, N, K, p(A), p(W), p(bias), p(resid), epi, p(C), st2)
                assert rc == 0, lib.seqdiff_last_error()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best[bn + 1000 * cg2] = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * M * N * K
    bb = min(best, key=best.get)
    tot += cnt * best[bb]
    print(f"{name:18s} x{cnt:2d} M={M:6d} N={N:5d} K={K:5d} {kind:>2s}  us/bn: { {k: round(v, 1) for k, v in best.items()} }  best {fl / best[bb] / 1e6:7.1f} TF/s  ({cnt*best[bb]:.0f} us/fwd)")
print(f"sum over one forward (best tile per shape): {tot/1e3:.3f} ms")
