"""Fixed cost per kernel inside a replayed CUDA graph: chains of 40 dependent launches of (a) the tcgen05 GEMM at M = 128 .. 8192,
(b) the LayerNorm kernel at the same row counts -- the floor a 90-kernel sampling step pays per launch."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seqdiff_b200 as sd
lib = sd.lib(); dev = "cuda:0"
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
def timed(fn, n=40):
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn(st); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cs = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(n): fn(cs)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
N, K = 768, 768
W = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16(); bias = torch.randn(N, device=dev)
lw, lb = torch.ones(N, device=dev), torch.zeros(N, device=dev)
for M in (128, 512, 2048, 8192):
    A = torch.randn(M, K, device=dev).bfloat16(); C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    X = torch.randn(M, N, device=dev); O32 = torch.empty(M, N, device=dev); O16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16); stt = torch.empty(M, 2, device=dev)
    def gemm(st): assert lib.seqdiff_op_gemm(1, M, N, K, p(A), p(W), p(bias), None, 0, p(C), st) == 0
    def ln(st): assert lib.seqdiff_op_layernorm(1, M, N, p(X), p(lw), p(lb), 1e-12, p(O32), p(O16), p(stt), st) == 0
    print(f"M={M:5d}: GEMM {N}x{K} {timed(gemm):6.2f} us per launch in a graph ({2e-6*M*N*K/timed(gemm):7.1f} TFLOP/s)   LayerNorm {timed(ln):6.2f} us")
