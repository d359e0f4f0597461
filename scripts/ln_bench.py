"""LayerNorm kernel alone (post-LN of the decoder: fp32 in, bf16 out + row statistics), 20 launches in a captured graph."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seqdiff_b200 as sd
lib = sd.lib(); dev = "cuda:0"; p = sd._cabi.ptr
H = 768
w = torch.randn(H, device=dev); b = torch.randn(H, device=dev)
for M in (1024, 4096, 8192, 16384):
    x = torch.randn(M, H, device=dev)
    y = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    st = torch.empty(M, 2, device=dev)
    call = lambda s: lib.seqdiff_op_layernorm(1, M, H, p(x), p(w), p(b), 1e-12, None, p(y), p(st), s)
    s0 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): assert call(s0) == 0
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    n = 20
    with torch.cuda.graph(gr):
        s1 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(n): assert call(s1) == 0
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    ref = torch.nn.functional.layer_norm(x, (H,), w, b, 1e-12)
    err = (y.float() - ref).abs().max().item()
    print(f"M={M:6d}: {us:6.2f} us/launch  {M * (H * 4 + H * 2 + 8) / us / 1e3:7.0f} GB/s  max err {err:.3e}")
