"""reverse-step kernel alone at the cfg-3 size (256 graphs x 512 residues = 31.5 MB of algorithmic traffic per launch)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seqdiff_b200 as sd
lib = sd.lib(); dev = "cuda:0"
B, L, T = 256, 512, 50
x = torch.nn.functional.one_hot(torch.randint(0, 20, (B, L), device=dev), 20).float()
lg = torch.randn(B, L, 20, device=dev)
out = torch.empty_like(x)
sched, tr = sd.PredefinedNoiseScheduleDiscrete("cosine", T), sd.BlosumTransition(x_classes=20)
s = torch.full((B, 1), 25.0)
tabsB = sd.utils.step_tables((s + 1) / T, s / T, sched, tr).to(dev)
tabs1 = tabsB[:1].contiguous()
p = sd._cabi.ptr
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
n_iter = int(os.environ.get("ITERS", "50"))
for name, tabs in (("shared table (sampling loop)", tabs1), ("per-graph tables", tabsB)):
    call = lambda: lib.seqdiff_reverse_step(p(tabs), tabs.shape[0], B, L, p(x), p(lg), 1, None, 5, 0, 1, p(out), None, st)
    for _ in range(3): call()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()  # device time: launches captured in a graph (eager launches from Python are CPU-bound at this size)
    with torch.cuda.graph(gr):
        st2 = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(n_iter):
            rc = lib.seqdiff_reverse_step(p(tabs), tabs.shape[0], B, L, p(x), p(lg), 1, None, 5, 0, 1, p(out), None, st2)
            assert rc == 0
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n_iter * 1e3
    print(f"{name}: {us:.1f} us/launch  {B*L*240/us/1e3:.0f} GB/s algorithmic ({B*L} residues x 240 B)")
