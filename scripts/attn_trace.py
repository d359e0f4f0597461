"""Role timelines of CTA 0 of the pipelined attention kernel (debug): prints per-event deltas in SM cycles."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seqdiff_b200 as sd
lib = sd.lib(); dev = "cuda:0"
heads, H = 12, 768
B, L, P = int(os.environ.get("TB", 64)), int(os.environ.get("TL", 128)), int(os.environ.get("TL", 128))
rel = int(os.environ.get("TREL", 1))
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(B * L, 3 * H, device=dev).bfloat16()
E = (torch.randn(2 * P - 1, 64, device=dev) * 0.5).bfloat16() if rel else None
mask = torch.ones(B, L, device=dev); out = torch.empty(B * L, H, device=dev, dtype=torch.bfloat16)
def call():
    rc = lib.seqdiff_op_attention(1, B, heads, L, L, p(qkv), 3 * H, p(qkv[:, H:]), 3 * H, p(qkv[:, 2 * H:]), 3 * H, p(E), P, p(mask), p(out), stream)
    assert rc == 0, lib.seqdiff_last_error()
for _ in range(3): call()
buf = torch.zeros(4 * 1024, dtype=torch.int64, device=dev)
lib.seqdiff_debug_attn_trace(p(buf)); call(); torch.cuda.synchronize(); lib.seqdiff_debug_attn_trace(None)
t = buf.cpu().tolist()
names = {1: "tma ke_empty ok", 2: "tma v_empty ok", 10: "mma issue_s enter", 11: "mma s_empty ok", 12: "mma S issued", 13: "mma p_full ok", 14: "mma v_full ok",
         15: "mma PV issued", 20: "sm step begin", 21: "sm s_full ok", 22: "sm scores done", 23: "sm pre-bar", 24: "sm post-bar", 25: "sm o_full ok", 26: "sm rescale done", 27: "sm P done", 28: "sm epilogue done"}
ev = []
for role in range(4):
    n = t[role * 1024]
    for x in t[role * 1024 + 1: role * 1024 + 1 + n]:
        ev.append((x >> 8, role, x & 255))
ev.sort()
t0 = ev[0][0]
last = {}
lim = int(os.environ.get("TLIM", 260))
for c, role, e in ev[:lim]:
    d = c - last.get(role, c); last[role] = c
    print(f"{c - t0:8d}  {'   ' * role * 6}r{role} {names.get(e, e)} (+{d})")
